"""Development probe (GPU box): smoke + parity on a synthetic batch + first timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as ge
from reasoning_image_generation_b200 import Renderer
from reasoning_image_generation_b200.synth import synthetic_scene, scene_to_elements
from oracle import ref_render

ge.smoke()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sc = synthetic_scene(n, seed=1)
print('scene', sc.n_panels, 'panels', sc.n_elements, 'elements')
r = Renderer(512, 512, device=0, max_panels=n, max_elements=max(sc.n_elements, 1))
out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
r.upload(sc, 'bgr')
r.render_resident(out.data_ptr()); r.sync()
print('stats', r.stats())
# parity on a subset
be = ref_render.CBackend(); bad = 0; checked = 0
host = out[:256].cpu().numpy()
for p in range(256):
    els, ug = scene_to_elements(sc, p)
    ref = ref_render.render_panel(512, 512, (255, 255, 255), els, ug, 3, backend=be)
    d = (ref != host[p]).any(axis=2).sum(); checked += 1
    if d:
        bad += 1
        if bad < 6: print('panel', p, 'differs in', d, 'px', [(e['kind'], e['size'], e['fill'], e['stroke_width'], e['center'], e['angle']) for e in els], ug)
print('parity: checked', checked, 'bad', bad)
r.set_profiling(True)
ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream; assert st != 0
for rep in range(3):
    r.render_resident(out.data_ptr(), stream=st)
torch.cuda.synchronize()
ev0.record()
K = 5
for rep in range(K):
    r.render_resident(out.data_ptr(), stream=st)
ev1.record(); torch.cuda.synchronize(); r.sync()
ms = ev0.elapsed_time(ev1) / K
print(f'resident render: {ms:.3f} ms per {n} panels -> {n/ms*1e3:.0f} img/s, {n*786432/ms/1e6:.1f} GB/s')
t0 = time.time()
for rep in range(K): r.render_resident(out.data_ptr(), stream=st)
torch.cuda.synchronize(); wall = (time.time() - t0) / K * 1e3
print(f'  wall-clock cross-check: {wall:.3f} ms')
r.sync(); stt = r.stats(); nl = max(stt['timed_launches'], 1)
print(f"  split per render: expand {stt['expand_ms']/ (K*2+3):.3f} ms, raster {stt['raster_ms']/(K*2+3):.3f} ms over {stt['timed_launches']} chunk launches")
r.set_profiling(False)
# empty scene (pure template stores)
from reasoning_image_generation_b200.scene import Scene
emp = Scene(512, 512, (255,255,255), 3, np.zeros(n+1, np.uint32), np.zeros(n, np.uint8), *[np.zeros(0, t) for t in (np.uint8, np.uint8, np.uint8, np.int32, np.int32, np.int32, np.float64, np.float64, np.uint8)])
r.upload(emp, 'bgr')
for rep in range(3): r.render_resident(out.data_ptr(), stream=st)
torch.cuda.synchronize(); ev0.record()
for rep in range(K): r.render_resident(out.data_ptr(), stream=st)
ev1.record(); torch.cuda.synchronize(); r.sync()
ms = ev0.elapsed_time(ev1) / K
print(f'empty scene: {ms:.3f} ms -> {n/ms*1e3:.0f} img/s, {n*786432/ms/1e6:.1f} GB/s')
# e2e host path
sub = sc.slice(0, min(n, 2048))
pin = r.pinned_output(sub.n_panels)
t = time.time(); r.render(sub, 'bgr', out=pin); dt = time.time() - t
t = time.time(); r.render(sub, 'bgr', out=pin); dt = time.time() - t
print(f'e2e host (pinned out) {sub.n_panels} panels: {dt*1e3:.1f} ms -> {sub.n_panels/dt:.0f} img/s')
# fused hash variant
r.upload(sc, 'bgr')
hs = torch.empty((n, 4), dtype=torch.int64, device='cuda')
for rep in range(3): r.render_resident_hashed(out.data_ptr(), hs.data_ptr(), stream=st)
torch.cuda.synchronize(); ev0.record()
for rep in range(K): r.render_resident_hashed(out.data_ptr(), hs.data_ptr(), stream=st)
ev1.record(); torch.cuda.synchronize(); r.sync()
ms = ev0.elapsed_time(ev1) / K
print(f'render + fused hash: {ms:.3f} ms per {n} panels -> {n/ms*1e3:.0f} img/s')
from reasoning_image_generation_b200.dedup import DedupFilter
f = DedupFilter(r)
h2 = f.hash_images(out, 'bgr'); torch.cuda.synchronize(); r.sync()
ev0.record(); 
for rep in range(K): h2 = f.hash_images(out, 'bgr')
ev1.record(); torch.cuda.synchronize(); r.sync()
print(f'standalone hash: {ev0.elapsed_time(ev1)/K:.3f} ms per {n} panels; equal to fused: {bool(torch.equal(h2, hs))}')
