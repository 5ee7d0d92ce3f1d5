"""Tuning sweep (GPU box): resident render time for scheduling knobs; one process per configuration."""
import os, subprocess, sys, itertools
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, torch
sys.path.insert(0, %r)
from reasoning_image_generation_b200 import Renderer
from reasoning_image_generation_b200.synth import synthetic_scene
n = 16384
sc = synthetic_scene(n, seed=1)
r = Renderer(512, 512, device=0, max_panels=n, max_elements=sc.n_elements)
out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
r.upload(sc, 'bgr')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
for _ in range(3): r.render_resident(out.data_ptr(), stream=st)
torch.cuda.synchronize(); r.sync(); r.set_profiling(True)
for _ in range(5): r.render_resident(out.data_ptr(), stream=st)
torch.cuda.synchronize(); r.sync(); s = r.stats()
print("expand %%.3f raster %%.3f ms" %% (s['expand_ms']/5, s['raster_ms']/5))
''' % root
for lib, w in (('o6', 22), ('o8', 16)):
    env = dict(os.environ, RIG_LIB=os.path.join(root, 'reasoning_image_generation_b200', 'librig_%s.so' % lib), RIG_WARPS=str(w))
    o = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    om = lib; hd = w
    print('order mode', om, 'heavy_div', hd, '->', o.stdout.strip() or o.stderr[-300:], flush=True)
