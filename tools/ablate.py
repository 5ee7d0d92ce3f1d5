"""Ablation timing (GPU box): raster time with record types skipped (results are wrong on purpose)."""
import os, subprocess, sys
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
names = {1: 'LINE8', 2: 'LINEAA', 3: 'DISC', 4: 'CRESCENT', 5: 'CONVEX', 6: 'POLYFILL', 8: 'MINIPOLY', 9: 'STAMP'}
masks = [('none', 0), ('all records', 0xffffffff)] + [('skip ' + v, 1 << k) for k, v in names.items()] + [('only MINIPOLY+STAMP', 0xffffffff & ~((1 << 8) | (1 << 9))), ('only LINEAA', 0xffffffff & ~(1 << 2))]
for name, m in masks:
    env = dict(os.environ, RIG_SKIP_MASK=hex(m))
    o = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(f'{name:24s}', o.stdout.strip() or o.stderr[-300:], flush=True)
