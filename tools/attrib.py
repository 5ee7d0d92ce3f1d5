"""Cost attribution (GPU box): raster/expand time for scenes restricted to one feature class."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reasoning_image_generation_b200 import Renderer
from reasoning_image_generation_b200.synth import synthetic_scene
from reasoning_image_generation_b200.scene import Scene
n = 8192
base = synthetic_scene(n, seed=1)
def variant(kind=None, stroke=None, fill=None, grid=None):
    k = base.kind.copy(); s = base.stroke.copy(); f = base.fill.copy(); g = base.flags.copy()
    if kind is not None: k[:] = np.random.default_rng(0).choice(kind, len(k))
    if stroke is not None: s[:] = stroke
    if fill is not None: f[:] = fill
    if grid is not None: g[:] = grid
    return Scene(512, 512, (255,255,255), 3, base.begin, g, k, f, base.flip, s, base.size, base.angle, base.cx, base.cy, base.rgb)
r = Renderer(512, 512, device=0, max_panels=n, max_elements=base.n_elements)
out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
POLY = [0, 2, 3, 4, 5, 6]; 
cases = [('baseline', {}), ('no grid', dict(grid=0)), ('all grid', dict(grid=1)),
         ('thin nofill', dict(stroke=1, fill=0)), ('thin fill', dict(stroke=1, fill=1)), ('thick3 nofill', dict(stroke=3, fill=0)), ('thick2 nofill', dict(stroke=2, fill=0)),
         ('small polys thick nofill', dict(kind=POLY, stroke=3, fill=0)), ('small polys thin nofill', dict(kind=POLY, stroke=1, fill=0)), ('small polys thin fill', dict(kind=POLY, stroke=1, fill=1)),
         ('circles thick nofill', dict(kind=[1], stroke=3, fill=0)), ('circles thin nofill', dict(kind=[1], stroke=1, fill=0)), ('circles thin fill', dict(kind=[1], stroke=1, fill=1)),
         ('hearts thick nofill', dict(kind=[8], stroke=3, fill=0)), ('hearts thin fill', dict(kind=[8], stroke=1, fill=1)),
         ('crescent thick nofill', dict(kind=[9], stroke=3, fill=0)), ('rounded thick nofill', dict(kind=[10], stroke=3, fill=0)), ('plus thick fill', dict(kind=[7], stroke=3, fill=1))]
for name, kw in cases:
    sc = variant(**kw)
    r.upload(sc, 'bgr')
    for _ in range(2): r.render_resident(out.data_ptr(), stream=st)
    torch.cuda.synchronize(); r.sync(); r.set_profiling(True)
    for _ in range(3): r.render_resident(out.data_ptr(), stream=st)
    torch.cuda.synchronize(); r.sync(); s = r.stats(); r.set_profiling(False)
    print(f"{name:28s} expand {s['expand_ms']/3*1e3/n:7.3f} us/panel  raster {s['raster_ms']/3*1e3/n:7.3f} us/panel  slots/panel {s['records']/n:7.1f}", flush=True)
