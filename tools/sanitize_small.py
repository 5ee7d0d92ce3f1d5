"""A small pass over every kernel, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py` on a GPU box."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import numpy as np
import torch
from reasoning_image_generation_b200 import Renderer
from reasoning_image_generation_b200.compose import GridCompositor, GridLayout
from reasoning_image_generation_b200.dedup import DedupFilter
from reasoning_image_generation_b200.figure import FigureRenderer
from reasoning_image_generation_b200.png import PngEncoder
from reasoning_image_generation_b200.synth import synthetic_scene
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
from figure_stress_scene import stress_scene

# figures: device geometry (thread and warp routes), k_figure, PNG (RGBA) through the pipelined loop
sc = synthetic_figure_scene(96, 512, 512, seed=5, kind='single')
r = FigureRenderer(512, 512, max_images=32)
enc = PngEncoder(512, 512, max_images=32, channels=4)
tot = sum(int(off[-1]) for i0, buf, off in r.iter_png(sc, enc, 32))
print('figures -> png bytes', tot, r.stats()['geom_fallbacks'])
mu = synthetic_figure_scene(24, 512, 512, seed=6, kind='multi')
print('multi', int(r.render(mu).sum()))
enc.close(); r.close()
st = stress_scene().slice(0, 120)
r = FigureRenderer(st.W, st.H, max_images=128, max_artists=st.n_artists)
print('stress', int(r.render(st).sum()), r.stats()['geom_status'])
r.close()
# sequence panels: k_expand, k_raster (+ fused hash), PNG (RGB), hash + scan, overview compositor
ss = synthetic_scene(56, seed=3)
rr = Renderer(512, 512, max_panels=64, max_elements=max(ss.n_elements, 1))
out = torch.empty((56, 512, 512, 3), dtype=torch.uint8, device='cuda')
rr.upload(ss, 'rgb')
rr.render_resident(out.data_ptr()); rr.sync()
f = DedupFilter(rr)
h = f.hash_images(out, 'rgb')
drop, hits = f.scan(h, 10)
print('panels', int(out.sum().item()), int(hits.item()))
e3 = PngEncoder(512, 512, max_images=64)
buf, off = e3.encode(out.data_ptr(), 56, 'rgb')
print('panel png bytes', int(off[-1]))
GL = GridLayout(512, 512, n_states=3, num_options=4)
gc = GridCompositor(GL, device=0, max_samples=8, channel_order='rgb')
g = gc.compose(out.data_ptr(), np.arange(56, dtype=np.int32).reshape(8, 7))
print('grids', int(g.sum()))
gc.close(); e3.close(); rr.close()
print('ok')
