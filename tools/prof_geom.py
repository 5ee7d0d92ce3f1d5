"""Minimal loop for ncu: upload (device geometry) + REPS draw_resident of N synthetic figures.  python tools/prof_geom.py N REPS [kind]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reasoning_image_generation_b200.figure import FigureRenderer
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
n, reps = int(sys.argv[1]), int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else 'single'
sc = synthetic_figure_scene(n, 512, 512, seed=1000, kind=kind)
r = FigureRenderer(512, 512, device=0, max_images=n, max_artists=max(sc.n_artists, 1))
out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
r.upload(sc)
for _ in range(reps):
    r.draw_resident(out.data_ptr())
r.sync()
print('ok', r.stats())
