"""Minimal resident-render loop for ncu (GPU box): python tools/prof_run.py N_PANELS REPS [seed]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reasoning_image_generation_b200 import Renderer
from reasoning_image_generation_b200.synth import synthetic_scene
n = int(sys.argv[1]); reps = int(sys.argv[2]); seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sc = synthetic_scene(n, seed=seed)
r = Renderer(512, 512, device=0, max_panels=n, max_elements=max(sc.n_elements, 1))
out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
r.upload(sc, 'bgr')
for _ in range(reps):
    r.render_resident(out.data_ptr())
r.sync()
print('ok', r.stats())
