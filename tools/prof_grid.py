"""Minimal compositor loop for ncu: python tools/prof_grid.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reasoning_image_generation_b200.compose import GridCompositor, GridLayout
n = 7 * 1024
out = torch.randint(0, 256, (n, 512, 512, 3), dtype=torch.uint8, device='cuda')
GL = GridLayout(512, 512, n_states=3, num_options=4)
ns = n // 7
gc = GridCompositor(GL, device=0, max_samples=ns, channel_order='rgb')
gout = torch.empty((ns, GL.grid_h, GL.grid_w, 3), dtype=torch.uint8, device='cuda')
idx = np.arange(ns * 7, dtype=np.int32).reshape(ns, 7)
for _ in range(3):
    gc.compose_to_device(out.data_ptr(), idx, gout.data_ptr())
gc.sync()
print('ok')
