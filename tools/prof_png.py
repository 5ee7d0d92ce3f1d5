"""One RGBA encode of 4096 rendered figures, for `ncu -k regex:k_png_pieces` captures."""
import sys
sys.path.insert(0, '.')
import torch
from reasoning_image_generation_b200.figure import FigureRenderer
from reasoning_image_generation_b200.png import PngEncoder
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
n = 4096
sc = synthetic_figure_scene(n, 512, 512, seed=1, kind='single')
r = FigureRenderer(512, 512, max_images=n)
fig = torch.empty(n * 512 * 512 * 3, dtype=torch.uint8, device='cuda')
r.set_blank_flags(True)
r.render_to_device(sc, fig.data_ptr()); r.sync()
enc = PngEncoder(512, 512, max_images=n, channels=4)
for _ in range(3):
    enc.set_blank_hint(*r.blank_bands())
    enc.encode(fig.data_ptr(), n)
print('ok')
