"""How the PNG kernels' time splits between blank and drawn rows: all-white images vs rendered figures (resident, kernels only)."""
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
r.render_to_device(sc, fig.data_ptr()); r.sync()
white = torch.full_like(fig, 255)
torch.cuda.synchronize()
for ch in (3, 4):
    enc = PngEncoder(512, 512, max_images=n, channels=ch)
    for name, t in (('white', white), ('figures', fig)):
        for _ in range(3):
            enc.encode(t.data_ptr(), n)
        enc.set_profiling(True)
        for _ in range(8):
            enc.encode(t.data_ptr(), n)
        st = enc.stats()
        enc.set_profiling(False)
        print('channels=%d %-8s %.3f ms per %d images' % (ch, name, st['encode_ms'] / 8, n))
    enc.close()
