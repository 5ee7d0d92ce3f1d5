"""k_figure launch time against the number of images per launch (resident scenes)."""
import sys
sys.path.insert(0, '.')
import torch
from reasoning_image_generation_b200.figure import FigureRenderer
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
big = synthetic_figure_scene(32768, 512, 512, seed=1, kind='single')
out = torch.empty(32768 * 512 * 512 * 3, dtype=torch.uint8, device='cuda')
r = FigureRenderer(512, 512, max_images=32768, max_artists=1 << 20)
import os
sizes = [int(x) for x in os.environ['SIZES'].split(',')] if os.environ.get('SIZES') else None
for n in (sizes or ((32768,) if os.environ.get('ONLY_BIG') else (1024, 2048, 4096, 8192, 16384, 32768))):
    r.upload(big.slice(0, n) if n < 32768 else big)
    for _ in range(5):
        r.render_resident(out.data_ptr())
    r.sync()
    r.set_profiling(True)
    for _ in range(10):
        r.render_resident(out.data_ptr())
    r.sync()
    st = r.stats()
    r.set_profiling(False)
    ms = st['cover_ms'] / st['timed_launches']
    print('%6d images: %.3f ms per launch, %.2f M img/s' % (n, ms, n / ms / 1e3))
