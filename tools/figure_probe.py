"""GPU probe of the figure path: resident render rate of synthetic config-2 / config-3 scenes (CUDA events inside the ABI)."""
import sys
import time

import torch

sys.path.insert(0, '.')
from reasoning_image_generation_b200.figure import FigureRenderer  # noqa: E402
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene, tile_scene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
kind = sys.argv[2] if len(sys.argv) > 2 else 'single'
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
distinct = min(n, 2048)
t = time.time()
base = synthetic_figure_scene(distinct, 512, 512, seed=1, kind=kind)
sc = tile_scene(base, n // distinct)
t_scene = time.time() - t
r = FigureRenderer(512, 512, max_images=sc.n_images)
t = time.time()
r.upload(sc)
t_up = time.time() - t
out = torch.empty((sc.n_images, 512, 512, 3), dtype=torch.uint8, device='cuda:0')
for _ in range(3):
    r.render_resident(out.data_ptr())
r.sync()
r.set_profiling(True)
for _ in range(reps):
    r.render_resident(out.data_ptr())
r.sync()
st = r.stats()
ms = st['cover_ms'] / st['timed_launches']
print(f'{kind}: n={sc.n_images} scene {t_scene:.1f}s upload {t_up:.2f}s (prepare {st["prepare_ms"]:.0f} ms) edges/img {st["edges"] / sc.n_images:.0f} '
      f'layers/img {st["layers"] / sc.n_images:.1f} k_figure {ms:.3f} ms  {sc.n_images / ms / 1e3:.3f} M img/s  {sc.n_images * 786432 / ms / 1e6:.0f} GB/s')
