"""GPU probe: device PNG encode rate and the savefig-style end-to-end rate (host artists in -> PNG files in host memory)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from reasoning_image_generation_b200.figure import FigureRenderer  # noqa: E402
from reasoning_image_generation_b200.png import PngEncoder  # noqa: E402
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sc = synthetic_figure_scene(n, 512, 512, seed=1, kind='single')
r = FigureRenderer(512, 512, max_images=chunk)
enc = PngEncoder(512, 512, max_images=chunk, channels=4 if len(sys.argv) > 4 else 3)
for _ in range(2):
    for i0, buf, off in r.iter_png(sc, enc, chunk, overlap=True):
        pass
enc.set_profiling(True)
t = time.perf_counter()
reps, total = 4, 0
for _ in range(reps):
    for i0, buf, off in r.iter_png(sc, enc, chunk, overlap=True):
        total += int(off[-1])
dt = time.perf_counter() - t
st = enc.stats()
print(f'iter_png: {n * reps / dt / 1e3:.1f} k img/s end to end, {total / (n * reps) / 1e3:.1f} KB/file, '
      f'encode kernels {st["encode_ms"] / reps:.2f} ms per {n} images = {st["timed_images"] / st["encode_ms"] / 1e3:.2f} M img/s, '
      f'host prepare {r.stats()["prepare_ms"]:.2f} ms per chunk')
