"""Renderer.iter_png (sequence panels -> PNG files in host memory) against the chunk size."""
import sys
import time
sys.path.insert(0, '.')
from reasoning_image_generation_b200 import Renderer
from reasoning_image_generation_b200.png import PngEncoder
from reasoning_image_generation_b200.synth import synthetic_scene
n = 16384
sc = synthetic_scene(n, seed=1000)
r = Renderer(512, 512, max_panels=8192, max_elements=max(sc.n_elements, 1))
for chunk in (2048, 4096, 8192):
    enc = PngEncoder(512, 512, max_images=chunk)
    for rep in range(3):
        t0 = time.perf_counter()
        tot = 0
        for p0, buf, off in r.iter_png(sc, enc, chunk, 'rgb'):
            tot += int(off[-1])
        dt = time.perf_counter() - t0
    print('chunk %d: %.0f k img/s (%.1f KB/file)' % (chunk, n / dt / 1e3, tot / n / 1e3))
    enc.close()
