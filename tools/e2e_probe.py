"""Where the end-to-end time of FigureRenderer.iter_png goes: per-chunk wall time of the render side (host path pipeline +
upload + k_figure) and of the encode side (PNG kernels + D2H), alone and overlapped."""
import sys
import threading
import time

sys.path.insert(0, '.')
from reasoning_image_generation_b200.figure import FigureRenderer  # noqa: E402
from reasoning_image_generation_b200.png import PngEncoder  # noqa: E402
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene  # noqa: E402

n, chunk = 32768, 4096
sc = synthetic_figure_scene(n, 512, 512, seed=1, kind='single')
subs = [sc.slice(i, i + chunk) for i in range(0, n, chunk)]
r = FigureRenderer(512, 512, max_images=chunk)
enc = PngEncoder(512, 512, max_images=chunk)
img = 512 * 512 * 3
base = enc.staging(2 * chunk)


def render(k):
    t = time.perf_counter()
    r.render_to_device(subs[k % len(subs)], base + (k & 1) * chunk * img, 'rgb', 0)
    r.sync()
    return time.perf_counter() - t


def encode(k):
    t = time.perf_counter()
    enc.encode(base + (k & 1) * chunk * img, chunk, 'rgb', 0)
    return time.perf_counter() - t


for k in range(4):
    render(k); encode(k)
tr = [render(k) for k in range(8)]
te = [encode(k) for k in range(8)]
print('alone: render %.2f ms  encode %.2f ms per %d images' % (1e3 * sum(tr) / 8, 1e3 * sum(te) / 8, chunk))
res = {}


def worker(k):
    res[k] = render(k)


t0 = time.perf_counter()
render(0)
tes, trs = [], []
for k in range(16):
    w = threading.Thread(target=worker, args=(k + 1,))
    w.start()
    tes.append(encode(k))
    w.join()
    trs.append(res[k + 1])
dt = time.perf_counter() - t0
print('overlapped: render %.2f ms  encode %.2f ms  per-chunk wall %.2f ms -> %.0f k img/s' % (1e3 * sum(trs) / 16, 1e3 * sum(tes) / 16, 1e3 * dt / 16, 16 * chunk / dt / 1e3))
st = r.stats()
print('host prepare %.2f ms' % st['prepare_ms'])

# the product loop itself (iter_png: worker thread + two encoder slots)
real_blank = r.blank_bands
for ch, hint in ((3, True), (4, True), (4, False)):
    r.blank_bands = real_blank if hint else (lambda: (0, 2))
    e2 = PngEncoder(512, 512, max_images=chunk, channels=ch)
    for rep in range(3):
        t0 = time.perf_counter()
        tot = 0
        for i0, buf, off in r.iter_png(sc, e2, chunk=chunk):
            tot += int(off[-1])
        dt = time.perf_counter() - t0
    print('iter_png channels=%d hint=%s: %.0f k img/s (%.1f KB/file)' % (ch, hint, n / dt / 1e3, tot / n / 1e3))
    e2.close()
