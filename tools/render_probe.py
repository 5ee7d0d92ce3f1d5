import sys, time
sys.path.insert(0, '.')
from reasoning_image_generation_b200.figure import FigureRenderer
from reasoning_image_generation_b200.png import PngEncoder
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
n, chunk = 16384, 4096
sc = synthetic_figure_scene(n, 512, 512, seed=1, kind='single')
subs = [sc.slice(i, i + chunk) for i in range(0, n, chunk)]
r = FigureRenderer(512, 512, max_images=chunk)
enc = PngEncoder(512, 512, max_images=chunk)
base = enc.staging(2 * chunk)
for k in range(8):
    t0 = time.perf_counter()
    r.render_to_device(subs[k % 4], base, 'rgb', 0)
    t1 = time.perf_counter()
    r.sync()
    t2 = time.perf_counter()
    print('render_to_device %.3f ms  sync %.3f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr)
