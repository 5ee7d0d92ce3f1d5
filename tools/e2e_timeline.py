"""GPU timeline of FigureRenderer.iter_png (RIG_TIMELINE=1): busy time per kernel, gaps, overlap.
usage: RIG_TIMELINE=1 python tools/e2e_timeline.py [channels]  -> gpurun_out/e2e_timeline.csv + a summary"""
import collections
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reasoning_image_generation_b200 import _abi
from reasoning_image_generation_b200.figure import FigureRenderer
from reasoning_image_generation_b200.png import PngEncoder
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene

n, chunk = 32768, 4096
ch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sc = synthetic_figure_scene(n, 512, 512, seed=1, kind='single')
r = FigureRenderer(512, 512, max_images=chunk)
enc = PngEncoder(512, 512, max_images=chunk, channels=ch)
for rep in range(3):
    for i0, buf, off in r.iter_png(sc, enc, chunk):
        pass
os.makedirs('gpurun_out', exist_ok=True)
path = 'gpurun_out/e2e_timeline.csv'
lib = _abi.load()
lib.rig_timeline_dump.argtypes = [ctypes.c_char_p]
assert lib.rig_timeline_dump(path.encode()) == 0
rows = [l.strip().split(',') for l in open(path)][1:]
rows = [(a, b, float(c), float(d)) for a, b, c, d in rows]
# last repetition only: the final 8 k_figure launches
figs = [x for x in rows if x[0] == 'k_figure']
t0 = figs[-8][2] - 1.0
rows = [x for x in rows if x[2] >= t0]
t_end = max(x[3] for x in rows)
busy = collections.defaultdict(float)
for name, st, a, b in rows:
    busy[name] += b - a
print('window %.2f ms for %d chunks (%.2f ms per chunk)' % (t_end - t0, 8, (t_end - t0) / 8))
for k, v in sorted(busy.items(), key=lambda kv: -kv[1]):
    print('  %-16s %.2f ms total, %.3f per chunk' % (k, v, v / 8))
# union of kernel intervals (not the copies)
iv = sorted((a, b) for name, st, a, b in rows if name.startswith('k_'))
u, cur_a, cur_b = 0.0, None, None
for a, b in iv:
    if cur_b is None or a > cur_b:
        if cur_b is not None:
            u += cur_b - cur_a
        cur_a, cur_b = a, b
    else:
        cur_b = max(cur_b, b)
u += cur_b - cur_a
print('  some kernel running: %.2f ms of %.2f (%.0f %%)' % (u, t_end - t0, 100 * u / (t_end - t0)))
for name, st, a, b in rows[:40]:
    print('   %-16s %9.3f -> %9.3f  (%.3f)' % (name, a - t0, b - t0, b - a))
