"""k_figure_path probe (one GPU): python tools/geom_probe.py [n_images] [kind] [reps] -> per-launch times of the geometry and coverage kernels."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reasoning_image_generation_b200.figure import FigureRenderer
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
kind = sys.argv[2] if len(sys.argv) > 2 else 'single'
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sc = synthetic_figure_scene(n, 512, 512, seed=1000, kind=kind)
r = FigureRenderer(512, 512, device=0, max_images=n, max_artists=max(sc.n_artists, 1))
out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
res = {'n': n, 'kind': kind, 'artists': sc.n_artists, 'vertices': sc.n_vertices}
for mode in ('device', 'host'):
    r.set_geometry(mode)
    t = time.perf_counter()
    r.upload(sc)
    r.sync()
    res['upload_%s_ms' % mode] = (time.perf_counter() - t) * 1e3
    t = time.perf_counter()
    r.upload(sc)
    r.sync()
    res['upload_%s_ms_2nd' % mode] = (time.perf_counter() - t) * 1e3
    res['prepare_%s_ms' % mode] = r.stats()['prepare_ms']
r.set_geometry('device')
r.upload(sc)
st = r.stats()
res.update(edges=st['edges'], layers=st['layers'], geom_status=st['geom_status'], geom_resident=st['geom_resident'])
for _ in range(3):
    r.draw_resident(out.data_ptr())
r.sync()
r.set_profiling(True)
for _ in range(reps):
    r.draw_resident(out.data_ptr())
r.sync()
st = r.stats()
res['geom_ms'] = st['geom_ms'] / max(1, st['geom_launches'])
res['cover_ms'] = st['cover_ms'] / max(1, st['timed_launches'])
res['images_per_s_draw'] = n / ((res['geom_ms'] + res['cover_ms']) / 1e3)
res['hbm_frac_draw'] = res['images_per_s_draw'] * 786432 / 1e9 / 6545.9
print(json.dumps(res, indent=1))
