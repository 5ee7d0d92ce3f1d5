"""A random scene that exercises every branch of the host path pipeline (all cap / join / snap values, curves, axis-aligned and
degenerate polylines, closed polygons, fills) -- used to pin the pipeline's output by digest (tests/test_figure_host_digest.py)."""
import numpy as np

from reasoning_image_generation_b200 import figure_paths as fp
from reasoning_image_generation_b200.figure import FigureSceneBuilder


def stress_scene():
    W = H = 256
    sb = FigureSceneBuilder(W, H)
    caps, joins = ['butt', 'projecting', 'round'], ['miter', 'round', 'bevel']
    rng = np.random.RandomState(5)
    for img in range(600):
        for a in range(rng.randint(1, 5)):
            kind = rng.randint(0, 6)
            cx, cy, r = rng.uniform(20, 236), rng.uniform(20, 236), rng.uniform(3, 90)
            if kind == 0:
                v, c = fp.unit_circle()
                aff = (r, 0.0, 0.0, r * rng.uniform(0.3, 1.0), cx, cy)
            elif kind == 1:
                v, c = fp.unit_rectangle()
                ang = rng.choice([0.0, 0.0, np.pi / 2, rng.uniform(0, np.pi)])
                aff = (r * np.cos(ang), r * np.sin(ang), -r * np.sin(ang), r * np.cos(ang), float(round(cx)) if rng.rand() < 0.5 else cx, float(round(cy)) if rng.rand() < 0.5 else cy)
            elif kind == 2:
                v, c = fp.unit_regular_polygon(int(rng.randint(3, 9)))
                aff = (r, 0.0, 0.0, r, cx, cy)
            elif kind == 3:     # axis-aligned and degenerate polylines: zero offsets, repeated points, reversals
                n = rng.randint(2, 8)
                pts = np.round(rng.uniform(10, 246, (n, 2)))
                for k in range(1, n):
                    if rng.rand() < 0.5: pts[k, rng.randint(0, 2)] = pts[k - 1, rng.randint(0, 2)] if rng.rand() < 0.2 else pts[k - 1][rng.randint(0, 2)]
                if rng.rand() < 0.3: pts[-1] = pts[-2]
                if rng.rand() < 0.3 and n > 2: pts[2] = pts[0]
                v, c = pts, None
                aff = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
            elif kind == 4:
                v, c = fp.polyline(rng.uniform(0, 256, (rng.randint(2, 12), 2)))
                aff = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
            else:
                v, c = fp.wedge_patch((0.0, 0.0), r, rng.uniform(0, 180), rng.uniform(181, 400))
                aff = (1.0, 0.0, 0.0, 1.0, cx, cy)
            closed = kind in (0, 1, 2, 5)
            if kind == 3 and rng.rand() < 0.4 and len(v) > 2:
                c = np.array([1] + [2] * (len(v) - 2) + [79], np.uint8)
            sb.add_artist(v, c, aff, linewidth=float(rng.choice([0.5, 1.0, 1.5, 2.0, 2.7, 4.0, 9.0])), cap=caps[rng.randint(0, 3)], join=joins[rng.randint(0, 3)],
                          snap=[None, True, False][rng.randint(0, 3)], edgecolor=(0.1, 0.2, 0.3, 0.9),
                          facecolor=(0.5, 0.5, 0.9, 0.8) if (closed and rng.rand() < 0.5) else None)
        sb.end_image()
    return sb.finish()
