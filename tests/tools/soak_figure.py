"""Soak test of the figure path on a GPU box: many synthetic and random scenes, k_figure (through the C ABI) against
oracle/aggdraw.c, plus PNG round trips of a subset.  Prints one line per batch; exits non-zero on the first differing image.
    python tests/tools/soak_figure.py [images_per_kind]"""
import os
import sys
from multiprocessing import Pool

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from reasoning_image_generation_b200 import figure_paths as fp  # noqa: E402
from reasoning_image_generation_b200.figure import FigureRenderer, FigureSceneBuilder  # noqa: E402
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene  # noqa: E402


def oracle_chunk(args):
    from oracle.agg_oracle import AggOracle
    sc, i0, i1, W, H = args
    o = AggOracle()
    return np.stack([o.render_rgba(sc, i, W, H)[:, :, :3] for i in range(i0, i1)])


def check(name, sc, W, H, pool, cores):
    r = FigureRenderer(W, H, max_images=sc.n_images, max_artists=max(1, sc.n_artists))
    got = r.render(sc)
    r.close()
    step = (sc.n_images + cores - 1) // cores
    parts = pool.map(oracle_chunk, [(sc, i, min(sc.n_images, i + step), W, H) for i in range(0, sc.n_images, step)])
    ref = np.concatenate(parts)
    bad = np.nonzero((got != ref).any(axis=(1, 2, 3)))[0]
    print(f'{name:28s} {sc.n_images:6d} images {W}x{H}: {len(bad)} differ', flush=True)
    if len(bad):
        i = int(bad[0])
        d = (got[i] != ref[i]).any(axis=2)
        ys, xs = np.nonzero(d)
        print(f'  first: image {i}, {int(d.sum())} pixels, rows {ys.min()}..{ys.max()}, cols {xs.min()}..{xs.max()}')
        sys.exit(1)


def fuzz_scene(n, W, H, seed):
    rng = np.random.RandomState(seed)
    sb = FigureSceneBuilder(W, H)
    caps, joins = ('butt', 'projecting', 'round'), ('miter', 'round', 'bevel')
    for i in range(n):
        for _ in range(rng.randint(1, 7)):
            k = rng.randint(2, 9)
            pts = rng.uniform(-0.1 * W, 1.1 * W, size=(k, 2)) * [1.0, H / W]
            closed = bool(rng.randint(2)) and k >= 3
            v, c = fp.polyline(pts, closed)
            if rng.randint(3) == 0 and k >= 4:
                c = c.copy()
                c[1:4] = 4
            elif rng.randint(4) == 0 and k >= 3:
                c = c.copy()
                c[1:3] = 3
            col = tuple(rng.uniform(0, 1, 3)) + (float(rng.choice([1.0, 0.9, 0.5, 0.1])),)
            face = tuple(rng.uniform(0, 1, 4)) if rng.randint(3) == 0 else None
            clip = None
            if rng.randint(5) == 0:
                x0, y0 = rng.uniform(0, W / 2), rng.uniform(0, H / 2)
                clip = (x0, y0, x0 + rng.uniform(1, W / 2), y0 + rng.uniform(1, H / 2))
            sb.add_artist(vertices=v, codes=c, linewidth=float(rng.choice([0.0, 0.3, 1.0, 1.37, 2.5, 9.0, 31.0])), cap=caps[rng.randint(3)],
                          join=joins[rng.randint(3)], edgecolor=col, facecolor=face, snap=[None, True, False][rng.randint(3)], clip=clip)
        sb.end_image()
    return sb.finish()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    cores = os.cpu_count() or 1
    from oracle.agg_oracle import build_agg_oracle
    build_agg_oracle()
    with Pool(cores) as pool:
        for kind in ('single', 'multi'):
            check(f'synthetic {kind}', synthetic_figure_scene(n, 512, 512, seed=4242, kind=kind), 512, 512, pool, cores)
        check('synthetic single 1600', synthetic_figure_scene(max(8, n // 64), 1600, 1600, seed=7, kind='single'), 1600, 1600, pool, cores)
        for (W, H, s) in ((512, 512, 1), (320, 200, 2), (1024, 96, 3), (64, 64, 4)):
            check(f'fuzz seed {s}', fuzz_scene(max(64, n // 8), W, H, s), W, H, pool, cores)
    print('soak: all equal')


if __name__ == '__main__':
    main()
