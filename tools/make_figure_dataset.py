"""Writes a synthetic multigraph data set to disk with the GPU path, in the reference's output layout
(multigraph_generation/cli.py:22-23: output/images/{i}_{mode}.png, output/params/{i}_{mode}.json):

    python tools/make_figure_dataset.py --out /tmp/figs --n 4096 [--kind single|multi] [--size 512] [--device 0]

Scenes come from synth_figure (the reference's own generator needs matplotlib + shapely); rendering, PNG encoding and the
file writes are the product path: FigureRenderer.iter_png -> png.write_files."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reasoning_image_generation_b200.figure import FigureRenderer  # noqa: E402
from reasoning_image_generation_b200.png import PngEncoder, write_files  # noqa: E402
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', required=True)
    ap.add_argument('--n', type=int, default=1024)
    ap.add_argument('--kind', default='single', choices=['single', 'multi'])
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--device', type=int, default=0)
    ap.add_argument('--chunk', type=int, default=4096)
    a = ap.parse_args()
    mode = 'random' if a.kind == 'single' else 'multi'
    os.makedirs(os.path.join(a.out, 'images'), exist_ok=True)
    os.makedirs(os.path.join(a.out, 'params'), exist_ok=True)
    t0 = time.time()
    scene = synthetic_figure_scene(a.n, a.size, a.size, seed=a.seed, kind=a.kind)
    t1 = time.time()
    chunk = min(a.chunk, a.n)
    r = FigureRenderer(a.size, a.size, device=a.device, max_images=chunk)
    enc = PngEncoder(a.size, a.size, device=a.device, max_images=chunk, channels=4)      # RGBA, as fig.savefig writes
    total = 0
    for i0, buf, off in r.iter_png(scene, enc, chunk):
        n = len(off) - 1
        write_files(buf, off, [os.path.join(a.out, 'images', f'{i0 + k}_{mode}.png') for k in range(n)])
        for k in range(n):
            i = i0 + k
            a0, a1 = int(scene.image_artist_begin[i]), int(scene.image_artist_begin[i + 1])
            with open(os.path.join(a.out, 'params', f'{i}_{mode}.json'), 'w') as f:
                json.dump({'seed': a.seed * 1000003 + i, 'mode': mode, 'n_artists': a1 - a0,
                           'linewidths_px': [float(x) for x in scene.linewidth[a0:a1]]}, f)
        total += int(off[-1])
    t2 = time.time()
    print(f'{a.n} figures -> {a.out}: scene emission {t1 - t0:.2f} s, render + encode + write {t2 - t1:.2f} s '
          f'({a.n / (t2 - t1):.0f} files/s), {total / a.n / 1e3:.1f} KB per PNG')


if __name__ == '__main__':
    main()
