"""The Agg path against a REAL matplotlib.

Two ways in, both skipped in this image (no matplotlib, no fixture: SURVEY H3) -- until one of them runs, the Agg oracle is
UNPINNED and every figure number in bench.py is "oracle-exact, matplotlib parity unverified":
  * tests/golden/figure_mpl.npz, written by `python tests/golden/make_figure_golden.py` on any machine with matplotlib: pixels of
    ~35 figures rendered by matplotlib's Agg backend + the draw_path argument lists read back from them.  oracle/aggdraw.c must
    reproduce the pixels within the north_star tolerance (mean |diff| <= 1/255, >= 99.5 % of pixels within 2 levels); the exact-match
    fraction is printed (the goal is 100 %).  tests/test_gpu_figure_parity.py::test_matplotlib_golden holds the CUDA path to
    the same file.
  * a live matplotlib: the walker (figure_mpl.extract_figure) and the oracle against fig.canvas.buffer_rgba().
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'figure_mpl.npz')


def check_against_matplotlib(got_rgb, want_rgba, what):
    want = want_rgba[..., :3].astype(np.int16)
    assert (want_rgba[..., 3] == 255).all(), 'savefig canvases are opaque'
    diff = np.abs(got_rgb.astype(np.int16) - want)
    per_image_mean = diff.reshape(len(diff), -1).mean(axis=1)
    within2 = (diff <= 2).reshape(len(diff), -1).mean(axis=1)
    exact = float((diff == 0).all(axis=-1).mean())
    print(f'{what}: exact-match pixels {100 * exact:.4f} %, worst image mean |diff| {per_image_mean.max():.4f}, worst within-2 {100 * within2.min():.3f} %')
    assert per_image_mean.max() <= 1.0 and within2.min() >= 0.995            # levels of 0..255: 1/255 of full scale


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason='tests/golden/figure_mpl.npz not generated (needs matplotlib: tests/golden/make_figure_golden.py)')
def test_oracle_against_matplotlib_golden(built):
    from oracle.agg_oracle import AggOracle
    from reasoning_image_generation_b200.figure import scene_from_arrays
    d = np.load(GOLDEN)
    print('fixture rendered by matplotlib', str(d['matplotlib_version']))
    scene = scene_from_arrays(d, 'scene_')
    got = AggOracle().render(scene, scene.W, scene.H)
    check_against_matplotlib(got, d['pixels'], 'oracle/aggdraw.c vs matplotlib ' + str(d['matplotlib_version']))


def test_extracted_figure_renders_like_savefig(built):
    mpl = pytest.importorskip('matplotlib')
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from make_figure_golden import build_figures
    from oracle.agg_oracle import AggOracle
    from reasoning_image_generation_b200.figure import FigureSceneBuilder
    from reasoning_image_generation_b200.figure_mpl import extract_figure
    print('matplotlib', mpl.__version__)
    figs = build_figures()
    sb = FigureSceneBuilder(512, 512)
    want = []
    for fig in figs:
        fig.canvas.draw()
        want.append(np.asarray(fig.canvas.buffer_rgba()).copy())
        sb.add_image(extract_figure(fig))
    got = AggOracle().render(sb.finish(), 512, 512)
    check_against_matplotlib(got, np.stack(want), 'oracle/aggdraw.c vs live matplotlib ' + mpl.__version__)
