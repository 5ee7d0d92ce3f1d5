"""The matplotlib-side walker (reasoning_image_generation_b200/figure_mpl.py) against a real matplotlib: skipped in images
without matplotlib (this one: SURVEY H3).  Where matplotlib exists it is also the test that would PIN oracle/aggdraw.c."""
import numpy as np
import pytest

mpl = pytest.importorskip('matplotlib')


def test_extracted_figure_renders_like_savefig(built):
    mpl.use('Agg')
    import matplotlib.pyplot as plt
    from matplotlib.patches import Circle, Rectangle
    from oracle.agg_oracle import AggOracle
    from reasoning_image_generation_b200.figure import FigureSceneBuilder
    from reasoning_image_generation_b200.figure_mpl import extract_figure
    fig, ax = plt.subplots(figsize=(8, 8), dpi=64)
    ax.set_xlim(-5, 5); ax.set_ylim(-5, 5); ax.set_aspect('equal'); ax.axis('off')
    ax.add_patch(Circle((0, 0), 1.95, facecolor='none', edgecolor='black', linewidth=1.7, alpha=0.9))
    ax.add_patch(Rectangle((-1, -2), 2.5, 1.2, facecolor='none', edgecolor='black', linewidth=1.5, angle=30))
    ax.plot([-1, 1.5], [0.3, -0.8], linewidth=1.0, alpha=0.85)
    fig.canvas.draw()
    want = np.asarray(fig.canvas.buffer_rgba())[:, :, :3]
    sb = FigureSceneBuilder(512, 512)
    sb.add_image(extract_figure(fig))
    got = AggOracle().render(sb.finish(), 512, 512)[0]
    diff = np.abs(got.astype(int) - want.astype(int))
    assert diff.mean() <= 1.0 and (diff <= 2).mean() >= 0.995
