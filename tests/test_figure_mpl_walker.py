"""Control flow of the matplotlib-side walker (figure_mpl.extract_figure) against a STAND-IN for matplotlib
(tests/fake_matplotlib): draw order (children in insertion order, stably sorted by zorder), what is read from each artist, and
that everything the walker cannot express raises instead of being dropped.  This executes the walker's code;
it says nothing about matplotlib's own behaviour -- the pin for that is tests/golden/make_figure_golden.py, which needs the
real library.  The stand-in is only put on sys.path inside this module's fixture and removed again."""
import importlib
import os
import sys

import numpy as np
import pytest

FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fake_matplotlib')


def _ours(k):
    return k == 'matplotlib' or k.startswith('matplotlib.') or k.endswith('.figure_mpl')


@pytest.fixture()
def walker():
    try:
        import matplotlib
        if not getattr(matplotlib, 'IS_STAND_IN', False):
            pytest.skip('a real matplotlib is installed: tests/test_figure_mpl.py covers the walker there')
    except ImportError:
        pass
    saved = {k: v for k, v in sys.modules.items() if _ours(k)}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, FAKE)
    try:
        yield importlib.import_module('reasoning_image_generation_b200.figure_mpl')
    finally:
        sys.path.remove(FAKE)
        for k in [k for k in sys.modules if _ours(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
        importlib.invalidate_caches()


def scene():
    from matplotlib.figure import Figure
    from matplotlib.lines import Line2D
    from matplotlib.patches import Patch
    from matplotlib.path import Path
    from matplotlib.transforms import Affine2D
    fig = Figure(dpi=64.0)
    ax = fig.add_axes()
    data = Affine2D().scale(39.4, 39.4).translate(256.0, 256.0)
    circle = ax.add(Patch(Path([[0, 1], [1, 1], [1, 0], [0, 1]], [1, 4, 4, 4]), data, facecolor='none', edgecolor='black', linewidth=1.8, alpha=0.9))
    late_low = ax.add(Line2D(Path([[0, 0], [1, 1]]), data, color='black', linewidth=1.7, solid_capstyle='butt'))
    late_low.set_zorder(1)                       # same zorder as the patch, inserted after it: drawn after it
    fill = ax.add(Patch(Path([[0, 0], [1, 0], [1, 1], [0, 0]], [1, 2, 2, 79]), data, facecolor='white', edgecolor='none', linewidth=1.0))
    fill.set_zorder(0.5)                         # inserted last, lowest zorder: drawn first
    deco = ax.add(Line2D(Path([[-1, 0], [1, 0]]), data, linewidth=0.9, alpha=0.85))
    return fig, ax, (circle, late_low, fill, deco)


def test_draw_order_and_what_is_read(walker):
    fig, ax, (circle, late_low, fill, deco) = scene()
    arts = walker.extract_figure(fig)
    assert ax.aspect_applied == 1
    assert len(arts) == 4
    # zorder 0.5 (fill), then zorder 1 in insertion order (patch, line), then zorder 2 (decoration line)
    assert [a['facecolor'] is not None for a in arts] == [True, False, False, False]
    assert [round(a['linewidth'], 6) for a in arts] == [0.0, round(1.8 * 64 / 72, 6), round(1.7 * 64 / 72, 6), round(0.9 * 64 / 72, 6)]
    assert arts[1]['edgecolor'] == (0.0, 0.0, 0.0, 0.9) and arts[1]['cap'] == 'butt' and arts[1]['join'] == 'miter'
    assert arts[2]['cap'] == 'butt' and arts[2]['join'] == 'round' and arts[2]['edgecolor'] == (0.0, 0.0, 0.0, 1.0)
    assert arts[3]['cap'] == 'projecting' and arts[3]['edgecolor'][3] == 0.85
    assert arts[0]['facecolor'] == (1.0, 1.0, 1.0, 1.0)
    assert arts[1]['codes'].tolist() == [1, 4, 4, 4] and arts[3]['codes'] is None
    m = np.asarray(arts[1]['affine'])
    assert m.shape == (3, 3) and m[0, 0] == 39.4 and m[0, 2] == 256.0
    assert late_low.recached + deco.recached == 2                       # Line2D.draw's recache happened
    # and the list goes through the scene builder unchanged
    from reasoning_image_generation_b200.figure import FigureSceneBuilder
    sb = FigureSceneBuilder(512, 512)
    sb.add_image(arts)
    sc = sb.finish()
    assert sc.n_artists == 4 and sc.has_face.tolist() == [1, 0, 0, 0] and sc.join.tolist() == [0, 0, 1, 1]


def test_dpi_argument_scales_line_widths(walker):
    fig, ax, _ = scene()
    arts = walker.extract_figure(fig, dpi=200.0)
    assert fig.dpi == 200.0 and round(arts[1]['linewidth'], 6) == round(1.8 * 200 / 72, 6)


@pytest.mark.parametrize('what', ['text', 'collection', 'dashed', 'marker', 'hatch', 'clip_path', 'axes_on', 'figure_artist', 'grey_figure'])
def test_what_cannot_be_expressed_raises(walker, what):
    from matplotlib.artist import Artist
    from matplotlib.lines import Line2D
    from matplotlib.patches import Patch
    from matplotlib.path import Path
    from matplotlib.text import Text
    from matplotlib.transforms import Affine2D
    fig, ax, (circle, late_low, fill, deco) = scene()
    if what == 'text':
        ax.add(Text('label'))
    elif what == 'collection':
        ax.add(Artist())                          # any artist type that is neither a patch nor a line
    elif what == 'dashed':
        ax.add(Line2D(Path([[0, 0], [1, 1]]), Affine2D(), linestyle='--'))
    elif what == 'marker':
        ax.add(Line2D(Path([[0, 0], [1, 1]]), Affine2D(), marker='o'))
    elif what == 'hatch':
        ax.add(Patch(Path([[0, 0], [1, 0], [1, 1]]), Affine2D(), hatch='//'))
    elif what == 'clip_path':
        circle._clip_path = object()
    elif what == 'axes_on':
        ax.axison = True
    elif what == 'figure_artist':
        fig.texts.append(Text('suptitle'))
    elif what == 'grey_figure':
        fig.patch.get_facecolor = lambda: (0.97, 0.97, 0.98, 1.0)
    with pytest.raises(NotImplementedError):
        walker.extract_figure(fig)


def test_invisible_artists_are_skipped(walker):
    fig, ax, (circle, late_low, fill, deco) = scene()
    deco.set_visible(False)
    assert len(walker.extract_figure(fig)) == 3
