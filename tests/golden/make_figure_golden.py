"""One command that PINS the Agg path to a real matplotlib (run it wherever matplotlib is installed; this image has none):

    python tests/golden/make_figure_golden.py            # writes tests/golden/figure_mpl.npz

It draws a fixed set of figures with matplotlib's Agg backend -- every cap / join / snap value, quadratic and cubic curves, alpha,
the patch types and styles of the reference's generator (multigraph_generation/shapes.py:21-72, style.py:29-66,
single_variants.py:285-374, 552-633), clipping at the canvas edge -- at 512 x 512 (figsize 8 x 8, dpi 64, as bench.py's workload),
keeps the RGBA canvas of each, and reads the draw_path argument lists back with reasoning_image_generation_b200.figure_mpl.
tests/test_figure_mpl.py (oracle/aggdraw.c against the pixels) and tests/test_gpu_figure_parity.py::test_matplotlib_golden
(k_figure_path + k_figure against the pixels) consume the file when it exists.  The matplotlib version is stored in the file and
printed by the tests: the three choices oracle/aggdraw.c states in its header (svn-Agg mult_cover rounding, mpl_round_to_int
snapping, the cp > vertex_dist_epsilon inner-join test) were read from matplotlib 3.8-3.10 sources.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def build_figures():
    import matplotlib
    matplotlib.use('Agg')
    import matplotlib.pyplot as plt
    from matplotlib.patches import Circle, Ellipse, FancyBboxPatch, PathPatch, Polygon, Rectangle, RegularPolygon, Wedge
    from matplotlib.path import Path
    from matplotlib.transforms import Affine2D
    rng = np.random.RandomState(20261018)
    figs = []

    def new():
        fig, ax = plt.subplots(figsize=(8, 8), dpi=64)      # the reference: plt.subplots(figsize=(8, 8)), savefig(dpi=...)
        ax.set_aspect('equal'); ax.axis('off'); ax.set_facecolor('#f8f9fa')
        ax.set_xlim(-5, 5); ax.set_ylim(-5, 5)
        figs.append(fig)
        return ax

    def styled(p, ax, rot=None):                              # style.py:29-66: no fill, black edge, 1.5-2.0 pt, alpha 0.9
        p.set_facecolor('none'); p.set_edgecolor('black'); p.set_linewidth(float(rng.uniform(1.5, 2.0))); p.set_alpha(0.9)
        if rot is not None:
            p.set_transform(Affine2D().rotate_deg_around(0.0, 0.0, rot) + ax.transData)
        ax.add_patch(p)

    # 1-8: one base patch each, as generator.py:105-128 builds them
    styled(Circle((0.2, -0.1), 1.95), new())
    ax = new(); styled(Ellipse((0, 0), 3.9, 2.34, angle=float(rng.uniform(-180, 180))), ax)
    ax = new(); styled(Rectangle((-1.95, -1.17), 3.9, 2.34), ax)
    ax = new(); styled(Rectangle((-1.95, -1.17), 3.9, 2.34), ax, rot=33.0)
    ax = new(); styled(FancyBboxPatch((-1.95, -1.17), 3.9, 2.34, boxstyle='round,pad=0,rounding_size=0.35'), ax)
    for n in (3, 5, 8):
        ax = new(); styled(RegularPolygon((0.1, 0.2), n, radius=1.95, orientation=float(rng.uniform(0, 1))), ax)
    ax = new(); styled(Wedge((0, 0), 1.95, 25.0, 290.0), ax)
    ax = new(); styled(Wedge((0.3, -0.2), 1.95, 100.0, 200.0), ax, rot=-70.0)
    # decoration lines: ax.plot defaults (projecting caps, round joins, colour cycle), alpha 0.8-0.9, 0.6-1.4 pt
    ax = new(); styled(Circle((0, 0), 1.95), ax)
    for k in range(5):
        a = 2 * np.pi * k / 5
        ax.plot([0, 1.95 * np.cos(a)], [0, 1.95 * np.sin(a)], linewidth=float(rng.uniform(0.6, 1.4)), alpha=float(rng.uniform(0.8, 0.9)))
    ax = new(); styled(Rectangle((-1.95, -1.17), 3.9, 2.34), ax)
    for x in (-1.0, 0.0, 1.0):
        ax.plot([x, x], [-1.17, 1.17], linewidth=1.0, alpha=0.85)
    for y in (-0.4, 0.4):
        ax.plot([-1.95, 1.95], [y, y], linewidth=0.8, alpha=0.8)
    ax = new(); styled(RegularPolygon((0, 0), 6, radius=1.95), ax)
    t = np.linspace(0, 2 * np.pi, 6)
    ax.plot(1.1 * np.cos(t), 1.1 * np.sin(t), linewidth=1.2, alpha=0.9)                     # an inner polygon through ax.plot: round joins
    # occlusion masks: white PathPatch fills (zorder 1) + black boundary segments, butt caps (single_variants.py:552-633)
    ax = new(); styled(Circle((0, 0), 1.95), ax)
    th = np.linspace(0, 2 * np.pi, 40)
    ring = np.c_[0.9 + 0.8 * np.cos(th), 0.7 + 0.8 * np.sin(th)]
    ax.add_patch(PathPatch(Path(ring, [Path.MOVETO] + [Path.LINETO] * (len(ring) - 2) + [Path.CLOSEPOLY]), facecolor='white', edgecolor='none', zorder=1))
    ax.plot(ring[:15, 0], ring[:15, 1], color='black', linewidth=1.7, solid_capstyle='butt', zorder=3)
    ax = new(); styled(Rectangle((-1.95, -1.17), 3.9, 2.34), ax)
    ax.add_patch(Polygon([[0.5, -2], [2.5, -2], [2.5, 0.2], [0.5, 0.2]], closed=True, facecolor='white', edgecolor='none', zorder=1))
    ax.plot([0.5, 0.5, 1.95], [-1.17, 0.2, 0.2], color='black', linewidth=1.6, solid_capstyle='butt', zorder=4)
    # multi-shape combinations (multi_combinator.py:685,774,1097): nested, adjacent, intersecting outlines
    ax = new()
    for i, r in enumerate((2.6, 2.6 * 0.4, 2.6 * 0.16)):
        styled(RegularPolygon((0, 0), 5 + i, radius=r), ax)
    ax = new(); styled(Circle((-1.6, 0), 1.6), ax); styled(Rectangle((0.0, -1.2), 2.6, 2.4), ax)
    ax = new(); styled(Ellipse((-0.6, 0), 3.4, 2.2, angle=20), ax); styled(Wedge((0.8, 0.3), 1.7, 10, 250), ax)
    # every cap x join on an open polyline and a closed one, wide enough for the joins to show; snap on / off / auto
    pts = np.array([[-3.5, -2.0], [-1.0, 2.5], [0.5, -1.5], [3.2, 1.8], [3.2, -2.8]])
    for cap in ('butt', 'projecting', 'round'):
        for join in ('miter', 'round', 'bevel'):
            ax = new()
            ax.plot(pts[:, 0], pts[:, 1], color=(0.1, 0.3, 0.7), alpha=0.8, linewidth=7.0, solid_capstyle=cap, solid_joinstyle=join)
            ax.add_patch(Polygon(pts[:4] * 0.45 + [0.3, -3.2], closed=True, facecolor=(0.9, 0.4, 0.1, 0.5), edgecolor=(0, 0, 0, 0.9), linewidth=3.0,
                                 joinstyle=join, capstyle=cap))
    for snap in (True, False, None):
        ax = new()
        r = Rectangle((-3.03, -2.21), 5.17, 3.38, facecolor=(0.2, 0.7, 0.3, 0.6), edgecolor='black', linewidth=1.0)
        r.set_snap(snap); ax.add_patch(r)
        (ln,) = ax.plot([-4.0, 4.0, 4.0], [3.3, 3.3, -3.9], color='black', linewidth=1.0)
        ln.set_snap(snap)
    # quadratic and cubic curves, a shape that leaves the canvas, thin lines
    ax = new()
    verts = [(-3, -1), (-1, 4), (2, 2), (3, -3), (0, -4), (-3, -1)]
    ax.add_patch(PathPatch(Path(verts, [Path.MOVETO, Path.CURVE3, Path.CURVE3, Path.CURVE3, Path.CURVE3, Path.LINETO]), facecolor='none', edgecolor='black',
                           linewidth=2.0, alpha=0.9))
    verts = [(-2, -2), (-3, 1), (1, 3), (2, 0), (3, -3), (-1, -3), (-2, -2)]
    ax.add_patch(PathPatch(Path(verts, [Path.MOVETO] + [Path.CURVE4] * 6), facecolor=(0.3, 0.3, 0.9, 0.3), edgecolor=(0.5, 0, 0, 1), linewidth=1.3))
    ax = new(); styled(Circle((3.9, 3.9), 2.5), ax); styled(Rectangle((-6.5, -6.2), 3.0, 3.0), ax, rot=12.0)
    ax.plot([-7, 7], [0.313, 0.877], linewidth=0.6, alpha=0.8)
    return figs


def main():
    import matplotlib
    from reasoning_image_generation_b200.figure import FigureSceneBuilder, scene_to_arrays
    from reasoning_image_generation_b200.figure_mpl import extract_figure
    figs = build_figures()
    sb = FigureSceneBuilder(512, 512)
    pixels = []
    for fig in figs:
        fig.canvas.draw()
        px = np.asarray(fig.canvas.buffer_rgba()).copy()
        assert px.shape == (512, 512, 4), px.shape
        pixels.append(px)
        sb.add_image(extract_figure(fig))
    scene = sb.finish()
    out = os.path.join(HERE, 'figure_mpl.npz')
    np.savez_compressed(out, pixels=np.stack(pixels), matplotlib_version=np.array(matplotlib.__version__), **scene_to_arrays(scene, 'scene_'))
    print('wrote', out, ':', len(figs), 'figures,', scene.n_artists, 'artists, matplotlib', matplotlib.__version__)


if __name__ == '__main__':
    main()
