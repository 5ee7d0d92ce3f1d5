"""oracle/aggdraw.c (the restated matplotlib/Agg pipeline) against answers that follow from Agg's definitions.

Parity with a real matplotlib is UNPINNED (matplotlib is not installable here): these tests pin the integer
arithmetic -- cell coverage, the blender, snapping, joins and caps -- to values derived by hand from the published algorithm.
"""
import numpy as np
import pytest

from reasoning_image_generation_b200 import figure_paths as fp
from reasoning_image_generation_b200.figure import FigureSceneBuilder


@pytest.fixture(scope='module')
def agg(built):
    from oracle.agg_oracle import AggOracle
    return AggOracle()


def _one(agg, W, H, **artist):
    sb = FigureSceneBuilder(W, H)
    sb.add_artist(**artist)
    return agg.render(sb.finish(), W, H)[0]


def test_pixel_aligned_fill_is_exact(agg):
    v, c = fp.unit_rectangle()
    aff = fp.compose(fp.affine_scale(10, 20), fp.affine_translate(10, 10))
    im = _one(agg, 64, 64, vertices=v, codes=c, affine=aff, linewidth=0, facecolor=(1, 0, 0, 1))
    red = (im == (255, 0, 0)).all(2)
    assert red.sum() == 200 and (im[~red] == 255).all()
    ys, xs = np.where(red)
    assert (ys.min(), ys.max(), xs.min(), xs.max()) == (34, 53, 10, 19)     # y flips: rows H-30 .. H-11


def test_half_covered_cells_and_the_plain_blender(agg):
    """A box edge at x.5 covers half a cell: area = 256 * 128 -> coverage 128 -> alpha = round(255 * 128 / 255) = 128
    -> black over white = (255 * 255 * (256 - 128)) // (65280 + 128) = 127 (fixed_blender_rgba_plain, dst alpha 255)."""
    v, c = fp.unit_rectangle()
    aff = fp.compose(fp.affine_scale(10, 20), fp.affine_translate(10.5, 10.25))
    im = _one(agg, 64, 64, vertices=v, codes=c, affine=aff, linewidth=0, facecolor=(0, 0, 0, 1), snap=False)[:, :, 0]
    assert list(im[40, 9:22]) == [255, 127] + [0] * 9 + [127, 255]
    # top edge at y = 64 - 30.25 = 33.75: 3/4 of row 33 is outside -> coverage 64 -> 191; bottom at 53.75 -> 63
    assert im[33, 15] == 191 and im[53, 15] == 63 and im[32, 15] == 255 and im[54, 15] == 255
    # corner cell: a quarter-by-half cell -> 256 * 256 / 8 -> coverage 32
    assert im[33, 10] == 255 - 32 or im[33, 10] == 223


def test_translucent_colour_uses_mult_cover(agg):
    """alpha 0.9 -> rgba8 230; full coverage 255: mult_cover(230, 255) = 230; black over white:
    (255 * 255 * (256 - 230)) // (65280 + 230) = 25."""
    v, c = fp.unit_rectangle()
    aff = fp.compose(fp.affine_scale(8, 8), fp.affine_translate(4, 4))
    im = _one(agg, 16, 16, vertices=v, codes=c, affine=aff, linewidth=0, facecolor=(0, 0, 0, 0.9))
    assert im[8, 8, 0] == (255 * 255 * (256 - 230)) // (65280 + 230) == 25


def test_rectilinear_paths_snap_to_pixel_centres(agg):
    """PathSnapper: a 1 px wide horizontal line at display y = 10.3 (Agg y = 21.7) snaps to the pixel
    centre floor(21.7 + 0.5) + 0.5 = 22.5 and covers exactly row 22; x snaps to 4.5 .. 21.5."""
    im = _one(agg, 32, 32, vertices=[[4.2, 10.3], [20.7, 10.3]], codes=None, linewidth=1.0, edgecolor=(0, 0, 0, 1))[:, :, 0]
    rows = np.where((im != 255).any(1))[0]
    assert list(rows) == [22]
    assert (im[22, 5:21] == 0).all() and im[22, 3] == 255 and im[22, 22] == 255 and im[22, 4] == 127 and im[22, 21] == 127
    # without snapping the same line straddles two rows
    im2 = _one(agg, 32, 32, vertices=[[4.2, 10.3], [20.7, 10.3]], codes=None, linewidth=1.0, edgecolor=(0, 0, 0, 1), snap=False)[:, :, 0]
    assert len(np.where((im2 != 255).any(1))[0]) == 2


def test_stroke_area_matches_geometry(agg):
    """Ink of a stroked circle = 2 pi r w; of a butt-capped segment = L w; projecting caps add w / 2 at each end."""
    v, c = fp.unit_circle()
    aff = fp.compose(fp.affine_scale(20, 20), fp.affine_translate(32, 32))
    im = _one(agg, 64, 64, vertices=v, codes=c, affine=aff, linewidth=2.0, edgecolor=(0, 0, 0, 1))[:, :, 0]
    ink = (255 - im.astype(int)).sum() / 255.0
    assert abs(ink - 2 * np.pi * 20 * 2) / (2 * np.pi * 20 * 2) < 0.01
    assert np.abs(im.astype(int) - im[:, ::-1]).max() <= 2 and np.abs(im.astype(int) - im[::-1]).max() <= 2
    L = np.hypot(45, 35)
    for cap, extra in (('butt', 0.0), ('projecting', 1.5), ('round', 1.5 * np.pi / 4)):
        im = _one(agg, 64, 64, vertices=[[5, 5], [50, 40]], codes=None, linewidth=1.5, cap=cap, edgecolor=(0, 0, 0, 1))[:, :, 0]
        ink = (255 - im.astype(int)).sum() / 255.0
        assert abs(ink - (L + extra) * 1.5) < 1.2, (cap, ink)


def test_miter_limit_is_the_line_width(agg):
    """_draw_path sets miter_limit = line width in pixels, so a right-angle corner (miter length w / sqrt 2) is mitred
    for w >= sqrt 2 and bevelled (miter_join_revert) below that."""
    sb = FigureSceneBuilder(64, 64)
    pts = [[10, 10], [40, 10], [40, 40]]
    for w in (1.2, 3.0):
        sb.add_artist(vertices=pts, codes=None, linewidth=w, join='miter', snap=False, edgecolor=(0, 0, 0, 1))
        sb.end_image()
    sc = sb.finish()
    n_bevel = len(agg.outline(sc, 0, 0, 64))
    n_miter = len(agg.outline(sc, 1, 0, 64))
    assert n_bevel == n_miter + 1      # the bevel replaces the one outer miter point by two


def test_round_join_and_cap_point_counts(agg):
    """calc_arc / round caps: da = 2 acos(w / (w + 0.125)); a half circle of half-width 2 gets int(pi / da) inner points."""
    sb = FigureSceneBuilder(64, 64)
    sb.add_artist(vertices=[[10, 10], [40, 10]], codes=None, linewidth=4.0, cap='round', snap=False, edgecolor=(0, 0, 0, 1))
    sc = sb.finish()
    da = 2 * np.arccos(2.0 / (2.0 + 0.125))
    n = int(np.pi / da)
    ol = agg.outline(sc, 0, 0, 64)
    assert len(ol) == 2 * (n + 2) + 1          # two caps of n + 2 points, one end_poly


def test_curves_flatten_to_the_distance_tolerance(agg):
    """curve4_div with approximation_scale 1: every flattened vertex lies on the circle to < 0.04 px (the unit-circle
    Beziers deviate by up to 4.1e-4 r), chord sagitta < 0.5 px."""
    v, c = fp.unit_circle()
    r = 77.0
    aff = fp.compose(fp.affine_scale(r, r), fp.affine_translate(256, 256))
    sb = FigureSceneBuilder(512, 512)
    sb.add_artist(vertices=v, codes=c, affine=aff, linewidth=1.0, edgecolor=(0, 0, 0, 1))
    sc = sb.finish()
    flat = agg.outline(sc, 0, 0, 512, stroke=False)
    pts = flat[flat[:, 2] <= 4][:, :2]
    rad = np.hypot(pts[:, 0] - 256, pts[:, 1] - 256)
    assert np.abs(rad - r).max() < 0.04
    seg = np.hypot(*(pts[1:] - pts[:-1]).T)
    sag = r - np.sqrt(r * r - (seg / 2) ** 2)
    assert sag.max() < 0.5 and len(pts) >= 32


def test_layers_blend_in_artist_order(agg):
    v, c = fp.unit_rectangle()
    a1 = fp.compose(fp.affine_scale(10, 10), fp.affine_translate(2, 2))
    a2 = fp.compose(fp.affine_scale(10, 10), fp.affine_translate(6, 6))
    sb = FigureSceneBuilder(20, 20)
    sb.add_artist(vertices=v, codes=c, affine=a1, linewidth=0, facecolor=(1, 0, 0, 1))
    sb.add_artist(vertices=v, codes=c, affine=a2, linewidth=0, facecolor=(0, 0, 1, 0.5))
    im = agg.render(sb.finish(), 20, 20)[0]
    # overlap pixel: blue at alpha 128 over red: r = (255*255*(256-128)) // (65280+128) = 127, b = (255*256*128) // 65408 = 127
    assert tuple(im[20 - 9, 8]) == (127, 0, 127)
    assert tuple(im[20 - 4, 4]) == (255, 0, 0)


def test_random_fills_cover_their_analytic_area(agg):
    """Independent of any reading of Agg's sources: an area-coverage rasteriser must lay down as much ink as the polygon has
    area (shoelace formula), for convex and self-avoiding non-convex polygons at arbitrary sub-pixel positions; and a
    translucent fill must lay down alpha times that."""
    rng = np.random.RandomState(11)
    worst = 0.0
    for k in range(60):
        n = int(rng.randint(3, 9))
        ang = 2 * np.pi * (np.arange(n) + rng.uniform(-0.2, 0.2, n)) / n + rng.uniform(0, 6.28)     # gaps below pi: a simple polygon
        rad = rng.uniform(8, 28, n) if k % 2 else np.full(n, rng.uniform(8, 28))       # star-shaped (non-convex) / convex
        cx, cy = rng.uniform(30, 34, 2)
        pts = np.c_[cx + rad * np.cos(ang), cy + rad * np.sin(ang)]
        x, y = pts[:, 0], pts[:, 1]
        area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
        v, c = fp.polyline(pts, True)
        alpha = 1.0 if k % 3 else 0.5
        im = _one(agg, 64, 64, vertices=v, codes=c, linewidth=0, facecolor=(0, 0, 0, alpha), snap=False)[:, :, 0]
        ink = (255 - im.astype(int)).sum() / 255.0
        want = area * (round(alpha * 255) / 255.0)
        worst = max(worst, abs(ink - want) / want)
        assert abs(ink - want) <= 0.01 * want + 0.5, (k, ink, want)
    assert worst < 0.012
