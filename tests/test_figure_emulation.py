"""Host path pipeline (librig.so: rig_figure_debug_edges, no GPU needed) + the per-lane cell arithmetic of k_figure compiled
for the CPU (tests/cpu_emul/figure_emul.cpp) against oracle/aggdraw.c: zero differing pixels, at every band height."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT
from reasoning_image_generation_b200 import figure_paths as fp
from reasoning_image_generation_b200.figure import FigureSceneBuilder, host_edges
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene, tile_scene


@pytest.fixture(scope='module')
def agg(built):
    from oracle.agg_oracle import AggOracle
    return AggOracle()


@pytest.fixture(scope='module')
def femul(built):
    lib = ctypes.CDLL(os.path.join(ROOT, 'tests', 'cpu_emul', 'libfigure_emul.so'))
    lib.figure_emul_render.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.figure_emul_render.restype = None
    return lib


def emul_render(femul, scene, band_rows, acc_cells=256):
    layers, edges = host_edges(scene)
    out = np.zeros((scene.n_images, scene.H, scene.W, 3), np.uint8)
    for i in range(scene.n_images):
        L = np.ascontiguousarray(layers[layers[:, 0] == i])
        femul.figure_emul_render(scene.W, scene.H, len(L), L.ctypes.data, edges.ctypes.data, band_rows, acc_cells, out[i].ctypes.data)
    return out


@pytest.mark.parametrize('kind', ['single', 'multi'])
def test_synthetic_figures_match_the_oracle(agg, femul, kind):
    sc = synthetic_figure_scene(40, 512, 512, seed=11, kind=kind)
    ref = agg.render(sc, 512, 512)
    assert (ref != 255).any(axis=(1, 2, 3)).all()
    for rows in (1, 4, 7):
        got = emul_render(femul, sc, rows)
        assert int((got != ref).any(axis=3).sum()) == 0, (kind, rows)


def test_other_canvas_sizes(agg, femul):
    for (W, H) in ((256, 256), (640, 480), (1600, 1600), (48, 33), (16, 3), (2048, 37)):
        sc = synthetic_figure_scene(3, W, H, seed=5, kind='single')
        ref = agg.render(sc, W, H)
        got = emul_render(femul, sc, 4)
        assert int((got != ref).any(axis=3).sum()) == 0, (W, H)


def test_geometry_crossing_the_canvas_and_clip_rectangles(agg, femul):
    """The rasteriser's clipper: shapes leaving the canvas on every side, a clip rectangle smaller than the canvas, fills that
    reach the right edge (running cover still non-zero at x = W)."""
    W = H = 128
    sb = FigureSceneBuilder(W, H)
    v, c = fp.unit_circle()
    for (cx, cy) in ((0, 64), (128, 64), (64, 0), (64, 128), (5, 5), (125, 125), (-30, 64), (64, 200)):
        aff = fp.compose(fp.affine_scale(40, 25), fp.affine_rotate_deg(30), fp.affine_translate(cx, cy))
        sb.add_artist(vertices=v, codes=c, affine=aff, linewidth=3.0, edgecolor=(0, 0, 0, 0.9), facecolor=(0.2, 0.6, 0.9, 0.7))
        sb.end_image()
    aff = fp.compose(fp.affine_scale(60, 60), fp.affine_translate(64, 64))
    sb.add_artist(vertices=v, codes=c, affine=aff, linewidth=5.0, edgecolor=(0.8, 0.1, 0.1, 1.0), facecolor=(0, 0.5, 0, 1), clip=(20.3, 30.6, 100.2, 90.9))
    sb.end_image()
    sb.add_artist(vertices=[[-50, -20], [300, 90]], codes=None, linewidth=7.5, cap='round', edgecolor=(0, 0, 0, 1))
    sb.add_artist(vertices=[[64, -50], [64.0, 300]], codes=None, linewidth=2.0, cap='projecting', edgecolor=(0, 0, 1, 1))
    sb.end_image()
    sc = sb.finish()
    ref = agg.render(sc, W, H)
    for rows in (1, 4, 5):
        got = emul_render(femul, sc, rows)
        assert int((got != ref).any(axis=3).sum()) == 0, rows
    # the clipped artist leaves everything outside its clip box white
    img = ref[8]
    assert (img[:128 - 91] == 255).all() and (img[:, :20] == 255).all() and (img[:, 101:] == 255).all()


def test_random_polylines_and_joins(agg, femul):
    """Fuzz: random open / closed polylines and Bezier paths with every cap / join, widths 0.3 .. 9 px, translucent colours."""
    rng = np.random.RandomState(3)
    W = H = 160
    sb = FigureSceneBuilder(W, H)
    caps, joins = ('butt', 'projecting', 'round'), ('miter', 'round', 'bevel')
    for i in range(60):
        for _ in range(rng.randint(1, 5)):
            n = rng.randint(2, 7)
            pts = rng.uniform(-20, 180, size=(n, 2))
            closed = bool(rng.randint(2)) and n >= 3
            vv, cc = fp.polyline(pts, closed)
            if rng.randint(3) == 0 and n >= 4:
                cc = cc.copy()
                cc[1:4] = 4                      # one cubic segment
            col = tuple(rng.uniform(0, 1, 3)) + (float(rng.choice([1.0, 0.9, 0.5])),)
            face = tuple(rng.uniform(0, 1, 4)) if rng.randint(3) == 0 else None
            sb.add_artist(vertices=vv, codes=cc, linewidth=float(rng.choice([0.3, 1.0, 1.37, 2.5, 9.0])), cap=caps[rng.randint(3)],
                          join=joins[rng.randint(3)], edgecolor=col, facecolor=face, snap=[None, True, False][rng.randint(3)])
        sb.end_image()
    sc = sb.finish()
    ref = agg.render(sc, W, H)
    for acc_cells in (256, 64, 32):          # layers wider than the accumulators: column chunks with a carry-in
        got = emul_render(femul, sc, 4, acc_cells)
        assert int((got != ref).any(axis=3).sum()) == 0, acc_cells


def test_empty_and_degenerate_inputs(agg, femul):
    sb = FigureSceneBuilder(64, 64)
    sb.end_image()                                                        # no artists
    sb.add_artist(vertices=np.zeros((0, 2)), codes=np.zeros(0, np.uint8))  # empty path
    sb.end_image()
    sb.add_artist(vertices=[[10, 10]], codes=None, linewidth=3.0)          # a single point strokes to nothing
    sb.add_artist(vertices=[[10, 10], [10, 10], [10, 10]], codes=None, linewidth=3.0, cap='round')   # coincident points
    sb.add_artist(vertices=[[5, 5], [30, 5], [30, 5], [30, 30]], codes=None, linewidth=2.0, facecolor=(0, 0, 0, 0))  # transparent face
    sb.end_image()
    sc = sb.finish()
    ref = agg.render(sc, 64, 64)
    got = emul_render(femul, sc, 4)
    assert (ref[0] == 255).all() and (ref[1] == 255).all()
    assert int((got != ref).any(axis=3).sum()) == 0


def test_scene_slicing_and_tiling_keep_images(agg):
    sc = synthetic_figure_scene(6, 128, 128, seed=2)
    ref = agg.render(sc, 128, 128)
    part = sc.slice(2, 5)
    assert part.n_images == 3 and (agg.render(part, 128, 128) == ref[2:5]).all()
    t = tile_scene(sc, 3)
    assert t.n_images == 18 and (agg.render(t, 128, 128, images=[7, 13]) == ref[[1, 1]]).all()
    # shards generated independently equal slices of the whole
    again = synthetic_figure_scene(3, 128, 128, seed=2, first=2)
    assert (agg.render(again, 128, 128) == ref[2:5]).all()


def test_division_free_blend_is_exact(femul):
    """k_figure blends with a multiply by ceil(2^40 / (65280 + alpha)) instead of three divisions per pixel: checked
    exhaustively (every destination value x colour value x alpha) against fixed_blender_rgba_plain's division."""
    femul.figure_emul_check_magic.restype = ctypes.c_long
    assert femul.figure_emul_check_magic() == 0


def test_grouped_resolve_as_the_kernel_organises_it(agg, femul):
    """k_figure visits only the touched 16-cell groups of a row and fills the runs between them at constant coverage (its two
    half-warps take the band's two rows; rows are independent).  The same organisation on the CPU, for the kernel's group size and
    two others, with accumulators narrower than the shapes (column chunks with a carry-in) and fills that run to the canvas edge."""
    femul.figure_emul_render_grouped.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_void_p]
    femul.figure_emul_render_grouped.restype = None
    rng = np.random.RandomState(9)
    sb = FigureSceneBuilder(320, 96)
    for i in range(24):
        for _ in range(int(rng.randint(1, 5))):
            n = int(rng.randint(3, 7))
            pts = rng.uniform(-30, 350, size=(n, 2)) * [1.0, 0.3]
            v, c = fp.polyline(pts, True)
            sb.add_artist(vertices=v, codes=c, linewidth=float(rng.choice([0.0, 1.3, 4.0])), edgecolor=tuple(rng.uniform(0, 1, 3)) + (0.9,),
                          facecolor=tuple(rng.uniform(0, 1, 4)) if rng.randint(2) else None, snap=False)
        sb.end_image()
    scenes = [sb.finish(), synthetic_figure_scene(12, 512, 512, seed=3, kind='single'), synthetic_figure_scene(8, 512, 512, seed=4, kind='multi')]
    for sc in scenes:
        ref = agg.render(sc, sc.W, sc.H)
        layers, edges = host_edges(sc)
        for group, acc_cells in ((16, 256), (16, 64), (32, 256), (8, 32)):
            out = np.zeros((sc.n_images, sc.H, sc.W, 3), np.uint8)
            for i in range(sc.n_images):
                L = np.ascontiguousarray(layers[layers[:, 0] == i])
                femul.figure_emul_render_grouped(sc.W, sc.H, len(L), L.ctypes.data, edges.ctypes.data, 2, acc_cells, group, out[i].ctypes.data)
            assert int((out != ref).any(axis=3).sum()) == 0, (sc.W, group, acc_cells)
