"""GPU parity of the figure (matplotlib/Agg) path: k_figure through the C ABI against oracle/aggdraw.c, bit-exact.

The north_star tolerance for this path is mean |diff| <= 1/255 and >= 99.5 % of pixels within 2 levels; the bar held here is
zero differing pixels against the oracle (integer work end to end).  The oracle itself is UNPINNED against a real matplotlib
(not installable here), see oracle/aggdraw.c.
"""
import numpy as np
import pytest

from reasoning_image_generation_b200 import figure_paths as fp
from reasoning_image_generation_b200.figure import FigureRenderer, FigureSceneBuilder
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene, tile_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def agg(built):
    from oracle.agg_oracle import AggOracle
    return AggOracle()


def check(got, ref):
    diff = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    # the stated tolerance (north_star: mean |diff| <= 1/255 of full scale = 1 level, >= 99.5 % of pixels within 2 levels) ...
    assert diff.mean() <= 1.0 and (diff <= 2).mean() >= 0.995
    # ... and the bar actually held
    assert int((diff != 0).any(axis=-1).sum()) == 0


def test_matplotlib_golden(built):
    """k_figure_path + k_figure against pixels rendered by a real matplotlib (tests/golden/make_figure_golden.py); skipped until
    someone with matplotlib has generated the fixture -- until then the figure path is oracle-exact, matplotlib parity unverified."""
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'figure_mpl.npz')
    if not os.path.exists(golden):
        pytest.skip('tests/golden/figure_mpl.npz not generated (needs matplotlib)')
    from test_figure_mpl import check_against_matplotlib
    from reasoning_image_generation_b200.figure import scene_from_arrays
    d = np.load(golden)
    scene = scene_from_arrays(d, 'scene_')
    r = FigureRenderer(scene.W, scene.H, max_images=scene.n_images, max_artists=max(scene.n_artists, 1))
    got = r.render(scene, layout='rgb')
    r.close()
    check_against_matplotlib(got, d['pixels'], 'k_figure vs matplotlib ' + str(d['matplotlib_version']))


@pytest.mark.parametrize('kind', ['single', 'multi'])
def test_synthetic_figures(agg, kind):
    sc = synthetic_figure_scene(96, 512, 512, seed=21, kind=kind)
    r = FigureRenderer(512, 512, max_images=128)
    got = r.render(sc, layout='rgb')
    ref = agg.render(sc, 512, 512)
    check(got, ref)
    # BGR layout is the same image with the channels swapped
    check(r.render(sc, layout='bgr')[..., ::-1], ref)
    st = r.stats()
    assert st['kernel_launches'] >= 2 and st['edges'] > 0
    r.close()


def test_more_images_than_one_staging_chunk(agg):
    """Host destination: 512-image staging chunks, double buffered; 1100 images cross two chunk boundaries."""
    base = synthetic_figure_scene(50, 256, 256, seed=4, kind='single')
    sc = tile_scene(base, 22)
    r = FigureRenderer(256, 256, max_images=2048)
    got = r.render(sc)
    ref = agg.render(base, 256, 256)
    for k in range(22):
        assert (got[k * 50:(k + 1) * 50] == ref).all(), k
    r.close()


def test_resident_render_into_device_memory(agg):
    import torch
    sc = synthetic_figure_scene(64, 512, 512, seed=8, kind='multi')
    r = FigureRenderer(512, 512, max_images=64)
    out = torch.zeros((64, 512, 512, 3), dtype=torch.uint8, device='cuda:0')
    r.upload(sc)
    r.set_profiling(True)
    for _ in range(3):
        out.zero_()
        torch.cuda.synchronize()
        r.render_resident(out.data_ptr())
        r.sync()
    check(out.cpu().numpy(), agg.render(sc, 512, 512))
    st = r.stats()
    assert st['timed_launches'] == 3 and st['cover_ms'] > 0 and st['timed_images'] == 192
    r.close()


def test_clipping_canvas_edges_and_clip_rectangles(agg):
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
    r = FigureRenderer(W, H, max_images=16)
    check(r.render(sc), agg.render(sc, W, H))
    r.close()


def test_fuzz_paths_caps_joins(agg):
    rng = np.random.RandomState(3)
    W, H = 320, 200
    sb = FigureSceneBuilder(W, H)
    caps, joins = ('butt', 'projecting', 'round'), ('miter', 'round', 'bevel')
    for i in range(150):
        for _ in range(rng.randint(1, 40) if i % 10 == 0 else rng.randint(1, 6)):    # some images with more than 32 layers
            n = rng.randint(2, 7)
            pts = rng.uniform(-20, 340, size=(n, 2))
            closed = bool(rng.randint(2)) and n >= 3
            vv, cc = fp.polyline(pts, closed)
            if rng.randint(3) == 0 and n >= 4:
                cc = cc.copy()
                cc[1:4] = 4
            col = tuple(rng.uniform(0, 1, 3)) + (float(rng.choice([1.0, 0.9, 0.5])),)
            face = tuple(rng.uniform(0, 1, 4)) if rng.randint(3) == 0 else None
            sb.add_artist(vertices=vv, codes=cc, linewidth=float(rng.choice([0.3, 1.0, 1.37, 2.5, 9.0])), cap=caps[rng.randint(3)],
                          join=joins[rng.randint(3)], edgecolor=col, facecolor=face, snap=[None, True, False][rng.randint(3)])
        sb.end_image()
    sc = sb.finish()
    r = FigureRenderer(W, H, max_images=256, max_artists=8192)
    check(r.render(sc), agg.render(sc, W, H))
    r.close()


@pytest.mark.parametrize('size', [(256, 256), (640, 480), (1600, 1600), (2400, 2400), (48, 33), (16, 3), (3072, 37)])
def test_other_canvas_sizes(agg, size):
    W, H = size
    sc = synthetic_figure_scene(5, W, H, seed=5, kind='single')
    r = FigureRenderer(W, H, max_images=8)
    check(r.render(sc), agg.render(sc, W, H))
    r.close()


def test_empty_scene_and_blank_images(agg):
    r = FigureRenderer(64, 64, max_images=8)
    sb = FigureSceneBuilder(64, 64)
    assert r.render(sb.finish()).shape == (0, 64, 64, 3)
    sb = FigureSceneBuilder(64, 64)
    sb.end_image()
    sb.add_artist(vertices=[[10, 10]], codes=None, linewidth=3.0)
    sb.end_image()
    got = r.render(sb.finish())
    assert got.shape == (2, 64, 64, 3) and (got == 255).all()
    r.close()


def test_capacity_and_bad_arguments_are_errors():
    from reasoning_image_generation_b200.figure import FigureError
    r = FigureRenderer(64, 64, max_images=2, max_artists=4)
    sc = synthetic_figure_scene(3, 64, 64, seed=1)
    with pytest.raises(FigureError) as e:
        r.render(sc)
    assert e.value.code == -3
    sb = FigureSceneBuilder(64, 64)
    sb.add_artist(vertices=[[1, 1], [float('nan'), 2]], codes=None)
    with pytest.raises(FigureError):
        r.render(sb.finish())
    r.close()


def test_roundtrip_properties_at_full_size(agg):
    """Size-independent checks on a BASELINE-sized batch (8192 images at 512^2 = 6.4 GB stay on the device): every image
    of a tiled scene equals its first copy, and a checksum of checksums equals the oracle's on the distinct images."""
    import torch
    base = synthetic_figure_scene(64, 512, 512, seed=33, kind='single')
    sc = tile_scene(base, 128)
    r = FigureRenderer(512, 512, max_images=8192)
    out = torch.empty((8192, 512, 512, 3), dtype=torch.uint8, device='cuda:0')
    r.upload(sc)
    r.render_resident(out.data_ptr())
    r.sync()
    v = out.view(128, 64, -1)
    assert bool((v == v[0:1]).all())
    ref = agg.render(base, 512, 512)
    sums = out[:64].to(torch.int64).sum(dim=(1, 2, 3)).cpu().numpy()
    assert (sums == ref.astype(np.int64).sum(axis=(1, 2, 3))).all()
    r.close()
