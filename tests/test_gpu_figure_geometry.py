"""k_figure_path (device geometry: transform, snap, curve flattening, stroking, clipping) against the host path pipeline.

The host pipeline (csrc/rig_figure_path.cpp, the libm the CPU renderer uses) is the definition; it is itself bit-exact against
oracle/aggdraw.c (tests/test_gpu_figure_parity.py, tests/test_figure_emulation.py) and pinned by digest
(tests/test_figure_host_digest.py).  The bar here: the device produces the SAME layers and the SAME edge arrays -- every 24.8
fixed-point coordinate, in the same order -- or hands the batch back to the host pipeline.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))

from reasoning_image_generation_b200.figure import FigureRenderer, FigureSceneBuilder, host_edges
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene

pytestmark = pytest.mark.gpu


def canon(layers, edges):
    """Edges of every layer in a canonical order: layers of at most 32 edges keep the emission order on the device (one chunk of
    k_figure's edge loop: the order inside it is irrelevant) while the host sorts them by first row; larger layers are sorted by
    first row, stably, on both sides."""
    out = edges.copy()
    for L in layers:
        b, n = int(L[1]), int(L[2])
        if n <= 32:
            e = out[b:b + n]
            out[b:b + n] = e[np.lexsort((e[:, 3], e[:, 2], e[:, 1], e[:, 0]))]
    return out


def same_geometry(scene, max_artists=None):
    r = FigureRenderer(scene.W, scene.H, max_images=max(scene.n_images, 1), max_artists=max_artists or max(scene.n_artists, 1))
    r.set_geometry('device')
    r.upload(scene)
    r.sync()
    st = r.stats()
    assert st['geom_resident'] == 1 and st['geom_status'] == 0, st
    dl, de = r.resident_edges()
    hl, he = host_edges(scene)
    r.close()
    assert dl.shape == hl.shape
    assert (dl == hl).all(), np.argwhere(dl != hl)[:5]
    assert de.shape == he.shape and (canon(dl, de) == canon(hl, he)).all()
    return st


@pytest.mark.parametrize('kind', ['single', 'multi'])
def test_synthetic_scenes_same_edges(built, kind):
    st = same_geometry(synthetic_figure_scene(1500, 512, 512, seed=33, kind=kind))
    assert st['edges'] > 0


def test_stress_scene_same_edges(built):
    """Every cap / join / snap value, curves, axis-aligned and degenerate polylines, fills (600 images)."""
    from figure_stress_scene import stress_scene
    sc = stress_scene()
    r = FigureRenderer(sc.W, sc.H, max_images=sc.n_images, max_artists=sc.n_artists)
    r.upload(sc)
    r.sync()
    st = r.stats()
    dl, de = r.resident_edges()
    hl, he = host_edges(sc)
    r.close()
    # round joins / caps go through libm: the device may hand the batch back (then the resident edges ARE the host's)
    assert st['geom_status'] in (0, 8), st
    assert dl.shape == hl.shape and (dl == hl).all()
    assert de.shape == he.shape and (canon(dl, de) == canon(hl, he)).all()


def test_large_canvas_same_edges(built):
    same_geometry(synthetic_figure_scene(64, 1600, 1600, seed=5, kind='single'))


def test_host_and_device_geometry_same_pixels(built):
    sc = synthetic_figure_scene(200, 512, 512, seed=77, kind='single')
    r = FigureRenderer(512, 512, max_images=256)
    r.set_geometry('device')
    a = r.render(sc).copy()
    assert r.stats()['geom_resident'] == 1
    r.set_geometry('host')
    b = r.render(sc)
    assert r.stats()['geom_resident'] == 0
    r.close()
    assert (a == b).all()


def test_draw_resident_reruns_the_geometry(built):
    import torch
    sc = synthetic_figure_scene(300, 512, 512, seed=9, kind='multi')
    r = FigureRenderer(512, 512, max_images=300)
    r.upload(sc)
    ref = torch.zeros((300, 512, 512, 3), dtype=torch.uint8, device='cuda:0')
    r.render_resident(ref.data_ptr())
    r.sync()
    out = torch.zeros_like(ref)
    r.set_profiling(True)
    for _ in range(3):
        out.zero_()
        torch.cuda.synchronize()
        r.draw_resident(out.data_ptr())
        r.sync()
        assert torch.equal(out, ref)
    st = r.stats()
    assert st['geom_launches'] == 3 and st['geom_ms'] > 0 and st['timed_launches'] == 3
    r.close()


def test_scratch_overflow_hands_the_batch_back(built, monkeypatch):
    """A stroke longer than the thread-private vertex storage: GEOM_SEQ_OVERFLOW, the host pipeline renders the batch."""
    monkeypatch.setenv('RIG_GEOM_SEQ_CAP', '16')
    sc = synthetic_figure_scene(40, 512, 512, seed=3, kind='single')
    r = FigureRenderer(512, 512, max_images=64)
    got = r.render(sc).copy()
    st = r.stats()
    assert st['geom_fallbacks'] == 1 and st['geom_status'] & 1 and st['geom_resident'] == 0, st
    monkeypatch.delenv('RIG_GEOM_SEQ_CAP')
    r2 = FigureRenderer(512, 512, max_images=64)
    ref = r2.render(sc)
    assert r2.stats()['geom_fallbacks'] == 0
    r.close(); r2.close()
    assert (got == ref).all()


def test_edge_arena_grows(built):
    """Wide round caps and joins: far more edges per vertex than the first estimate of the edge arena."""
    W = H = 512
    rng = np.random.RandomState(1)
    sb = FigureSceneBuilder(W, H)
    for i in range(64):
        pts = rng.uniform(100, 400, (3, 2))
        sb.add_artist(vertices=pts, codes=None, linewidth=float(rng.uniform(60, 120)), cap='round', join='round', edgecolor=(0, 0, 0, 0.8))
        sb.end_image()
    sc = sb.finish()
    r = FigureRenderer(W, H, max_images=64)
    r.upload(sc)
    r.sync()
    st = r.stats()
    dl, de = r.resident_edges()
    hl, he = host_edges(sc)
    r.close()
    assert st['edges'] > 12 * sc.n_vertices + 16 * sc.n_artists or st['geom_status'] == 8
    assert (dl == hl).all() and (canon(dl, de) == canon(hl, he)).all()


def test_bad_input_is_named_by_the_host_pipeline(built):
    from reasoning_image_generation_b200.figure import FigureError
    sb = FigureSceneBuilder(64, 64)
    sb.add_artist(vertices=[[1.0, 2.0], [float('nan'), 5.0]], codes=None, linewidth=1.0)
    sb.end_image()
    r = FigureRenderer(64, 64, max_images=4)
    with pytest.raises(FigureError) as e:
        r.render(sb.finish())
    assert 'non-finite' in str(e.value)
    r.close()
