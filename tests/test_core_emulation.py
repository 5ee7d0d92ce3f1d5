"""The product's per-lane code (csrc/rig_core.h, rig_shapes.h, rig_expand.h, rig_raster.h) compiled for the
host and run serially (tests/cpu_emul/emul.cpp) against fixtures rendered by the live reference.  This is the
no-GPU check of the arithmetic the kernels execute; the GPU tests repeat it through librig.so."""
import ctypes

import numpy as np
import pytest

from conftest import emul_render_scene, restore_rng
from reasoning_image_generation_b200.scene import SceneBuilder
from reasoning_image_generation_b200.synth import scene_to_elements, synthetic_scene
from oracle import ref_render


def _scene(c):
    sb = SceneBuilder(c['W'], c['H'], c['bg'], c['grid_size'])
    sb.add_panel(c['elements'], c['use_grid'])
    return sb.finish()


@pytest.mark.parametrize('band_rows', [8, 1, 512])
def test_sequence_goldens_bit_exact(emul_lib, golden_seq, band_rows):
    panels, pix = golden_seq
    step = 1 if band_rows == 8 else 4
    for c, want in list(zip(panels, pix))[::step]:
        restore_rng(c)
        rc, out, nrec, err = emul_render_scene(emul_lib, _scene(c), band_rows)
        assert rc == 0, err
        assert (out[0] == want).all(), [e['kind'] for e in c['elements']]


def test_kind_goldens_bit_exact(emul_lib, golden_kinds):
    panels, pix = golden_kinds
    for c, want in zip(panels, pix):
        restore_rng(c)
        rc, out, nrec, err = emul_render_scene(emul_lib, _scene(c), 8)
        assert rc == 0, err
        assert (out[0] == want).all(), [(e['kind'], e['center'], e['angle']) for e in c['elements']]


def test_synthetic_scene_matches_oracle(emul_lib):
    sc = synthetic_scene(120, seed=5)
    rc, out, nrec, err = emul_render_scene(emul_lib, sc, 8)
    assert rc == 0, err
    be = ref_render.CBackend()
    for p in range(sc.n_panels):
        els, ug = scene_to_elements(sc, p)
        ref = ref_render.render_panel(512, 512, (255, 255, 255), els, ug, 3, backend=be)
        assert (ref == out[p]).all(), p


def test_rgb_layout_is_channel_reversed_bgr(emul_lib):
    sc = synthetic_scene(6, seed=2)
    _, bgr, _, _ = emul_render_scene(emul_lib, sc, 8, layout=1)
    _, rgb, _, _ = emul_render_scene(emul_lib, sc, 8, layout=0)
    assert (bgr[..., ::-1] == rgb).all()


def test_bresenham_closed_form(emul_lib):
    for dx in range(0, 150):
        for dy in range(0, dx + 1):
            assert emul_lib.emul_selftest_line8(dx, dy) == 0, (dx, dy)


def test_unsupported_inputs_are_rejected(emul_lib):
    sb = SceneBuilder()
    sb.add_panel([dict(kind='square', size=100, fill=True, stroke_width=9, center=(1, 1), angle=0, color=(1, 2, 3))])
    rc, _, _, err = emul_render_scene(emul_lib, sb.finish(), 8)
    assert rc == -4 and 'stroke_width' in err


def test_fuzzed_elements_match_oracle(emul_lib):
    """Elements far outside the sampler's range (sizes 1..300, stroke 0..4, any angle, centres off-canvas)."""
    import random
    rng = random.Random(77)
    sb = SceneBuilder(512, 512, bg_color=(250, 240, 230))
    panels = []
    for p in range(60):
        els = []
        for _ in range(rng.randint(0, 6)):
            els.append(dict(kind=rng.choice(ref_render.KINDS), size=rng.choice([1, 2, 7, 30, 31, 60, 99, 128, 170, 255, 300]),
                            fill=rng.random() < 0.5, stroke_width=rng.randint(0, 4),
                            center=(rng.uniform(-300, 800), rng.uniform(-300, 800)) if rng.random() < 0.5
                            else (rng.randint(-40, 552), rng.randint(-40, 552)),
                            angle=rng.choice([rng.randint(-720, 720), rng.uniform(0, 360), 0, 45, 90]),
                            color=(rng.randint(0, 255), rng.randint(0, 255), rng.randint(0, 255))))
        panels.append((els, rng.random() < 0.5))
        sb.add_panel(els, panels[-1][1])
    rc, out, nrec, err = emul_render_scene(emul_lib, sb.finish(), 5)
    assert rc == 0, err
    be = ref_render.CBackend()
    for p, (els, ug) in enumerate(panels):
        ref = ref_render.render_panel(512, 512, (250, 240, 230), els, ug, 3, backend=be)
        assert (ref == out[p]).all(), (p, els)


@pytest.mark.parametrize('W,H', [(256, 192), (640, 361), (64, 48)])
def test_other_canvas_sizes(emul_lib, W, H):
    """GenConfig.canvas_size is configurable (/root/reference/src/config.py:26): nothing may assume 512 x 512."""
    import random
    rng = random.Random(W * 1000 + H)
    sb = SceneBuilder(W, H)
    panels = []
    for p in range(24):
        els = [dict(kind=rng.choice(ref_render.KINDS), size=rng.choice([20, 40, 64, 90]), fill=rng.random() < 0.6,
                    stroke_width=rng.randint(1, 3), center=(rng.randint(-10, W + 10), rng.randint(-10, H + 10)),
                    angle=rng.choice([0, 30, 45, 90, 200]), color=(rng.randint(0, 255), rng.randint(0, 255), rng.randint(0, 255)))
               for _ in range(rng.randint(0, 4))]
        panels.append((els, p % 2 == 0))
        sb.add_panel(els, p % 2 == 0)
    rc, out, nrec, err = emul_render_scene(emul_lib, sb.finish(), 5)
    assert rc == 0, err
    be = ref_render.CBackend()
    for p, (els, ug) in enumerate(panels):
        ref = ref_render.render_panel(W, H, (255, 255, 255), els, ug, 3, backend=be)
        assert (ref == out[p]).all(), (p, els)
