"""Host-side scene emitter mirrors the argument handling of the reference method it replaces
(/root/reference/src/generator.py:85-94, /root/reference/src/shapes.py:233)."""
import copy
import random

import numpy as np
import pytest

from reasoning_image_generation_b200.scene import KINDS, SceneBuilder
from oracle import ref_render


def el(**kw):
    d = dict(kind='square', size=100, fill=True, stroke_width=2, center=(100, 120), angle=45.0, color=(10, 20, 30))
    d.update(kw)
    return d


def test_kind_order_is_the_reference_tuple():
    assert KINDS == ref_render.KINDS == ('square', 'circle', 'triangle', 'diamond', 'star', 'pentagon', 'hexagon',
                                         'plus', 'heart', 'crescent', 'rounded_square')


def test_fields_defaults_and_coercions():
    sb = SceneBuilder(512, 512)
    sb.add_panel([el(size=100.9, angle=44.9, stroke_width=2.7, fill=1, center=(10.7, -3.2)),
                  {'kind': 'circle', 'size': 50, 'fill': False}], use_grid=True)
    s = sb.finish()
    assert s.size.tolist() == [100, 50] and s.angle.tolist() == [44, 0] and s.stroke.tolist() == [2, 1]
    assert s.fill.tolist() == [1, 0] and s.kind.tolist() == [0, 1]
    assert s.cx.tolist() == [10.7, 256.0] and s.cy.tolist() == [-3.2, 256.0]   # default centre (W//2, H//2)
    assert s.flags.tolist() == [1] and s.begin.tolist() == [0, 2]


def test_elements_are_not_mutated():
    els = [el(color=None), el(kind='heart')]
    before = copy.deepcopy(els)
    SceneBuilder().add_panel(els)
    assert els == before


def test_colourless_elements_consume_the_global_stream_like_the_reference():
    """rand_color() = 3 x random.uniform(30, 220) per colourless element, in element order, at call time."""
    els = [el(color=None), el(), el(color=None, kind='circle')]
    random.seed(1234)
    sb = SceneBuilder(); sb.add_panel(els); s = sb.finish()
    after_ours = random.random()
    random.seed(1234)
    ref_render.panel_ops(512, 512, els)
    after_oracle = random.random()
    assert after_ours == after_oracle
    random.seed(1234)
    want = [tuple(int(random.uniform(30, 220)) for _ in range(3)) for _ in range(2)]
    rgb = s.rgb.reshape(-1, 3)
    assert tuple(rgb[0]) == want[0] and tuple(rgb[2]) == want[1] and tuple(rgb[1]) == (10, 20, 30)


def test_error_behaviour_matches_reference():
    with pytest.raises(KeyError):
        SceneBuilder().add_panel([{'size': 10, 'fill': True}])          # el['kind'] (generator.py:87)
    with pytest.raises(KeyError):
        SceneBuilder().add_panel([{'kind': 'square', 'fill': True}])    # el['size'] (generator.py:88)
    with pytest.raises(AssertionError, match='unsupported shape'):
        SceneBuilder().add_panel([el(kind='octagon')])                  # Shape.__init__ assert (shapes.py:203)


def test_scene_slice_roundtrip():
    sb = SceneBuilder()
    for k in range(5):
        sb.add_panel([el(size=60 + k)] * k, use_grid=bool(k % 2))
    s = sb.finish()
    part = s.slice(2, 5)
    assert part.n_panels == 3 and part.begin.tolist() == [0, 2, 5, 9] and part.flags.tolist() == [0, 1, 0]
    assert part.size.tolist() == [62, 62, 63, 63, 63, 64, 64, 64, 64]
