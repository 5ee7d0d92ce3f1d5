"""The restated reference renderer (oracle/ref_render.py) against fixtures rendered by the live reference
(tests/golden/make_golden.py): 0 differing pixels, for the cv2 backend and for the cv2-free C backend."""
import pytest

from conftest import restore_rng
from oracle import ref_render


def _check(panels, pix, backend):
    for c, want in zip(panels, pix):
        restore_rng(c)
        got = ref_render.render_panel(c['W'], c['H'], c['bg'], c['elements'], c['use_grid'], c['grid_size'], backend=backend)
        assert (got == want).all(), [e['kind'] for e in c['elements']]


def test_c_backend_sequence_panels(golden_seq, built):
    _check(*golden_seq, ref_render.CBackend())


def test_c_backend_kind_coverage(golden_kinds, built):
    _check(*golden_kinds, ref_render.CBackend())


def test_cv2_backend(golden_seq, golden_kinds):
    pytest.importorskip('cv2')
    be = ref_render.Cv2Backend()
    _check(golden_seq[0][::2], golden_seq[1][::2], be)
    _check(golden_kinds[0][::2], golden_kinds[1][::2], be)


def test_golden_covers_every_kind_and_wraps(golden_kinds):
    kinds = {e['kind'] for c in golden_kinds[0] for e in c['elements']}
    assert kinds == set(ref_render.KINDS)
    assert any(c['rng_draws'] for c in golden_kinds[0])
    assert any(e['center'][0] < 0 or e['center'][0] > 512 for c in golden_kinds[0] for e in c['elements'])
