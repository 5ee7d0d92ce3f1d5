"""Differential pinning of oracle/cvdraw.c against the installed cv2 wheel (the reference's actual
rasteriser: /root/reference/src/shapes.py:289-302).  Zero differing pixels is required."""
import ctypes

import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')
u8p = ctypes.POINTER(ctypes.c_uint8)
i32p = ctypes.POINTER(ctypes.c_int32)
W = H = 128


def P(a):
    return a.ctypes.data_as(u8p)


def col3(c):
    return (ctypes.c_uint8 * 3)(*c)


def bg(rng):
    return rng.integers(0, 256, (H, W, 3), dtype=np.uint8) if rng.random() < 0.5 else np.full((H, W, 3), 255, np.uint8)


def rc(rng):
    return tuple(int(v) for v in rng.integers(0, 256, 3))


@pytest.mark.parametrize('seed', [0, 1])
def test_lines(oracle_lib, seed):
    rng = np.random.default_rng(seed)
    for _ in range(400):
        a = bg(rng); b = a.copy()
        shift = int(rng.choice([0, 0, 4, 16])); s = 1 << shift
        p = [int(v) for v in rng.integers(-40 * s, (W + 40) * s, 4)]
        c = rc(rng); th = int(rng.choice([1, 1, 2, 3, 4]))
        cv2.line(a, (p[0], p[1]), (p[2], p[3]), c, th, cv2.LINE_AA, shift)
        oracle_lib.cvd_line(P(b), W, H, *p, col3(c), th, 16, shift)
        assert (a == b).all(), (p, c, th, shift)
        a = bg(rng); b = a.copy()
        p = [int(v) for v in rng.integers(-60, W + 60, 4)]
        cv2.line(a, (p[0], p[1]), (p[2], p[3]), c, 1, cv2.LINE_8)
        oracle_lib.cvd_line(P(b), W, H, *p, col3(c), 1, 8, 0)
        assert (a == b).all(), p


@pytest.mark.parametrize('seed', [0, 1])
def test_fillpoly_and_polylines(oracle_lib, seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(400):
        a = bg(rng); b = a.copy()
        n = int(rng.integers(3, 12))
        pts = rng.integers(-50, W + 50, (n, 2)).astype(np.int32)
        c = rc(rng)
        cv2.fillPoly(a, [pts], c)
        oracle_lib.cvd_fill_poly(P(b), W, H, pts.ctypes.data_as(i32p), n, col3(c))
        assert (a == b).all(), pts.tolist()
        a = bg(rng); b = a.copy()
        n = int(rng.integers(2, 9))
        pts = rng.integers(-30, W + 30, (n, 2)).astype(np.int32)
        th = int(rng.choice([1, 2, 3]))
        cv2.polylines(a, [pts], True, c, th, cv2.LINE_AA)
        oracle_lib.cvd_polylines(P(b), W, H, pts.ctypes.data_as(i32p), n, 1, col3(c), th, 16)
        assert (a == b).all(), (pts.tolist(), th)


def test_fillpoly_border_degenerate_clip(oracle_lib):
    """An edge that clips to a single border point becomes a vertical edge on the border column."""
    for pts in ([[43, 25], [31, 30], [43, 30]], [[43, 5], [31, 30], [43, 30]], [[-10, 5], [0, 18], [-10, 30]],
                [[43, -5], [31, 18], [43, 40]], [[43, 5], [29, 30], [43, 30]]):
        a = np.full((32, 32, 3), 255, np.uint8); b = a.copy()
        p = np.array(pts, np.int32)
        cv2.fillPoly(a, [p], (0, 0, 0))
        oracle_lib.cvd_fill_poly(P(b), 32, 32, p.ctypes.data_as(i32p), len(p), col3((0, 0, 0)))
        assert (a == b).all(), pts


def test_circles(oracle_lib):
    rng = np.random.default_rng(7)
    for _ in range(400):
        a = bg(rng); b = a.copy()
        cx, cy = [int(v) for v in rng.integers(-40, W + 40, 2)]
        r = int(rng.integers(0, 90)); c = rc(rng); th = int(rng.choice([1, 2, 3]))
        cv2.circle(a, (cx, cy), r, c, -1)
        oracle_lib.cvd_circle(P(b), W, H, cx, cy, r, col3(c), -1, 8)
        assert (a == b).all(), (cx, cy, r)
        cv2.circle(a, (cx, cy), r, (0, 0, 0), th, cv2.LINE_AA)
        oracle_lib.cvd_circle(P(b), W, H, cx, cy, r, col3((0, 0, 0)), th, 16)
        assert (a == b).all(), (cx, cy, r, th)


def test_filled_circle_is_exact_disc_for_all_radii():
    """The closed form the CUDA DISC record uses: Circle(fill) == {dx^2 + dy^2 <= r^2} for r in 0..399."""
    for r in list(range(0, 130)) + [170, 256, 399]:
        S = 2 * r + 5; c = r + 2
        a = np.zeros((S, S), np.uint8)
        cv2.circle(a, (c, c), r, 255, -1)
        yy, xx = np.mgrid[0:S, 0:S]
        assert ((a > 0) == ((xx - c) ** 2 + (yy - c) ** 2 <= r * r)).all(), r


def test_convex_aa(oracle_lib):
    rng = np.random.default_rng(9)
    for _ in range(400):
        a = bg(rng); b = a.copy()
        shift = int(rng.choice([0, 16])); s = 1 << shift
        ang = np.sort(rng.random(int(rng.integers(3, 7))) * 2 * np.pi)
        cx, cy = rng.integers(-10, W + 10, 2); r = rng.integers(2, 50)
        pts = np.stack([(cx + r * np.cos(ang)) * s, (cy + r * np.sin(ang)) * s], 1).astype(np.int32)
        c = rc(rng)
        cv2.fillConvexPoly(a, pts, c, cv2.LINE_AA, shift)
        oracle_lib.cvd_fill_convex_poly(P(b), W, H, pts.ctypes.data_as(i32p), len(pts), col3(c), 16, shift)
        assert (a == b).all(), (pts.tolist(), shift)


def test_tables_match_installed_wheel():
    """cvdraw_tables.h == the tables inside the cv2 binary the other tests run against."""
    import re, os
    from oracle.tools.extract_cv_tables import extract
    slope, filt, sin = extract()
    txt = open(os.path.join(os.path.dirname(__file__), '..', 'oracle', 'cvdraw_tables.h')).read()
    got_slope = [int(v) for v in re.search(r'CVD_SLOPE_CORR\[32\] = \{([^}]*)\}', txt).group(1).split(',')]
    got_filt = [int(v) for v in re.search(r'CVD_FILTER\[64\] = \{([^}]*)\}', txt).group(1).split(',')]
    assert got_slope == slope and got_filt == filt
    body = re.search(r'CVD_SIN\[451\] = \{([^}]*)\}', txt, re.S).group(1)
    vals = [np.float32(v.strip().rstrip('f')) for v in body.split(',') if v.strip()]
    assert len(vals) == 451 and all(np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32) or (a == 0 and b == 0) for a, b in zip(vals, sin))
