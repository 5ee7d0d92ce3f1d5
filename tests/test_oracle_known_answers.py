"""Known-answer vectors of SURVEY.md section 8(c) (captured from cv2 4.13.0) against oracle/cvdraw.c.
No cv2 needed: these pin the C restatement on boxes without OpenCV."""
import ctypes

import numpy as np

u8p = ctypes.POINTER(ctypes.c_uint8)
i32p = ctypes.POINTER(ctypes.c_int32)
BLACK = (ctypes.c_uint8 * 3)(0, 0, 0)


def canvas(h, w, v=255):
    return np.full((h, w, 3), v, np.uint8)


def P(a):
    return a.ctypes.data_as(u8p)


def test_ka1_thin_aa_line_profile(oracle_lib):
    im = canvas(24, 140)
    oracle_lib.cvd_line(P(im), 140, 24, 0, 10, 128, 12, BLACK, 1, 16, 0)
    a = 255 - im[:, :, 0].astype(int)
    table = {2: (53, 232, 58), 4: (47, 232, 65), 8: (37, 230, 79), 16: (21, 221, 109), 24: (12, 204, 143),
             30: (6, 188, 167), 32: (6, 181, 173)}
    for x, taps in table.items():
        assert tuple(a[9:12, x]) == taps, (x, a[8:13, x])
    assert tuple(a[10:13, 34]) == (173, 181, 6)
    assert tuple(a[10:13, 64]) == (58, 232, 53)


def test_ka1_axis_aligned_and_colour_blend(oracle_lib):
    im = canvas(20, 40)
    oracle_lib.cvd_line(P(im), 40, 20, 5, 10, 30, 10, BLACK, 1, 16, 0)
    assert list(im[8:13, 15, 0]) == [255, 202, 23, 197, 255]
    im = np.empty((20, 40, 3), np.uint8)
    im[:] = (200, 100, 50)
    oracle_lib.cvd_line(P(im), 40, 20, 5, 10, 30, 10, BLACK, 1, 16, 0)
    assert tuple(im[10, 15]) == (18, 9, 5)
    assert tuple(im[9, 15]) == (159, 79, 40) and tuple(im[11, 15]) == (155, 77, 39)
    # direction independence
    a, b = canvas(40, 80), canvas(40, 80)
    for t in (1, 3):
        oracle_lib.cvd_line(P(a), 80, 40, 5, 20, 58, 27, BLACK, t, 16, 0)
        oracle_lib.cvd_line(P(b), 80, 40, 58, 27, 5, 20, BLACK, t, 16, 0)
        assert (a == b).all()


def test_ka2_thick_aa_cross_sections(oracle_lib):
    for t, want in ((2, [202, 0, 0, 0, 197]), (3, [202, 0, 0, 0, 0, 0, 197])):
        im = canvas(30, 60)
        oracle_lib.cvd_line(P(im), 60, 30, 5, 15, 50, 15, BLACK, t, 16, 0)
        col = list(im[:, 25, 0])
        ink = [v for v in col if v != 255]
        assert ink == want, (t, ink)
    im = canvas(30, 60)
    oracle_lib.cvd_line(P(im), 60, 30, 5, 15, 50, 15, BLACK, 3, 16, 0)
    assert list(im[15, 0:5, 0]) == [255, 255, 255, 60, 0]   # round (diamond) cap reaches x = 3


def test_ka3_filled_circle_is_exact_disc(oracle_lib):
    for r, area in ((3, 29), (5, 81), (43, 5789), (51, 8173), (85, 22701)):
        S = 2 * r + 7
        im = canvas(S, S)
        c = r + 3
        oracle_lib.cvd_circle(P(im), S, S, c, c, r, BLACK, -1, 8)
        got = im[:, :, 0] == 0
        yy, xx = np.mgrid[0:S, 0:S]
        assert (got == ((xx - c) ** 2 + (yy - c) ** 2 <= r * r)).all()
        assert got.sum() == area


def test_ka4_fillpoly_includes_boundary(oracle_lib):
    im = canvas(16, 16)
    pts = np.array([[3, 3], [12, 3], [12, 12], [3, 12]], np.int32)
    oracle_lib.cvd_fill_poly(P(im), 16, 16, pts.ctypes.data_as(i32p), 4, BLACK)
    assert (im[:, :, 0] == 0).sum() == 100
    im = canvas(16, 16)
    pts = np.array([[2, 12], [8, 2], [14, 12]], np.int32)
    oracle_lib.cvd_fill_poly(P(im), 16, 16, pts.ctypes.data_as(i32p), 3, BLACK)
    widths = [(im[y, :, 0] == 0).sum() for y in range(2, 13)]
    assert widths == [1, 3, 3, 5, 5, 7, 9, 9, 11, 11, 13]


def test_ka5_aa_circle_keeps_subpixel_vertices(oracle_lib):
    """cv2.circle(r=51,t=2,LINE_AA) is not the int-rounded 72-gon: EllipseEx keeps 16.16 vertices."""
    im = canvas(128, 128)
    oracle_lib.cvd_circle(P(im), 128, 128, 64, 64, 51, BLACK, 2, 16)
    ang = np.deg2rad(np.arange(0, 360, 5))
    poly = np.stack([np.round(64 + 51 * np.cos(ang)), np.round(64 + 51 * np.sin(ang))], 1).astype(np.int32)
    im2 = canvas(128, 128)
    oracle_lib.cvd_polylines(P(im2), 128, 128, poly.ctypes.data_as(i32p), len(poly), 1, BLACK, 2, 16)
    assert (im != im2).any(axis=2).sum() > 500
