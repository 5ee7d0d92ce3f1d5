"""figure_paths.py -- the vertex/code tables matplotlib's patches hand to its renderer, restated.

The reference's multigraph generator builds matplotlib patches (multigraph_generation/shapes.py:21-72) and lets
``fig.savefig`` draw them (multigraph_generation/generator.py:250-259).  What reaches the rasteriser per artist is a
``Path`` (vertices + codes) and an affine; matplotlib is not installed in this image, so the few tables the synthetic
figure scenes need are restated here from matplotlib's published definitions (``matplotlib.path.Path.unit_circle``,
``unit_rectangle``, ``unit_regular_polygon``, ``arc``/``wedge``; ``patches.BoxStyle.Round``).  When matplotlib is
available, ``figure_mpl.extract_figure`` reads the real paths instead and none of this is used.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

STOP, MOVETO, LINETO, CURVE3, CURVE4, CLOSEPOLY = 0, 1, 2, 3, 4, 79

PathT = Tuple[np.ndarray, np.ndarray]   # (float64[n, 2], uint8[n])


def unit_circle() -> PathT:
    """Path.unit_circle(): 8 cubic Beziers, MAGIC = 0.2652031."""
    m = 0.2652031
    s = math.sqrt(0.5)
    k = s * m
    v = [[0.0, -1.0],
         [m, -1.0], [s - k, -s - k], [s, -s],
         [s + k, -s + k], [1.0, -m], [1.0, 0.0],
         [1.0, m], [s + k, s - k], [s, s],
         [s - k, s + k], [m, 1.0], [0.0, 1.0],
         [-m, 1.0], [-s + k, s + k], [-s, s],
         [-s - k, s - k], [-1.0, m], [-1.0, 0.0],
         [-1.0, -m], [-s - k, -s + k], [-s, -s],
         [-s + k, -s - k], [-m, -1.0], [0.0, -1.0],
         [0.0, -1.0]]
    c = [CURVE4] * 26
    c[0] = MOVETO
    c[-1] = CLOSEPOLY
    return np.array(v, dtype=np.float64), np.array(c, dtype=np.uint8)


def unit_rectangle() -> PathT:
    v = [[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.0, 0.0]]
    return np.array(v, dtype=np.float64), np.array([MOVETO, LINETO, LINETO, LINETO, CLOSEPOLY], dtype=np.uint8)


def unit_regular_polygon(n: int) -> PathT:
    theta = (2 * np.pi / n * np.arange(n + 1)) + np.pi / 2
    v = np.column_stack((np.cos(theta), np.sin(theta)))
    c = np.full(n + 1, LINETO, dtype=np.uint8)
    c[0] = MOVETO
    c[-1] = CLOSEPOLY
    return v.astype(np.float64), c


def arc(theta1: float, theta2: float, is_wedge: bool = False) -> PathT:
    """Path.arc / Path.wedge on the unit circle (degrees)."""
    halfpi = np.pi * 0.5
    eta1 = theta1
    eta2 = theta2 - 360 * np.floor((theta2 - theta1) / 360)
    if theta2 != theta1 and eta2 <= eta1:
        eta2 += 360
    eta1, eta2 = np.deg2rad([eta1, eta2])
    n = int(2 ** np.ceil((eta2 - eta1) / halfpi))
    n = max(n, 1)
    deta = (eta2 - eta1) / n
    t = np.tan(0.5 * deta)
    alpha = np.sin(deta) * (np.sqrt(4.0 + 3.0 * t * t) - 1) / 3.0
    steps = np.linspace(eta1, eta2, n + 1, True)
    cos_eta, sin_eta = np.cos(steps), np.sin(steps)
    xA, yA = cos_eta[:-1], sin_eta[:-1]
    xB, yB = cos_eta[1:], sin_eta[1:]
    if is_wedge:
        length = n * 3 + 4
        v = np.zeros((length, 2), np.float64)
        c = np.full(length, CURVE4, dtype=np.uint8)
        v[1] = [xA[0], yA[0]]
        c[0:2] = [MOVETO, LINETO]
        c[-2:] = [LINETO, CLOSEPOLY]
        off, end = 2, length - 2
    else:
        length = n * 3 + 1
        v = np.empty((length, 2), np.float64)
        c = np.full(length, CURVE4, dtype=np.uint8)
        v[0] = [xA[0], yA[0]]
        c[0] = MOVETO
        off, end = 1, length
    v[off:end:3, 0] = xA + alpha * -yA
    v[off:end:3, 1] = yA + alpha * xA
    v[off + 1:end:3, 0] = xB - alpha * -yB
    v[off + 1:end:3, 1] = yB - alpha * xB
    v[off + 2:end:3, 0] = xB
    v[off + 2:end:3, 1] = yB
    return v, c


def wedge_patch(center: Sequence[float], r: float, theta1: float, theta2: float) -> PathT:
    """patches.Wedge._recompute_path with width=None: arc + LINETO centre + CLOSEPOLY, scaled and moved in data space."""
    v1, c1 = arc(theta1, theta2)
    full = abs((theta2 - theta1) - 360) <= 1e-12
    v = np.concatenate([v1, [(0.0, 0.0), (0.0, 0.0)]])
    c = np.concatenate([c1, [MOVETO if full else LINETO, CLOSEPOLY]]).astype(np.uint8)
    v = v * r + np.asarray(center, dtype=np.float64)
    return v, c


def round_box(x0: float, y0: float, width: float, height: float, pad: float, rounding_size: float | None = None) -> PathT:
    """BoxStyle.Round.__call__ (FancyBboxPatch(boxstyle='round')), mutation_size 1."""
    dr = rounding_size if rounding_size else pad
    width, height = width + 2 * pad, height + 2 * pad
    x0, y0 = x0 - pad, y0 - pad
    x1, y1 = x0 + width, y0 + height
    cp = [(x0 + dr, y0), (x1 - dr, y0), (x1, y0), (x1, y0 + dr), (x1, y1 - dr), (x1, y1), (x1 - dr, y1), (x0 + dr, y1),
          (x0, y1), (x0, y1 - dr), (x0, y0 + dr), (x0, y0), (x0 + dr, y0), (x0 + dr, y0)]
    com = [MOVETO, LINETO, CURVE3, CURVE3, LINETO, CURVE3, CURVE3, LINETO, CURVE3, CURVE3, LINETO, CURVE3, CURVE3, CLOSEPOLY]
    return np.array(cp, dtype=np.float64), np.array(com, dtype=np.uint8)


def polyline(points: Sequence[Sequence[float]], closed: bool = False) -> PathT:
    """Path(points) as Line2D / Polygon build it: MOVETO + LINETOs (+ CLOSEPOLY on a repeated first point)."""
    v = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    if closed:
        v = np.concatenate([v, v[:1]])
    c = np.full(len(v), LINETO, dtype=np.uint8)
    c[0] = MOVETO
    if closed:
        c[-1] = CLOSEPOLY
    return v, c


# ---- affine helpers: 3x3 matrices composed the way matplotlib.transforms does (np.dot of float64 matrices) -----------
def affine_identity() -> np.ndarray:
    return np.identity(3, dtype=np.float64)


def affine_scale(sx: float, sy: float | None = None) -> np.ndarray:
    m = affine_identity()
    m[0, 0] = sx
    m[1, 1] = sx if sy is None else sy
    return m


def affine_rotate_deg(deg: float) -> np.ndarray:
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def affine_translate(tx: float, ty: float) -> np.ndarray:
    m = affine_identity()
    m[0, 2] = tx
    m[1, 2] = ty
    return m


def compose(*ms: np.ndarray) -> np.ndarray:
    """compose(a, b, c) applies a first, then b, then c (matplotlib's ``a + b + c``)."""
    out = affine_identity()
    for m in ms:
        out = np.dot(m, out)
    return out


def to_values(m: np.ndarray) -> Tuple[float, float, float, float, float, float]:
    """Affine2D.to_values(): (a, b, c, d, e, f) of [[a, c, e], [b, d, f], [0, 0, 1]]."""
    return (float(m[0, 0]), float(m[1, 0]), float(m[0, 1]), float(m[1, 1]), float(m[0, 2]), float(m[1, 2]))


def data_to_display(W: int, H: int, lim: float = 5.0) -> np.ndarray:
    """transData of the reference's figure: ``plt.subplots(figsize=(8, 8))``, ``set_xlim/ylim(-5, 5)``,
    ``set_aspect('equal')`` (multigraph_generation/generator.py:488-493): default subplot slot [0.125, 0.9] x
    [0.11, 0.88] of the canvas, the axes box shrunk to a square of side 0.77 * H and centred in it."""
    l, r, b, t = 0.125 * W, 0.9 * W, 0.11 * H, 0.88 * H
    side = min(r - l, t - b)
    x0 = l + ((r - l) - side) / 2.0
    y0 = b + ((t - b) - side) / 2.0
    inw = 1.0 / (2 * lim)
    limits = np.array([[inw, 0.0, lim * inw], [0.0, inw, lim * inw], [0.0, 0.0, 1.0]], dtype=np.float64)
    axes = np.array([[side, 0.0, x0], [0.0, side, y0], [0.0, 0.0, 1.0]], dtype=np.float64)
    return np.dot(axes, limits)
