"""oracle/ref_render.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's sequence-panel renderer:

    RPMGenerator._render_elements_to_canvas      /root/reference/src/generator.py:66-124
    Shape.draw (antialias_mode='fast', scale=1)  /root/reference/src/shapes.py:209-667
    rand_color                                   /root/reference/src/utils.py:13-15

The reference's algorithm is "build float vertices, rotate by -angle, round half-even to int32, then
call cv2 drawing primitives in a fixed order".  This file restates that as two steps so each can be
checked on its own:

  1. `panel_ops`   element dicts -> ordered list of integer drawing ops (pure Python/float64 math);
  2. a backend executes the ops:  `Cv2Backend` calls the installed cv2 wheel exactly the way the
     reference does (this is the CPU baseline "port": same third-party rasteriser, same per-element
     canvas copy);  `CBackend` calls oracle/cvdraw.c, the C restatement of that cv2 arithmetic.

Parity status: PINNED.  tools/make_golden.py imports the live reference from /root/reference/src in
the build container and freezes (elements, use_grid) -> pixels fixtures under tests/golden/; the
tests check this file (both backends) against those fixtures with zero differing pixels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path never does.
"""
from __future__ import annotations

import ctypes
import math
import os
import random
import subprocess
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

KINDS = ('square', 'circle', 'triangle', 'diamond', 'star', 'pentagon', 'hexagon', 'plus',
         'heart', 'crescent', 'rounded_square')

Op = Tuple  # ('fill_poly', pts, bgr) | ('stroke_poly', pts, t) | ('disc', cx, cy, r, bgr) |
#             ('ring', cx, cy, r, t) | ('crescent', ocx, ocy, orad, icx, icy, irad, bgr) |
#             ('line8', x1, y1, x2, y2)


def rand_color(min_v=30, max_v=220):
    """utils.py:13-15 -- three draws from the process-global `random` stream."""
    return tuple(int(random.uniform(min_v, max_v)) for _ in range(3))


# ----------------------------------------------------------------------------------------------
# step 1: geometry.  Every expression keeps the reference's operation order (float64, no fusing).
# ----------------------------------------------------------------------------------------------
def _regular(n: int, half: float, step_deg: float) -> List[Tuple[float, float]]:
    # shapes.py:431-435 ('star', step 72) and :456-460 (pentagon/hexagon, step 360.0/n)
    out = []
    for i in range(n):
        a = math.radians(i * step_deg - 90)
        out.append((half * math.cos(a), half * math.sin(a)))
    return out


def outlines(kind: str, s: int) -> List[List[Tuple[float, float]]]:
    """Centre-relative, unrotated vertex loops of one polygon kind (shapes.py:358-375,390-544,596-656)."""
    half = s / 2.0
    if kind == 'square':
        return [[(-half, -half), (half, -half), (half, half), (-half, half)]]
    if kind == 'triangle':
        return [[(-half, half), (0, -half), (half, half)]]
    if kind == 'diamond':
        return [[(0, -half), (half, 0), (0, half), (-half, 0)]]
    if kind == 'star':
        return [_regular(5, half, 72)]
    if kind == 'pentagon':
        return [_regular(5, half, 360.0 / 5)]
    if kind == 'hexagon':
        return [_regular(6, half, 360.0 / 6)]
    if kind == 'plus':
        arm = int(round(s * 0.25))
        length = int(round(s * 0.9))
        v = [(-arm / 2, -length / 2), (arm / 2, -length / 2), (arm / 2, length / 2), (-arm / 2, length / 2)]
        h = [(-length / 2, -arm / 2), (length / 2, -arm / 2), (length / 2, arm / 2), (-length / 2, arm / 2)]
        return [v, h]
    if kind == 'heart':
        r = half * 0.8
        pts = []
        for t in np.linspace(0, 2 * math.pi, num=60):
            x = 16 * (math.sin(t) ** 3)
            y = 13 * math.cos(t) - 5 * math.cos(2 * t) - 2 * math.cos(3 * t) - math.cos(4 * t)
            scale = r / 16
            pts.append((x * scale, -y * scale))
        return [pts]
    if kind == 'rounded_square':
        r = int(round(half * 0.4))
        r = min(r, int(round(half - 1)))
        lo, hi = -half + r, half - r
        pts = []
        for th in np.linspace(math.pi, 3 * math.pi / 2, num=12):
            pts.append((lo + r * math.cos(th), lo + r * math.sin(th)))
        pts.append((hi, lo - r))
        for th in np.linspace(3 * math.pi / 2, 2 * math.pi, num=12):
            pts.append((hi + r * math.cos(th), lo + r * math.sin(th)))
        pts.append((hi + r, hi))
        for th in np.linspace(0, math.pi / 2, num=12):
            pts.append((hi + r * math.cos(th), hi + r * math.sin(th)))
        pts.append((lo, hi + r))
        for th in np.linspace(math.pi / 2, math.pi, num=12):
            pts.append((lo + r * math.cos(th), hi + r * math.sin(th)))
        pts.append((lo - r, lo))
        return [pts]
    raise ValueError(kind)


def _place(points, angle_deg, flip_mode, cx, cy) -> np.ndarray:
    # shapes.py:238-258 (rotate by -angle, optional flip) and :278-280 (round half-even to int32)
    a = math.radians(-angle_deg)
    ca, sa = math.cos(a), math.sin(a)
    out = []
    for (x, y) in points:
        xr, yr = x * ca - y * sa, x * sa + y * ca
        if flip_mode in ('horizontal', 'both'):
            xr = -xr
        if flip_mode in ('vertical', 'both'):
            yr = -yr
        out.append((int(round(cx + xr)), int(round(cy + yr))))
    return np.array(out, dtype=np.int32).reshape(-1, 2)


def _copies(minx, maxx, miny, maxy, W, H, origin_always: bool):
    """Wrap-around copy offsets in the reference's loop order (shapes.py:260-266,369-375,385-388)."""
    for ox in (-W, 0, W):
        for oy in (-H, 0, H):
            if origin_always and ox == 0 and oy == 0:
                yield 0, 0
                continue
            if maxx + ox < 0 or minx + ox >= W:
                continue
            if maxy + oy < 0 or miny + oy >= H:
                continue
            yield ox, oy


def element_ops(kind: str, size: int, fill: bool, stroke_width: int, cx: int, cy: int, angle: int,
                rgb: Sequence[int], flip_mode: Optional[str], W: int, H: int) -> List[Op]:
    """One element -> ordered integer drawing ops (shapes.py:269-656 with ratio == 1)."""
    if kind not in KINDS:
        raise AssertionError(f'unsupported shape: {kind}')
    bgr = (int(rgb[2]), int(rgb[1]), int(rgb[0]))
    t = max(1, int(round(stroke_width * 1.0)))
    s = int(round(size * 1.0))
    half = s / 2.0
    ops: List[Op] = []

    if kind == 'circle':
        rad = max(1, int(round(half)))
        for ox, oy in _copies(cx - rad, cx + rad, cy - rad, cy + rad, W, H, origin_always=False):
            x, y = int(round(cx + ox)), int(round(cy + oy))
            if fill:
                ops.append(('disc', x, y, rad, bgr))
            ops.append(('ring', x, y, rad, t))
        return ops

    if kind == 'crescent':
        orad = max(1, int(round(half)))
        irad = int(round(orad * 0.65))
        off = int(round(orad * 0.35))
        a = math.radians(-angle)
        rx = off * math.cos(a) - 0 * math.sin(a)
        ry = off * math.sin(a) + 0 * math.cos(a)
        icx, icy = int(round(cx + rx)), int(round(cy + ry))
        if fill:
            ops.append(('crescent', cx, cy, orad, icx, icy, irad, bgr))
        ops.append(('ring', cx, cy, orad, t))
        ops.append(('ring', icx, icy, irad, t))
        for ox, oy in _copies(cx - orad, cx + orad, cy - orad, cy + orad, W, H, origin_always=False):
            if ox == 0 and oy == 0:
                continue
            ocx, ocy = cx + ox, cy + oy
            # wrap copies are blended even for unfilled crescents, and get no outline
            ops.append(('crescent', ocx, ocy, orad, int(round(ocx + rx)), int(round(ocy + ry)), irad, bgr))
        return ops

    loops = [_place(p, angle, flip_mode, cx, cy) for p in outlines(kind, s)]
    allp = np.concatenate(loops, axis=0)
    minx, maxx = int(allp[:, 0].min()), int(allp[:, 0].max())
    miny, maxy = int(allp[:, 1].min()), int(allp[:, 1].max())
    for ox, oy in _copies(minx, maxx, miny, maxy, W, H, origin_always=True):
        for loop in loops:
            pts = loop + np.array([ox, oy], dtype=np.int32)
            if fill:
                ops.append(('fill_poly', pts, bgr))
            ops.append(('stroke_poly', pts, t))
    return ops


def panel_ops(W: int, H: int, elements: Iterable[Dict[str, Any]], use_grid: bool = False,
              grid_size: int = 3) -> List[List[Op]]:
    """generator.py:85-122 -> per-element op lists (+ a final list holding the grid lines).

    Consumes the global `random` stream exactly where the reference does (colourless elements)."""
    cell_w = W / grid_size
    cell_h = H / grid_size if use_grid else (0, 0)
    per_element: List[List[Op]] = []
    for el in elements:
        kind = el['kind']
        size = int(el['size'])
        fill = bool(el['fill'])
        stroke = int(el.get('stroke_width', 1))
        cx, cy = el.get('center', (W // 2, H // 2))
        color = el.get('color')
        angle = int(el.get('angle', 0))
        flip_mode = el.get('flip_mode')
        if use_grid:
            col = min(grid_size - 1, max(0, int(cx // cell_w)))
            row = min(grid_size - 1, max(0, int(cy // cell_h)))
            cx, cy = int((col + 0.5) * cell_w), int((row + 0.5) * cell_h)
        cx, cy = int(cx), int(cy)
        rgb = color if color is not None else rand_color()
        per_element.append(element_ops(kind, size, fill, stroke, cx, cy, angle, rgb, flip_mode, W, H))
    grid: List[Op] = []
    if use_grid:
        for i in range(1, grid_size):
            x = int(round(i * W / grid_size))
            grid.append(('line8', x, 0, x, H))
        for j in range(1, grid_size):
            y = int(round(j * H / grid_size))
            grid.append(('line8', 0, y, W, y))
    per_element.append(grid)
    return per_element


# ----------------------------------------------------------------------------------------------
# step 2: backends
# ----------------------------------------------------------------------------------------------
class Cv2Backend:
    """Executes ops with the installed cv2 wheel, call for call as the reference does."""

    name = 'cv2'

    def __init__(self):
        import cv2
        self.cv2 = cv2

    def begin_element(self, canvas: np.ndarray) -> np.ndarray:
        return canvas.copy()  # shapes.py:229 -> to_cv2() copies the whole canvas per element

    def run(self, canvas: np.ndarray, ops: List[Op]) -> np.ndarray:
        cv2 = self.cv2
        black = (0, 0, 0)
        for op in ops:
            k = op[0]
            if k == 'fill_poly':
                cv2.fillPoly(canvas, [op[1]], op[2])
            elif k == 'stroke_poly':
                cv2.polylines(canvas, [op[1]], isClosed=True, color=black, thickness=op[2], lineType=cv2.LINE_AA)
            elif k == 'disc':
                cv2.circle(canvas, (op[1], op[2]), op[3], op[4], thickness=-1)
            elif k == 'ring':
                cv2.circle(canvas, (op[1], op[2]), op[3], black, thickness=op[4], lineType=cv2.LINE_AA)
            elif k == 'crescent':
                # shapes.py:563-570 -- uint8 mask from two LINE_8 discs, then a float32 select
                _, ocx, ocy, orad, icx, icy, irad, bgr = op
                mask = np.zeros(canvas.shape[:2], dtype=np.uint8)
                cv2.circle(mask, (ocx, ocy), orad, 255, thickness=-1)
                cv2.circle(mask, (icx, icy), irad, 0, thickness=-1)
                layer = np.full_like(canvas, bgr, dtype=np.uint8)
                alpha = (mask.astype(np.float32) / 255.0)[:, :, None]
                canvas[:] = (alpha * layer + (1.0 - alpha) * canvas).astype(np.uint8)
            elif k == 'line8':
                cv2.line(canvas, (op[1], op[2]), (op[3], op[4]), black, 1)
            else:
                raise ValueError(k)
        return canvas


_HERE = os.path.dirname(os.path.abspath(__file__))


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/cvdraw.c -> oracle/_build/libcvdraw.so (gcc only; building is not using)."""
    out = os.path.join(_HERE, '_build', 'libcvdraw.so')
    src = os.path.join(_HERE, 'cvdraw.c')
    hdr = os.path.join(_HERE, 'cvdraw_tables.h')
    if force or not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-fno-fast-math', '-ffp-contract=off',
                               '-o', out, src, '-lm'])
    return out


class CBackend:
    """Executes ops with oracle/cvdraw.c (no cv2 anywhere)."""

    name = 'c'

    def __init__(self):
        self.lib = ctypes.CDLL(build_c_oracle())
        self._u8p = ctypes.POINTER(ctypes.c_uint8)
        self._i32p = ctypes.POINTER(ctypes.c_int32)

    def begin_element(self, canvas: np.ndarray) -> np.ndarray:
        return canvas

    def _img(self, a):
        return a.ctypes.data_as(self._u8p), a.shape[1], a.shape[0]

    def run(self, canvas: np.ndarray, ops: List[Op]) -> np.ndarray:
        lib = self.lib
        black = (ctypes.c_uint8 * 3)(0, 0, 0)
        p, w, h = self._img(canvas)
        for op in ops:
            k = op[0]
            if k == 'fill_poly':
                pts = np.ascontiguousarray(op[1], dtype=np.int32)
                lib.cvd_fill_poly(p, w, h, pts.ctypes.data_as(self._i32p), len(pts), (ctypes.c_uint8 * 3)(*op[2]))
            elif k == 'stroke_poly':
                pts = np.ascontiguousarray(op[1], dtype=np.int32)
                lib.cvd_polylines(p, w, h, pts.ctypes.data_as(self._i32p), len(pts), 1, black, op[2], 16)
            elif k == 'disc':
                lib.cvd_circle(p, w, h, op[1], op[2], op[3], (ctypes.c_uint8 * 3)(*op[4]), -1, 8)
            elif k == 'ring':
                lib.cvd_circle(p, w, h, op[1], op[2], op[3], black, op[4], 16)
            elif k == 'crescent':
                _, ocx, ocy, orad, icx, icy, irad, bgr = op
                mask = np.zeros((h, w, 3), dtype=np.uint8)
                mp = mask.ctypes.data_as(self._u8p)
                lib.cvd_circle(mp, w, h, ocx, ocy, orad, (ctypes.c_uint8 * 3)(255, 255, 255), -1, 8)
                lib.cvd_circle(mp, w, h, icx, icy, irad, black, -1, 8)
                canvas[mask[:, :, 0] == 255] = bgr  # alpha is exactly 0 or 1
            elif k == 'line8':
                lib.cvd_line(p, w, h, op[1], op[2], op[3], op[4], black, 1, 8, 0)
            else:
                raise ValueError(k)
        return canvas


def render_panel(W: int, H: int, bg_color: Sequence[int], elements: Iterable[Dict[str, Any]],
                 use_grid: bool = False, grid_size: int = 3, backend=None) -> np.ndarray:
    """Restated RPMGenerator._render_elements_to_canvas -> uint8[H, W, 3] BGR."""
    backend = backend or Cv2Backend()
    bgr = (int(bg_color[2]), int(bg_color[1]), int(bg_color[0]))
    canvas = np.full((H, W, 3), bgr, dtype=np.uint8)
    groups = panel_ops(W, H, elements, use_grid, grid_size)
    for ops in groups[:-1]:
        canvas = backend.begin_element(canvas)
        canvas = backend.run(canvas, ops)
    return backend.run(canvas, groups[-1])
