"""Host-side scene emitter: reference element dicts -> struct-of-arrays scene description.

Mirrors the argument handling of RPMGenerator._render_elements_to_canvas
(/root/reference/src/generator.py:85-94) and the colour fallback of Shape.draw
(/root/reference/src/shapes.py:233): the fields read, their defaults, the int()/bool() coercions and
the rand_color() draw for colourless elements all happen here, at call time and in element order, so
the process-global `random` stream advances exactly as it does in the reference.
Everything else (grid snap, vertex math, rasterisation) happens behind the C ABI.
"""
from __future__ import annotations

import ctypes
import random
from typing import Any, Dict, Iterable, List, Sequence

import numpy as np

from ._abi import RigScene

KINDS = ('square', 'circle', 'triangle', 'diamond', 'star', 'pentagon', 'hexagon', 'plus', 'heart',
         'crescent', 'rounded_square')
_KIND_INDEX = {k: i for i, k in enumerate(KINDS)}
_FLIP_BITS = {None: 0, '': 0, 'horizontal': 1, 'vertical': 2, 'both': 3}


def rand_color(min_v=30, max_v=220):
    """utils.rand_color (/root/reference/src/utils.py:13-15): 3 draws from the global stream."""
    return tuple(int(random.uniform(min_v, max_v)) for _ in range(3))


def _sat8(v) -> int:
    v = int(v)
    return 0 if v < 0 else 255 if v > 255 else v


class SceneBuilder:
    """Accumulates panels; `finish()` returns a `Scene` whose arrays back a `rig_scene` struct."""

    def __init__(self, W: int = 512, H: int = 512, bg_color: Sequence[int] = (255, 255, 255), grid_size: int = 3):
        self.W, self.H = int(W), int(H)
        self.bg_color = tuple(int(c) for c in bg_color)
        self.grid_size = int(grid_size)
        self._begin: List[int] = [0]
        self._flags: List[int] = []
        self._kind: List[int] = []
        self._fill: List[int] = []
        self._flip: List[int] = []
        self._stroke: List[int] = []
        self._size: List[int] = []
        self._angle: List[int] = []
        self._cx: List[float] = []
        self._cy: List[float] = []
        self._rgb: List[int] = []

    def add_panel(self, elements: Iterable[Dict[str, Any]], use_grid: bool = False) -> int:
        W, H = self.W, self.H
        for el in elements:
            kind = el['kind']
            size = int(el['size'])
            fill = bool(el['fill'])
            stroke = int(el.get('stroke_width', 1))
            cx, cy = el.get('center', (W // 2, H // 2))
            color = el.get('color')
            angle = int(el.get('angle', 0))
            flip_mode = el.get('flip_mode')
            if kind not in _KIND_INDEX:
                raise AssertionError(f'unsupported shape: {kind}. supported={KINDS}')
            rgb = color if color is not None else rand_color()
            self._kind.append(_KIND_INDEX[kind])
            self._size.append(size)
            self._fill.append(1 if fill else 0)
            self._stroke.append(stroke)
            self._cx.append(float(cx))
            self._cy.append(float(cy))
            self._angle.append(angle)
            self._flip.append(_FLIP_BITS.get(flip_mode, 0))
            self._rgb.extend((_sat8(rgb[0]), _sat8(rgb[1]), _sat8(rgb[2])))
        self._begin.append(len(self._kind))
        self._flags.append(1 if use_grid else 0)
        return len(self._flags) - 1

    def finish(self) -> 'Scene':
        return Scene(self.W, self.H, self.bg_color, self.grid_size,
                     np.asarray(self._begin, dtype=np.uint32), np.asarray(self._flags, dtype=np.uint8),
                     np.asarray(self._kind, dtype=np.uint8), np.asarray(self._fill, dtype=np.uint8),
                     np.asarray(self._flip, dtype=np.uint8), np.asarray(self._stroke, dtype=np.int32),
                     np.asarray(self._size, dtype=np.int32), np.asarray(self._angle, dtype=np.int32),
                     np.asarray(self._cx, dtype=np.float64), np.asarray(self._cy, dtype=np.float64),
                     np.asarray(self._rgb, dtype=np.uint8))


class Scene:
    """Owns the numpy arrays of one batch and exposes them as a ctypes `rig_scene`."""

    def __init__(self, W, H, bg_color, grid_size, begin, flags, kind, fill, flip, stroke, size, angle, cx, cy, rgb):
        self.W, self.H, self.bg_color, self.grid_size = W, H, tuple(bg_color), grid_size
        self.begin = np.ascontiguousarray(begin, dtype=np.uint32)
        self.flags = np.ascontiguousarray(flags, dtype=np.uint8)
        self.kind = np.ascontiguousarray(kind, dtype=np.uint8)
        self.fill = np.ascontiguousarray(fill, dtype=np.uint8)
        self.flip = np.ascontiguousarray(flip, dtype=np.uint8)
        self.stroke = np.ascontiguousarray(stroke, dtype=np.int32)
        self.size = np.ascontiguousarray(size, dtype=np.int32)
        self.angle = np.ascontiguousarray(angle, dtype=np.int32)
        self.cx = np.ascontiguousarray(cx, dtype=np.float64)
        self.cy = np.ascontiguousarray(cy, dtype=np.float64)
        self.rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1)
        self.n_panels = int(len(self.flags))
        self.n_elements = int(len(self.kind))
        if len(self.begin) != self.n_panels + 1 or int(self.begin[-1]) != self.n_elements:
            raise ValueError('panel_elem_begin does not describe the element arrays')
        if len(self.rgb) != 3 * self.n_elements:
            raise ValueError('rgb must hold 3 bytes per element')

    def as_struct(self) -> RigScene:
        def p(a, t):
            return a.ctypes.data_as(ctypes.POINTER(t))
        s = RigScene()
        s.n_panels, s.n_elements, s.grid_size = self.n_panels, self.n_elements, self.grid_size
        s.bg_rgb[0], s.bg_rgb[1], s.bg_rgb[2] = (_sat8(c) for c in self.bg_color)
        s.panel_elem_begin = p(self.begin, ctypes.c_uint32)
        s.panel_flags = p(self.flags, ctypes.c_uint8)
        s.kind = p(self.kind, ctypes.c_uint8)
        s.fill = p(self.fill, ctypes.c_uint8)
        s.flip = p(self.flip, ctypes.c_uint8)
        s.stroke = p(self.stroke, ctypes.c_int32)
        s.size = p(self.size, ctypes.c_int32)
        s.angle = p(self.angle, ctypes.c_int32)
        s.cx = p(self.cx, ctypes.c_double)
        s.cy = p(self.cy, ctypes.c_double)
        s.rgb = p(self.rgb, ctypes.c_uint8)
        return s

    def slice(self, p0: int, p1: int) -> 'Scene':
        """Panels [p0, p1) as a new Scene (used to shard a batch across ranks)."""
        e0, e1 = int(self.begin[p0]), int(self.begin[p1])
        return Scene(self.W, self.H, self.bg_color, self.grid_size, self.begin[p0:p1 + 1] - self.begin[p0],
                     self.flags[p0:p1], self.kind[e0:e1], self.fill[e0:e1], self.flip[e0:e1], self.stroke[e0:e1],
                     self.size[e0:e1], self.angle[e0:e1], self.cx[e0:e1], self.cy[e0:e1], self.rgb[3 * e0:3 * e1])
