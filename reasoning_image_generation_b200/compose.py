"""compose.py -- host side of the overview-image compositor (mirror of /root/reference/src/layout.py).

`compose_grid` (layout.py:49-192) pastes every panel of a sample, down-scaled with cv2.resize(INTER_AREA) (fit_into_cell,
layout.py:17-31), into a two-row overview and draws borders, labels and the '?' query cell around them.  Only the resize
depends on the sample; `GridLayout` therefore runs the reference's own drawing calls ONCE per layout on the host (cv2 is a
host dependency of the reference) to get a constant overlay and a per-pixel cell map, and `GridCompositor` (librig.so,
include/rig_grid.h, kernel k_grid) composes whole batches on the device from panels that are already in HBM.
`cells_meta` keeps the reference's metadata layout (bbox per cell).
"""
from __future__ import annotations

import ctypes
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi


def _make_query_image(cell_size: int, color=(0, 0, 0)) -> np.ndarray:
    """layout.make_query_image (layout.py:33-47): the same cv2 calls, in the same order."""
    import cv2
    im = np.full((cell_size, cell_size, 3), 255, dtype=np.uint8)
    font = cv2.FONT_HERSHEY_SIMPLEX
    thickness = max(1, cell_size // 80)
    font_scale = cell_size / 100.0
    (tw, th), _ = cv2.getTextSize('?', font, font_scale, thickness)
    if tw > cell_size * 0.8:
        font_scale = font_scale * (cell_size * 0.8 / tw)
        (tw, th), _ = cv2.getTextSize('?', font, font_scale, thickness)
    x = (cell_size - tw) // 2
    y = (cell_size + th) // 2
    cv2.putText(im, '?', (x, y), font, font_scale, color, thickness, lineType=cv2.LINE_AA)
    return im


class GridLayout:
    """Geometry + constant pixels of compose_grid for `n_states` sequence panels (+ the query cell) over `num_options`
    option panels on a W-wide canvas.  Cell k < n_states is state k, cell n_states + j is option j."""

    def __init__(self, W: int = 512, H: int = 512, n_states: int = 4, num_options: int = 4, margin: int = 20, padding_v: int = 20,
                 show_labels: bool = True, show_border: bool = True, bg_color: Sequence[int] = (255, 255, 255), panel_size: Optional[int] = None):
        self.W, self.H, self.n_states, self.num_options = int(W), int(H), int(n_states), int(num_options)
        self.panel_size = int(panel_size if panel_size is not None else W)
        cols_seq, cols_opt = self.n_states + 1, self.num_options
        max_cell_w = (W - 2 * margin) // max(1, max(cols_seq, cols_opt))
        max_cell_h = (H - 2 * margin - padding_v) // 2
        self.cell_size = cell = max(1, min(max_cell_w, max_cell_h))
        self.grid_h = 2 * cell + padding_v + 2 * margin
        self.grid_w = W
        seq_off, opt_off = (W - cols_seq * cell) // 2, (W - cols_opt * cell) // 2
        top_y, bottom_y = margin, margin + cell + padding_v
        self.cells: List[Tuple[int, int]] = [(seq_off + i * cell, top_y) for i in range(self.n_states)] + \
                                            [(opt_off + i * cell, bottom_y) for i in range(cols_opt)]
        self.query_xy = (seq_off + self.n_states * cell, top_y)
        self.cells_meta: List[Dict] = []
        # the reference's drawing, with the panel cells filled by a sentinel: run twice with different sentinels to tell
        # constant pixels (overlay) from panel pixels, and to refuse layouts where a label blends over panel content
        runs = [self._draw(s, margin, show_labels, show_border, bg_color, record_meta=(k == 0)) for k, s in enumerate(((1, 2, 3), (250, 251, 252)))]
        a, b = runs
        is_a, is_b = (a == (1, 2, 3)).all(axis=2), (b == (250, 251, 252)).all(axis=2)
        same = (a == b).all(axis=2)
        if not ((is_a & is_b) | same).all():
            raise ValueError('this layout draws anti-aliased text over panel pixels; the overlay is not constant')
        self.overlay = a.copy()                                     # BGR, like the reference's canvas
        self.overlay[is_a & is_b] = 255
        self.cell_of_pixel = np.full((self.grid_h, self.grid_w), -1, dtype=np.int8)
        for k, (cx, cy) in enumerate(self.cells):
            sub = (is_a & is_b)[cy:cy + cell, cx:cx + cell]
            self.cell_of_pixel[cy:cy + cell, cx:cx + cell][sub] = k
        if (self.cell_of_pixel >= 0).sum() != (is_a & is_b).sum():
            raise ValueError('panel pixels outside the panel cells')

    def _draw(self, sentinel, margin, show_labels, show_border, bg_color, record_meta):
        import cv2
        cell = self.cell_size
        bgr = (bg_color[2], bg_color[1], bg_color[0])
        grid = np.full((self.grid_h, self.grid_w, 3), bgr, dtype=np.uint8)
        font = cv2.FONT_HERSHEY_SIMPLEX
        font_scale = max(0.25, min(0.8, cell / 240.0))
        patch = np.empty((cell, cell, 3), np.uint8)
        patch[:] = sentinel
        items = [(xy, f'S{i}', patch) for i, xy in enumerate(self.cells[:self.n_states])]
        items.append((self.query_xy, f'S{self.n_states}', _make_query_image(cell)))
        items += [(xy, chr(65 + i), patch) for i, xy in enumerate(self.cells[self.n_states:])]
        for idx, ((x, y), label, p) in enumerate(items):
            grid[y:y + cell, x:x + cell] = p
            if show_border:
                cv2.rectangle(grid, (x, y), (x + cell - 1, y + cell - 1), (0, 0, 0), 1)
            if show_labels and label:
                cv2.putText(grid, label, (x + 3, y + cell + int(12 * font_scale) + 6), font, font_scale, (0, 0, 0), 1, cv2.LINE_AA)
            if record_meta:
                row = 0 if idx <= self.n_states else 1
                col = idx if row == 0 else idx - self.n_states - 1
                self.cells_meta.append({'r': row, 'c': col, 'label': label if show_labels else '', 'bbox': [int(x), int(y), int(cell), int(cell)],
                                        'is_query': idx == self.n_states})
        return grid


class RigGridLayout(ctypes.Structure):
    """struct rig_grid_layout (include/rig_grid.h)"""
    _vp = ctypes.c_void_p
    _fields_ = [('panel_w', ctypes.c_int32), ('panel_h', ctypes.c_int32), ('grid_w', ctypes.c_int32), ('grid_h', ctypes.c_int32),
                ('cell_size', ctypes.c_int32), ('n_cells', ctypes.c_int32), ('cell_x', _vp), ('cell_y', _vp), ('overlay', _vp),
                ('cell_of_pixel', _vp)]


class RigGridStats(ctypes.Structure):
    _fields_ = [('kernel_launches', ctypes.c_uint64), ('compose_ms', ctypes.c_double), ('timed_samples', ctypes.c_uint64)]


GRID_EXPORTS = ('rig_grid_create', 'rig_grid_destroy', 'rig_grid_last_error', 'rig_grid_compose', 'rig_grid_sync', 'rig_grid_get_stats',
                'rig_grid_set_profiling')
_bound = False


def _lib():
    global _bound
    lib = _abi.load()
    if not _bound:
        vp = ctypes.c_void_p
        lib.rig_grid_create.argtypes = [ctypes.c_int, ctypes.POINTER(RigGridLayout), ctypes.c_int, ctypes.POINTER(vp)]
        lib.rig_grid_create.restype = ctypes.c_int
        lib.rig_grid_destroy.argtypes = [vp]
        lib.rig_grid_destroy.restype = None
        lib.rig_grid_last_error.argtypes = [vp]
        lib.rig_grid_last_error.restype = ctypes.c_char_p
        lib.rig_grid_compose.argtypes = [vp, vp, vp, ctypes.c_int, vp, ctypes.c_int, vp]
        lib.rig_grid_compose.restype = ctypes.c_int
        lib.rig_grid_sync.argtypes = [vp]
        lib.rig_grid_sync.restype = ctypes.c_int
        lib.rig_grid_get_stats.argtypes = [vp, ctypes.POINTER(RigGridStats)]
        lib.rig_grid_get_stats.restype = ctypes.c_int
        lib.rig_grid_set_profiling.argtypes = [vp, ctypes.c_int]
        lib.rig_grid_set_profiling.restype = ctypes.c_int
        _bound = True
    return lib


class GridError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'librig error {code}: {message}')
        self.code = code


class GridCompositor:
    """One CUDA context-bound compositor for one layout.  `channel_order='bgr'` matches panels rendered with layout='bgr'
    (what the reference's canvas is); 'rgb' flips the overlay's channels to match RGB panels."""

    def __init__(self, layout: GridLayout, device: int = 0, max_samples: int = 4096, channel_order: str = 'bgr'):
        self._lib = _lib()
        self.layout, self.device, self.max_samples = layout, int(device), int(max_samples)
        ov = layout.overlay if channel_order == 'bgr' else layout.overlay[:, :, ::-1]
        self._ov = np.ascontiguousarray(ov, dtype=np.uint8)
        self._cm = np.ascontiguousarray(layout.cell_of_pixel, dtype=np.int8)
        self._cx = np.array([c[0] for c in layout.cells], dtype=np.int32)
        self._cy = np.array([c[1] for c in layout.cells], dtype=np.int32)
        L = RigGridLayout(layout.panel_size, layout.panel_size, layout.grid_w, layout.grid_h, layout.cell_size, len(layout.cells),
                          self._cx.ctypes.data, self._cy.ctypes.data, self._ov.ctypes.data, self._cm.ctypes.data)
        h = ctypes.c_void_p()
        rc = self._lib.rig_grid_create(self.device, ctypes.byref(L), self.max_samples, ctypes.byref(h))
        if rc:
            raise GridError(rc, (self._lib.rig_grid_last_error(None) or b'').decode())
        self._h = h
        self._lock = threading.Lock()

    def _check(self, rc: int):
        if rc:
            raise GridError(rc, (self._lib.rig_grid_last_error(self._h) or b'').decode())

    def close(self):
        if getattr(self, '_h', None):
            self._lib.rig_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compose(self, panels_ptr: int, panel_index, out: Optional[np.ndarray] = None) -> np.ndarray:
        """panel_index: int array [n_samples, n_cells] of indices into the panels at device pointer `panels_ptr` (-1: blank)
        -> uint8[n_samples, grid_h, grid_w, 3] in host memory."""
        idx = np.ascontiguousarray(panel_index, dtype=np.int32).reshape(-1, len(self.layout.cells))
        n = idx.shape[0]
        if out is None:
            out = np.empty((n, self.layout.grid_h, self.layout.grid_w, 3), dtype=np.uint8)
        with self._lock:
            self._check(self._lib.rig_grid_compose(self._h, panels_ptr, idx.ctypes.data, n, out.ctypes.data, 0, None))
        return out

    def compose_to_device(self, panels_ptr: int, panel_index, out_ptr: int, stream: int = 0):
        idx = np.ascontiguousarray(panel_index, dtype=np.int32).reshape(-1, len(self.layout.cells))
        with self._lock:
            self._check(self._lib.rig_grid_compose(self._h, panels_ptr, idx.ctypes.data, idx.shape[0], out_ptr, 1, stream or None))

    def sync(self):
        self._check(self._lib.rig_grid_sync(self._h))

    def set_profiling(self, on: bool):
        self._check(self._lib.rig_grid_set_profiling(self._h, 1 if on else 0))

    def stats(self) -> Dict[str, float]:
        st = RigGridStats()
        self._check(self._lib.rig_grid_get_stats(self._h, ctypes.byref(st)))
        return dict(kernel_launches=int(st.kernel_launches), compose_ms=float(st.compose_ms), timed_samples=int(st.timed_samples))
