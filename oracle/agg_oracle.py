"""oracle/agg_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of oracle/aggdraw.c, the CPU restatement of matplotlib's Agg rendering of the reference's
multigraph figures (/root/reference/multigraph_generation/generator.py:250-259).  PARITY STATUS: UNPINNED -- see the
header of aggdraw.c.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build_agg_oracle(force: bool = False) -> str:
    """gcc oracle/aggdraw.c -> oracle/_build/libaggdraw.so (building is not using)."""
    out = os.path.join(_HERE, '_build', 'libaggdraw.so')
    src = os.path.join(_HERE, 'aggdraw.c')
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-fno-fast-math', '-ffp-contract=off', '-o', out, src, '-lm'])
    return out


class _AggFigure(ctypes.Structure):
    _fields_ = [('n_artists', ctypes.c_int32),
                ('vertex_begin', ctypes.c_void_p), ('vertices', ctypes.c_void_p), ('codes', ctypes.c_void_p),
                ('affine', ctypes.c_void_p), ('linewidth', ctypes.c_void_p),
                ('cap', ctypes.c_void_p), ('join', ctypes.c_void_p), ('snap', ctypes.c_void_p),
                ('edge_rgba', ctypes.c_void_p), ('face_rgba', ctypes.c_void_p), ('has_face', ctypes.c_void_p),
                ('clip', ctypes.c_void_p), ('has_clip', ctypes.c_void_p)]


class AggOracle:
    """Renders figure scenes (the arrays of reasoning_image_generation_b200.figure.FigureScene) on the CPU."""

    def __init__(self):
        self.lib = ctypes.CDLL(build_agg_oracle())
        self.lib.agg_render_figure.argtypes = [ctypes.POINTER(_AggFigure), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        self.lib.agg_render_figure.restype = None
        self.lib.agg_debug_outline.argtypes = [ctypes.POINTER(_AggFigure), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_int]
        self.lib.agg_debug_outline.restype = ctypes.c_int

    @staticmethod
    def _figure(scene, i):
        """The sub-scene of image i as an _AggFigure plus the arrays that must outlive it."""
        a0, a1 = int(scene.image_artist_begin[i]), int(scene.image_artist_begin[i + 1])
        v0 = int(scene.artist_vertex_begin[a0])
        keep = dict(
            vb=np.ascontiguousarray(scene.artist_vertex_begin[a0:a1 + 1] - v0, dtype=np.uint32),
            vx=np.ascontiguousarray(scene.vertices[v0:int(scene.artist_vertex_begin[a1])], dtype=np.float64),
            cd=np.ascontiguousarray(scene.codes[v0:int(scene.artist_vertex_begin[a1])], dtype=np.uint8),
            af=np.ascontiguousarray(scene.affine[a0:a1], dtype=np.float64),
            lw=np.ascontiguousarray(scene.linewidth[a0:a1], dtype=np.float64),
            cap=np.ascontiguousarray(scene.cap[a0:a1], dtype=np.uint8),
            join=np.ascontiguousarray(scene.join[a0:a1], dtype=np.uint8),
            snap=np.ascontiguousarray(scene.snap[a0:a1], dtype=np.uint8),
            edge=np.ascontiguousarray(scene.edge_rgba[a0:a1], dtype=np.uint8),
            face=np.ascontiguousarray(scene.face_rgba[a0:a1], dtype=np.uint8),
            hf=np.ascontiguousarray(scene.has_face[a0:a1], dtype=np.uint8),
            clip=np.ascontiguousarray(scene.clip[a0:a1], dtype=np.float64),
            hc=np.ascontiguousarray(scene.has_clip[a0:a1], dtype=np.uint8),
        )
        f = _AggFigure(a1 - a0, keep['vb'].ctypes.data, keep['vx'].ctypes.data, keep['cd'].ctypes.data, keep['af'].ctypes.data,
                       keep['lw'].ctypes.data, keep['cap'].ctypes.data, keep['join'].ctypes.data, keep['snap'].ctypes.data,
                       keep['edge'].ctypes.data, keep['face'].ctypes.data, keep['hf'].ctypes.data, keep['clip'].ctypes.data,
                       keep['hc'].ctypes.data)
        return f, keep

    def render_rgba(self, scene, i, W, H):
        """Image i of the scene as uint8[H, W, 4] (savefig's opaque white figure patch underneath)."""
        f, keep = self._figure(scene, i)
        img = np.full((H, W, 4), 255, dtype=np.uint8)
        self.lib.agg_render_figure(ctypes.byref(f), W, H, img.ctypes.data)
        del keep
        return img

    def render(self, scene, W, H, images=None):
        """uint8[n, H, W, 3] RGB of the given images (default: all)."""
        idx = range(scene.n_images) if images is None else images
        return np.stack([self.render_rgba(scene, i, W, H)[:, :, :3] for i in idx]) if len(idx) else np.zeros((0, H, W, 3), np.uint8)

    def outline(self, scene, i, a, H, stroke=True, cap=1 << 16):
        """Debug: the stroked outline (or flattened path) of artist a of image i as float64[n, 3] (x, y, cmd)."""
        f, keep = self._figure(scene, i)
        buf = np.zeros((cap, 3), dtype=np.float64)
        n = self.lib.agg_debug_outline(ctypes.byref(f), a, H, 1 if stroke else 0, buf.ctypes.data, cap)
        del keep
        return buf[:min(n, cap)].copy()
