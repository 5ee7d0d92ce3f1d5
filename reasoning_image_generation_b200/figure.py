"""figure.py -- scene description + GPU renderer of the reference's multigraph figures.

The reference renders every multigraph sample with ``fig.savefig`` (multigraph_generation/generator.py:250-259), i.e.
matplotlib's Agg backend drawing ``ax.patches`` and ``ax.lines``.  Per artist the backend receives exactly what
``RendererAgg.draw_path(gc, path, transform, rgbFace)`` receives: a Path (vertices + codes), an affine to display
coordinates, and the graphics-context state (line width in pixels, cap, join, snap, RGBA colours, clip rectangle).
`FigureScene` is that list, as a struct of arrays, for a whole batch of figures; `FigureRenderer` hands it to librig.so
(`rig_figure_*`, include/rig_figure.h), whose CUDA kernel accumulates Agg's 24.8 fixed-point coverage cells per band of
rows in shared memory and blends them with matplotlib's non-premultiplied blender.

There is no CPU rendering path in this package (oracle/aggdraw.c is test infrastructure).
"""
from __future__ import annotations

import ctypes
import threading
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _abi

CAPS = {'butt': 0, 'projecting': 1, 'round': 2}
JOINS = {'miter': 0, 'round': 1, 'bevel': 2}
SNAP_FALSE, SNAP_TRUE, SNAP_AUTO = 0, 1, 2


def rgba8(color: Sequence[float]) -> tuple:
    """agg::rgba8(const rgba&): uround(c * 255) per component (colours arrive as matplotlib floats in 0..1)."""
    c = tuple(color)
    if len(c) == 3:
        c = c + (1.0,)
    return tuple(int(float(v) * 255.0 + 0.5) for v in c)


class FigureScene:
    """Struct-of-arrays artist list of a batch of W x H figures (all arrays C-contiguous)."""

    def __init__(self, W, H, image_artist_begin, artist_vertex_begin, vertices, codes, affine, linewidth, cap, join, snap,
                 edge_rgba, face_rgba, has_face, clip, has_clip):
        self.W, self.H = int(W), int(H)
        self.image_artist_begin = np.ascontiguousarray(image_artist_begin, dtype=np.uint32)
        self.artist_vertex_begin = np.ascontiguousarray(artist_vertex_begin, dtype=np.uint32)
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 2)
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.affine = np.ascontiguousarray(affine, dtype=np.float64).reshape(-1, 6)
        self.linewidth = np.ascontiguousarray(linewidth, dtype=np.float64)
        self.cap = np.ascontiguousarray(cap, dtype=np.uint8)
        self.join = np.ascontiguousarray(join, dtype=np.uint8)
        self.snap = np.ascontiguousarray(snap, dtype=np.uint8)
        self.edge_rgba = np.ascontiguousarray(edge_rgba, dtype=np.uint8).reshape(-1, 4)
        self.face_rgba = np.ascontiguousarray(face_rgba, dtype=np.uint8).reshape(-1, 4)
        self.has_face = np.ascontiguousarray(has_face, dtype=np.uint8)
        self.clip = np.ascontiguousarray(clip, dtype=np.float64).reshape(-1, 4)
        self.has_clip = np.ascontiguousarray(has_clip, dtype=np.uint8)

    @property
    def n_images(self) -> int:
        return len(self.image_artist_begin) - 1

    @property
    def n_artists(self) -> int:
        return len(self.artist_vertex_begin) - 1

    @property
    def n_vertices(self) -> int:
        return len(self.codes)

    def as_struct(self) -> 'RigFigureScene':
        s = RigFigureScene()
        s.n_images, s.n_artists, s.n_vertices = self.n_images, self.n_artists, self.n_vertices
        p = lambda a: a.ctypes.data if a.size else None   # noqa: E731
        s.image_artist_begin = p(self.image_artist_begin)
        s.artist_vertex_begin = p(self.artist_vertex_begin)
        s.vertices, s.codes, s.affine, s.linewidth = p(self.vertices), p(self.codes), p(self.affine), p(self.linewidth)
        s.cap, s.join, s.snap = p(self.cap), p(self.join), p(self.snap)
        s.edge_rgba, s.face_rgba, s.has_face = p(self.edge_rgba), p(self.face_rgba), p(self.has_face)
        s.clip, s.has_clip = p(self.clip), p(self.has_clip)
        return s

    def slice(self, i0: int, i1: int) -> 'FigureScene':
        """Images [i0, i1) as a scene of their own (used to shard a scene across ranks)."""
        a0, a1 = int(self.image_artist_begin[i0]), int(self.image_artist_begin[i1])
        v0, v1 = int(self.artist_vertex_begin[a0]), int(self.artist_vertex_begin[a1])
        return FigureScene(self.W, self.H, self.image_artist_begin[i0:i1 + 1] - a0, self.artist_vertex_begin[a0:a1 + 1] - v0,
                           self.vertices[v0:v1], self.codes[v0:v1], self.affine[a0:a1], self.linewidth[a0:a1], self.cap[a0:a1],
                           self.join[a0:a1], self.snap[a0:a1], self.edge_rgba[a0:a1], self.face_rgba[a0:a1], self.has_face[a0:a1],
                           self.clip[a0:a1], self.has_clip[a0:a1])


_SCENE_ARRAYS = ('image_artist_begin', 'artist_vertex_begin', 'vertices', 'codes', 'affine', 'linewidth', 'cap', 'join', 'snap',
                 'edge_rgba', 'face_rgba', 'has_face', 'clip', 'has_clip')


def scene_to_arrays(scene: FigureScene, prefix: str = '') -> Dict[str, np.ndarray]:
    """The scene as a flat dict of arrays (np.savez-able; `scene_from_arrays` is the inverse)."""
    d = {prefix + k: getattr(scene, k) for k in _SCENE_ARRAYS}
    d[prefix + 'WH'] = np.array([scene.W, scene.H], dtype=np.int64)
    return d


def scene_from_arrays(d, prefix: str = '') -> FigureScene:
    W, H = (int(v) for v in d[prefix + 'WH'])
    return FigureScene(W, H, *[d[prefix + k] for k in _SCENE_ARRAYS])


class FigureSceneBuilder:
    """Collects figures artist by artist, in draw order (matplotlib: sorted by zorder, then insertion)."""

    def __init__(self, W: int = 512, H: int = 512):
        self.W, self.H = int(W), int(H)
        self._img_begin: List[int] = [0]
        self._vtx_begin: List[int] = [0]
        self._verts: List[np.ndarray] = []
        self._codes: List[np.ndarray] = []
        self._affine: List[tuple] = []
        self._lw: List[float] = []
        self._cap: List[int] = []
        self._join: List[int] = []
        self._snap: List[int] = []
        self._edge: List[tuple] = []
        self._face: List[tuple] = []
        self._has_face: List[int] = []
        self._clip: List[tuple] = []
        self._has_clip: List[int] = []
        self._nv = 0

    def add_artist(self, vertices, codes, affine=(1.0, 0.0, 0.0, 1.0, 0.0, 0.0), linewidth: float = 1.0, cap='butt', join='miter',
                   snap=None, edgecolor=(0.0, 0.0, 0.0, 1.0), facecolor=None, clip=None):
        """One ``draw_path`` call.  vertices: [n, 2] in the artist's own coordinates; codes: matplotlib Path codes (None:
        MOVETO + LINETOs); affine: Affine2D.to_values() (a, b, c, d, e, f) or a 3 x 3 matrix to display coordinates (origin
        bottom-left, pixels); linewidth in PIXELS (points * dpi / 72), 0 for no stroke; colours as 0..1 floats; clip:
        (x0, y0, x1, y1) in display coordinates or None."""
        v = np.asarray(vertices, dtype=np.float64).reshape(-1, 2)
        if codes is None:
            c = np.full(len(v), 2, dtype=np.uint8)
            if len(c):
                c[0] = 1
        else:
            c = np.asarray(codes, dtype=np.uint8)
        if len(c) != len(v):
            raise ValueError('codes and vertices differ in length')
        a = np.asarray(affine, dtype=np.float64)
        if a.shape == (3, 3):
            a = np.array([a[0, 0], a[1, 0], a[0, 1], a[1, 1], a[0, 2], a[1, 2]])
        if a.shape != (6,):
            raise ValueError('affine must be 6 values or a 3x3 matrix')
        self._verts.append(v)
        self._codes.append(c)
        self._nv += len(v)
        self._vtx_begin.append(self._nv)
        self._affine.append(tuple(float(x) for x in a))
        self._lw.append(float(linewidth))
        self._cap.append(CAPS[cap] if isinstance(cap, str) else int(cap))
        self._join.append(JOINS[join] if isinstance(join, str) else int(join))
        self._snap.append(SNAP_AUTO if snap is None else (SNAP_TRUE if snap else SNAP_FALSE))
        self._edge.append(rgba8(edgecolor))
        self._face.append(rgba8(facecolor) if facecolor is not None else (0, 0, 0, 0))
        self._has_face.append(1 if facecolor is not None else 0)
        self._clip.append(tuple(float(x) for x in clip) if clip is not None else (0.0, 0.0, 0.0, 0.0))
        self._has_clip.append(1 if clip is not None else 0)

    def end_image(self):
        self._img_begin.append(len(self._lw))

    def add_image(self, artists: Sequence[Dict[str, Any]]):
        for a in artists:
            self.add_artist(**a)
        self.end_image()

    def finish(self) -> FigureScene:
        if self._img_begin[-1] != len(self._lw):
            self.end_image()
        verts = np.concatenate(self._verts) if self._verts else np.zeros((0, 2))
        codes = np.concatenate(self._codes) if self._codes else np.zeros(0, np.uint8)
        na = len(self._lw)
        return FigureScene(self.W, self.H, self._img_begin, self._vtx_begin, verts, codes,
                           np.array(self._affine, dtype=np.float64).reshape(na, 6), self._lw, self._cap, self._join, self._snap,
                           np.array(self._edge, dtype=np.uint8).reshape(na, 4), np.array(self._face, dtype=np.uint8).reshape(na, 4),
                           self._has_face, np.array(self._clip, dtype=np.float64).reshape(na, 4), self._has_clip)


class RigFigureScene(ctypes.Structure):
    """struct rig_figure_scene (include/rig_figure.h)"""
    _vp = ctypes.c_void_p
    _fields_ = [('n_images', ctypes.c_int32), ('n_artists', ctypes.c_int32), ('n_vertices', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('image_artist_begin', _vp), ('artist_vertex_begin', _vp), ('vertices', _vp), ('codes', _vp), ('affine', _vp),
                ('linewidth', _vp), ('cap', _vp), ('join', _vp), ('snap', _vp), ('edge_rgba', _vp), ('face_rgba', _vp),
                ('has_face', _vp), ('clip', _vp), ('has_clip', _vp)]


class RigFigureStats(ctypes.Structure):
    """struct rig_figure_stats (include/rig_figure.h)"""
    _fields_ = [('kernel_launches', ctypes.c_uint64), ('edges', ctypes.c_uint64), ('layers', ctypes.c_uint64),
                ('timed_launches', ctypes.c_uint32), ('reserved', ctypes.c_uint32), ('cover_ms', ctypes.c_double),
                ('timed_images', ctypes.c_uint64), ('prepare_ms', ctypes.c_double), ('geom_ms', ctypes.c_double),
                ('geom_launches', ctypes.c_uint32), ('geom_fallbacks', ctypes.c_uint32), ('geom_status', ctypes.c_uint32),
                ('geom_resident', ctypes.c_uint32)]


FIGURE_EXPORTS = ('rig_figure_create', 'rig_figure_destroy', 'rig_figure_last_error', 'rig_figure_render_batch',
                  'rig_figure_upload', 'rig_figure_render_resident', 'rig_figure_sync', 'rig_figure_get_stats',
                  'rig_figure_set_profiling', 'rig_figure_debug_edges', 'rig_figure_stream', 'rig_figure_fence',
                  'rig_figure_blank_bands', 'rig_figure_set_blank_flags', 'rig_figure_set_geometry', 'rig_figure_draw_resident',
                  'rig_figure_debug_resident')
GEOM_HOST, GEOM_DEVICE = 0, 1

_bound = False


def _lib():
    global _bound
    lib = _abi.load()
    if not _bound:
        vp = ctypes.c_void_p
        lib.rig_figure_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        lib.rig_figure_create.restype = ctypes.c_int
        lib.rig_figure_destroy.argtypes = [vp]
        lib.rig_figure_destroy.restype = None
        lib.rig_figure_last_error.argtypes = [vp]
        lib.rig_figure_last_error.restype = ctypes.c_char_p
        lib.rig_figure_render_batch.argtypes = [vp, ctypes.POINTER(RigFigureScene), vp, ctypes.c_int, ctypes.c_int, vp]
        lib.rig_figure_render_batch.restype = ctypes.c_int
        lib.rig_figure_upload.argtypes = [vp, ctypes.POINTER(RigFigureScene), ctypes.c_int, vp]
        lib.rig_figure_upload.restype = ctypes.c_int
        lib.rig_figure_render_resident.argtypes = [vp, vp, vp]
        lib.rig_figure_render_resident.restype = ctypes.c_int
        lib.rig_figure_sync.argtypes = [vp]
        lib.rig_figure_sync.restype = ctypes.c_int
        lib.rig_figure_stream.argtypes = [vp]
        lib.rig_figure_stream.restype = vp
        lib.rig_figure_fence.argtypes = [vp, ctypes.POINTER(vp)]
        lib.rig_figure_fence.restype = ctypes.c_int
        lib.rig_figure_blank_bands.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int)]
        lib.rig_figure_blank_bands.restype = ctypes.c_int
        lib.rig_figure_set_blank_flags.argtypes = [vp, ctypes.c_int]
        lib.rig_figure_set_blank_flags.restype = ctypes.c_int
        lib.rig_figure_get_stats.argtypes = [vp, ctypes.POINTER(RigFigureStats)]
        lib.rig_figure_get_stats.restype = ctypes.c_int
        lib.rig_figure_set_profiling.argtypes = [vp, ctypes.c_int]
        lib.rig_figure_set_profiling.restype = ctypes.c_int
        lib.rig_figure_debug_edges.argtypes = [ctypes.POINTER(RigFigureScene), ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int64, vp]
        lib.rig_figure_debug_edges.restype = ctypes.c_int64
        lib.rig_figure_set_geometry.argtypes = [vp, ctypes.c_int]
        lib.rig_figure_set_geometry.restype = ctypes.c_int
        lib.rig_figure_draw_resident.argtypes = [vp, vp, vp]
        lib.rig_figure_draw_resident.restype = ctypes.c_int
        lib.rig_figure_debug_resident.argtypes = [vp, vp, vp, ctypes.c_int64, vp]
        lib.rig_figure_debug_resident.restype = ctypes.c_int64
        _bound = True
    return lib


_LAYOUTS = {'rgb': _abi.RIG_LAYOUT_RGB, 'bgr': _abi.RIG_LAYOUT_BGR}


class FigureError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'librig error {code}: {message}')
        self.code = code


def host_edges(scene: FigureScene, cap: int = 1 << 22):
    """Debug / CPU-test hook (needs no GPU): runs the library's host path pipeline (transform, snap, curve flattening,
    stroking, clipping) and returns (layers, edges): layers uint32[n, 8] = (image, edge_begin, n_edges, rgba, xmin, xmax,
    ymin, ymax) and edges int32[m, 4] = (x1, y1, x2, y2) in 24.8 fixed point, exactly what is uploaded to the device."""
    lib = _lib()
    s = scene.as_struct()
    layers = np.zeros((max(1, 2 * scene.n_artists), 8), dtype=np.uint32)
    edges = np.zeros((cap, 4), dtype=np.int32)
    nl = ctypes.c_int64(0)
    n = lib.rig_figure_debug_edges(ctypes.byref(s), scene.W, scene.H, layers.ctypes.data, edges.ctypes.data, cap, ctypes.byref(nl))
    if n < 0:
        raise FigureError(int(n), 'rig_figure_debug_edges failed (edge capacity?)')
    return layers[:nl.value].copy(), edges[:n].copy()


class FigureRenderer:
    """One CUDA context-bound figure rasteriser (one per process per GPU)."""

    def __init__(self, W: int = 512, H: int = 512, device: int = 0, max_images: int = 4096, max_artists: Optional[int] = None):
        self._lib = _lib()
        self.W, self.H, self.device = int(W), int(H), int(device)
        self.max_images = int(max_images)
        self.max_artists = int(max_artists if max_artists is not None else 16 * max_images)
        h = ctypes.c_void_p()
        rc = self._lib.rig_figure_create(self.device, self.W, self.H, self.max_images, self.max_artists, ctypes.byref(h))
        if rc:
            raise FigureError(rc, (self._lib.rig_figure_last_error(None) or b'').decode())
        self._h = h
        self._lock = threading.Lock()

    def _check(self, rc: int):
        if rc:
            raise FigureError(rc, (self._lib.rig_figure_last_error(self._h) or b'').decode())

    def close(self):
        if getattr(self, '_h', None):
            self._lib.rig_figure_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_scene(self, scene: FigureScene):
        if (scene.W, scene.H) != (self.W, self.H):
            raise ValueError(f'scene is {scene.W}x{scene.H}, renderer is {self.W}x{self.H}')

    def render(self, scene: FigureScene, layout: str = 'rgb', out: Optional[np.ndarray] = None) -> np.ndarray:
        """Scene -> uint8[n_images, H, W, 3] in host memory (rig_figure_render_batch, host destination)."""
        self._check_scene(scene)
        n = scene.n_images
        if out is None:
            out = np.empty((n, self.H, self.W, 3), dtype=np.uint8)
        if out.dtype != np.uint8 or not out.flags['C_CONTIGUOUS'] or out.size != n * self.H * self.W * 3:
            raise ValueError('out must be a C-contiguous uint8 array of n_images*H*W*3 bytes')
        s = scene.as_struct()
        with self._lock:
            self._check(self._lib.rig_figure_render_batch(self._h, ctypes.byref(s), out.ctypes.data, 0, _LAYOUTS[layout], None))
        return out

    def render_to_device(self, scene: FigureScene, out_ptr: int, layout: str = 'rgb', stream: int = 0):
        self._check_scene(scene)
        s = scene.as_struct()
        with self._lock:
            self._check(self._lib.rig_figure_render_batch(self._h, ctypes.byref(s), out_ptr, 1, _LAYOUTS[layout], stream or None))

    def stream(self) -> int:
        """The context's own CUDA stream (what asynchronous calls run on when no stream is given)."""
        return int(self._lib.rig_figure_stream(self._h) or 0)

    def fence(self) -> int:
        """An event handle behind everything enqueued on the context's stream so far (valid for the next 7 calls)."""
        e = ctypes.c_void_p()
        self._check(self._lib.rig_figure_fence(self._h, ctypes.byref(e)))
        return int(e.value)

    def set_blank_flags(self, on: bool):
        """Makes renders into device memory also write one flag per band of rows (see `blank_bands`)."""
        self._check(self._lib.rig_figure_set_blank_flags(self._h, 1 if on else 0))

    def blank_bands(self):
        """(device pointer, rows per band) of the blank-band flags k_figure wrote during the last render into device memory: one byte
        per band and image, 1 = only background pixels (rig_figure_blank_bands).  Valid until the fourth next render; (0, rows)
        unless `set_blank_flags(True)` was in force."""
        p, rows = ctypes.c_void_p(), ctypes.c_int()
        self._check(self._lib.rig_figure_blank_bands(self._h, ctypes.byref(p), ctypes.byref(rows)))
        return int(p.value or 0), int(rows.value)

    def iter_png(self, scene: FigureScene, encoder, chunk: int = 4096, overlap: bool = True):
        """``fig.savefig`` for a whole batch (multigraph_generation/generator.py:256): renders `chunk` figures at a time into
        the encoder's device staging buffer and encodes them there, so only PNG bytes cross PCIe.  Yields
        (first_image_index, buffer, offsets): file i of the chunk is buffer[offsets[i]:offsets[i + 1]] (pinned host memory,
        valid until the next iteration; the copy of one chunk's files overlaps the encode kernels of the next).  With `overlap` a
        worker thread runs the host path pipeline and launches k_figure
        for the chunks ahead (two halves of the staging buffer) while this thread encodes and copies; the encoder is ordered
        after each render on the GPU (an event recorded right behind it), not by a host synchronisation."""
        from .png import chunk_plan, pipelined_png
        self._check_scene(scene)
        chunk = max(1, min(int(chunk), encoder.max_images, self.max_images))

        def render_chunk(i0, i1, ptr):
            sub = scene if (i0 == 0 and i1 == scene.n_images) else scene.slice(i0, i1)
            self.render_to_device(sub, ptr, 'rgb', 0)            # asynchronous: returns once k_figure is enqueued
            return sub.n_images, self.fence(), self.blank_bands()

        self.set_blank_flags(True)                               # the encoder skips the bands k_figure left blank
        try:
            yield from pipelined_png(encoder, chunk_plan(scene.n_images, chunk), chunk, self.W * self.H * 3, render_chunk, 'rgb', overlap)
        finally:
            self.set_blank_flags(False)

    def upload(self, scene: FigureScene, layout: str = 'rgb', stream: int = 0):
        self._check_scene(scene)
        s = scene.as_struct()
        with self._lock:
            self._check(self._lib.rig_figure_upload(self._h, ctypes.byref(s), _LAYOUTS[layout], stream or None))

    def render_resident(self, out_ptr: int, stream: int = 0):
        with self._lock:
            self._check(self._lib.rig_figure_render_resident(self._h, out_ptr, stream or None))

    def draw_resident(self, out_ptr: int, stream: int = 0):
        """The whole draw_path of every artist of the resident scene: k_figure_path (transform, snap, flattening, stroking,
        clipping, from the artist arrays in HBM) followed by k_figure."""
        with self._lock:
            self._check(self._lib.rig_figure_draw_resident(self._h, out_ptr, stream or None))

    def set_geometry(self, mode: str):
        """'device' (default): k_figure_path; 'host': the host path pipeline.  Same integers either way."""
        self._check(self._lib.rig_figure_set_geometry(self._h, {'host': GEOM_HOST, 'device': GEOM_DEVICE}[mode]))

    def resident_edges(self, cap: int = 1 << 22):
        """(layers, edges) of the resident scene copied back from the device, in the format of `host_edges` (test hook)."""
        layers = np.zeros((max(1, int(self.stats()['layers'])), 8), dtype=np.uint32)
        edges = np.zeros((cap, 4), dtype=np.int32)
        nl = ctypes.c_int64(0)
        with self._lock:
            n = self._lib.rig_figure_debug_resident(self._h, layers.ctypes.data, edges.ctypes.data, cap, ctypes.byref(nl))
        if n < 0:
            raise FigureError(int(n), 'rig_figure_debug_resident failed')
        return layers[:nl.value].copy(), edges[:n].copy()

    def sync(self):
        with self._lock:
            self._check(self._lib.rig_figure_sync(self._h))

    def set_profiling(self, on: bool):
        self._check(self._lib.rig_figure_set_profiling(self._h, 1 if on else 0))

    def stats(self) -> Dict[str, Any]:
        st = RigFigureStats()
        self._check(self._lib.rig_figure_get_stats(self._h, ctypes.byref(st)))
        return dict(kernel_launches=int(st.kernel_launches), edges=int(st.edges), layers=int(st.layers),
                    timed_launches=int(st.timed_launches), cover_ms=float(st.cover_ms), timed_images=int(st.timed_images),
                    prepare_ms=float(st.prepare_ms), geom_ms=float(st.geom_ms), geom_launches=int(st.geom_launches),
                    geom_fallbacks=int(st.geom_fallbacks), geom_status=int(st.geom_status), geom_resident=int(st.geom_resident))
