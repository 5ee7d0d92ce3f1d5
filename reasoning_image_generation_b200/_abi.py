"""ctypes view of include/rig.h (the C ABI of librig.so).  No torch types cross this boundary."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RIG_LIB') or os.path.join(_HERE, 'librig.so')

RIG_OK = 0
RIG_E_INVALID, RIG_E_CUDA, RIG_E_CAPACITY, RIG_E_UNSUPPORTED, RIG_E_NODEVICE = -1, -2, -3, -4, -5
RIG_LAYOUT_RGB, RIG_LAYOUT_BGR = 0, 1

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)


class RigScene(ctypes.Structure):
    """struct rig_scene (include/rig.h)"""
    _fields_ = [
        ('n_panels', ctypes.c_int32),
        ('n_elements', ctypes.c_int32),
        ('grid_size', ctypes.c_int32),
        ('bg_rgb', ctypes.c_uint8 * 3),
        ('reserved', ctypes.c_uint8),
        ('panel_elem_begin', _u32p),
        ('panel_flags', _u8p),
        ('kind', _u8p),
        ('fill', _u8p),
        ('flip', _u8p),
        ('stroke', _i32p),
        ('size', _i32p),
        ('angle', _i32p),
        ('cx', _f64p),
        ('cy', _f64p),
        ('rgb', _u8p),
    ]


class RigStats(ctypes.Structure):
    """struct rig_stats (include/rig.h)"""
    _fields_ = [
        ('kernel_launches', ctypes.c_uint64),
        ('records', ctypes.c_uint32),
        ('aux_units', ctypes.c_uint32),
        ('overflow', ctypes.c_uint32),
        ('timed_launches', ctypes.c_uint32),
        ('expand_ms', ctypes.c_double),
        ('raster_ms', ctypes.c_double),
        ('timed_panels', ctypes.c_uint64),
    ]


# every symbol include/rig.h declares; tests check that the built library exports all of them
EXPORTS = (
    'rig_create', 'rig_destroy', 'rig_last_error', 'rig_render_batch', 'rig_scene_upload',
    'rig_render_resident', 'rig_render_resident_hashed', 'rig_sync', 'rig_get_stats', 'rig_set_profiling', 'rig_debug_shape_tables',
    'rig_hash_batch', 'rig_dedup', 'rig_hash_words', 'rig_int_pipe_peak', 'rig_pinned_alloc', 'rig_pinned_free', 'rig_timeline_dump',
)

_lib = None


def load() -> ctypes.CDLL:
    """Loads librig.so (built in-tree by __graft_entry__.build()).  There is no fallback: a missing
    library is an error, not a reason to render on the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(nvcc, sm_100a).  This package has no CPU rendering path.')
    lib = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    lib.rig_create.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(vp)]
    lib.rig_create.restype = ctypes.c_int
    lib.rig_destroy.argtypes = [vp]
    lib.rig_destroy.restype = None
    lib.rig_last_error.argtypes = [vp]
    lib.rig_last_error.restype = ctypes.c_char_p
    lib.rig_render_batch.argtypes = [vp, ctypes.POINTER(RigScene), vp, ctypes.c_int, ctypes.c_int, vp]
    lib.rig_render_batch.restype = ctypes.c_int
    lib.rig_scene_upload.argtypes = [vp, ctypes.POINTER(RigScene), ctypes.c_int, vp]
    lib.rig_scene_upload.restype = ctypes.c_int
    lib.rig_render_resident.argtypes = [vp, vp, vp]
    lib.rig_render_resident.restype = ctypes.c_int
    lib.rig_render_resident_hashed.argtypes = [vp, vp, vp, vp]
    lib.rig_render_resident_hashed.restype = ctypes.c_int
    lib.rig_sync.argtypes = [vp]
    lib.rig_sync.restype = ctypes.c_int
    lib.rig_get_stats.argtypes = [vp, ctypes.POINTER(RigStats)]
    lib.rig_get_stats.restype = ctypes.c_int
    lib.rig_set_profiling.argtypes = [vp, ctypes.c_int]
    lib.rig_set_profiling.restype = ctypes.c_int
    lib.rig_debug_shape_tables.argtypes = [vp, ctypes.c_size_t]
    lib.rig_debug_shape_tables.restype = ctypes.c_size_t
    lib.rig_hash_words.argtypes = []
    lib.rig_hash_words.restype = ctypes.c_int
    lib.rig_hash_batch.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.rig_hash_batch.restype = ctypes.c_int
    lib.rig_dedup.argtypes = [vp, vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, vp, vp, vp]
    lib.rig_dedup.restype = ctypes.c_int
    lib.rig_int_pipe_peak.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.rig_int_pipe_peak.restype = ctypes.c_int
    lib.rig_pinned_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.rig_pinned_alloc.restype = ctypes.c_int
    lib.rig_pinned_free.argtypes = [vp]
    lib.rig_pinned_free.restype = None
    _lib = lib
    return lib
