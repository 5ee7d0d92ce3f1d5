"""Batched GPU replacement for the reference's render stage.

`Renderer.render(scene)` is the batched form; `render_elements_to_canvas` has the exact signature,
return type and error behaviour of RPMGenerator._render_elements_to_canvas
(/root/reference/src/generator.py:66-124) and `install(RPMGenerator)` swaps it in, so the reference's
generator.py / cli.py run unchanged with the pixels coming from the B200.

There is no CPU rendering path in this package: if librig.so or a CUDA device is missing the calls
raise RuntimeError (which the reference already catches per sample, generator.py:226-230).
"""
from __future__ import annotations

import ctypes
import threading
from typing import Any, Dict, Iterable, Optional, Sequence

import numpy as np

from . import _abi
from .scene import Scene, SceneBuilder

_LAYOUTS = {'rgb': _abi.RIG_LAYOUT_RGB, 'bgr': _abi.RIG_LAYOUT_BGR}


class RigError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'librig error {code}: {message}')
        self.code = code


class PinnedBuffer:
    """Page-locked host memory owned by the library (rig_pinned_alloc), viewed as a numpy array."""

    def __init__(self, nbytes: int):
        self._lib = _abi.load()
        p = ctypes.c_void_p()
        rc = self._lib.rig_pinned_alloc(nbytes, ctypes.byref(p))
        if rc:
            raise RigError(rc, (self._lib.rig_last_error(None) or b'').decode())
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self._lib.rig_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """One CUDA context-bound rasteriser (one per process per GPU)."""

    def __init__(self, W: int = 512, H: int = 512, device: int = 0, max_panels: int = 4096,
                 max_elements: Optional[int] = None):
        self._lib = _abi.load()
        self.W, self.H, self.device = int(W), int(H), int(device)
        self.max_panels = int(max_panels)
        self.max_elements = int(max_elements if max_elements is not None else 8 * max_panels)
        h = ctypes.c_void_p()
        rc = self._lib.rig_create(self.device, self.W, self.H, self.max_panels, self.max_elements, ctypes.byref(h))
        if rc:
            raise RigError(rc, (self._lib.rig_last_error(None) or b'').decode())
        self._h = h
        self._lock = threading.Lock()
        self._pinned: Optional[PinnedBuffer] = None

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc:
            raise RigError(rc, (self._lib.rig_last_error(self._h) or b'').decode())

    def close(self):
        if getattr(self, '_h', None):
            self._lib.rig_destroy(self._h)
            self._h = None
        if self._pinned is not None:
            self._pinned.close()
            self._pinned = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_scene(self, scene: Scene):
        if (scene.W, scene.H) != (self.W, self.H):
            raise ValueError(f'scene is {scene.W}x{scene.H}, renderer is {self.W}x{self.H}')

    def pinned_output(self, n_panels: int) -> np.ndarray:
        """A reusable page-locked uint8[n_panels, H, W, 3] view (valid until the next call)."""
        need = n_panels * self.H * self.W * 3
        if self._pinned is None or self._pinned.nbytes < need:
            if self._pinned is not None:
                self._pinned.close()
            self._pinned = PinnedBuffer(max(need, 1))
        return self._pinned.array[:need].reshape(n_panels, self.H, self.W, 3)

    # -- rendering --------------------------------------------------------------------------
    def render(self, scene: Scene, layout: str = 'bgr', out: Optional[np.ndarray] = None) -> np.ndarray:
        """Scene -> uint8[n_panels, H, W, 3] in host memory (rig_render_batch, host destination)."""
        self._check_scene(scene)
        n = scene.n_panels
        if out is None:
            out = np.empty((n, self.H, self.W, 3), dtype=np.uint8)
        if out.dtype != np.uint8 or not out.flags['C_CONTIGUOUS'] or out.size != n * self.H * self.W * 3:
            raise ValueError('out must be a C-contiguous uint8 array of n_panels*H*W*3 bytes')
        s = scene.as_struct()
        with self._lock:
            self._check(self._lib.rig_render_batch(self._h, ctypes.byref(s), out.ctypes.data, 0, _LAYOUTS[layout], None))
        return out

    def render_to_device(self, scene: Scene, out_ptr: int, layout: str = 'bgr', stream: int = 0):
        """Scene -> device memory at `out_ptr` (e.g. a torch uint8 tensor's data_ptr()); asynchronous."""
        self._check_scene(scene)
        s = scene.as_struct()
        with self._lock:
            self._check(self._lib.rig_render_batch(self._h, ctypes.byref(s), out_ptr, 1, _LAYOUTS[layout], stream or None))

    def render_to_device_checked(self, scene: Scene, out_ptr: int, layout: str = 'bgr', attempts: int = 6):
        """render_to_device + sync; a draw-record arena overflow (RIG_E_CAPACITY from rig_sync, which also halves the context's
        sticky chunk budget) is retried here instead of reaching the caller in the middle of a batch."""
        for k in range(attempts):
            self.render_to_device(scene, out_ptr, layout, 0)
            try:
                self.sync()
                return
            except RigError as e:
                if e.code != _abi.RIG_E_CAPACITY or k == attempts - 1:
                    raise

    def iter_png(self, scene: Scene, encoder, chunk: int = 4096, layout: str = 'bgr', overlap: bool = True):
        """``save_image(path, img)`` = ``cv2.imwrite`` (src/utils.py:18-21) for a whole batch of panels: renders up to `chunk` panels at
        a time into the encoder's device staging buffer and encodes them there, so only PNG bytes cross PCIe.  `layout` is the
        channel order the panels are rendered in ('bgr' = what the reference hands to cv2.imwrite; the files hold RGB either
        way).  Yields (first_panel_index, buffer, offsets); file i of the chunk is buffer[offsets[i]:offsets[i + 1]] (pinned
        host memory, valid until the next iteration).  With `overlap` a worker thread renders the chunks ahead while this thread
        encodes, and the copy of one chunk's files overlaps the kernels of the next."""
        from .png import chunk_plan, pipelined_png
        self._check_scene(scene)
        chunk = max(1, min(int(chunk), encoder.max_images, self.max_panels))

        def render_chunk(p0, p1, ptr):
            sub = scene if (p0 == 0 and p1 == scene.n_panels) else scene.slice(p0, p1)
            self.render_to_device_checked(sub, ptr, layout)   # context's own stream: finished before the encoder reads
            return sub.n_panels, 0

        yield from pipelined_png(encoder, chunk_plan(scene.n_panels, chunk), chunk, self.W * self.H * 3, render_chunk, layout, overlap)

    def upload(self, scene: Scene, layout: str = 'bgr', stream: int = 0):
        self._check_scene(scene)
        s = scene.as_struct()
        with self._lock:
            self._check(self._lib.rig_scene_upload(self._h, ctypes.byref(s), _LAYOUTS[layout], stream or None))

    def render_resident(self, out_ptr: int, stream: int = 0):
        with self._lock:
            self._check(self._lib.rig_render_resident(self._h, out_ptr, stream or None))

    def render_resident_hashed(self, out_ptr: int, hashes_ptr: int, stream: int = 0):
        """Resident scene -> pixels at out_ptr and uint64[n, 4] perceptual hashes at hashes_ptr (hash fused into the rasteriser)."""
        with self._lock:
            self._check(self._lib.rig_render_resident_hashed(self._h, out_ptr, hashes_ptr, stream or None))

    def sync(self):
        with self._lock:
            self._check(self._lib.rig_sync(self._h))

    def stats(self) -> Dict[str, int]:
        st = _abi.RigStats()
        self._check(self._lib.rig_get_stats(self._h, ctypes.byref(st)))
        return dict(kernel_launches=int(st.kernel_launches), records=int(st.records), aux_units=int(st.aux_units),
                    overflow=int(st.overflow), timed_launches=int(st.timed_launches), expand_ms=float(st.expand_ms),
                    raster_ms=float(st.raster_ms), timed_panels=int(st.timed_panels))

    def set_profiling(self, on: bool):
        self._check(self._lib.rig_set_profiling(self._h, 1 if on else 0))

    # -- near-duplicate filter (device pointers in, device pointers out) --------------------
    def hash_batch(self, imgs_ptr: int, n: int, hashes_ptr: int, layout: str = 'bgr', stream: int = 0):
        with self._lock:
            self._check(self._lib.rig_hash_batch(self._h, imgs_ptr, n, _LAYOUTS[layout], hashes_ptr, stream or None))

    def dedup(self, hashes_ptr: int, n_total: int, row_begin: int, row_end: int, threshold: int, drop_ptr: int,
              hits_ptr: int = 0, stream: int = 0):
        with self._lock:
            self._check(self._lib.rig_dedup(self._h, hashes_ptr, n_total, row_begin, row_end, threshold, drop_ptr,
                                            hits_ptr or None, stream or None))


    def int_pipe_peak(self, which: str = 'popc') -> float:
        """Measured lane-operations per second of the POPC ('popc') or ALU ('lop3') pipe on this device (rig_int_pipe_peak)."""
        v = ctypes.c_double()
        with self._lock:
            self._check(self._lib.rig_int_pipe_peak(self._h, {'popc': 0, 'lop3': 1}[which], ctypes.byref(v)))
        return float(v.value)


# ---- drop-in for the reference method ---------------------------------------------------------
_default: Dict[tuple, Renderer] = {}
_default_lock = threading.Lock()


def default_renderer(W: int, H: int, device: int = 0) -> Renderer:
    """Lazily created per (W, H, device); the CUDA context is only touched here, i.e. in the
    process (and thread) that first renders -- never at import, so fork()ed pool workers are safe."""
    key = (int(W), int(H), int(device))
    with _default_lock:
        r = _default.get(key)
        if r is None:
            r = _default[key] = Renderer(W, H, device=device, max_panels=64, max_elements=1024)
        return r


def render_elements_to_canvas(W: int, H: int, bg_color: Sequence[int], elements: Iterable[Dict[str, Any]],
                              use_grid: bool = False, grid_size: int = 3, device: int = 0) -> np.ndarray:
    """Same contract as RPMGenerator._render_elements_to_canvas: uint8[H, W, 3] BGR, `elements` untouched,
    rand_color() draws for colourless elements consumed here in element order."""
    sb = SceneBuilder(W, H, bg_color, grid_size)
    sb.add_panel(elements, use_grid)
    return default_renderer(W, H, device).render(sb.finish(), layout='bgr')[0]


def install(generator_cls, device: int = 0):
    """Replaces generator_cls._render_elements_to_canvas (generator.py:66) by the GPU path."""
    def _render(self, W, H, bg_color, elements, use_grid=False, grid_size=3):
        return render_elements_to_canvas(W, H, bg_color, elements, use_grid, grid_size, device=device)
    generator_cls._render_elements_to_canvas = _render
    return generator_cls
