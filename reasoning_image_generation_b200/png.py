"""png.py -- device-side PNG writer (ctypes view of include/rig_png.h).

Drop-in for the reference's host-side PNG step -- ``cv2.imwrite(path, img)`` in src/utils.py:18-21 and ``fig.savefig`` in
multigraph_generation/generator.py:256 -- for images that are already on the GPU: ``PngEncoder.encode(ptr, n)`` returns the
finished PNG files (one bytes-like buffer + offsets), so only compressed bytes cross PCIe.  `write_files` stores them under
the reference's output layout.  The files are standard 8-bit RGB PNGs that decode to exactly the rendered pixels; their
bytes differ from libpng's (different deflate strategy).  No CPU encoding path exists here.
"""
from __future__ import annotations

import ctypes
import re
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from .render import PinnedBuffer


class RigPngStats(ctypes.Structure):
    """struct rig_png_stats (include/rig_png.h)"""
    _fields_ = [('kernel_launches', ctypes.c_uint64), ('images', ctypes.c_uint64), ('png_bytes', ctypes.c_uint64),
                ('encode_ms', ctypes.c_double), ('timed_images', ctypes.c_uint64)]


PNG_EXPORTS = ('rig_png_create', 'rig_png_destroy', 'rig_png_last_error', 'rig_png_bound', 'rig_png_encode', 'rig_png_get_stats',
               'rig_png_set_profiling', 'rig_png_staging', 'rig_png_wait_event', 'rig_png_encode_begin', 'rig_png_encode_end',
               'rig_png_set_blank_hint', 'rig_png_debug_code')

_bound = False
_LAYOUTS = {'rgb': _abi.RIG_LAYOUT_RGB, 'bgr': _abi.RIG_LAYOUT_BGR}


def _lib():
    global _bound
    lib = _abi.load()
    if not _bound:
        vp = ctypes.c_void_p
        lib.rig_png_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        lib.rig_png_create.restype = ctypes.c_int
        lib.rig_png_destroy.argtypes = [vp]
        lib.rig_png_destroy.restype = None
        lib.rig_png_last_error.argtypes = [vp]
        lib.rig_png_last_error.restype = ctypes.c_char_p
        lib.rig_png_bound.argtypes = [vp, ctypes.c_int]
        lib.rig_png_bound.restype = ctypes.c_size_t
        lib.rig_png_encode.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_size_t, vp, ctypes.c_int, vp]
        lib.rig_png_encode.restype = ctypes.c_int
        lib.rig_png_encode_begin.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_size_t, vp]
        lib.rig_png_encode_begin.restype = ctypes.c_int
        lib.rig_png_encode_end.argtypes = [vp, ctypes.c_int, vp]
        lib.rig_png_encode_end.restype = ctypes.c_int
        lib.rig_png_set_blank_hint.argtypes = [vp, vp, ctypes.c_int]
        lib.rig_png_set_blank_hint.restype = ctypes.c_int
        lib.rig_png_wait_event.argtypes = [vp, vp]
        lib.rig_png_wait_event.restype = ctypes.c_int
        lib.rig_png_staging.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp)]
        lib.rig_png_staging.restype = ctypes.c_int
        lib.rig_png_get_stats.argtypes = [vp, ctypes.POINTER(RigPngStats)]
        lib.rig_png_get_stats.restype = ctypes.c_int
        lib.rig_png_set_profiling.argtypes = [vp, ctypes.c_int]
        lib.rig_png_set_profiling.restype = ctypes.c_int
        _bound = True
    return lib


class PngError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'librig error {code}: {message}')
        self.code = code


class PngEncoder:
    """One CUDA context-bound PNG writer for W x H x 3 images (one per process per GPU)."""

    def __init__(self, W: int = 512, H: int = 512, device: int = 0, max_images: int = 4096, host_capacity: Optional[int] = None,
                 channels: int = 3):
        """channels=3: RGB files (cv2.imwrite of a 3-channel array); channels=4: RGBA files with alpha 255, the format
        fig.savefig writes.  The device images are 3 bytes per pixel either way."""
        self._lib = _lib()
        self.W, self.H, self.device, self.max_images, self.channels = int(W), int(H), int(device), int(max_images), int(channels)
        h = ctypes.c_void_p()
        rc = self._lib.rig_png_create(self.device, self.W, self.H, self.max_images, self.channels, ctypes.byref(h))
        if rc:
            raise PngError(rc, (self._lib.rig_png_last_error(None) or b'').decode())
        self._h = h
        self._lock = threading.Lock()
        # pinned destination: sized for typical puzzle images (64 KB each), grown on RIG_E_CAPACITY
        # (one per output slot of the library: the copy of one batch overlaps the kernels of the next)
        cap = int(host_capacity if host_capacity else max(1 << 20, self.max_images * 65536))
        self._pins = [PinnedBuffer(cap), PinnedBuffer(cap)]
        self._offs = [np.zeros(self.max_images + 1, dtype=np.uint64) for _ in range(2)]
        self._slot_n = [0, 0]
        self._begun = 0

    def _check(self, rc: int):
        if rc:
            raise PngError(rc, (self._lib.rig_png_last_error(self._h) or b'').decode())

    def close(self):
        if getattr(self, '_h', None):
            self._lib.rig_png_destroy(self._h)
            self._h = None
        for pin in getattr(self, '_pins', None) or ():
            pin.close()
        self._pins = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bound(self, n: int) -> int:
        return int(self._lib.rig_png_bound(self._h, int(n)))

    def encode_begin(self, imgs_ptr: int, n: int, layout: str = 'rgb', stream: int = 0) -> int:
        """First half of `encode` (rig_png_encode_begin): runs the kernels and starts the copy of the files to pinned host memory.
        Returns the slot to hand to `encode_end`.  At most two batches may be in flight."""
        if n > self.max_images:
            raise ValueError(f'{n} images > max_images {self.max_images}')
        with self._lock:
            s = self._begun & 1
            for _ in range(3):
                pin = self._pins[s]
                rc = self._lib.rig_png_encode_begin(self._h, imgs_ptr, n, _LAYOUTS[layout], pin.ptr, pin.nbytes, stream or None)
                msg = (self._lib.rig_png_last_error(self._h) or b'') if rc == _abi.RIG_E_CAPACITY else b''
                if b'out_capacity' in msg:
                    # the message carries the exact size ("the PNG files need N bytes"): grow to it in one step (with slack for the
                    # next batches), never past the worst case; the library keeps the blank-band hint for the retry
                    m = re.search(rb'need (\d+) bytes', msg)
                    need = int(m.group(1)) if m else 2 * pin.nbytes
                    pin.close()
                    self._pins[s] = PinnedBuffer(min(self.bound(n), max(need + need // 4, 2 * pin.nbytes)))
                    continue
                break
            if rc < 0:
                self._check(rc)
            assert rc == s, (rc, s)
            self._slot_n[s] = n
            self._begun += 1
        return s

    def encode_end(self, slot: int) -> Tuple[np.ndarray, np.ndarray]:
        """Second half of `encode`: waits for the copy of `slot`; the views are valid until that slot is begun again (two
        `encode_begin` calls later)."""
        with self._lock:
            self._check(self._lib.rig_png_encode_end(self._h, slot, self._offs[slot].ctypes.data))
        off = self._offs[slot][:self._slot_n[slot] + 1]
        return self._pins[slot].array[:int(off[-1])], off

    def encode(self, imgs_ptr: int, n: int, layout: str = 'rgb', stream: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """n images at device pointer `imgs_ptr` -> (uint8 buffer view in pinned host memory, uint64 offsets[n + 1]).
        File i is buffer[offsets[i]:offsets[i + 1]].  The views are valid until the next call."""
        return self.encode_end(self.encode_begin(imgs_ptr, n, layout, stream))

    def set_blank_hint(self, flags_ptr: int, rows_per_band: int):
        """Blank-band flags (device pointer, e.g. FigureRenderer.blank_bands()) for the next encode / encode_begin only: rows marked 1
        are white and are encoded without being read."""
        self._check(self._lib.rig_png_set_blank_hint(self._h, flags_ptr or None, int(rows_per_band)))

    def wait_event(self, event: int):
        """The next `encode` starts only after `event` (a cudaEvent_t handle, e.g. FigureRenderer.fence()) has completed on the GPU."""
        self._check(self._lib.rig_png_wait_event(self._h, event))

    def staging(self, n: int) -> int:
        """Device pointer of a context-owned buffer for n images (render into it, then `encode` it)."""
        p = ctypes.c_void_p()
        self._check(self._lib.rig_png_staging(self._h, int(n), ctypes.byref(p)))
        return int(p.value)

    def set_profiling(self, on: bool):
        self._check(self._lib.rig_png_set_profiling(self._h, 1 if on else 0))

    def stats(self) -> Dict[str, float]:
        st = RigPngStats()
        self._check(self._lib.rig_png_get_stats(self._h, ctypes.byref(st)))
        return dict(kernel_launches=int(st.kernel_launches), images=int(st.images), png_bytes=int(st.png_bytes),
                    encode_ms=float(st.encode_ms), timed_images=int(st.timed_images))


def chunk_plan(n: int, chunk: int) -> List[Tuple[int, int]]:
    """[i0, i1) ranges of `chunk` items (the last one shorter) covering n.  (Shorter first and last chunks, to shrink the two
    pipeline stages nothing overlaps, measured no better: launches of 1-2 k images run at half the rate of 4 k ones.)"""
    n, chunk = int(n), max(1, int(chunk))
    return [(i, min(n, i + chunk)) for i in range(0, n, chunk)]


def pipelined_png(encoder: 'PngEncoder', ranges: Sequence[Tuple[int, int]], chunk: int, img_bytes: int, render_chunk, layout: str = 'rgb',
                  overlap: bool = True):
    """The loop behind FigureRenderer.iter_png and Renderer.iter_png.  `ranges`: (i0, i1) item ranges of at most `chunk` items
    (chunk_plan).  render_chunk(i0, i1, device_ptr) renders those items into device_ptr
    and returns (n_images, event[, (flags_ptr, rows_per_band)]): `event` is a cudaEvent_t handle recorded behind the render, or
    0 when the render has already finished; the optional third item is the renderer's blank-band flags for that chunk
    (PngEncoder.set_blank_hint).  With `overlap` a worker thread renders the chunks ahead into the two halves of the encoder's
    staging buffer while the calling thread encodes; the copy of one chunk's files overlaps the kernels of the next (two encoder
    slots).  Yields (i0, buffer, offsets)."""
    import queue
    starts = list(ranges)
    if not starts:
        return
    two = overlap and len(starts) > 1
    # staging slots the worker may render into: with three, the render of chunk k + 2 is queued on the GPU while chunk k + 1 is
    # being encoded, so a slow host wake-up (many ranks on one box) leaves no hole in the GPU's queue
    n_slots = 3 if len(starts) > 2 else 2
    base = encoder.staging(n_slots * chunk if two else chunk)
    if not two:
        for i0, i1 in starts:
            n_img, ev, *hint = render_chunk(i0, i1, base)
            if ev:
                encoder.wait_event(ev)
            if hint and hint[0][0]:
                encoder.set_blank_hint(*hint[0])
            buf, off = encoder.encode(base, n_img, layout, 0)
            yield i0, buf, off
        return
    free = threading.Semaphore(n_slots)
    ready: 'queue.Queue' = queue.Queue()
    stop = threading.Event()

    def worker():
        try:
            for k, (i0, i1) in enumerate(starts):
                free.acquire()
                if stop.is_set():
                    return
                ptr = base + (k % n_slots) * chunk * img_bytes
                n_img, ev, *hint = render_chunk(i0, i1, ptr)
                ready.put((i0, n_img, ptr, ev, hint[0] if hint else None))
        except BaseException as e:                               # re-raised on the consumer's thread
            ready.put(e)

    t = threading.Thread(target=worker, daemon=True)
    t.start()
    pending = None                                               # (first item, encoder slot) whose copy is in flight
    try:
        for _ in starts:
            item = ready.get()
            if isinstance(item, BaseException):
                raise item
            i0, n_img, ptr, ev, hint = item
            if ev:
                encoder.wait_event(ev)                           # exactly this chunk's render, not what the worker enqueues next
            if hint and hint[0]:
                encoder.set_blank_hint(*hint)
            slot = encoder.encode_begin(ptr, n_img, layout, 0)   # kernels run while the previous chunk's files are copied
            free.release()                                       # the encoder has read this half
            if pending is not None:
                p0, pslot = pending
                pending = None
                buf, off = encoder.encode_end(pslot)
                pending = (i0, slot)
                yield p0, buf, off
            else:
                pending = (i0, slot)
        if pending is not None:
            p0, pslot = pending
            pending = None
            buf, off = encoder.encode_end(pslot)
            yield p0, buf, off
    finally:
        stop.set()
        if pending is not None:                                  # closed early: retire the slot in flight
            encoder.encode_end(pending[1])
        for _ in range(n_slots):
            free.release()
        t.join()


def split_files(buf: np.ndarray, offsets: np.ndarray) -> List[bytes]:
    return [bytes(buf[int(offsets[i]):int(offsets[i + 1])]) for i in range(len(offsets) - 1)]


def write_files(buf: np.ndarray, offsets: np.ndarray, paths: Sequence[str]):
    """Stores file i at paths[i] (e.g. out_dir/samples/sample_%06d/state_0.png, the reference's layout)."""
    for i, p in enumerate(paths):
        os.makedirs(os.path.dirname(p) or '.', exist_ok=True)
        with open(p, 'wb') as f:
            f.write(memoryview(buf[int(offsets[i]):int(offsets[i + 1])]))


def decode_png_strict(data: bytes) -> np.ndarray:
    """A small, strict PNG reader used by the tests (not a product path): checks the signature, every chunk CRC, the zlib
    Adler-32 (via zlib) and undoes the row filters; returns uint8[H, W, 3] (RGB files) or uint8[H, W, 4] (RGBA files)."""
    import struct
    import zlib
    assert data[:8] == b'\x89PNG\r\n\x1a\n', 'bad signature'
    pos, idat, W = 8, b'', None
    H, ch = 0, 3
    while pos < len(data):
        (ln,), typ = struct.unpack('>I', data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + ln]
        (crc,) = struct.unpack('>I', data[pos + 8 + ln:pos + 12 + ln])
        assert zlib.crc32(typ + body) & 0xffffffff == crc, f'bad CRC in {typ!r}'
        if typ == b'IHDR':
            W, H, depth, ctype, comp, flt, inter = struct.unpack('>IIBBBBB', body)
            assert (depth, comp, flt, inter) == (8, 0, 0, 0) and ctype in (2, 6)
            ch = 3 if ctype == 2 else 4
        elif typ == b'IDAT':
            idat += body
        elif typ == b'IEND':
            assert pos + 12 + ln == len(data), 'bytes after IEND'
        pos += 12 + ln
    raw = zlib.decompress(idat)
    stride = 1 + ch * W
    assert len(raw) == H * stride
    rows = np.frombuffer(raw, np.uint8).reshape(H, stride)
    out = np.zeros((H, W * ch), np.uint8)
    for y in range(H):
        ft, line = rows[y, 0], rows[y, 1:].astype(np.int64)
        if ft == 0:
            out[y] = line
        elif ft == 1:
            out[y] = (np.cumsum(line.reshape(W, ch), axis=0) % 256).reshape(-1)
        elif ft == 2:
            out[y] = (line + (out[y - 1] if y else 0)) % 256
        else:
            raise AssertionError(f'filter type {ft} not written by this encoder')
    return out.reshape(H, W, ch)
