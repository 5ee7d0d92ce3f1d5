"""oracle/dedup_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED: the reference contains no cross-sample near-duplicate filter.  Its README mentions
"filtering semantically duplicate samples" (/root/reference/README.md:8) but the only code with that
word in mind, multigraph_generation/check.py:82-140, is an in-bounds check whose result is printed
(multigraph_generation/generator.py:554-563), and the only duplicate logic is the in-sample option
equality of src/generator.py:148-177,465-469.  The perceptual-hash filter is therefore NEW functionality
whose semantics this file DEFINES (numpy, integer-exact) and the CUDA kernels must reproduce bit for bit:

  hash   Y = (77 R + 150 G + 29 B) >> 8 per pixel; S_b = sum of Y over block b of a 16 x 16 grid
         (blocks of H/16 x W/16 pixels, b = 16*by + bx); T = sum_b S_b; bit b = (256 * S_b > T).
         256 bits -> 4 little-endian uint64 words, block b in word b // 64 at bit b % 64.
  filter sample j is dropped iff there is an i < j with popcount(h_i xor h_j) <= threshold.
         (Order-independent and parallel; NOT sequential-greedy: a dropped i still drops later j.)

The one duplicate rule the reference does define -- `_elements_set` -- is restated in `elements_set`
below and pinned against decisions recorded from the live reference (tests/golden/dedup_keys.json).
"""
from __future__ import annotations

import numpy as np

GRID = 16
WORDS = GRID * GRID // 64


def phash(img: np.ndarray, layout: str = 'bgr') -> np.ndarray:
    """uint8[H, W, 3] -> uint64[4]"""
    H, W = img.shape[:2]
    a = img.astype(np.uint32)
    r, g, b = (a[..., 2], a[..., 1], a[..., 0]) if layout == 'bgr' else (a[..., 0], a[..., 1], a[..., 2])
    y = (77 * r + 150 * g + 29 * b) >> 8
    s = y.reshape(GRID, H // GRID, GRID, W // GRID).sum(axis=(1, 3)).astype(np.uint64).reshape(-1)
    t = s.sum()
    bits = (np.uint64(256) * s > t)
    out = np.zeros(WORDS, dtype=np.uint64)
    for bidx in np.nonzero(bits)[0]:
        out[bidx // 64] |= np.uint64(1) << np.uint64(bidx % 64)
    return out


def hamming(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """popcount(a xor b) summed over the last axis (uint64 words)"""
    return np.bitwise_count(a ^ b).sum(axis=-1)


def drop_flags(hashes: np.ndarray, threshold: int, row_begin: int = 0, row_end=None, block: int = 1024) -> np.ndarray:
    """hashes uint64[n, 4] -> uint8[n]; rows i restricted to [row_begin, row_end)"""
    n = len(hashes)
    row_end = n if row_end is None else row_end
    drop = np.zeros(n, dtype=np.uint8)
    for i0 in range(row_begin, row_end, block):
        i1 = min(i0 + block, row_end)
        for j0 in range(i0, n, block):
            j1 = min(j0 + block, n)
            d = hamming(hashes[i0:i1, None, :], hashes[None, j0:j1, :])
            ii = np.arange(i0, i1)[:, None]
            jj = np.arange(j0, j1)[None, :]
            hit = (d <= threshold) & (jj > ii)
            drop[j0:j1] |= hit.any(axis=0).astype(np.uint8)
    return drop


def count_pairs(hashes: np.ndarray, threshold: int, block: int = 1024) -> int:
    n = len(hashes)
    total = 0
    for i0 in range(0, n, block):
        for j0 in range(i0, n, block):
            d = hamming(hashes[i0:i0 + block, None, :], hashes[None, j0:j0 + block, :])
            ii = np.arange(i0, min(i0 + block, n))[:, None]
            jj = np.arange(j0, min(j0 + block, n))[None, :]
            total += int(((d <= threshold) & (jj > ii)).sum())
    return total


# ---- the reference's own in-sample duplicate rule (src/generator.py:148-177) --------------------
def _hashable(v):
    if isinstance(v, dict):
        return tuple(sorted((k, _hashable(x)) for k, x in v.items()))
    if isinstance(v, (list, tuple)):
        return tuple(_hashable(x) for x in v)
    if isinstance(v, set):
        return tuple(sorted(_hashable(x) for x in v))
    if isinstance(v, (np.integer, np.floating)):
        return v.item()
    try:
        hash(v)
        return v
    except TypeError:
        return str(v)


def element_key(el):
    return tuple(sorted((k, _hashable(v)) for k, v in el.items() if k not in ('timestamp', 'temp_id')))


def elements_set(elements):
    return frozenset(element_key(e) for e in elements)
