"""Near-duplicate stage: host-side pieces.

* `elements_set` / `element_key`: the in-sample option de-dup of the reference
  (/root/reference/src/generator.py:148-177, used at :465-469).  It is exact structural equality of
  element dicts and stays on the host; it is restated here so the drop-in generator does not need the
  reference module for it.
* `DedupFilter`: cross-sample perceptual-hash filter on the GPU (new functionality, see include/rig.h).
  Hashes live in HBM as uint64[n, 4]; with several ranks every rank hashes its own panels, the hash
  tables are all-gathered (NCCL over NVLink under torch.distributed), each rank scans a balanced block
  of rows of the upper-triangular pair matrix, and the drop flags are max-reduced.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

HASH_WORDS = 4


def _hashable(v):
    if isinstance(v, dict):
        return tuple(sorted((k, _hashable(val)) for k, val in v.items()))
    if isinstance(v, (list, tuple)):
        return tuple(_hashable(item) for item in v)
    if isinstance(v, set):
        return tuple(sorted(_hashable(item) for item in v))
    if isinstance(v, (np.integer, np.floating)):
        return v.item()
    try:
        hash(v)
        return v
    except TypeError:
        return str(v)


def element_key(elem) -> Tuple:
    """RPMGenerator._element_to_key: order-insensitive key of one element dict."""
    return tuple(sorted((k, _hashable(v)) for k, v in elem.items() if k not in ['timestamp', 'temp_id']))


def elements_set(elements):
    """RPMGenerator._elements_set: frozenset of element keys (multiplicity- and order-insensitive)."""
    return frozenset(element_key(e) for e in elements)


def row_blocks(n_total: int, world_size: int):
    """Row ranges with (nearly) equal pair counts: rank g scans rows [r_g, r_{g+1}) of the upper triangle,
    r_g = n * (1 - sqrt(1 - g / G)).  Returns world_size (begin, end) tuples covering [0, n)."""
    edges = [int(round(n_total * (1.0 - math.sqrt(max(0.0, 1.0 - g / world_size))))) for g in range(world_size)]
    edges.append(n_total)
    for g in range(1, world_size + 1):
        edges[g] = max(edges[g], edges[g - 1])
    return [(edges[g], edges[g + 1]) for g in range(world_size)]


def pair_count(begin: int, end: int, n_total: int) -> int:
    """Pairs (i, j), i < j, with begin <= i < end."""
    rows = end - begin
    return rows * n_total - (begin + end - 1) * rows // 2 - rows


def shard_panels(n_panels: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of sample indices per rank (render path: no collective)."""
    base, rem = divmod(n_panels, world_size)
    p0 = rank * base + min(rank, rem)
    return p0, p0 + base + (1 if rank < rem else 0)


class DedupFilter:
    """GPU hash + all-pairs Hamming scan bound to one `Renderer` (device buffers are torch tensors)."""

    def __init__(self, renderer):
        import torch
        self.torch = torch
        self.r = renderer
        self.dev = torch.device('cuda', renderer.device)

    def hash_images(self, imgs, layout: str = 'bgr'):
        """imgs: uint8 CUDA tensor [n, H, W, 3] -> int64 CUDA tensor [n, 4] holding the uint64 words."""
        torch = self.torch
        n = imgs.shape[0]
        out = torch.empty((n, HASH_WORDS), dtype=torch.int64, device=self.dev)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        self.r.hash_batch(imgs.data_ptr(), n, out.data_ptr(), layout=layout, stream=st)
        if st == 0:
            self.r.sync()
        return out

    def scan(self, hashes, threshold: int, row_begin: int = 0, row_end: Optional[int] = None, drop=None):
        """hashes: int64 CUDA [n, 4] (all samples).  Returns (drop uint8 CUDA [n], hit-pair count tensor)."""
        torch = self.torch
        n = hashes.shape[0]
        row_end = n if row_end is None else row_end
        if drop is None:
            drop = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        hits = torch.zeros(1, dtype=torch.int64, device=self.dev)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        self.r.dedup(hashes.data_ptr(), n, row_begin, row_end, threshold, drop.data_ptr(), hits.data_ptr(), stream=st)
        if st == 0:
            self.r.sync()
        return drop, hits

    def scan_distributed(self, local_hashes, threshold: int):
        """All ranks call this with the hashes of their own (contiguous, rank-ordered) samples.
        all_gather over NCCL -> balanced row block per rank -> MAX all-reduce of the flags."""
        import torch.distributed as dist
        torch = self.torch
        world, rank = dist.get_world_size(), dist.get_rank()
        n_local = int(local_hashes.shape[0])
        # one collective + one host read for the slice sizes; equal slices (the usual case) are gathered straight into place
        counts_t = torch.empty(world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(counts_t, torch.tensor([n_local], dtype=torch.int64, device=self.dev))
        counts = counts_t.tolist()
        n_total, n_max = sum(counts), max(counts)
        local = local_hashes.contiguous()
        if n_max * world == n_total:
            allh = torch.empty((n_total, HASH_WORDS), dtype=torch.int64, device=self.dev)
            dist.all_gather_into_tensor(allh, local)
        else:
            padded = torch.zeros((n_max, HASH_WORDS), dtype=torch.int64, device=self.dev)
            padded[:n_local] = local
            slab = torch.empty((world, n_max, HASH_WORDS), dtype=torch.int64, device=self.dev)
            dist.all_gather_into_tensor(slab, padded)
            allh = torch.cat([slab[g, :c] for g, c in enumerate(counts)], dim=0)
        b, e = row_blocks(n_total, world)[rank]
        drop, hits = self.scan(allh, threshold, b, e)           # on torch's current stream: the collectives below follow it
        dist.all_reduce(drop, op=dist.ReduceOp.MAX)
        dist.all_reduce(hits, op=dist.ReduceOp.SUM)
        return drop, hits, allh
