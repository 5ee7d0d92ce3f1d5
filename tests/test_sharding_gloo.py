"""N>1 host logic on CPU: world_size-2 gloo run of the sharding / exchange steps the GPU ranks perform
(panel blocks per rank with no collective on the render path; all-gather of hash tables, balanced row
blocks, MAX-reduce of drop flags for the filter).  The per-rank 'kernels' here are the numpy oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reasoning_image_generation_b200.dedup import pair_count, row_blocks, shard_panels
from oracle import dedup_oracle


def test_panel_shards_cover_everything_once():
    for n in (0, 1, 7, 1000, 100003):
        for world in (1, 2, 3, 8):
            spans = [shard_panels(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_row_blocks_balance_pairs():
    n = 1_000_000
    for world in (2, 4, 8):
        blocks = row_blocks(n, world)
        pairs = [pair_count(b, e, n) for b, e in blocks]
        assert sum(pairs) == n * (n - 1) // 2
        assert max(pairs) / (sum(pairs) / world) < 1.01


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, thr, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(0)                         # same table on every rank, each owns a slice
    h = rng.integers(-2 ** 63, 2 ** 63 - 1, (n, 4), dtype=np.int64)
    h[n - 5] = h[3]; h[n // 2 + 1] = h[n // 2 - 1]; h[7] = h[2]
    p0, p1 = shard_panels(n, world, rank)
    local = torch.from_numpy(h[p0:p1].copy())
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([p1 - p0]))
    parts = [torch.empty((int(c), 4), dtype=torch.int64) for c in counts]
    dist.all_gather(parts, local)
    allh = torch.cat(parts).numpy().view(np.uint64)
    b, e = row_blocks(n, world)[rank]
    drop = torch.from_numpy(dedup_oracle.drop_flags(allh, thr, b, e, block=512))
    dist.all_reduce(drop, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((drop.numpy().copy(), allh.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_equals_single_rank_scan():
    n, thr = 1500, 6
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, thr, q)) for r in range(2)]
    for p in procs:
        p.start()
    drop, allh = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = dedup_oracle.drop_flags(allh, thr, block=512)
    assert (drop == want).all() and set(np.nonzero(want)[0]) >= {7, n // 2 + 1, n - 5}


# ---- figure path: image blocks per rank, no collective on the render path; each rank's "kernel" here is the host path
# pipeline (librig.so, no GPU needed) + the CPU emulation of k_figure's per-lane arithmetic
def _figure_worker(rank, world, port, n, q):
    import ctypes
    from conftest import ROOT
    from reasoning_image_generation_b200.figure import host_edges
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    i0, i1 = shard_panels(n, world, rank)
    sc = synthetic_figure_scene(i1 - i0, 128, 128, seed=5, kind='single', first=i0)     # shards are generated independently
    layers, edges = host_edges(sc)
    lib = ctypes.CDLL(os.path.join(ROOT, 'tests', 'cpu_emul', 'libfigure_emul.so'))
    lib.figure_emul_render.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    out = np.zeros((i1 - i0, 128, 128, 3), np.uint8)
    for i in range(i1 - i0):
        L = np.ascontiguousarray(layers[layers[:, 0] == i])
        lib.figure_emul_render(128, 128, len(L), L.ctypes.data, edges.ctypes.data, 2, 256, out[i].ctypes.data)
    sums = torch.from_numpy(out.reshape(i1 - i0, -1).astype(np.int64).sum(axis=1))
    cap = (n + world - 1) // world                   # ragged shards: pad to the largest one for the gather
    padded = torch.full((cap,), -1, dtype=torch.int64)
    padded[:i1 - i0] = sums
    parts = [torch.empty(cap, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(parts, padded)                   # only the test gathers: the render path itself has no exchange
    if rank == 0:
        allv = torch.cat(parts).numpy()
        q.put(allv[allv >= 0].copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_figure_shards_equal_the_whole_scene(built):
    from oracle.agg_oracle import AggOracle
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    n = 21
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_figure_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = AggOracle().render(synthetic_figure_scene(n, 128, 128, seed=5, kind='single'), 128, 128)
    assert (sums == ref.reshape(n, -1).astype(np.int64).sum(axis=1)).all()
