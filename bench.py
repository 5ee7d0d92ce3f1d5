#!/usr/bin/env python
"""bench.py -- rendered images/s at 512x512 on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path (oracle port)

A "step" is one pass of the hot path (k_expand + k_raster through rig_render_resident) over one batch of
synthetic scene descriptions of the sequence-panel shape, per GPU.
  value      whole-job images/s, scenes already resident in HBM, CUDA events on the launching stream,
             barrier + synchronize on both sides, max over ranks
  e2e        the same metric through the public API (Renderer.render) with HOST buffers: scene packing,
             H2D of the scene, rendering and the D2H copy of every pixel into pinned host memory are inside
             the timed region
  roofline   k_raster (the dominant kernel): algorithmic bytes (786,432 per image) / its CUDA-event time,
             against the measured HBM figure in MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference renderer (oracle/ref_render.py over the installed cv2 wheel)
             on this box's host cores, bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_BYTES = 512 * 512 * 3
PANELS_PER_STEP = 16384        # 12.9 GB of pixels per step per GPU: far larger than the 126 MB L2
E2E_PANELS = 4096              # 3.2 GB pinned host destination per e2e step
DEDUP_N = 1 << 20              # hashes in the all-pairs timing input (5.5e11 pairs)
METRIC = 'rendered images/sec at 512x512 (sequence-panel scenes)'
WORKLOAD = ('sequence-panel synthetic scenes (SURVEY 8d: 1.85 elements/panel, 11 kinds, AA strokes 1-3 px, '
            'grid/wrap-around mix), %d panels per step per GPU, 512x512x3 uint8' % PANELS_PER_STEP)


def fallback_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference renderer on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, p0, p1 = args
    from oracle import ref_render
    from reasoning_image_generation_b200.synth import scene_to_elements, synthetic_scene
    try:
        import cv2  # noqa: F401
        cv2.setNumThreads(1)
        be = ref_render.Cv2Backend()
    except Exception:
        be = ref_render.CBackend()
    sc = synthetic_scene(PANELS_PER_STEP, seed=seed)
    t = time.time()
    acc = 0
    for p in range(p0, p1):
        els, ug = scene_to_elements(sc, p)
        img = ref_render.render_panel(512, 512, (255, 255, 255), els, ug, 3, backend=be)
        acc += int(img[0, 0, 0])
    return p1 - p0, time.time() - t, be.name


def cpu_sample(seed, budget_s=14.0, cores=None):
    """Renders consecutive slices of the step's scene on all host cores until ~budget_s has passed."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    per = 24
    done, t0, backend = 0, time.time(), '?'
    with mp.get_context('fork').Pool(cores) as pool:
        while time.time() - t0 < budget_s and done + per * cores <= PANELS_PER_STEP:
            jobs = [(seed, done + i * per, done + (i + 1) * per) for i in range(cores)]
            res = pool.map(_cpu_worker, jobs)
            done += sum(r[0] for r in res)
            backend = res[0][2]
    dt = time.time() - t0
    return done, dt, cores, backend


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_panels, total_s = 0, 0.0
    steps = max(1, args.steps)
    budget = max(4.0, min(20.0, 90.0 / (steps + args.warmup)))
    backend, cores = '?', os.cpu_count() or 1
    for s in range(args.warmup + steps):
        n, dt, cores, backend = cpu_sample(seed=1000, budget_s=budget)
        if s >= args.warmup:
            sample_panels += n
            total_s += dt
    v = sample_panels / total_s
    kind = 'port'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total_s / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'panels_per_step_per_gpu': PANELS_PER_STEP},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': kind,
                         'sample': '%d panels of the step (first slices of the seed-1000 scene) per timed step, '
                                   'oracle/ref_render.py (%s backend), one process per core' % (sample_panels // steps, backend)},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from reasoning_image_generation_b200 import Renderer
    from reasoning_image_generation_b200.dedup import DedupFilter
    from reasoning_image_generation_b200.synth import synthetic_scene

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    # stdout carries exactly one JSON line: anything libraries print there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; this repository has no CPU rendering path (use --impl reference '
                         'for the CPU baseline)')
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n = PANELS_PER_STEP
    scene = synthetic_scene(n, seed=1000 + rank)
    r = Renderer(512, 512, device=local, max_panels=n, max_elements=max(scene.n_elements, 1))
    out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    r.upload(scene, 'rgb')

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n_warm = 0
    t_warm = time.perf_counter()
    while n_warm < max(3, args.warmup) or time.perf_counter() - t_warm < 0.6:   # >= 3 steps and long enough for clocks / sampler
        r.render_resident(out.data_ptr(), stream=st)
        torch.cuda.synchronize()
        n_warm += 1
    barrier()
    r.sync()
    r.set_profiling(True)
    launches0 = r.stats()['kernel_launches']
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        r.render_resident(out.data_ptr(), stream=st)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    r.sync()
    stats = r.stats()
    r.set_profiling(False)
    launches = stats['kernel_launches'] - launches0
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max / 1e3)

    # ---- e2e: public API, host scene in, pinned host pixels out
    sub = scene.slice(0, E2E_PANELS)
    pinned = r.pinned_output(E2E_PANELS)
    for _ in range(2):
        r.render(sub, 'rgb', out=pinned)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 4))
    for _ in range(e2e_steps):
        r.render(sub, 'rgb', out=pinned)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * E2E_PANELS * e2e_steps / float(t.item())
    h2d = 40 * sub.n_elements + 4 * (sub.n_panels + 1) + sub.n_panels
    d2h = E2E_PANELS * IMG_BYTES
    checksum = int(pinned[::257, 100, 100, 0].astype('int64').sum())

    # ---- near-duplicate stage (reported beside `value`, not part of it)
    # (1) perceptual hash of the rendered step (HBM-read-bound: 786,432 B per image)
    f = DedupFilter(r)
    hashes = f.hash_images(out, 'rgb')                            # warm-up (lazy module load, scratch allocation)
    torch.cuda.synchronize(); r.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    hashes = f.hash_images(out, 'rgb')
    e1.record(stream)
    barrier(); r.sync()
    hash_ms = e0.elapsed_time(e1)
    # (1b) the same hashes from the fused variant (block sums taken from the band while it is in shared memory)
    r.upload(scene, 'rgb')                                        # the e2e leg left its 4096-panel sub-scene resident
    fused = torch.zeros_like(hashes)
    r.render_resident_hashed(out.data_ptr(), fused.data_ptr(), stream=st)
    barrier(); r.sync()
    e0.record(stream)
    r.render_resident_hashed(out.data_ptr(), fused.data_ptr(), stream=st)
    e1.record(stream)
    barrier(); r.sync()
    fused_ms = e0.elapsed_time(e1)
    fused_equal = bool(torch.equal(fused, hashes))
    drop_r, hits_r = f.scan(hashes, 10)                           # also the warm-up of the scan kernel
    torch.cuda.synchronize(); r.sync()
    # (2) all-pairs scan on the SURVEY 8(d) timing input: uniform random 256-bit hashes (seed 0) with 1 % planted
    #     near-duplicates; every rank owns a contiguous slice, tables are all-gathered over NCCL, rows split evenly
    n_total = DEDUP_N
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    table = torch.randint(-2 ** 62, 2 ** 62, (n_total, 4), dtype=torch.int64, device='cuda', generator=g)
    idx = torch.randint(0, n_total, (n_total // 100, 2), device='cuda', generator=g)
    table[idx[:, 1]] = table[idx[:, 0]] ^ (1 << torch.randint(0, 62, (n_total // 100, 1), device='cuda', generator=g))
    from reasoning_image_generation_b200.dedup import shard_panels
    s0, s1 = shard_panels(n_total, world, rank)
    local = table[s0:s1].contiguous()
    barrier()
    e0.record(stream)
    if dist is not None:
        drop, hits, allh = f.scan_distributed(local, 10)
    else:
        drop, hits = f.scan(table, 10)
    e1.record(stream)
    barrier(); r.sync()
    dd_ms = e0.elapsed_time(e1)
    t = torch.tensor([dd_ms, hash_ms], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dd_ms, hash_ms = float(t[0].item()), float(t[1].item())
    pairs = n_total * (n_total - 1) // 2

    if rank == 0:
        peak, peak_src = fallback_peak()
        raster_s = stats['raster_ms'] / 1e3
        achieved = stats['timed_panels'] * IMG_BYTES / raster_s / 1e9 if raster_s > 0 else 0.0
        per_launch = stats['timed_panels'] / max(1, stats['timed_launches'])
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'raster_traffic.json')
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp))['dram_bytes_per_panel'] * per_launch
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            done, cdt, cores, backend = cpu_sample(seed=1000, budget_s=14.0)
            cpu = {'value': done / cdt, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                   'sample': 'first %d panels of the step scene (seed 1000), oracle/ref_render.py (%s backend), '
                             'one process per core, %.1f s' % (done, backend, cdt)}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': n_warm, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'panels_per_step_per_gpu': n, 'l2': 'outputs (12.9 GB/step) exceed L2; no flush needed',
                       'elements_per_step_per_gpu': scene.n_elements, 'out_layout': 'RGB'},
            'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'panels_per_step_per_gpu': E2E_PANELS, 'note': 'Renderer.render: host SoA scene in, every pixel copied to '
                    'pinned host memory (PCIe-bound); checksum %d' % checksum},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'kernel': 'k_raster', 'peak_source': peak_src,
                         'launches': stats['timed_launches'], 'panels_per_launch': per_launch,
                         'expand_ms_per_step': stats['expand_ms'] / args.steps, 'raster_ms_per_step': stats['raster_ms'] / args.steps,
                         'whole_step_frac': value / world * IMG_BYTES / 1e9 / peak},
            'cpu_baseline': cpu,
            'dedup': {'pairs_per_s': pairs / (dd_ms / 1e3), 'n_hashes': n_total, 'pairs': pairs, 'ms': dd_ms, 'threshold': 10,
                      'hash_bits': 256, 'hit_pairs': int(hits.item()), 'dropped': int(drop.sum().item()),
                      'input': 'uniform random 256-bit hashes, seed 0, 1 % planted near-duplicates (SURVEY 8d)',
                      'exchange': 'nccl all_gather + max all_reduce' if dist is not None else 'none (1 rank)',
                      'hash_images_per_s': world * n / (hash_ms / 1e3), 'hash_read_gbs_per_gpu': n * IMG_BYTES / (hash_ms / 1e3) / 1e9,
                      'render_plus_fused_hash_ms_per_step': fused_ms, 'fused_hash_equals_standalone': fused_equal,
                      'rendered_step_dropped': int(drop_r.sum().item())},
            'clocks': clocks,
        }
        os.write(json_fd, (json.dumps(line) + '\n').encode())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    r.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
