#!/usr/bin/env python
"""bench.py -- rendered images/s at 512x512 on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path (oracle port)

Headline workload (BASELINE.json configs[1]): multigraph single-shape variants (one outlined base patch with internal
lines or occlusion masks) at 512x512, the figure path.  A "step" is one pass of the hot path -- the whole of
RendererAgg.draw_path for every artist -- over one batch of synthetic figure scenes, per GPU: k_figure_path
(transform, snap, curve flattening, stroking, clipping: artist arrays -> polygon edges) followed by k_figure (coverage, blend,
store).
  value      whole-job images/s, ARTIST ARRAYS already resident in HBM (what draw_path receives: path vertices + codes, affine,
             line width, cap / join / snap, colours), CUDA events on the launching stream, barrier + synchronize on both sides,
             max over ranks
  e2e        the same metric through the public API (FigureRenderer.iter_png = fig.savefig for a batch) with HOST buffers:
             packing + H2D of the artist arrays, both kernels, the device-side PNG encode and the D2H copy of the files into
             pinned host memory are inside the timed region
  roofline   k_figure (the dominant kernel): algorithmic bytes (786,432 per image) / its CUDA-event time, against the
             measured HBM figure in MEASURED_PEAKS.json; whole_step_frac covers both kernels
  cpu_baseline  oracle/aggdraw.c (the restated matplotlib/Agg renderer) on this box's host cores, bounded sample
  sequence_panels  the same quantities for the sequence-panel path (k_expand + k_raster, BASELINE configs[3] shape), with
             its own roofline, e2e and cpu_baseline, plus the near-duplicate stage (`dedup`)
`--workload sequence` makes the sequence-panel path the headline instead (and skips the figure leg).
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
DEDUP_N = 10_000_000           # BASELINE configs[4]: hashes in the all-pairs timing input (5.0e13 pairs)
FIG_PER_STEP = 32768           # 25.8 GB of pixels per step per GPU
FIG_E2E = 32768               # images per e2e step (the whole step scene) ...
FIG_E2E_CHUNK = 4096          # ... rendered + encoded 4096 at a time (--e2e-chunk)
FIG_METRIC = 'rendered images/sec at 512x512 (multigraph single-shape variants)'
FIG_WORKLOAD = ('BASELINE configs[1]: multigraph single-shape variants, synthetic (SURVEY 8d: one outlined base patch -- circle / '
                'ellipse / rectangle / rounded box / regular polygon / wedge, 1.5-2.0 pt, alpha 0.9 -- plus 1-6 internal lines or 1-3 '
                'occlusion masks), %d images per step per GPU, 512x512x3 uint8' % FIG_PER_STEP)
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
        if be.name == 'cv2':
            ok, png = cv2.imencode('.png', img)      # save_image = cv2.imwrite (src/utils.py:18-21): the PNG is the deliverable
            acc += len(png)
    return p1 - p0, time.time() - t, be.name + (' + cv2.imencode(.png)' if be.name == 'cv2' else '')


def cpu_sample(seed, budget_s=14.0, cores=None):
    """Renders consecutive slices of the step's scene on all host cores until ~budget_s has passed."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    per = 24
    done, t0, backend = 0, time.time(), '?'
    # 'spawn': the workers must not inherit this process's CUDA context (the GPU legs ran before the CPU baseline)
    with mp.get_context('spawn').Pool(cores) as pool:
        while time.time() - t0 < budget_s and done + per * cores <= PANELS_PER_STEP:
            jobs = [(seed, done + i * per, done + (i + 1) * per) for i in range(cores)]
            res = pool.map(_cpu_worker, jobs)
            done += sum(r[0] for r in res)
            backend = res[0][2]
    dt = time.time() - t0
    return done, dt, cores, backend


def _fig_cpu_worker(args):
    seed, i0, i1 = args
    from oracle.agg_oracle import AggOracle
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    sc = synthetic_figure_scene(i1 - i0, 512, 512, seed=seed, kind='single', first=i0)
    o = AggOracle()
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        cv2 = None
    t_render = t_png = 0.0
    acc = 0
    for i in range(sc.n_images):
        t = time.time()
        img = o.render_rgba(sc, i, 512, 512)
        t_render += time.time() - t
        if cv2 is not None:
            # fig.savefig writes the RGBA canvas as a PNG (generator.py:256); libpng through cv2.imencode, default settings
            t = time.time()
            ok, png = cv2.imencode('.png', img)
            t_png += time.time() - t
            acc += len(png)
    return i1 - i0, t_render, t_png, acc


def fig_cpu_sample(seed, budget_s=14.0, cores=None, pool=None):
    """Renders consecutive slices of the step's figure scene on all host cores until ~budget_s has passed
    (scene emission is outside each worker's own timer but inside the wall clock used for the rate)."""
    import contextlib
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    per = 48
    done, render_s, total_s, png_bytes = 0, 0.0, 0.0, 0
    # (not fork: see cpu_sample; the reference arm hands in one pool for all its steps)
    with (contextlib.nullcontext(pool) if pool is not None else mp.get_context('spawn').Pool(cores)) as pool:
        pool.map(_fig_cpu_worker, [(seed, 0, 1)] * cores)                 # workers up (interpreter + imports) before the clock starts
        t0 = time.time()
        while time.time() - t0 < budget_s:
            first = done % FIG_PER_STEP
            jobs = [(seed, first + i * per, first + (i + 1) * per) for i in range(cores)]
            res = pool.map(_fig_cpu_worker, jobs)
            done += sum(r[0] for r in res)
            render_s += max(r[1] for r in res)
            total_s += max(r[1] + r[2] for r in res)
            png_bytes += sum(r[3] for r in res)
    return {'images': done, 'render_s': render_s, 'total_s': total_s, 'cores': cores, 'png_bytes': png_bytes}


def fig_cpu_dict(c):
    return {'value': c['images'] / c['total_s'], 'unit': 'images/s', 'cores': c['cores'], 'kind': 'port',
            'render_only_value': c['images'] / c['render_s'], 'png_bytes_per_image': c['png_bytes'] / max(1, c['images']),
            'sample': '%d images of the step scene (seed 1000), one process per core: oracle/aggdraw.c (restated matplotlib/Agg '
                      'pipeline; matplotlib itself is not installable here) + the PNG encode savefig ends with (libpng via '
                      'cv2.imencode, RGBA, default settings); value = render + PNG, %.1f s; render_only_value = rasterisation alone, '
                      '%.1f s; scene emission is outside both' % (c['images'], c['total_s'], c['render_s'])}


def run_reference_figure(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle.agg_oracle import build_agg_oracle
    build_agg_oracle()
    steps = max(1, args.steps)
    budget = max(4.0, min(20.0, 90.0 / (steps + args.warmup)))
    tot = None
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    with mp.get_context('spawn').Pool(cores) as pool:
        for s in range(args.warmup + steps):
            c = fig_cpu_sample(seed=1000, budget_s=budget, cores=cores, pool=pool)
            if s >= args.warmup:
                if tot is None:
                    tot = c
                else:
                    for k in ('images', 'render_s', 'total_s', 'png_bytes'):
                        tot[k] += c[k]
    cpu = fig_cpu_dict(tot)
    v = cpu['value']
    line = {
        'impl': 'reference', 'metric': FIG_METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * tot['total_s'] / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': FIG_WORKLOAD, 'images_per_step_per_gpu': FIG_PER_STEP,
                   'deliverable': 'PNG files in host memory (what fig.savefig produces), as in the GPU arm\'s e2e'},
        'cpu_baseline': cpu,
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


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
class Env:
    """Process-wide state of the GPU arm: ranks, the launching stream, barrier and max-over-ranks helpers."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device; this repository has no CPU rendering path (use --impl reference '
                             'for the CPU baseline)')
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
            self.dist = dist
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.st = self.stream.cuda_stream

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device='cuda')
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]


def timed_steps(env, render, sync, steps, warmup):
    """>= max(3, warmup) untimed steps (and >= 0.6 s, so clocks and the sampler settle), then exactly `steps` steps between
    CUDA events on the launching stream, barrier + synchronize on both sides; returns (ms max over ranks, warm-up steps)."""
    torch = env.torch
    n_warm, t_warm = 0, time.perf_counter()
    while n_warm < max(3, warmup) or time.perf_counter() - t_warm < 0.6:
        render()
        torch.cuda.synchronize()
        n_warm += 1
    env.barrier()
    sync()
    return n_warm


def figure_leg(env, args, sampler):
    torch = env.torch
    from reasoning_image_generation_b200.figure import FigureRenderer
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    n = FIG_PER_STEP
    scene = synthetic_figure_scene(n, 512, 512, seed=1000, kind='single', first=env.rank * n)
    r = FigureRenderer(512, 512, device=env.local, max_images=n, max_artists=max(scene.n_artists, 1))
    out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
    r.upload(scene, 'rgb')
    if not r.stats()['geom_resident']:
        raise SystemExit('bench.py: the device geometry handed the step scene back to the host pipeline (status 0x%x); the headline '
                         'is defined on k_figure_path + k_figure' % r.stats()['geom_status'])
    render = lambda: r.draw_resident(out.data_ptr(), stream=env.st)   # noqa: E731
    if sampler is not None:
        sampler.start()
    n_warm = timed_steps(env, render, r.sync, args.steps, args.warmup)
    r.set_profiling(True)
    launches0 = r.stats()['kernel_launches']
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ev0.record(env.stream)
    for _ in range(args.steps):
        render()
    ev1.record(env.stream)
    env.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler is not None else None
    r.sync()
    stats = r.stats()
    r.set_profiling(False)
    (ms_max,) = env.max_over_ranks(ms)
    value = env.world * n * args.steps / (ms_max / 1e3)
    checksum_dev = int(out[::509, 200:312, 200:312, 0].to(torch.int64).sum().item())
    # ---- BASELINE configs[2] shape: multi-shape combinations (nested / adjacent / intersecting), resident render only
    n_multi = n // 2
    multi = synthetic_figure_scene(n_multi, 512, 512, seed=2000, kind='multi', first=env.rank * n_multi)
    r.upload(multi, 'rgb')
    for _ in range(3):
        r.draw_resident(out.data_ptr(), stream=env.st)
    r.sync()
    r.set_profiling(True)
    for _ in range(max(2, args.steps // 2)):
        r.draw_resident(out.data_ptr(), stream=env.st)
    r.sync()
    mst = r.stats()
    r.set_profiling(False)
    multi_cover, multi_geom = env.max_over_ranks(mst['cover_ms'] / max(1, mst['timed_launches']), mst['geom_ms'] / max(1, mst['geom_launches']))
    multi_ms = multi_cover + multi_geom
    multi_info = {'value': env.world * n_multi / (multi_ms / 1e3), 'unit': 'images/s', 'images_per_step_per_gpu': n_multi,
                  'ms_per_step': multi_ms, 'figure_ms': multi_cover, 'geometry_ms': multi_geom,
                  'edges_per_image': mst['edges'] / n_multi, 'layers_per_image': mst['layers'] / n_multi,
                  'hbm_frac': n_multi * IMG_BYTES / (multi_ms / 1e3) / 1e9 / fallback_peak()[0],
                  'hbm_frac_k_figure': n_multi * IMG_BYTES / (multi_cover / 1e3) / 1e9 / fallback_peak()[0],
                  'workload': 'BASELINE configs[2] shape: 2-3 outlined base patches, nested / adjacent / intersecting (synthetic), '
                              'k_figure_path + k_figure, artist arrays resident in HBM'}
    # ---- the reference CLI's real output size: figsize (8, 8) x dpi 200 = 1600 x 1600 (multigraph_generation/cli.py:24), resident draw
    big_info = None
    try:
        nb = 2048
        big = synthetic_figure_scene(nb, 1600, 1600, seed=3000, kind='single', first=env.rank * nb)
        rb = FigureRenderer(1600, 1600, device=env.local, max_images=nb, max_artists=max(big.n_artists, 1))
        outb = out.view(-1)[:nb * 1600 * 1600 * 3]
        rb.upload(big, 'rgb')
        if rb.stats()['geom_resident']:
            for _ in range(3):
                rb.draw_resident(outb.data_ptr(), stream=env.st)
            rb.sync()
            rb.set_profiling(True)
            for _ in range(max(2, args.steps // 2)):
                rb.draw_resident(outb.data_ptr(), stream=env.st)
            rb.sync()
            bst = rb.stats()
            b_cover, b_geom = env.max_over_ranks(bst['cover_ms'] / max(1, bst['timed_launches']), bst['geom_ms'] / max(1, bst['geom_launches']))
            b_ms = b_cover + b_geom
            big_bytes = 1600 * 1600 * 3
            big_info = {'value': env.world * nb / (b_ms / 1e3), 'unit': 'images/s', 'canvas': '1600x1600', 'images_per_step_per_gpu': nb,
                        'ms_per_step': b_ms, 'figure_ms': b_cover, 'geometry_ms': b_geom, 'edges_per_image': bst['edges'] / nb,
                        'bytes_per_image': big_bytes, 'hbm_frac': nb * big_bytes / (b_ms / 1e3) / 1e9 / fallback_peak()[0],
                        'hbm_frac_k_figure': nb * big_bytes / (b_cover / 1e3) / 1e9 / fallback_peak()[0],
                        'workload': 'single-shape variants at the reference CLI\'s output size (8 in x 200 dpi), k_figure_path + k_figure, '
                                    'artist arrays resident in HBM'}
        rb.close()
    except Exception as e:                                   # (reported, not fatal: the headline is the 512 x 512 workload)
        big_info = {'error': str(e)}
    del out
    torch.cuda.empty_cache()

    # ---- e2e (headline): fig.savefig for a batch -- host artist arrays in, PNG files in pinned host memory out
    from reasoning_image_generation_b200.png import PngEncoder
    from reasoning_image_generation_b200.render import PinnedBuffer
    sub = scene.slice(0, FIG_E2E)
    e2e_chunk = args.e2e_chunk
    enc = PngEncoder(512, 512, device=env.local, max_images=e2e_chunk, channels=4)      # RGBA files, as fig.savefig writes them
    r2 = FigureRenderer(512, 512, device=env.local, max_images=e2e_chunk, max_artists=max(sub.n_artists, 1))
    png_bytes = 0
    for _ in range(2):
        for i0, buf, off in r2.iter_png(sub, enc, e2e_chunk):
            pass
    env.barrier()
    l0 = r2.stats()['kernel_launches'] + enc.stats()['kernel_launches']
    t0 = time.perf_counter()
    e2e_steps = args.e2e_steps                      # 48 x 32 768 images: a timed region of more than a second
    digest = 0
    for _ in range(e2e_steps):
        for i0, buf, off in r2.iter_png(sub, enc, e2e_chunk):
            png_bytes += int(off[-1])
            digest ^= int(buf[int(off[1]) - 16]) if len(off) > 1 else 0
    torch.cuda.synchronize()
    (dt,) = env.max_over_ranks(time.perf_counter() - t0)
    st2 = r2.stats()
    e2e_launches = st2['kernel_launches'] + enc.stats()['kernel_launches'] - l0
    e2e_value = env.world * FIG_E2E * e2e_steps / dt
    # the artist arrays as the C ABI receives them + the three index tables the library adds (work order, layer slots)
    h2d = int(sum(a.nbytes for a in (sub.image_artist_begin, sub.artist_vertex_begin, sub.vertices, sub.codes, sub.affine, sub.linewidth, sub.cap,
                                      sub.join, sub.snap, sub.edge_rgba, sub.face_rgba, sub.has_face)) + 8 * sub.n_artists + 4 * (FIG_E2E + 1))
    geom_fallbacks = st2['geom_fallbacks']
    enc.close()
    r2.close()
    # ---- the same through the raw-pixel API: every pixel copied to pinned host memory (PCIe-bound)
    pin = PinnedBuffer(FIG_E2E_CHUNK * IMG_BYTES)
    pinned = pin.array.reshape(FIG_E2E_CHUNK, 512, 512, 3)
    sub_px = scene.slice(0, FIG_E2E_CHUNK)
    for _ in range(2):
        r.render(sub_px, 'rgb', out=pinned)
    env.barrier()
    t0 = time.perf_counter()
    px_steps = 4
    for _ in range(px_steps):
        r.render(sub_px, 'rgb', out=pinned)
    torch.cuda.synchronize()
    (dt_px,) = env.max_over_ranks(time.perf_counter() - t0)
    px_value = env.world * FIG_E2E_CHUNK * px_steps / dt_px
    per_img_edges = stats['edges'] / n
    checksum = int(pinned[::257, 200:312, 200:312, 0].astype('int64').sum())
    res = {
        'value': value, 'ms_per_step': ms_max / args.steps, 'warmup': n_warm, 'clocks': clocks,
        'launches': stats['kernel_launches'] - launches0, 'stats': stats, 'n': n, 'multi_shape': multi_info, 'dpi200': big_info,
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': png_bytes // e2e_steps,
                'images_per_step_per_gpu': FIG_E2E, 'chunk': e2e_chunk,
                'png_bytes_per_image': png_bytes / (e2e_steps * FIG_E2E), 'gpu_launches': e2e_launches,
                'timed_s': dt, 'steps': e2e_steps, 'geom_fallbacks': geom_fallbacks,
                'note': 'FigureRenderer.iter_png = fig.savefig for a batch: host artist arrays in (packed, one H2D copy per chunk), '
                        'k_figure_path + k_figure, device-side PNG encode (RGBA files), finished files copied to pinned host memory; '
                        'chunk k + 1 is prepared and rendered while chunk k is encoded and copied; digest %d' % digest},
        'e2e_pixels': {'value': px_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d // max(1, FIG_E2E // FIG_E2E_CHUNK),
                       'd2h_bytes_per_step': FIG_E2E_CHUNK * IMG_BYTES, 'images_per_step_per_gpu': FIG_E2E_CHUNK,
                       'note': 'FigureRenderer.render: the same with every raw pixel copied to pinned host memory instead of PNG '
                               'files (PCIe-bound: 786,432 B per image); checksum %d' % checksum},
        'config': {'workload': FIG_WORKLOAD, 'images_per_step_per_gpu': n, 'l2': 'outputs (25.8 GB/step) exceed L2; no flush needed',
                   'artists_per_step_per_gpu': scene.n_artists, 'vertices_per_step_per_gpu': scene.n_vertices, 'edges_per_image': per_img_edges,
                   'resident_input': 'artist arrays (draw_path arguments), %d bytes per image' % (h2d // FIG_E2E),
                   'layers_per_image': stats['layers'] / n, 'out_layout': 'RGB', 'checksum': checksum_dev,
                   'parity': 'pixels bit-exact against oracle/aggdraw.c (the restated matplotlib -> Agg pipeline) in tests/; oracle-exact, '
                             'matplotlib parity ' + ('checked against tests/golden/figure_mpl.npz' if os.path.exists(os.path.join(ROOT, 'tests', 'golden', 'figure_mpl.npz'))
                                                     else 'UNVERIFIED (no matplotlib in the build image; tests/golden/make_figure_golden.py is the pin)')},
    }
    pin.close()
    r.close()
    return res


def sequence_leg(env, args, sampler, with_cpu):
    torch = env.torch
    from reasoning_image_generation_b200 import Renderer
    from reasoning_image_generation_b200.dedup import DedupFilter
    from reasoning_image_generation_b200.synth import synthetic_scene
    world, rank, dist, stream, st = env.world, env.rank, env.dist, env.stream, env.st
    barrier = env.barrier
    n = PANELS_PER_STEP
    scene = synthetic_scene(n, seed=1000 + rank)
    r = Renderer(512, 512, device=env.local, max_panels=n, max_elements=max(scene.n_elements, 1))
    out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device='cuda')
    r.upload(scene, 'rgb')
    render = lambda: r.render_resident(out.data_ptr(), stream=st)   # noqa: E731
    if sampler is not None:
        sampler.start()
    n_warm = timed_steps(env, render, r.sync, args.steps, args.warmup)
    r.set_profiling(True)
    launches0 = r.stats()['kernel_launches']
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        render()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler is not None else None
    r.sync()
    stats = r.stats()
    r.set_profiling(False)
    launches = stats['kernel_launches'] - launches0
    (ms_max,) = env.max_over_ranks(ms)
    value = world * n * args.steps / (ms_max / 1e3)

    # ---- e2e: public API, host scene in, pinned host pixels out
    sub = scene.slice(0, E2E_PANELS)
    pinned = r.pinned_output(E2E_PANELS)
    for _ in range(2):
        r.render(sub, 'rgb', out=pinned)
    barrier()
    t0 = time.perf_counter()
    px_steps = 4
    for _ in range(px_steps):
        r.render(sub, 'rgb', out=pinned)
    torch.cuda.synchronize()
    (dt,) = env.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * E2E_PANELS * px_steps / dt
    e2e_steps = args.e2e_steps
    h2d = 40 * sub.n_elements + 4 * (sub.n_panels + 1) + sub.n_panels
    d2h = E2E_PANELS * IMG_BYTES
    checksum = int(pinned[::257, 100, 100, 0].astype('int64').sum())
    # ---- the same with the PNG files cv2.imwrite would produce as the deliverable (encoded on the device)
    from reasoning_image_generation_b200.png import PngEncoder
    enc = PngEncoder(512, 512, device=env.local, max_images=E2E_PANELS)
    # (the whole step scene, E2E_PANELS at a time: a worker thread renders chunk k + 1 while chunk k is encoded and copied)
    for _ in range(2):
        for p0, buf, off in r.iter_png(scene, enc, E2E_PANELS, 'rgb'):
            pass
    barrier()
    t0 = time.perf_counter()
    seq_png_bytes = 0
    for _ in range(e2e_steps):
        for p0, buf, off in r.iter_png(scene, enc, E2E_PANELS, 'rgb'):
            seq_png_bytes += int(off[-1])
    torch.cuda.synchronize()
    (dt_png,) = env.max_over_ranks(time.perf_counter() - t0)
    e2e_png_value = world * n * e2e_steps / dt_png
    h2d_png = 40 * scene.n_elements + 4 * (scene.n_panels + 1) + scene.n_panels
    enc.close()

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
    dd_ms, hash_ms = env.max_over_ranks(dd_ms, hash_ms)
    pairs = n_total * (n_total - 1) // 2
    # roofline of the scan: integer pipes, against throughputs MEASURED on this device by the same library (k_int_pipe_peak).
    # Per pair the filter issues 1 POPC and 2.5 ALU-pipe instructions (2 LOP3 + half a 3-input minimum); see rig_dedup.cuh.
    popc_peak, alu_peak = r.int_pipe_peak('popc'), r.int_pipe_peak('lop3')
    pairs_per_gpu_s = pairs / world / (dd_ms / 1e3)
    t_popc, t_alu = 1.0 / popc_peak, 2.5 / alu_peak              # seconds per pair if that pipe alone were the limit
    bound_pipe = 'popc' if t_popc >= t_alu else 'alu'
    dd_roofline = {'bound': 'int_pipe', 'pipe': bound_pipe, 'achieved': pairs_per_gpu_s * (1.0 if bound_pipe == 'popc' else 2.5) / 1e12,
                   'peak': (popc_peak if bound_pipe == 'popc' else alu_peak) / 1e12, 'unit': 'T lane-ops/s per GPU',
                   'frac': pairs_per_gpu_s * max(t_popc, t_alu), 'popc_peak_ops_per_s': popc_peak, 'alu_peak_ops_per_s': alu_peak,
                   'popc_per_pair': 1.0, 'alu_per_pair': 2.5, 'kernel': 'k_hamming',
                   'peak_source': 'measured on this device: rig_int_pipe_peak (dependent POPC / LOP3 chains, 64 warps per SM)',
                   'note': 'pairs counted algorithmically (N(N-1)/2); the kernel also evaluates the below-diagonal part of the diagonal '
                           'tiles and the all_gather / all_reduce are inside the timed region'}

    # ---- overview compositor (layout.compose_grid): 7 consecutive panels of the rendered step per sample
    from reasoning_image_generation_b200.compose import GridCompositor, GridLayout
    GL = GridLayout(512, 512, n_states=3, num_options=4)
    n_samples = n // 7
    gidx = torch.arange(n_samples * 7, dtype=torch.int32).reshape(n_samples, 7).numpy()
    gc = GridCompositor(GL, device=env.local, max_samples=n_samples, channel_order='rgb')
    gout = torch.empty((n_samples, GL.grid_h, GL.grid_w, 3), dtype=torch.uint8, device='cuda')
    for _ in range(3):
        gc.compose_to_device(out.data_ptr(), gidx, gout.data_ptr(), stream=st)
    gc.sync()
    gc.set_profiling(True)
    gsum = 0.0
    for _ in range(3):
        gc.compose_to_device(out.data_ptr(), gidx, gout.data_ptr(), stream=st)
        gc.sync()
    gst = gc.stats()
    (grid_ms,) = env.max_over_ranks(gst['compose_ms'] / 3)
    gc.close()
    del gout

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
    if rank == 0 and world == 1 and with_cpu:
        done, cdt, cores, backend = cpu_sample(seed=1000, budget_s=args.seq_cpu_budget)
        cpu = {'value': done / cdt, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
               'sample': 'first %d panels of the step scene (seed 1000), oracle/ref_render.py (%s backend), '
                         'one process per core, %.1f s' % (done, backend, cdt)}
    res = {
        'metric': METRIC, 'value': value, 'unit': 'images/s', 'ms_per_step': ms_max / args.steps, 'warmup': n_warm,
        'config': {'workload': WORKLOAD, 'panels_per_step_per_gpu': n, 'l2': 'outputs (12.9 GB/step) exceed L2; no flush needed',
                   'elements_per_step_per_gpu': scene.n_elements, 'out_layout': 'RGB'},
        'e2e_pixels': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                       'panels_per_step_per_gpu': E2E_PANELS, 'note': 'Renderer.render: host SoA scene in, every pixel copied to '
                       'pinned host memory (PCIe-bound); checksum %d' % checksum},
        'e2e': {'value': e2e_png_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d_png, 'd2h_bytes_per_step': seq_png_bytes // e2e_steps,
                    'png_bytes_per_image': seq_png_bytes / (e2e_steps * n), 'panels_per_step_per_gpu': n, 'chunk': E2E_PANELS,
                    'timed_s': dt_png, 'steps': e2e_steps,
                    'note': 'Renderer.iter_png = save_image / cv2.imwrite for a batch: host SoA scene in, panels rendered and '
                            'PNG-encoded on the device %d at a time, finished files copied to pinned host memory; the render of '
                            'chunk k + 1 overlaps the encode and copy of chunk k' % E2E_PANELS},
        'gpu_launches': launches,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'kernel': 'k_raster', 'peak_source': peak_src,
                     'launches': stats['timed_launches'], 'panels_per_launch': per_launch,
                     'expand_ms_per_step': stats['expand_ms'] / args.steps, 'raster_ms_per_step': stats['raster_ms'] / args.steps,
                     'whole_step_frac': value / world * IMG_BYTES / 1e9 / peak},
        'cpu_baseline': cpu,
        'grid': {'samples_per_s': world * n_samples / (grid_ms / 1e3), 'panels_per_s': world * n_samples * 7 / (grid_ms / 1e3),
                 'ms': grid_ms, 'samples': n_samples, 'panels_per_sample': 7, 'cell_size': GL.cell_size,
                 'panel_read_gbs_per_gpu': n_samples * 7 * IMG_BYTES / (grid_ms / 1e3) / 1e9,
                 'note': 'k_grid: layout.compose_grid for a batch (INTER_AREA 512 -> 118 per cell + constant overlay), panels '
                         'already in HBM; bound: HBM read of the panels'},
        'dedup': {'pairs_per_s': pairs / (dd_ms / 1e3), 'roofline': dd_roofline, 'n_hashes': n_total, 'pairs': pairs, 'ms': dd_ms, 'threshold': 10,
                  'hash_bits': 256, 'hit_pairs': int(hits.item()), 'dropped': int(drop.sum().item()),
                  'input': 'uniform random 256-bit hashes, seed 0, 1 % planted near-duplicates (SURVEY 8d)',
                  'exchange': 'nccl all_gather + max all_reduce' if dist is not None else 'none (1 rank)',
                  'hash_images_per_s': world * n / (hash_ms / 1e3), 'hash_read_gbs_per_gpu': n * IMG_BYTES / (hash_ms / 1e3) / 1e9,
                  'render_plus_fused_hash_ms_per_step': fused_ms, 'fused_hash_equals_standalone': fused_equal,
                  'rendered_step_dropped': int(drop_r.sum().item())},
        'clocks': clocks,
    }
    r.close()
    del out, table
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    # stdout carries exactly one JSON line: anything libraries print there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    env = Env()
    rank, world = env.rank, env.world
    line = None
    if args.workload == 'figure':
        fig = figure_leg(env, args, ClockSampler(env.local) if rank == 0 else None)
        seq = None if args.no_sequence else sequence_leg(env, args, None, with_cpu=not args.no_cpu)
        if rank == 0:
            peak, peak_src = fallback_peak()
            stats, n = fig['stats'], fig['n']
            cover_s = stats['cover_ms'] / 1e3
            achieved = stats['timed_images'] * IMG_BYTES / cover_s / 1e9 if cover_s > 0 else 0.0
            per_launch = stats['timed_images'] / max(1, stats['timed_launches'])
            traffic = None
            tp = os.path.join(ROOT, 'profiles', 'figure_traffic.json')
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp))['dram_bytes_per_image'] * per_launch
                except Exception:
                    traffic = None
            cpu = None
            if world == 1 and not args.no_cpu:
                cpu = fig_cpu_dict(fig_cpu_sample(seed=1000, budget_s=14.0))
            line = {
                'metric': FIG_METRIC, 'value': fig['value'], 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': fig['warmup'], 'ms_per_step': fig['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic', 'config': fig['config'], 'e2e': fig['e2e'],
                'e2e_pixels': fig['e2e_pixels'], 'multi_shape': fig['multi_shape'], 'dpi200': fig['dpi200'], 'gpu_launches': fig['launches'],
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                             'traffic': traffic, 'kernel': 'k_figure', 'peak_source': peak_src, 'launches': stats['timed_launches'],
                             'images_per_launch': per_launch, 'figure_ms_per_step': stats['cover_ms'] / args.steps,
                             'geometry_ms_per_step': stats['geom_ms'] / max(1, stats['geom_launches']), 'geometry_kernel': 'k_figure_path',
                             'whole_step_frac': fig['value'] / world * IMG_BYTES / 1e9 / peak,
                             'note': 'achieved / frac: k_figure alone (the dominant kernel); whole_step_frac: k_figure_path + k_figure, '
                                     'i.e. the whole draw_path from artist arrays, = value x 786,432 B / peak'},
                'cpu_baseline': cpu, 'clocks': fig['clocks'],
            }
            if seq:
                line['dedup'] = seq.pop('dedup')
                seq.pop('clocks', None)
                line['sequence_panels'] = seq
    else:
        seq = sequence_leg(env, args, ClockSampler(env.local) if rank == 0 else None, with_cpu=not args.no_cpu)
        if rank == 0:
            line = {'metric': seq['metric'], 'value': seq['value'], 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
                    'warmup': seq['warmup'], 'ms_per_step': seq['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                    'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic'}
            line.update({k: v for k, v in seq.items() if k not in line})
    if rank == 0:
        os.write(json_fd, (json.dumps(line) + '\n').encode())
    if env.dist is not None:
        env.dist.barrier()
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline legs')
    ap.add_argument('--workload', default='figure', choices=['figure', 'sequence'],
                    help='figure: BASELINE configs[1] (headline) + the sequence-panel path as a sub-object; sequence: that path only')
    ap.add_argument('--no-sequence', action='store_true', help='figure workload: skip the sequence-panel leg')
    ap.add_argument('--e2e-steps', type=int, default=48, help='timed passes of the e2e legs over the step scene (a timed region of 1 s or more)')
    ap.add_argument('--e2e-chunk', type=int, default=FIG_E2E_CHUNK, help='images rendered + encoded per chunk of the figure e2e leg')
    ap.add_argument('--seq-cpu-budget', type=float, default=10.0, help='seconds of CPU work for the sequence-panel cpu_baseline')
    args = ap.parse_args()
    if args.impl == 'reference':
        if args.workload == 'figure':
            run_reference_figure(args)
        else:
            run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
