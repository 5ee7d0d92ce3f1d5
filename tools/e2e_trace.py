"""Timeline of FigureRenderer.iter_png's two threads (same logic, with clocks): where each side waits."""
import queue
import sys
import threading
import time

sys.path.insert(0, '.')
from reasoning_image_generation_b200.figure import FigureRenderer  # noqa: E402
from reasoning_image_generation_b200.png import PngEncoder  # noqa: E402
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene  # noqa: E402

import os
DEV = int(os.environ.get('LOCAL_RANK', '0')) if os.environ.get('E2E_USE_LOCAL_RANK') else 0
if os.environ.get('E2E_DIST'):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(DEV)
    dist.init_process_group('nccl', device_id=torch.device('cuda', DEV))
    dist.barrier()
n, chunk = 32768, (int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
ch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
assert n % chunk == 0, 'chunk must divide n (a worker exception would leave the main loop waiting)'
sc = synthetic_figure_scene(n, 512, 512, seed=1, kind='single')
r = FigureRenderer(512, 512, device=DEV, max_images=chunk)
enc = PngEncoder(512, 512, device=DEV, max_images=chunk, channels=ch)
img = 512 * 512 * 3
base = enc.staging(2 * chunk)
starts = list(range(0, n, chunk))
for rep in range(3):
    free = threading.Semaphore(2)
    ready = queue.Queue()
    W = []
    M = []

    def worker():
        for k, i0 in enumerate(starts):
            t0 = time.perf_counter()
            free.acquire()
            t1 = time.perf_counter()
            sub = sc.slice(i0, i0 + chunk)
            r.render_to_device(sub, base + (k & 1) * chunk * img, 'rgb', 0)
            ev = r.fence()
            t2 = time.perf_counter()
            W.append((t1 - t0, t2 - t1))
            ready.put((i0, chunk, base + (k & 1) * chunk * img, ev))

    T0 = time.perf_counter()
    t = threading.Thread(target=worker)
    t.start()
    pending = None
    for _ in starts:
        t0 = time.perf_counter()
        i0, n_img, ptr, ev = ready.get()
        t1 = time.perf_counter()
        enc.wait_event(ev)
        slot = enc.encode_begin(ptr, n_img, 'rgb', 0)
        t2 = time.perf_counter()
        free.release()
        if pending is not None:
            enc.encode_end(pending)
        t3 = time.perf_counter()
        pending = slot
        M.append((t1 - t0, t2 - t1, t3 - t2))
    enc.encode_end(pending)
    t.join()
    dt = time.perf_counter() - T0
    print('rep %d: %.0f k img/s, %.2f ms per chunk' % (rep, n / dt / 1e3, dt / len(starts) * 1e3))
print('worker  (wait-free, render) ms:', ' '.join('(%.2f %.2f)' % (a * 1e3, b * 1e3) for a, b in W))
print('main (wait-ready, begin, end) ms:', ' '.join('(%.2f %.2f %.2f)' % (a * 1e3, b * 1e3, c * 1e3) for a, b, c in M))
