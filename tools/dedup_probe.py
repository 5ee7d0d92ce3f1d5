"""Scan-kernel probe: measured POPC / LOP3 peaks and k_hamming pairs/s at a few table sizes (one GPU)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reasoning_image_generation_b200 import Renderer
from reasoning_image_generation_b200.dedup import DedupFilter

r = Renderer(512, 512, device=0, max_panels=64, max_elements=512)
f = DedupFilter(r)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
out = {'popc_peak': r.int_pipe_peak('popc'), 'lop3_peak': r.int_pipe_peak('lop3')}
sm = torch.cuda.get_device_properties(0).multi_processor_count
out['popc_per_clk_per_sm_at_1965'] = out['popc_peak'] / sm / 1.965e9
out['lop3_per_clk_per_sm_at_1965'] = out['lop3_peak'] / sm / 1.965e9
for n in [int(a) for a in sys.argv[1:]] or [1 << 20, 1 << 22]:
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    table = torch.randint(-2 ** 62, 2 ** 62, (n, 4), dtype=torch.int64, device='cuda', generator=g)
    idx = torch.randint(0, n, (n // 100, 2), device='cuda', generator=g)
    table[idx[:, 1]] = table[idx[:, 0]] ^ (1 << torch.randint(0, 62, (n // 100, 1), device='cuda', generator=g))
    f.scan(table[:100000], 10); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    drop, hits = f.scan(table, 10)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    pairs = n * (n - 1) // 2
    out[f'n={n}'] = {'ms': ms, 'pairs_per_s': pairs / ms * 1e3, 'hits': int(hits.item()), 'dropped': int(drop.sum().item()),
                     'popc_frac': pairs / ms * 1e3 / out['popc_peak']}
print(json.dumps(out, indent=1))
