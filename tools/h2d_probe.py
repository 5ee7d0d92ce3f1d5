import torch, time
torch.cuda.init()
for mb in (1, 4, 16, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    for name, src, dst in (('h2d', h, d), ('d2h', d, h)):
        with torch.cuda.stream(st):
            for _ in range(3): dst.copy_(src, non_blocking=True)
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(10): dst.copy_(src, non_blocking=True)
            e1.record(st); st.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'{name} {mb} MB: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s')
