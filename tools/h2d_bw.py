import torch, time
for mb in (9.6, 14.4, 38.5, 154):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): d.copy_(h, non_blocking=True)
        s.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): d.copy_(h, non_blocking=True)
        e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1) / 10
    t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{mb} MB: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s; one-shot wall {1e3*(t1-t0):.3f} ms")
