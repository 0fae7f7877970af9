"""Developer tool (GPU box): host-to-device bandwidth from pinned memory at a few sizes (one stream / two streams)."""
import torch, time
dev = torch.device("cuda:0")
for mb in (1, 4, 16, 64, 256):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # two halves on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
        with torch.cuda.stream(s2): d[n // 2:].copy_(h[n // 2:], non_blocking=True)
    torch.cuda.synchronize()
    ms2 = (time.perf_counter() - t0) * 1e3 / 10
    print(f"{mb:4d} MB: H2D {ms:.3f} ms = {mb / 1024 / (ms * 1e-3):.1f} GB/s ; split on two streams {ms2:.3f} ms = {mb / 1024 / (ms2 * 1e-3):.1f} GB/s")
