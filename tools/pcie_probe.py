"""Pinned host <-> device copy bandwidth (the e2e figures are bound by the D2H of the result tensor)."""
import torch
dev = torch.device('cuda:0')
for mb in (3, 16, 64):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device=dev)
    for name, fn in (('D2H', lambda: h.copy_(d, non_blocking=True)), ('H2D', lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print('%s %d MB: %.1f GB/s' % (name, mb, 20 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
