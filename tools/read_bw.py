#!/usr/bin/env python
"""Developer tool: read-only HBM bandwidth of plain torch reductions, to put the streaming kernel's 6.4 TB/s in context."""
import torch

dev = torch.device("cuda", 0)
for mb in (619, 1238):
    n = mb * 1000 * 1000 // 4
    x = torch.randn(n, device=dev)
    y = torch.empty_like(x)
    for name, fn, bytes_ in (("sum (read)", lambda: x.sum(), 4 * n), ("amax (read)", lambda: x.amax(), 4 * n),
                             ("copy (read+write)", lambda: y.copy_(x), 8 * n)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(f"{mb} MB {name}: {ms * 1e3:.1f} us  {bytes_ / ms / 1e6:.0f} GB/s")
