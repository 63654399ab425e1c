#!/usr/bin/env python
"""Write-only and copy bandwidth of the device memory (what the frame-recording path can reach at best):
python benchmarks/hbm_write_peak.py"""
import json

import torch

dev = torch.device("cuda", 0)
n = 5 * 2**30 // 8
x = torch.empty(n, dtype=torch.float64, device=dev)
y = torch.empty(n, dtype=torch.float64, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


out = {"bytes": n * 8}
ms = timed(lambda: x.fill_(1.0))
out["fill_kernel_gbs"] = n * 8 / (ms * 1e-3) / 1e9
ms = timed(lambda: torch.cuda.current_stream().synchronize() or x.zero_())
out["zero_gbs"] = n * 8 / (ms * 1e-3) / 1e9
ms = timed(lambda: y.copy_(x))
out["copy_read_plus_write_gbs"] = 2 * n * 8 / (ms * 1e-3) / 1e9
print(json.dumps(out))
