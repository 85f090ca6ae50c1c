#!/usr/bin/env python
"""Device memory bandwidth probes (read-only, write-only, copy) with torch ops: context for
the roofline fractions of write-only (kscan) and read-mostly kernels."""
import json
import torch

def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

n = 1 << 31  # 8 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
b = torch.empty_like(a)
res = {}
ms = t(lambda: b.fill_(1.0)); res["write_only_fill_GBs"] = n * 4 / ms / 1e6
ms = t(lambda: b.copy_(a)); res["copy_read_plus_write_GBs"] = 2 * n * 4 / ms / 1e6
ms = t(lambda: a.sum()); res["read_only_sum_GBs"] = n * 4 / ms / 1e6
ms = t(lambda: torch.add(a, 1.0, out=b)); res["add_scalar_read_plus_write_GBs"] = 2 * n * 4 / ms / 1e6
print(json.dumps(res))
