import os, sys, json
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/benchmarks")
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit, peak
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def rep(name, nbytes, med):
    print(json.dumps({"path": name, "ms": round(med, 3), "frac": round(nbytes / (med * 1e-3) / 1e9 / peak(), 3), "kernel": _lib.last_kernel()}), flush=True)
for C, K in [(3, 5), (8, 5), (3, 7)]:
    n = 4096
    x = torch.randn((C, n, n), device=dev, generator=g).requires_grad_()
    w = torch.randn((C, K, K), device=dev, generator=g).requires_grad_()
    f = kex.kmap(kernel_size=(C, K, K), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    y = f(x, w)
    with torch.no_grad():
        med, _ = timeit(lambda: f(x, w), 3)
    rep(f"conv-as-kmap ({C},{K},{K}) forward on ({C},{n},{n})", x.numel() * 4 + y.numel() * 4, med)
    gy = torch.randn(y.shape, device=dev, generator=g)
    med, _ = timeit(lambda: torch.autograd.grad(y, w, gy, retain_graph=True), 3)
    rep(f"  its weight gradient", x.numel() * 4 + y.numel() * 4, med)
    med, _ = timeit(lambda: torch.autograd.grad(y, x, gy, retain_graph=True), 3)
    rep(f"  its input gradient", x.numel() * 4 + y.numel() * 4, med)
