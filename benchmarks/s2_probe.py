#!/usr/bin/env python
"""One stencil2d-family case per invocation (for ncu / A-B timing):  python s2_probe.py <case> [reps]
cases: ma7 (1x7 moving average), box5 (5x5 mean), max5, w4 (weighted 4x4), smap, lap (5-pt, 'same'), lap8k, d7 (dense 7x7),
fir15 (1-D, 2^27), avg3d (3-D average pooling 3^3 s2), convc16 ((16,4096^2) conv-as-kmap)"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit, peak

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n = 16384
bytes_ = None
if case in ("ma7", "box5", "max5", "w4", "lap", "d7", "smap", "lap8k"):
    if case == "lap8k":
        n = 8192
    x = torch.randn((n, n), device=dev, generator=g)
    rw = np.random.default_rng(3).standard_normal((7, 7)).astype(np.float32)
    lap = lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0]
    if case == "smap":
        f = kex.smap(kernel_size=(3, 3), offset=1, relative=True)(lap)
    else:
        ks, fn, rel = {"ma7": ((1, 7), lambda v: np.sum(v) / 7.0, False), "box5": ((5, 5), lambda v: np.mean(v), False),
                       "max5": ((5, 5), lambda v: np.max(v), False), "w4": ((4, 4), lambda v: np.sum(v * rw[:4, :4]), False),
                       "lap": ((3, 3), lap, True), "lap8k": ((3, 3), lap, True), "d7": ((7, 7), lambda v: np.sum(v * rw), False)}[case]
        f = kex.kmap(kernel_size=ks, padding="same", relative=rel)(fn)
elif case in ("patch5", "patch7", "patch5_old"):
    import os as _os
    k = 7 if case == "patch7" else 5
    n = 4096
    x = torch.randn((n, n), device=dev, generator=g)
    f = kex.kmap(kernel_size=(k, k), padding="same")(lambda v: v)
elif case == "wgrad5":
    x = torch.randn((8192, 8192), device=dev, generator=g)
    w = torch.randn((5, 5), device=dev, generator=g).requires_grad_()
    f0 = kex.kmap(kernel_size=(5, 5), padding="same")(lambda v, w: np.sum(v * w))
    yy = f0(x, w)
    gy = torch.randn(yy.shape, device=dev, generator=g)
    f = lambda t: torch.autograd.grad(yy, w, gy, retain_graph=True)[0]
elif case == "poolbwd":
    x = torch.randn((8192, 8192), device=dev, generator=g).requires_grad_()
    f0 = kex.kmap(kernel_size=(3, 3), strides=(2, 2))(lambda v: np.mean(v))
    yy = f0(x)
    gy = torch.randn(yy.shape, device=dev, generator=g)
    f = lambda t: torch.autograd.grad(yy, x, gy, retain_graph=True)[0]
elif case == "conv355":
    x = torch.randn((3, 4096, 4096), device=dev, generator=g)
    w = torch.randn((3, 5, 5), device=dev, generator=g)
    f0 = kex.kmap(kernel_size=(3, 5, 5), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    f = lambda t: f0(t, w)
elif case == "fir15":
    x = torch.randn((1 << 27,), device=dev, generator=g)
    taps = np.hanning(15).astype(np.float32)
    f = kex.kmap(kernel_size=(15,), padding="same")(lambda v: np.sum(v * taps))
elif case == "avg3d":
    x = torch.randn((1024, 1024, 1024), device=dev, generator=g)
    f = kex.kmap(kernel_size=(3, 3, 3), strides=(2, 2, 2), padding=1)(lambda v: np.mean(v))
elif case == "convc16":
    x = torch.randn((16, 4096, 4096), device=dev, generator=g)
    w = torch.randn((16, 3, 3), device=dev, generator=g)
    f0 = kex.kmap(kernel_size=(16, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    f = lambda t: f0(t, w)
else:
    raise SystemExit(f"unknown case {case}")
y = f(x)
med, mn = timeit(lambda: f(x), reps)
nb = x.numel() * 4 + y.numel() * 4
print(json.dumps({"case": case, "us": round(med * 1e3, 2), "min_us": round(mn * 1e3, 2), "glups": round(y.numel() / med / 1e6, 1),
                  "frac": round(nb / (med * 1e-3) / 1e9 / peak(), 3), "kernel": _lib.last_kernel()}), flush=True)
