#!/usr/bin/env python
"""Where the multi-function mesh kernel loses time against the single-function kernel (8192^2, fp32)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit, peak

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
n = 8192
x = torch.randn((n, n), device=dev, generator=g)
lap = lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0]
ident = lambda v: v[0, 0]
src = lambda v: 0.5 * v[0, 0] + 1.0


def mesh(kind):
    F = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)
    if kind == "aligned":          # region edges on band edges (th = 12): no tile is cut
        F[:, :] = lap
        F[0:4092, :] = src
    elif kind == "rows":           # boundary rows only
        F[:, :] = ident
        F[1:-1, :] = lap
    elif kind == "cols":           # boundary columns only
        F[:, :] = ident
        F[:, 1:-1] = lap
    elif kind == "frame":
        F[:, :] = ident
        F[1:-1, 1:-1] = lap
    else:                          # frame + source block (misc_paths.py)
        F[:, :] = ident
        F[1:-1, 1:-1] = lap
        F[n // 4: n // 2, n // 4: n // 2] = src
    return F


for kind in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["aligned", "rows", "cols", "frame", "full"]):
    F = mesh(kind)
    y = F(x)
    med, mn = timeit(lambda: F(x), 8)
    print(json.dumps({"mesh": kind, "us": round(med * 1e3, 2), "min_us": round(mn * 1e3, 2),
                      "frac": round(x.numel() * 8 / (med * 1e-3) / 1e9 / peak(), 3), "kernel": _lib.last_kernel()}), flush=True)
f1 = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lap)
y = f1(x)
med, mn = timeit(lambda: f1(x), 8)
print(json.dumps({"mesh": "single function", "us": round(med * 1e3, 2), "min_us": round(mn * 1e3, 2),
                  "frac": round(x.numel() * 8 / (med * 1e-3) / 1e9 / peak(), 3), "kernel": _lib.last_kernel()}), flush=True)

# 3-D: interior 7-point update + six Dirichlet faces (+ a source block) -- outside the 2-D mesh kernel: the dispatcher
# runs the interior function on the star kernel and recomputes the faces / the block with the generic kernel
if len(sys.argv) > 2 and sys.argv[2] == "3d":
    n = 768
    x3 = torch.randn((n, n, n), device=dev, generator=g)
    step = lambda v: v[0, 0, 0] + 0.1 * (v[1, 0, 0] + v[-1, 0, 0] + v[0, 1, 0] + v[0, -1, 0] + v[0, 0, 1] + v[0, 0, -1] - 6 * v[0, 0, 0])
    for kind in ("faces", "faces+block"):
        F = kex.kmap(kernel_size=(3, 3, 3), padding="same", relative=True)
        F[:, :, :] = lambda v: 1.0
        F[1:-1, 1:-1, 1:-1] = step
        if kind == "faces+block":
            F[100:200, 100:200, 100:200] = lambda v: v[0, 0, 0] + 2.0
        for env in ("", "1"):
            if env:
                os.environ["KX_NO_DECOMPOSE"] = "1"
            else:
                os.environ.pop("KX_NO_DECOMPOSE", None)
            y = F(x3)
            med, mn = timeit(lambda: F(x3), 4)
            print(json.dumps({"mesh": f"3-D heat equation {n}^3, {kind}", "decomposed": not env, "ms": round(med, 3),
                              "glups": round(y.numel() / med / 1e6, 1), "frac": round(x3.numel() * 8 / (med * 1e-3) / 1e9 / peak(), 3)}), flush=True)
    os.environ.pop("KX_NO_DECOMPOSE", None)
    f1 = kex.kmap(kernel_size=(3, 3, 3), padding="same", relative=True)(step)
    y = f1(x3)
    med, mn = timeit(lambda: f1(x3), 4)
    print(json.dumps({"mesh": f"single function {n}^3", "ms": round(med, 3), "glups": round(y.numel() / med / 1e6, 1),
                      "frac": round(x3.numel() * 8 / (med * 1e-3) / 1e9 / peak(), 3)}), flush=True)
