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
