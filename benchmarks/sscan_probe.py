#!/usr/bin/env python
"""Time-marching 2-D diffusion written as a 3-D sscan (the reference's tests/test_interface.py:28-97): plane n from plane n-1."""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
from kernex_b200 import _lib

dev = torch.device("cuda")
for nt, ny, nx in [(16, 512, 512), (16, 2048, 2048), (8, 8192, 8192)]:
    @kex.sscan(kernel_size=(3, 3, 3), offset=((1, 0), (1, 1), (1, 1)), named_axis={0: "n", 1: "i", 2: "j"}, relative=True)
    def diffusion(T, a, b):
        return (T["n-1", "i", "j"] + a * (T["n-1", "i+1", "j"] - 2 * T["n-1", "i", "j"] + T["n-1", "i-1", "j"])
                + b * (T["n-1", "i", "j+1"] - 2 * T["n-1", "i", "j"] + T["n-1", "i", "j-1"]))
    U = torch.ones((nt, ny, nx), device=dev)
    U[0, ny // 4: ny // 2, nx // 4: nx // 2] = 2.0
    out = diffusion(U, 0.1, 0.1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        out = diffusion(U, 0.1, 0.1)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    # reference by plain torch slicing (plane by plane)
    ref = U.clone()
    for n in range(1, nt):
        p = ref[n - 1]
        ref[n, 1:-1, 1:-1] = (p[1:-1, 1:-1] + 0.1 * (p[2:, 1:-1] - 2 * p[1:-1, 1:-1] + p[:-2, 1:-1])
                              + 0.1 * (p[1:-1, 2:] - 2 * p[1:-1, 1:-1] + p[1:-1, :-2]))
    err = float((out - ref).abs().max())
    print(json.dumps({"sscan diffusion": [nt, ny, nx], "ms": round(ms, 3), "us_per_plane": round(ms * 1e3 / (nt - 1), 1),
                      "glups": round((nt - 1) * (ny - 2) * (nx - 2) / ms / 1e6, 1), "max_err_vs_torch": err, "kernel": _lib.last_kernel()}), flush=True)

# kscan mesh: 2-D convection with boundary / initial conditions (the README's mesh with two space axes)
for nt, ny, nx in [(16, 2048, 2048), (8, 8192, 8192)]:
    for env in ("", "1"):
        if env:
            os.environ["KX_NO_PLANEMARCH"] = "1"
        else:
            os.environ.pop("KX_NO_PLANEMARCH", None)
        F = kex.kscan(kernel_size=(3, 3, 3), padding=((1, 1), (1, 1), (1, 1)), relative=True)
        F[:, 0, :] = F[:, -1, :] = lambda u: 1.0
        F[:, :, 0] = F[:, :, -1] = lambda u: 1.0
        F[0, 1:-1, 1:-1] = lambda u: 1.0
        F[0, ny // 4: ny // 2, nx // 4: nx // 2] = lambda u: 2.0
        F[1:, 1:-1, 1:-1] = lambda u: u[-1, 0, 0] - 0.3 * (u[-1, 0, 0] - u[-1, -1, 0]) - 0.2 * (u[-1, 0, 0] - u[-1, 0, -1])
        U = torch.ones((nt, ny, nx), device=dev)
        out = F(U)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = F(U)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        print(json.dumps({"kscan mesh convection 2-D": [nt, ny, nx], "planemarch": not env, "ms": round(ms, 3),
                          "glups": round(nt * ny * nx / ms / 1e6, 1), "kernel": _lib.last_kernel()}), flush=True)
os.environ.pop("KX_NO_PLANEMARCH", None)
