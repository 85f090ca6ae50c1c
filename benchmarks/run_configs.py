#!/usr/bin/env python
"""Measure every BASELINE.json config on ONE B200 (device-resident inputs, CUDA events).

Not the driver's bench (that is /bench.py, config 2 only): this produces the per-config
table quoted in DESIGN.md / profiles/.  One JSON object per line on stdout.

    python benchmarks/run_configs.py [--only C2,C5] [--reps 20]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import kernex_b200 as kex  # noqa: E402
from kernex_b200 import _lib  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, reps, warm=3, inner=None):
    """Median / min time per call; each sample brackets `inner` back-to-back calls with CUDA
    events (sustained rate: launch gaps of a single call are not charged to the kernel)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    if inner is None:
        inner = max(1, min(20, int(5.0 / one)))   # ~5 ms of work per sample
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / inner)
    times.sort()
    return times[len(times) // 2], times[0]


def report(name, lups, bytes_per_lup, med, best, note=""):
    pk = peak()
    gbs = lups * bytes_per_lup / (med * 1e-3) / 1e9
    print(json.dumps({"config": name, "glups": lups / (med * 1e-3) / 1e9, "ms_median": med, "ms_min": best,
                      "bytes_per_lup": bytes_per_lup, "achieved_gbs": gbs, "frac_of_measured_peak": gbs / pk,
                      "frac_of_8TBs": gbs / 8000.0, "kernel": _lib.last_kernel(), "note": note}), flush=True)


def lap2(x):
    return (0 * x[1, -1] + 1 * x[1, 0] + 0 * x[1, 1] + 1 * x[0, -1] + -4 * x[0, 0] + 1 * x[0, 1]
            + 0 * x[-1, -1] + 1 * x[-1, 0] + 0 * x[-1, 1])


def lap3(x):
    return (x[1, 0, 0] + x[-1, 0, 0] + x[0, 1, 0] + x[0, -1, 0] + x[0, 0, 1] + x[0, 0, -1] - 6 * x[0, 0, 0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--n3", type=int, default=2048)
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda c: not only or c in only
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)

    if want("C1"):
        x = torch.arange(1, 1_000_001, dtype=torch.int32, device=dev)
        f = kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))
        med, best = timeit(lambda: f(x), args.reps * 5)
        report("C1 1-D int32 sum k=3 valid N=1e6", 999_998, 8, med, best, "8 MB: launch-latency bound")

    if want("C2"):
        n = 16384
        x = torch.randn((n, n), device=dev, generator=g)
        f = kex.kmap(kernel_size=(3, 3), padding="valid", relative=True)(lap2)
        med, best = timeit(lambda: f(x), args.reps)
        report("C2 2-D 5-pt Laplacian fp32 16384^2 valid (fwd)", (n - 2) ** 2, 8, med, best)
        xg = x.requires_grad_()
        y = f(xg)
        gy = torch.randn(y.shape, device=dev, generator=g)
        med, best = timeit(lambda: torch.autograd.grad(y, xg, gy, retain_graph=True), max(args.reps // 2, 3))
        report("C2 VJP wrt x (read g, write dx)", n * n, 8, med, best)
        del x, xg, y, gy
        torch.cuda.empty_cache()

    if want("C3a"):
        n = 8192
        x = torch.randn((3, n, n), device=dev, generator=g)
        w = torch.randn((3, 3, 3), device=dev, generator=g)
        conv = kex.kmap(kernel_size=(3, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
        med, best = timeit(lambda: conv(x, w), args.reps)
        report("C3a conv-as-kmap (3,3,3) on (3,8192,8192), device weights", n * n, 16, med, best)
        wg = w.clone().requires_grad_()
        y = conv(x, wg)
        gy = torch.randn(y.shape, device=dev, generator=g)
        med, best = timeit(lambda: torch.autograd.grad(y, wg, gy, retain_graph=True), max(args.reps // 2, 3))
        report("C3a weight gradient (read x + g)", n * n, 16, med, best)
        del x, y, gy
        torch.cuda.empty_cache()

    if want("C3b"):
        n = 8192
        x = torch.randn((n, n), device=dev, generator=g)
        f = kex.kmap(kernel_size=(3, 3), padding="same")(lambda v: v)
        med, best = timeit(lambda: f(x), args.reps)
        report("C3b 3x3 patches same on 8192^2 fp32", n * n, 40, med, best)
        del x
        torch.cuda.empty_cache()

    if want("C4"):
        nt, nx = 4096, 1 << 21
        c, tmax, xmax = 0.5, 0.5, 2.0
        kappa = 0.5  # CFL kept at 0.5 (the README's dt/dx would give CFL 512 at this size)
        F = kex.kscan(kernel_size=(3, 3), padding=((1, 1), (1, 1)), named_axis={0: "n", 1: "i"}, relative=True)
        F[:, 0] = F[:, -1] = lambda u: 1
        F[:, : int((nx - 1) / 4) + 1] = F[:, int((nx - 1) / 2):] = lambda u: 1
        F[0:1, int((nx - 1) / 4) + 1: int((nx - 1) / 2)] = lambda u: 2
        F[1:, 1:-1] = lambda u: u["i", "n-1"] - kappa * (u["i", "n-1"] - u["i-1", "n-1"])
        u = torch.ones((nt, nx), device=dev)
        med, best = timeit(lambda: F(u), max(args.reps // 4, 3), warm=1)
        report("C4 kscan linear convection nt=4096 nx=2^21 (one GPU's share)", nt * nx, 4, med, best)
        del u
        torch.cuda.empty_cache()

    if want("C5"):
        n = args.n3
        x = torch.randn((n, n, n), device=dev, generator=g)
        f = kex.kmap(kernel_size=(3, 3, 3), padding="valid", relative=True)(lap3)
        med, best = timeit(lambda: f(x), max(args.reps // 2, 3), warm=2)
        report(f"C5 3-D 7-pt Laplacian fp32 {n}^3 valid (single GPU)", (n - 2) ** 3, 8, med, best)
        del x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
