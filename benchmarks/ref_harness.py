#!/usr/bin/env python
"""The reference's two benchmark harnesses, shape for shape, on this engine (device-resident inputs):

  tests_and_benchmarks/test_benchmark_conv2d.py:21-79  conv-as-kmap, kernel (C,3,3), padding (valid,same,same),
        C in {4,8,16,32,64} x H in {16,...,1024}, 50 calls each
  tests_and_benchmarks/test_benchmark_patch.py:19-61   1-D patch extraction, k=3, 'same', dim in {10,...,10^6}, 1000 calls

Timed the way the reference times itself (wall clock around ONE call + a device synchronisation, mean / stddev in
microseconds), plus the back-to-back device time per call (CUDA events) -- the first is launch latency + host overhead
for all but the largest shapes.  The reference publishes no numbers for these harnesses; the comparator
it prints (lax.conv / conv_general_dilated_patches) needs JAX.  One JSON line per shape."""
import json
import os
import sys
import time
from statistics import mean, stdev

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex  # noqa: E402
from kernex_b200 import _lib  # noqa: E402

dev = torch.device("cuda")


def wall(fn, iters):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        t1 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t1)
    return mean(ts) * 1e6, stdev(ts) * 1e6


def device_time(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


only = sys.argv[1] if len(sys.argv) > 1 else ""
if only in ("", "conv"):
    g = torch.Generator(device=dev).manual_seed(0)
    for C in (4, 8, 16, 32, 64):
        for H in (16, 32, 64, 128, 256, 512, 1024):
            conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"), relative=False)(lambda x, w: np.sum(x * w))
            x = torch.randn((C, H, H), device=dev, generator=g)
            w = torch.randn((C, 3, 3), device=dev, generator=g)
            y = conv(x, w)
            ref = torch.nn.functional.conv2d(x[None].double(), w[None].double(), padding=1)[0].float()   # sanity only (test_benchmark_conv2d.py:50)
            ok = bool(torch.allclose(y, ref, atol=1e-3))
            m, s = wall(lambda: conv(x, w), 50)
            d = device_time(lambda: conv(x, w), 50)
            print(json.dumps({"harness": "conv2d", "dim": [C, H, H], "wall_us_mean": round(m, 1), "wall_us_stddev": round(s, 1),
                              "device_us": round(d, 2), "allclose_1e-3": ok, "kernel": _lib.last_kernel()}), flush=True)
if only in ("", "patch"):
    for dim in (10, 100, 1_000, 10_000, 100_000, 1_000_000):
        get_patches = kex.kmap(kernel_size=(3,), padding="same", relative=False)(lambda x: x)
        x = torch.arange(1, dim + 1, dtype=torch.int32, device=dev)
        p = get_patches(x)
        xp = torch.nn.functional.pad(x, (1, 1))
        ok = bool(torch.equal(p, torch.stack([xp[:-2], xp[1:-1], xp[2:]], dim=1)))
        m, s = wall(lambda: get_patches(x), 1000)
        d = device_time(lambda: get_patches(x), 1000)
        print(json.dumps({"harness": "patch", "dim": dim, "wall_us_mean": round(m, 1), "wall_us_stddev": round(s, 1),
                          "device_us": round(d, 2), "bit_exact": ok, "kernel": _lib.last_kernel()}), flush=True)
