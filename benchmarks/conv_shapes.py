#!/usr/bin/env python
"""Probe: does conv-as-kmap throughput depend on the plane stride (2^28-byte planes at 8192^2)?"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
for (h, w) in [(8192, 8192), (8200, 8192), (8192, 8256), (8064, 8192), (4096, 16384), (16384, 4096)]:
    x = torch.randn((3, h, w), device=dev, generator=g)
    wt = torch.randn((3, 3, 3), device=dev, generator=g)
    conv = kex.kmap(kernel_size=(3, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    med, best = timeit(lambda: conv(x, wt), 10)
    gl = h * w / (med * 1e-3) / 1e9
    wg = wt.clone().requires_grad_()
    y = conv(x, wg)
    gy = torch.randn(y.shape, device=dev, generator=g)
    med2, _ = timeit(lambda: torch.autograd.grad(y, wg, gy, retain_graph=True), 6)
    print(json.dumps({"shape": [3, h, w], "fwd_glups": gl, "fwd_gbs": gl * 16, "wgrad_glups": h * w / (med2 * 1e-3) / 1e9,
                      "kernel": _lib.last_kernel()}), flush=True)
    del x, y, gy
    torch.cuda.empty_cache()
