#!/usr/bin/env python
"""Second batch of paths next to the tuned configs: weight gradients beyond 3x3, 3-D pooling gradients, decimating FIR."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit, peak
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def rep(name, nbytes, med):
    print(json.dumps({"path": name, "ms": round(med, 3), "frac": round(nbytes / (med * 1e-3) / 1e9 / peak(), 3), "kernel": _lib.last_kernel()}), flush=True)
def guard(name, fn):
    try:
        fn()
    except Exception as e:
        print(json.dumps({"path": name, "error": f"{type(e).__name__}: {str(e)[:160]}"}), flush=True)

def dw5():
    x = torch.randn((4096, 4096), device=dev, generator=g)
    w = torch.randn((5, 5), device=dev, generator=g).requires_grad_()
    f = kex.kmap(kernel_size=(5, 5), padding="same")(lambda v, w: np.sum(v * w))
    y = f(x, w); gy = torch.randn(y.shape, device=dev, generator=g)
    med, _ = timeit(lambda: torch.autograd.grad(y, w, gy, retain_graph=True), 4)
    rep("weight gradient of a 5x5 conv-as-kmap on 4096^2", x.numel() * 8, med)
    with torch.no_grad():
        med, _ = timeit(lambda: f(x, w), 4)
    rep("5x5 conv-as-kmap forward, device weights, 4096^2", x.numel() * 8, med)
guard("dw5", dw5)

def dw3d():
    x = torch.randn((256, 256, 256), device=dev, generator=g)
    w = torch.randn((3, 3, 3), device=dev, generator=g).requires_grad_()
    f = kex.kmap(kernel_size=(3, 3, 3), padding="same")(lambda v, w: np.sum(v * w))
    y = f(x, w); gy = torch.randn(y.shape, device=dev, generator=g)
    med, _ = timeit(lambda: torch.autograd.grad(y, w, gy, retain_graph=True), 4)
    rep("weight gradient of a 3x3x3 conv-as-kmap on 256^3", x.numel() * 8, med)
    xg = x.clone().requires_grad_()
    y = f(xg, w.detach())
    med, _ = timeit(lambda: torch.autograd.grad(y, xg, gy, retain_graph=True), 4)
    rep("input gradient of a 3x3x3 conv-as-kmap on 256^3", x.numel() * 8, med)
guard("dw3d", dw3d)

def pool3d_bwd():
    x = torch.randn((256, 256, 256), device=dev, generator=g).requires_grad_()
    f = kex.kmap(kernel_size=(2, 2, 2), strides=(2, 2, 2))(lambda v: np.mean(v))
    y = f(x); gy = torch.randn(y.shape, device=dev, generator=g)
    med, _ = timeit(lambda: torch.autograd.grad(y, x, gy, retain_graph=True), 4)
    rep("avgpool 2x2x2 s2 backward on 256^3", x.numel() * 4 + y.numel() * 4, med)
guard("pool3d_bwd", pool3d_bwd)

def fir_dec():
    x = torch.randn((1 << 26,), device=dev, generator=g)
    t = np.hanning(5).astype(np.float32)
    f = kex.kmap(kernel_size=(5,), strides=(2,), padding="same")(lambda v: np.sum(v * t))
    y = f(x); med, _ = timeit(lambda: f(x), 4)
    rep("decimating 5-tap FIR (stride 2) on 2^26", x.numel() * 4 + y.numel() * 4, med)
guard("fir_dec", fir_dec)

def lap9():
    x = torch.randn((8192, 8192), device=dev, generator=g).requires_grad_()
    f = kex.kmap(kernel_size=(3, 3), padding="same")(lambda v: np.sum(v) - 9 * v[1, 1])
    y = f(x); gy = torch.randn(y.shape, device=dev, generator=g)
    med, _ = timeit(lambda: torch.autograd.grad(y, x, gy, retain_graph=True), 4)
    rep("input gradient of a 9-point stencil on 8192^2", x.numel() * 8, med)
guard("lap9", lap9)
