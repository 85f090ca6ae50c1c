#!/usr/bin/env python
"""Throughput of the paths next to the BASELINE configs: pooling (strided mean / max),
batched depthwise conv through vmap, smap write-back, the wavefront kscan."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit, peak

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)


def rep(name, nbytes, med, lups):
    gbs = nbytes / (med * 1e-3) / 1e9
    print(json.dumps({"path": name, "ms": med, "glups": lups / (med * 1e-3) / 1e9, "alg_gbs": gbs,
                      "frac_of_measured_peak": gbs / peak(), "kernel": _lib.last_kernel()}), flush=True)


only = set(filter(None, (sys.argv[1] if len(sys.argv) > 1 else "").split(",")))
want = lambda c: not only or c in only

if want("pool"):
    x = torch.randn((8, 8192, 8192), device=dev, generator=g)
    for name, fn, k, s in [("avgpool 3x3 s2", lambda v: np.mean(v), 3, 2), ("maxpool 2x2 s2", lambda v: np.max(v), 2, 2),
                           ("maxpool 3x3 s2", lambda v: np.max(v), 3, 2), ("maxpool 3x3 s1 same", lambda v: np.max(v), 3, 1)]:
        f = kex.vmap(kex.kmap(kernel_size=(k, k), strides=(s, s), padding="same" if s == 1 else 0)(fn))
        y = f(x)
        med, _ = timeit(lambda: f(x), 6)
        rep(f"{name} on (8,8192,8192)", x.numel() * 4 + y.numel() * 4, med, y.numel())
    del x, y
    torch.cuda.empty_cache()

if want("wide"):
    # window shapes / reductions the stride-1 row-march kernel learnt in r1c (before: pool2d or the generic kernel)
    x = torch.randn((16384, 16384), device=dev, generator=g)
    gw = np.exp(-0.5 * (np.arange(-3, 4)[:, None] ** 2 + np.arange(-3, 4)[None, :] ** 2) / 2.0).astype(np.float32)
    gw /= gw.sum()
    def gauss(k):
        gg = np.exp(-0.5 * np.linspace(-(k - 1) / 2, (k - 1) / 2, k) ** 2 / 2.0)
        ww = np.outer(gg, gg)
        return (ww / ww.sum()).astype(np.float32)
    g11, g15 = gauss(11), gauss(15)
    rw = np.random.default_rng(3).standard_normal((7, 7)).astype(np.float32)
    cases = [("box blur 5x5 (np.mean) same", (5, 5), lambda v: np.mean(v)),
             ("Gaussian blur 7x7 (README.md:277-294, outer-product weights) same", (7, 7), lambda v: np.sum(v * gw)),
             ("Gaussian blur 11x11 same", (11, 11), lambda v: np.sum(v * g11)),
             ("Gaussian blur 15x15 same", (15, 15), lambda v: np.sum(v * g15)),
             ("dense 7x7 window (random weights) same", (7, 7), lambda v: np.sum(v * rw)),
             ("weighted 3x3 window same", (3, 3), lambda v: np.sum(v * gw[:3, :3])),
             ("max filter 3x3 same", (3, 3), lambda v: np.max(v)),
             ("max filter 5x5 same", (5, 5), lambda v: np.max(v)),
             ("weighted 4x4 window same", (4, 4), lambda v: np.sum(v * gw[:4, :4])),
             ("moving average k=7 along rows (1,7) same", (1, 7), lambda v: np.sum(v) / 7.0)]
    for name, ks, fn in cases:
        f = kex.kmap(kernel_size=ks, padding="same")(fn)
        y = f(x)
        med, _ = timeit(lambda: f(x), 4)
        rep(f"{name} on 16384^2", x.numel() * 8, med, y.numel())
    del x, y
    torch.cuda.empty_cache()

if want("fir"):
    # 1-D arrays: FIR filters / moving averages of up to 15 taps are one-row windows of the row-march kernel
    x = torch.randn((1 << 27,), device=dev, generator=g)
    w15 = np.random.default_rng(5).standard_normal(15).astype(np.float32)
    for name, k, fn in [("15-tap FIR filter", 15, lambda v: np.sum(v * w15)), ("moving average k=9", 9, lambda v: np.mean(v)),
                        ("running max k=11", 11, lambda v: np.max(v))]:
        f = kex.kmap(kernel_size=(k,), padding="same")(fn)
        y = f(x)
        med, _ = timeit(lambda: f(x), 4)
        rep(f"{name} on a 1-D array of 2^27 fp32" + (" (one-row launch, KX_NO_FOLD_1D)" if os.environ.get("KX_NO_FOLD_1D") else ""),
            x.numel() * 8, med, y.numel())
    del x, y
    torch.cuda.empty_cache()

if want("depthwise"):
    x = torch.randn((16, 4096, 4096), device=dev, generator=g)
    w = torch.randn((16, 3, 3), device=dev, generator=g)
    f = kex.vmap(kex.kmap(kernel_size=(3, 3), padding="same")(lambda v, w: np.sum(v * w)))
    y = f(x, w)
    med, _ = timeit(lambda: f(x, w), 6)
    rep("depthwise conv 3x3 same on (16,4096,4096), per-channel device weights", x.numel() * 8, med, y.numel())
    del x, y
    torch.cuda.empty_cache()

if want("smap"):
    x = torch.randn((16384, 16384), device=dev, generator=g)
    f = kex.smap(kernel_size=(3, 3), offset=1, relative=True)(lambda v: v[0, 1] + v[0, -1] + v[1, 0] + v[-1, 0] - 4 * v[0, 0])
    y = f(x)
    med, _ = timeit(lambda: f(x), 6)
    rep("smap 5-pt Laplacian 16384^2 (offset write-back)", x.numel() * 8, med, y.numel())
    del x, y
    torch.cuda.empty_cache()

if want("wave"):
    for n in (1024, 4096):
        x = torch.randn((n, n), device=dev, generator=g)
        f = kex.kscan(kernel_size=(3, 3), padding="same", relative=True)(
            lambda u: 0.25 * (u[0, -1] + u[-1, 0] + u[0, 1] + u[1, 0]))
        y = f(x)
        med, _ = timeit(lambda: f(x), 3, warm=1, inner=1)
        rep(f"Gauss-Seidel sweep kscan {n}^2 (wavefront)", x.numel() * 8, med, y.numel())
    del x, y
    torch.cuda.empty_cache()

if want("vjp"):
    n = 16384
    x = torch.randn((n, n), device=dev, generator=g).requires_grad_()
    def lap2(v):
        return v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0]
    for pad in ("valid", "same"):
        f = kex.kmap(kernel_size=(3, 3), padding=pad, relative=True)(lap2)
        y = f(x)
        gy = torch.randn(y.shape, device=dev, generator=g)
        med, _ = timeit(lambda: torch.autograd.grad(y, x, gy, retain_graph=True), 6)
        rep(f"C2 VJP wrt x, padding={pad}", n * n * 8, med, n * n)

if want("mesh"):
    # multi-function kmap mesh (F[slice] = fn): interior Laplacian + Dirichlet edges + a source block
    n = 8192
    x = torch.randn((n, n), device=dev, generator=g)
    F = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)
    F[:, :] = lambda v: v[0, 0]
    F[1:-1, 1:-1] = lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0]
    F[n // 4: n // 2, n // 4: n // 2] = lambda v: 0.5 * v[0, 0] + 1.0
    y = F(x)
    med, _ = timeit(lambda: F(x), 6)
    rep("kmap mesh, 3 region functions, 8192^2", x.numel() * 8, med, y.numel())
    f1 = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0])
    y = f1(x)
    med, _ = timeit(lambda: f1(x), 6)
    rep("single-function Laplacian 'same' on the same 8192^2 grid (calibration)", x.numel() * 8, med, y.numel())

if want("dense3d"):
    n = 1024
    x = torch.randn((n, n, n), device=dev, generator=g)
    w = np.random.default_rng(1).standard_normal((3, 3, 3)).astype(np.float32)
    f = kex.kmap(kernel_size=(3, 3, 3), padding="same")(lambda v: np.sum(v * w))
    y = f(x)
    med, _ = timeit(lambda: f(x), 4)
    rep("dense 27-point 3-D stencil (host weights) 'same' on 1024^3", x.numel() * 8, med, y.numel())
    os.environ["KX_NO_DENSE3D_PLANES"] = "1"      # the conv row ring with the three z planes as channels (device-weight path)
    y2 = f(x)
    med, _ = timeit(lambda: f(x), 4)
    rep("dense 27-point 3-D stencil on the conv row ring (KX_NO_DENSE3D_PLANES), 1024^3", x.numel() * 8, med, y2.numel())
    print(json.dumps({"max_abs_diff_vs_default": float((y2 - y).abs().max())}), flush=True)
    del os.environ["KX_NO_DENSE3D_PLANES"], y2
    xb = torch.randn((256, 2048, 2048), device=dev, generator=g)
    yb = f(xb)
    med, _ = timeit(lambda: f(xb), 4)
    rep("dense 27-point 3-D stencil (host weights) 'same' on 256x2048^2", xb.numel() * 8, med, yb.numel())
    del xb, yb
    f2 = kex.kmap(kernel_size=(3, 3, 3), strides=(2, 2, 2))(lambda v: np.max(v))
    y = f2(x)
    med, _ = timeit(lambda: f2(x), 3)
    rep("3-D maxpool 3x3x3 s2 on 1024^3", x.numel() * 4 + y.numel() * 4, med, y.numel())
    del x, y
    torch.cuda.empty_cache()

if want("pool3d"):
    # 3-D pooling (kx_pool3d): cubic windows, stride 2 / 3, whole-depth z march per warp
    x = torch.randn((1024, 1024, 1024), device=dev, generator=g)
    for name, fn, k, s in [("3-D maxpool 3x3x3 s2", lambda v: np.max(v), 3, 2), ("3-D avgpool 3x3x3 s2", lambda v: np.mean(v), 3, 2),
                           ("3-D maxpool 2x2x2 s2", lambda v: np.max(v), 2, 2), ("3-D avgpool 3x3x3 s3", lambda v: np.mean(v), 3, 3)]:
        f = kex.kmap(kernel_size=(k,) * 3, strides=(s,) * 3)(fn)
        y = f(x)
        med, _ = timeit(lambda: f(x), 4)
        rep(f"{name} on 1024^3", x.numel() * 4 + y.numel() * 4, med, y.numel())
    del x, y
    xb = torch.randn((16, 256, 512, 512), device=dev, generator=g)
    f = kex.vmap(kex.kmap(kernel_size=(3, 3, 3), strides=(2, 2, 2))(lambda v: np.mean(v)))
    y = f(xb)
    med, _ = timeit(lambda: f(xb), 4)
    rep("3-D avgpool 3x3x3 s2 under vmap on (16,256,512,512)", xb.numel() * 4 + y.numel() * 4, med, y.numel())
    del xb, y
    torch.cuda.empty_cache()
