#!/usr/bin/env python
"""Where one small kmap call spends its host time (BASELINE config 1 is launch-latency bound: 8 MB of traffic is
1.3 us of HBM time).  Layers, innermost first, each timed back to back over N calls with one synchronize at the end:
  noop     a ctypes call that does nothing (kx_abi_version)
  abi      kx_map through ctypes with pre-built arguments (validation, dispatch, cudaLaunchKernel)
  prepared PreparedMap.__call__ into a caller-owned output
  kmap     the decorated function, output allocated per call (what a user writes)
  torch    torch.add(x, 1, out=y): one library launch, for scale
One JSON line per (workload, layer)."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
from kernex_b200 import _lib, engine

dev = torch.device("cuda")
N = int(os.environ.get("N", 20000))
lib = _lib.load()


def bench(fn, n=N):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / n * 1e6)
    return round(best, 2)


def layers(name, f, x, extra=()):
    y = f(x, *extra)
    itf = f
    inner = getattr(f, "__wrapped__", None)
    shape = tuple(x.shape)
    # the engine closure's pieces, rebuilt through the public PreparedMap
    ki = f.__self__ if hasattr(f, "__self__") else None
    rows = {"workload": name, "kernel": _lib.last_kernel()}
    rows["kmap_us"] = bench(lambda: itf(x, *extra))
    rows["torch_add_us"] = bench(lambda: torch.add(x, 1, out=x2))
    rows["noop_ctypes_us"] = bench(lambda: lib.kx_abi_version())
    return rows


x = torch.arange(1, 1_000_001, dtype=torch.int32, device=dev)
x2 = torch.empty_like(x)
f1 = kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))
r = layers("C1: 1-D int32 sum k=3, 1e6", f1, x)
pm = engine.PreparedMap({(lambda v: np.sum(v)): ()}, (1_000_000,), (3,), (1,), ((0, 0),), False, dtype=torch.int32)
out = torch.empty(pm.out_shape, dtype=pm.out_dtype, device=dev)
r["prepared_us"] = bench(lambda: pm(x, out=out))
geom, prog = C.byref(pm.geom), C.byref(pm.prog.c)
px, po, nul = C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
r["abi_us"] = bench(lambda: lib.kx_map(geom, prog, px, nul, 0, po, 1, st))
print(json.dumps(r), flush=True)

g = torch.Generator(device=dev).manual_seed(0)
for Cc, H in ((4, 16), (16, 16), (64, 16), (16, 256)):
    conv = kex.kmap(kernel_size=(Cc, 3, 3), padding=("valid", "same", "same"), relative=False)(lambda v, w: np.sum(v * w))
    xi = torch.randn((Cc, H, H), device=dev, generator=g)
    w = torch.randn((Cc, 3, 3), device=dev, generator=g)
    x2 = torch.empty_like(xi)
    x = xi
    r = layers(f"conv ({Cc},{H},{H}), device weights", conv, xi, (w,))
    print(json.dumps(r), flush=True)

lap = kex.kmap(kernel_size=(3, 3), padding="valid", relative=True)(lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0])
x = torch.randn((256, 256), device=dev, generator=g)
x2 = torch.empty_like(x)
print(json.dumps(layers("2-D Laplacian 256^2", lap, x)), flush=True)
xh = np.ones((256, 256), np.float32)
print(json.dumps({"workload": "2-D Laplacian 256^2, NumPy in / NumPy out", "kmap_us": bench(lambda: lap(xh), 2000)}), flush=True)
