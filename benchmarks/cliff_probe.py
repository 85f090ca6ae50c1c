import os, sys, json
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/benchmarks")
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit, peak
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def rep(name, nbytes, med):
    print(json.dumps({"path": name, "ms": round(med, 3), "frac": round(nbytes / (med * 1e-3) / 1e9 / peak(), 3), "kernel": _lib.last_kernel()}), flush=True)
# 1. avg pool 3x3 s2 backward
x = torch.randn((8, 4096, 4096), device=dev, generator=g).requires_grad_()
pool = kex.vmap(kex.kmap(kernel_size=(3, 3), strides=(2, 2))(lambda v: np.mean(v)))
try:
    y = pool(x)
    gy = torch.randn(y.shape, device=dev, generator=g)
    med, _ = timeit(lambda: torch.autograd.grad(y, x, gy, retain_graph=True), 4)
    rep("avgpool 3x3 s2 backward (8,4096^2)", x.numel() * 4 + y.numel() * 4, med)
except Exception as e:
    print("avgpool backward:", type(e).__name__, str(e)[:200])
# non-vmapped single image
x1 = torch.randn((8192, 8192), device=dev, generator=g).requires_grad_()
p1 = kex.kmap(kernel_size=(3, 3), strides=(2, 2))(lambda v: np.mean(v))
y = p1(x1); gy = torch.randn(y.shape, device=dev, generator=g)
med, _ = timeit(lambda: torch.autograd.grad(y, x1, gy, retain_graph=True), 4)
rep("avgpool 3x3 s2 backward 8192^2", x1.numel() * 4 + y.numel() * 4, med)
# 2. int32 -> float results
xi = torch.randint(-100, 100, (8192, 8192), device=dev, dtype=torch.int32, generator=g)
f = kex.kmap(kernel_size=(3, 3), padding="same")(lambda v: np.mean(v))
y = f(xi); med, _ = timeit(lambda: f(xi), 4)
rep("np.mean 3x3 same on int32 8192^2 (float32 result)", xi.numel() * 8, med)
f = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lambda v: 0.25 * (v[1, 0] + v[-1, 0] + v[0, 1] + v[0, -1]))
y = f(xi); med, _ = timeit(lambda: f(xi), 4)
rep("0.25*(4 neighbours) on int32 8192^2 (float32 result)", xi.numel() * 8, med)
# 3. strided weighted window 5x5 s2
xf = torch.randn((8192, 8192), device=dev, generator=g)
w5 = np.random.default_rng(0).standard_normal((5, 5)).astype(np.float32)
f = kex.kmap(kernel_size=(5, 5), strides=(2, 2), padding="same")(lambda v: np.sum(v * w5))
y = f(xf); med, _ = timeit(lambda: f(xf), 4)
rep("weighted 5x5 s2 on 8192^2", xf.numel() * 4 + y.numel() * 4, med)
# 4. 3-D patches
x3 = torch.randn((256, 256, 256), device=dev, generator=g)
f = kex.kmap(kernel_size=(3, 3, 3), padding="same")(lambda v: v)
y = f(x3); med, _ = timeit(lambda: f(x3), 4)
rep("3x3x3 patches of 256^3", x3.numel() * 4 + y.numel() * 4, med)
# 5. unaligned width 3-D star
x3 = torch.randn((512, 512, 514), device=dev, generator=g)
f = kex.kmap(kernel_size=(3, 3, 3), padding="same", relative=True)(lambda v: v[1,0,0]+v[-1,0,0]+v[0,1,0]+v[0,-1,0]+v[0,0,1]+v[0,0,-1]-6*v[0,0,0])
y = f(x3); med, _ = timeit(lambda: f(x3), 4)
rep("3-D 7-pt on 512x512x514 (rows not 16-byte aligned)", x3.numel() * 8, med)
# 6. unaligned width 2-D
x2 = torch.randn((8192, 8191), device=dev, generator=g)
f = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lambda v: v[1,0]+v[-1,0]+v[0,1]+v[0,-1]-4*v[0,0])
y = f(x2); med, _ = timeit(lambda: f(x2), 4)
rep("2-D 5-pt on 8192x8191 (odd width)", x2.numel() * 8, med)
# 7. smap over a 3-D volume (a Jacobi step)
x3 = torch.randn((768, 768, 768), device=dev, generator=g)
f = kex.smap(kernel_size=(3, 3, 3), offset=1, relative=True)(lambda v: (v[1,0,0]+v[-1,0,0]+v[0,1,0]+v[0,-1,0]+v[0,0,1]+v[0,0,-1]) / 6.0)
y = f(x3); med, _ = timeit(lambda: f(x3), 4)
rep("smap 3-D 6-neighbour Jacobi step on 768^3", x3.numel() * 8, med)
# 8. 3-D max filter
x3 = torch.randn((512, 512, 512), device=dev, generator=g)
f = kex.kmap(kernel_size=(3, 3, 3), padding="same")(lambda v: np.max(v))
y = f(x3); med, _ = timeit(lambda: f(x3), 4)
rep("3-D max filter 3x3x3 same on 512^3", x3.numel() * 8, med)
