#!/usr/bin/env python
"""CPU baselines for every BASELINE.json config at sizes the host finishes in seconds, timed with the
oracle's C + OpenMP restatements on all host threads: `stream` = hand-written streaming stencil / patch / row-march
loops (the best CPU code), `views` = kernex's own algorithm (pad + materialised index views + gather + function; 2-D
kmap only).  One JSON object per line.  Reported next to the GPU numbers, never the optimisation target."""
import itertools, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kx_oracle_c as kc


def best(fn, reps=3):
    fn()
    t = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t = min(t, time.perf_counter() - t0)
    return t


def emit(config, kind, lups, dt, note=""):
    print(json.dumps({"config": config, "kind": kind, "glups": lups / dt / 1e9, "ms": dt * 1e3, "cores": kc.max_threads(),
                      "note": note}), flush=True)


rng = np.random.default_rng(0)
# C1: full size
x = np.arange(1, 1_000_001, dtype=np.int32)
out = np.empty(999_998, np.int32)
dt = best(lambda: kc.affine_map(x, (3,), (1,), ((0, 0),), [(0,), (1,), (2,)], [1, 1, 1], out=out))
emit("C1 1-D int32 sum k=3 N=1e6 (full size)", "stream", out.size, dt)
# C2: 4096^2
n = 4096
x = rng.standard_normal((n, n), dtype=np.float32)
taps, coefs = [(2, 1), (1, 0), (1, 1), (1, 2), (0, 1)], [1.0, 1.0, -4.0, 1.0, 1.0]
out = np.empty((n - 2, n - 2), np.float32)
dt = best(lambda: kc.affine_map(x, (3, 3), (1, 1), ((0, 0), (0, 0)), taps, coefs, out=out))
emit("C2 2-D 5-pt Laplacian 4096^2", "stream", out.size, dt)
ref_terms = [((1, 2), 0.0), ((1, 0), 1.0), ((1, 1), 0.0), ((0, 2), 1.0), ((0, 0), -4.0), ((0, 1), 1.0), ((2, 2), 0.0),
             ((2, 0), 1.0), ((2, 1), 0.0)]
dt = best(lambda: kc.map_views(x, (3, 3), (1, 1), ((0, 0), (0, 0)), [t for t, _ in ref_terms], [c for _, c in ref_terms],
                               relative=True, out=out), 2)
emit("C2 2-D 5-pt Laplacian 4096^2", "views", out.size, dt, "kernex's own algorithm restated")
# C3a: (3, 2048, 2048) conv
n = 2048
x = rng.standard_normal((3, n, n), dtype=np.float32)
w = rng.standard_normal((3, 3, 3)).astype(np.float32)
t3 = list(itertools.product(range(3), range(3), range(3)))
out = np.empty((1, n, n), np.float32)
dt = best(lambda: kc.affine_map(x, (3, 3, 3), (1, 1, 1), ((0, 0), (1, 1), (1, 1)), t3, [float(w[t]) for t in t3], out=out))
emit("C3a conv-as-kmap (3,3,3) on (3,2048,2048)", "stream", out.size, dt)
# C3b: patches 2048^2
x = rng.standard_normal((n, n), dtype=np.float32)
dt = best(lambda: kc.patch_map(x, (3, 3), (1, 1), ((1, 1), (1, 1))))
emit("C3b 3x3 patches same on 2048^2", "stream", n * n, dt)
# C4: kscan convection nt=512, nx=2^18
nt, nx, kappa = 512, 1 << 18, 0.5
q1, q2 = int((nx - 1) / 4) + 1, int((nx - 1) / 2)
regions = [(0, nt, 0, 1, (0, 0, 0), 1.0), (0, nt, nx - 1, nx, (0, 0, 0), 1.0), (0, nt, 0, q1, (0, 0, 0), 1.0),
           (0, nt, q2, nx, (0, 0, 0), 1.0), (0, 1, q1, q2, (0, 0, 0), 2.0), (1, nt, 1, nx - 1, (kappa, 1.0 - kappa, 0.0), 0.0)]
outs = np.zeros((nt, nx), np.float32)
dt = best(lambda: kc.rowmarch(nt, nx, regions, out=outs), 2)
emit("C4 kscan convection nt=512 nx=2^18", "stream", nt * nx, dt, "row-march restatement (rows sequential, columns parallel)")
# C5: 3-D 7-pt Laplacian 256^3
n = 256
x = rng.standard_normal((n, n, n), dtype=np.float32)
t7 = [(2, 1, 1), (0, 1, 1), (1, 2, 1), (1, 0, 1), (1, 1, 2), (1, 1, 0), (1, 1, 1)]
out = np.empty((n - 2,) * 3, np.float32)
dt = best(lambda: kc.affine_map(x, (3, 3, 3), (1, 1, 1), ((0, 0),) * 3, t7, [1, 1, 1, 1, 1, 1, -6], out=out))
emit("C5 3-D 7-pt Laplacian 256^3", "stream", out.size, dt)
