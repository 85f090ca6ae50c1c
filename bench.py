#!/usr/bin/env python
"""Headline benchmark of the kernex kmap hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, host cores)

Workload (BASELINE.json configs[1]): 2-D 5-point relative Laplacian ``kmap``,
16384 x 16384 float32, padding='valid'  ->  (16382, 16382) output, 8 algorithmic bytes
per lattice update (read the grid once, write the result once).  With N > 1 the grid is
(N*16384) x 16384, domain-decomposed in row slabs (axis 0) over the ranks: every step each
rank exchanges one halo row with each neighbour (NCCL send/recv) and applies the stencil to
its slab, so per-GPU work is fixed ("weak" scaling).

One "step" = one application of the stencil to the whole (per-rank) grid.  Prints ONE JSON
line (rank 0).  Inputs are 1.07 GB per rank, far larger than the 126 MB L2, so consecutive
steps cannot be served from cache (stated in config.cache).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 16384
TAPS = [(2, 1), (1, 0), (1, 1), (1, 2), (0, 1)]      # window-origin offsets of the 5 live taps
COEFS = [1.0, 1.0, -4.0, 1.0, 1.0]
BYTES_PER_LUP = 8.0                                     # 4 B read + 4 B written
METRIC = "kmap 2-D 5-pt Laplacian fp32 16384^2 (valid) throughput"
UNIT = "GLUP/s"


def laplacian(x):
    # README.md:122-133 verbatim, zero-coefficient terms included
    return (0 * x[1, -1] + 1 * x[1, 0] + 0 * x[1, 1] + 1 * x[0, -1] + -4 * x[0, 0] + 1 * x[0, 1]
            + 0 * x[-1, -1] + 1 * x[-1, 0] + 0 * x[-1, 1])


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): an NVML
    polling thread (5 ms period; `nvidia-smi -lms` is the fallback), samples are time-stamped
    and only those inside [mark_start, mark_end] are summarised."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t0, self.t1 = index, [], False, None, None
        self.thread = None

    def _poll_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        while not self.stop:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append((time.perf_counter(), float(sm), float(mx), int(bits)))
            time.sleep(0.005)

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        try:
            for line in proc.stdout:
                c = [v.strip() for v in line.split(",")]
                bits = 0
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                    if v.lower().startswith("active"):
                        bits |= self.REASONS[nm]
                try:
                    self.rows.append((time.perf_counter(), float(c[0]), float(c[1]), bits))
                except ValueError:
                    pass
                if self.stop:
                    break
        finally:
            proc.terminate()

    def _run(self):
        try:
            self._poll_nvml()
        except Exception:
            try:
                self._poll_smi()
            except Exception:
                pass

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *exc):
        self.stop = True
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self, t0=None, t1=None):
        t0 = self.t0 if t0 is None else t0
        t1 = self.t1 if t1 is None else t1
        rows = [r for r in self.rows if t0 is not None and t0 <= r[0] <= (t1 or 1e30)]
        window = "timed region"
        if not rows:
            rows, window = list(self.rows), "warm-up + timed region (no sample fell inside the timed region)"
        sm = [r[1] for r in rows]
        reasons = sorted(nm for nm, bit in self.REASONS.items() if any(r[3] & bit for r in rows))
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "reasons": reasons, "samples": len(sm), "window": window}


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def host_grid(rows, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((rows, N), dtype=np.float32)


# README.md:122-133 as the reference evaluates it: nine terms on the rolled 3x3 patch (relative
# index -1 is patch index 2), zero coefficients included
REF_TERMS = [((1, 2), 0.0), ((1, 0), 1.0), ((1, 1), 0.0), ((0, 2), 1.0), ((0, 0), -4.0), ((0, 1), 1.0),
             ((2, 2), 0.0), ((2, 0), 1.0), ((2, 1), 0.0)]
SAMPLE_ROWS = 4096   # the CPU arms time a (4096, 16384) slab of the grid: 1/4 of the workload per step


class CpuArms:
    """The two CPU restatements of oracle/kx_oracle_c.c on the host cores:
    `reference_algorithm` = kernex's own kmap algorithm (pad at the end, materialised int32 index
    views, per-window gather + roll + function: kxo_map_views_f32), the stand-in for kernex on JAX's
    CPU backend, which cannot be installed here; `streaming_stencil` = a hand-written OpenMP
    stencil (kxo_affine_f32), i.e. the best CPU code for the same result, reported next to it."""

    def __init__(self, x_host):
        from oracle import kx_oracle_c as kc
        self.kc, self.threads = kc, host_threads()
        self.x = np.ascontiguousarray(x_host[:SAMPLE_ROWS])
        self.geom = ((3, 3), (1, 1), ((0, 0), (0, 0)))
        self.out = np.empty((SAMPLE_ROWS - 2, N - 2), dtype=np.float32)
        self.scratch = (np.empty(SAMPLE_ROWS * N, np.float32), np.empty(self.out.size * 6, np.int32))

    def reference_algorithm(self):
        self.kc.map_views(self.x, *self.geom, [t for t, _ in REF_TERMS], [c for _, c in REF_TERMS], relative=True,
                          nthreads=self.threads, out=self.out, scratch=self.scratch)
        return self.out

    def streaming_stencil(self):
        self.kc.affine_map(self.x, *self.geom, TAPS, COEFS, out=self.out, nthreads=self.threads)
        return self.out

    def best_of(self, fn, repeats):
        fn()  # warm-up (page faults)
        best = float("inf")
        for _ in range(repeats):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return self.out.size / best / 1e9


CPU_SAMPLE = (f"({SAMPLE_ROWS}, {N}) slab of the grid (1/4 of a step), kernex's kmap algorithm restated in C + OpenMP: "
              "pad, materialised int32 views, per-window gather + roll + 9-term function (oracle/kx_oracle_c.c "
              "kxo_map_views_f32)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arms = CpuArms(host_grid(SAMPLE_ROWS))
    for _ in range(max(args.warmup, 1)):
        arms.reference_algorithm()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arms.reference_algorithm()
    dt = time.perf_counter() - t0
    glups = arms.out.size * args.steps / dt / 1e9
    stream = arms.best_of(arms.streaming_stencil, 3)
    line = {
        "impl": "reference", "metric": METRIC, "value": glups, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: 2-D 5-pt Laplacian kmap, 16384x16384 fp32, padding=valid",
                   "note": "kernex needs JAX, which is not installable in this image (no jax wheel, no network); this "
                           "arm times the oracle port of kernex's own algorithm on the host cores, rank 0 only; "
                           "each step is a bounded sample, throughput is per lattice update"},
        "cpu_baseline": {"value": glups, "unit": UNIT, "cores": arms.threads, "kind": "port", "sample": CPU_SAMPLE,
                         "streaming_stencil_glups": stream,
                         "streaming_stencil_note": "hand-written OpenMP stencil for the same result (kxo_affine_f32): "
                                                   "the best CPU code, not what the reference executes"},
        "e2e": {"value": glups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configs, measured in the same process after the headline (C2) numbers: `configs` in the
# JSON line.  Device-resident inputs, CUDA events around back-to-back applications after 3 warm-ups, max over ranks;
# every entry carries its own oracle spot check (`parity`) and the SM clock seen while it ran.
# --------------------------------------------------------------------------------------------------------------

def lap3(x):
    return (x[1, 0, 0] + x[-1, 0, 0] + x[0, 1, 0] + x[0, -1, 0] + x[0, 0, 1] + x[0, 0, -1] - 6 * x[0, 0, 0])


LAP3_TAPS = [(2, 1, 1), (0, 1, 1), (1, 2, 1), (1, 0, 1), (1, 1, 2), (1, 1, 0), (1, 1, 1)]
LAP3_COEFS = [1, 1, 1, 1, 1, 1, -6]
NVLINK_GBS = 900.0   # per direction per GPU (task statement)


def convection_mesh(kex, nx, kappa=0.5):
    """README.md:231-266 linear convection as a kscan mesh; CFL kept at 0.5."""
    F = kex.kscan(kernel_size=(3, 3), padding=((1, 1), (1, 1)), named_axis={0: "n", 1: "i"}, relative=True)
    q1, q2 = int((nx - 1) / 4) + 1, int((nx - 1) / 2)
    F[:, 0] = F[:, -1] = lambda u: 1
    F[:, :q1] = F[:, q2:] = lambda u: 1
    F[0:1, q1:q2] = lambda u: 2
    F[1:, 1:-1] = lambda u: u["i", "n-1"] - kappa * (u["i", "n-1"] - u["i-1", "n-1"])
    return F, q1, q2


def run_config_suite(world, rank, dev, clocks, peak, budget_s=60.0):
    import torch
    import torch.distributed as dist

    import kernex_b200 as kex
    from kernex_b200 import _lib
    from oracle import kx_oracle as ko        # checker only (never inside a timed region)
    from oracle import kx_oracle_c as kc

    t_suite = time.perf_counter()
    out = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, reps, warm=3):
        for _ in range(warm):
            step()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), (t0, t1)

    def total(v):
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def emit(name, lups_total, bpl, ms, window, parity, kernel=None, **extra):
        glups = lups_total / (ms * 1e-3) / 1e9
        per_gpu_gbs = glups / world * bpl
        ck = clocks.summary(*window)
        out.append({"config": name, "n_gpus": world, "glups": glups, "ms": ms, "bytes_per_lup": bpl,
                    "roofline_frac": per_gpu_gbs / peak, "frac_of_8TBs_spec": per_gpu_gbs / 8000.0,
                    "kernel": kernel or _lib.last_kernel(), "parity": parity,
                    "clocks": {"sm_mhz": ck["sm_mhz"], "reasons": ck["reasons"]}, **extra})

    def within(got, want, scale):
        return bool(np.all(np.abs(got.astype(np.float64) - want) <= 1e-5 * scale + 1e-30))

    def over_budget():
        return time.perf_counter() - t_suite > budget_s

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    if world == 1:
        # ---- C1: README sum, 1-D int32, k=3, valid, N=1e6 (bit exact; 8 MB: launch-latency bound) -------------
        x = torch.arange(1, 1_000_001, dtype=torch.int32, device=dev)
        f = kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))
        y = f(x)
        xh = np.arange(1, 1_000_001, dtype=np.int32)
        par = bool(np.array_equal(y.cpu().numpy(), xh[:-2] + xh[1:-1] + xh[2:]))
        ms, win = timed(lambda: f(x), 200)
        emit("C1 1-D int32 sum k=3 valid N=1e6", 999_998, 8, ms, win, par, note="8 MB: launch-latency bound, no roofline claim")

        # ---- C2 backward: dx = flipped-tap stencil of the (16382, 16382) cotangent --------------------------------
        n = N
        lap = kex.kmap(kernel_size=(3, 3), padding="valid", relative=True)(laplacian)
        x = torch.randn((n, n), device=dev, generator=gen).requires_grad_()
        y = lap(x)
        g = torch.randn((n - 2, n - 2), device=dev, generator=gen)
        (dx,) = torch.autograd.grad(y, x, g, retain_graph=True)
        gb = g[7998:8064].cpu().numpy()                   # dx rows [8000, 8064) = 'full' correlation of g rows [7998, 8064)
        flipped = [(2 - a, 2 - b) for a, b in TAPS]
        want = ko.affine_map(gb, (3, 3), (1, 1), ((0, 0), (2, 2)), flipped, COEFS)
        scale = ko.affine_map(np.abs(gb), (3, 3), (1, 1), ((0, 0), (2, 2)), flipped, np.abs(COEFS))
        par = within(dx[8000:8064].cpu().numpy(), want, scale)
        ms, win = timed(lambda: torch.autograd.grad(y, x, g, retain_graph=True), 30)
        emit("C2 backward (custom VJP wrt x) 16384^2", n * n, 8, ms, win, par)
        del x, y, g, dx
        torch.cuda.empty_cache()

        # ---- C3a: conv-as-kmap (3,3,3), padding (valid, same, same), (3, 8192, 8192), device weights + dW --------
        n = 8192
        x = torch.randn((3, n, n), device=dev, generator=gen)
        w = torch.randn((3, 3, 3), device=dev, generator=gen).requires_grad_()
        conv = kex.kmap(kernel_size=(3, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
        y = conv(x, w)
        taps3 = [(c, a, b) for c in range(3) for a in range(3) for b in range(3)]
        wh = w.detach().cpu().numpy()
        coefs3 = [float(wh[t]) for t in taps3]
        band = x[:, 3999:4033].cpu().numpy()
        bb = ((0, 0), (0, 0), (1, 1))
        want = ko.affine_map(band, (3, 3, 3), (1, 1, 1), bb, taps3, coefs3)
        scale = ko.affine_map(np.abs(band), (3, 3, 3), (1, 1, 1), bb, taps3, np.abs(coefs3))
        par = within(y[:, 4000:4032].detach().cpu().numpy(), want, scale)
        with torch.no_grad():
            ms, win = timed(lambda: conv(x, w), 30)
        emit("C3a conv-as-kmap (3,3,3) on (3,8192,8192), padding (valid,same,same)", n * n, 16, ms, win, par,
             kernel="conv_fwd_rows_kernel")
        g = torch.zeros((1, n, n), device=dev)
        g[0, 4000:4032].normal_(generator=gen)            # banded cotangent: the NumPy oracle can follow it
        (dw,) = torch.autograd.grad(y, w, g, retain_graph=True)
        want_dw = ko.weight_grad(band, g[:, 4000:4032].cpu().numpy(), (3, 3, 3), (1, 1, 1), bb)
        mag_dw = ko.weight_grad(np.abs(band), np.abs(g[:, 4000:4032].cpu().numpy()), (3, 3, 3), (1, 1, 1), bb)
        par = bool(np.all(np.abs(dw.cpu().numpy().astype(np.float64) - want_dw) <= 1e-5 * mag_dw + 1e-30))
        g.normal_(generator=gen)
        ms_ag, _ = timed(lambda: torch.autograd.grad(y, w, g, retain_graph=True), 30)
        # the same kernels without autograd's Python around them: one kx_map_vjp_coef call on a prepared plan
        from kernex_b200.engine import PreparedMap
        pm = PreparedMap({(lambda v, w: np.sum(v * w)): ()}, (3, n, n), (3, 3, 3), (1, 1, 1), ((0, 0), (1, 1), (1, 1)), False,
                         args=(w.detach(),))
        dw2 = pm.vjp_coef(x, g)
        par = par and bool(torch.equal(dw2.reshape(-1), torch.autograd.grad(y, w, g, retain_graph=True)[0].reshape(-1)))
        ms, win = timed(lambda: pm.vjp_coef(x, g), 30)
        emit("C3a weight gradient (fused window correlation)", n * n, 16, ms, win, par, kernel="conv_wgrad_rows_kernel",
             ms_through_autograd=ms_ag)
        del x, y, g, w
        torch.cuda.empty_cache()

        # ---- C3b: 3x3 patch extraction 'same' on 8192^2 (bit exact) --------------------------------------------------
        x = torch.randn((n, n), device=dev, generator=gen)
        fp = kex.kmap(kernel_size=(3, 3), padding="same")(lambda v: v)
        p = fp(x)
        par = True
        for (a, b) in ((0, 0), (1, 2), (2, 1)):
            ys, xs = a - 1, b - 1
            want = torch.zeros((n, n), device=dev)
            want[max(-ys, 0):n - max(ys, 0), max(-xs, 0):n - max(xs, 0)] = \
                x[max(ys, 0):n - max(-ys, 0), max(xs, 0):n - max(-xs, 0)]
            par = par and bool(torch.equal(p[:, :, a, b], want))
        del p, want
        ms, win = timed(lambda: fp(x), 20)
        emit("C3b 3x3 patches 'same' on 8192^2 fp32", n * n, 40, ms, win, par)
        del x
        torch.cuda.empty_cache()

    if world > 1 and not over_budget():
        # ---- C2 backward on the slabs: flipped-tap stencil per piece + the TRANSPOSED halo exchange (the cotangent
        # rows that landed in my halos are sent back to their owners and added there) -------------------------------
        from kernex_b200.dist import SlabStencil
        slab = SlabStencil(laplacian, kernel_size=(3, 3), relative=True, rows_per_rank=N, width=N, device=dev)
        slab.fill_random(gen)
        o = slab.step()
        g = torch.randn(o.shape, device=dev, generator=gen)
        dx, _ = slab.vjp(g)
        torch.cuda.synchronize()
        # rows 8000..8063 of my slab: interior, the float64-free oracle form of the forward check applies to g
        off = 0 if slab.layout.first else 1     # my output row j is the window starting at own row j - off
        gb = g[7998 + off:8064 + off].cpu().numpy()
        flipped = [(2 - a, 2 - b) for a, b in TAPS]
        want = ko.affine_map(gb, (3, 3), (1, 1), ((0, 0), (2, 2)), flipped, COEFS)
        scale = ko.affine_map(np.abs(gb), (3, 3), (1, 1), ((0, 0), (2, 2)), flipped, np.abs(COEFS))
        par = all_ok(within(dx[8000:8064].cpu().numpy(), want, scale))
        ms, win = timed(lambda: slab.vjp(g), 20)
        emit("C2 backward on axis-0 slabs (VJP wrt x, transposed halo exchange)", total(N * N), 8, ms, win, par,
             kernel="stencil2d_kernel",
             halo_transport=("p2p pull-add (the cotangent rows that landed in a neighbour's halo are pulled over NVLink and "
                             "added by one kernel)" if slab.__dict__.get("_gpeers") is not None else
                             "nccl (cotangent halo rows are sent back and added)"))
        slab.close()
        del slab, o, g, dx
        torch.cuda.empty_cache()

    # ---- C4: kscan linear convection, nt=4096, nx = 2^21 columns per GPU (2^24 on 8), axis-1 shards -------------
    if not over_budget():
        from kernex_b200.dist import ColumnShardedScan
        nt, nx = 4096, (1 << 21) * world
        F, q1, q2 = convection_mesh(kex, nx)
        sh = ColumnShardedScan(F, nt, nx, rank=rank, world=world, device=dev)
        res = sh.run()
        torch.cuda.synchronize()
        # the wave front (non-constant at every depth) against the C oracle, on the rank that owns it
        W = 3000
        c_lo = sh.col_base + sh.keep_lo
        par = bool(float(res.max()) <= 2.0 and float(res.min()) >= 1.0)
        if c_lo <= q1 - 1 and q1 - 1 + W <= c_lo + res.shape[1]:
            front = [(0, nt, 0, 1, (0, 0, 0), 1.0), (0, 1, 1, W, (0, 0, 0), 2.0), (1, nt, 1, W, (0.5, 0.5, 0.0), 0.0)]
            want = kc.rowmarch(nt, W, front)
            got = res[:, q1 - 1 - c_lo:q1 - 1 - c_lo + W].cpu().numpy()
            par = par and bool(np.allclose(got, want, rtol=1e-5, atol=1e-5))
        par = all_ok(par)
        del res
        ms, win = timed(sh.run, 5, warm=2)
        emit("C4 kscan linear convection nt=4096, nx=2^21 per GPU (row-march)", total(sh.local_windows), 4, ms, win, par,
             kernel="scan_rowmarch_kernel", nx_global=nx, sharding="axis 1 (columns); the nt*R-column dependency cone is "
             "recomputed instead of exchanged" if world > 1 else "single GPU",
             recomputed_columns=int(sh.keep_lo), halo_bytes_per_step=0)
        del sh
        torch.cuda.empty_cache()

    # ---- C5: 3-D 7-pt Laplacian; N=1: 2048^3; N>1: weak, 256 x 2048^2 planes per GPU, one-plane halo per neighbour ----
    if not over_budget():
        from kernex_b200.dist import SlabStencil
        planes = 2048 if world == 1 else 256
        slab = SlabStencil(lap3, kernel_size=(3, 3, 3), relative=True, rows_per_rank=planes, trailing_shape=(2048, 2048),
                           device=dev, padding="valid")
        slab.fill_random(gen)
        o = slab.step()
        torch.cuda.synchronize()
        par = True
        L = slab.layout
        for out0 in (0, o.shape[0] - 2):                   # first / last output planes: they read the exchanged halo planes
            in0 = out0 + (1 if L.first else 0) + L.halo_top - 1
            blk = slab.ext[in0:in0 + 4].cpu().numpy()
            want = ko.affine_map(blk, (3, 3, 3), (1, 1, 1), ((0, 0),) * 3, LAP3_TAPS, LAP3_COEFS)
            scale = ko.affine_map(np.abs(blk), (3, 3, 3), (1, 1, 1), ((0, 0),) * 3, LAP3_TAPS, np.abs(LAP3_COEFS))
            par = par and within(o[out0:out0 + 2].cpu().numpy(), want, scale)
        par = all_ok(par)
        ms, win = timed(slab.step, 8 if world == 1 else 30)
        halo = 0 if world == 1 else 2048 * 2048 * 4
        emit(f"C5 3-D 7-pt Laplacian fp32, {'2048^3' if world == 1 else f'{256 * world}x2048x2048 (256 planes per GPU, weak)'}",
             total(slab.local_windows), 8, ms, win, par, kernel="stencil3d_star_kernel",
             halo_bytes_per_neighbour=halo, halo_transport=slab.transport if world > 1 else None,
             nvlink_halo_frac=(halo / (NVLINK_GBS * 1e9)) / (ms * 1e-3) if world > 1 else 0.0)
        slab.close()
        del slab, o
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import kernex_b200 as kex
    from kernex_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL prints its version banner on stdout when the first communicator is created: everything this process
    # writes goes to stderr until rank 0 prints its ONE JSON line
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()  # no extension -> fail loudly, there is no fallback

    lap = kex.kmap(kernel_size=(3, 3), padding="valid", relative=True)(laplacian)

    # ---- device-resident arm ------------------------------------------------------
    gen = torch.Generator(device=dev).manual_seed(rank)
    if world > 1:
        from kernex_b200.dist import SlabStencil
        slab = SlabStencil(laplacian, kernel_size=(3, 3), relative=True, rows_per_rank=N, width=N, device=dev)
        slab.fill_random(gen)
        step = slab.step
        n_win = slab.local_windows
        halo_transport = slab.transport + (f" ({slab.transport_note})" if slab.transport_note else "")
    else:
        x = torch.randn((N, N), device=dev, dtype=torch.float32, generator=gen)
        holder = {}

        def step():
            holder["y"] = lap(x)
        n_win = (N - 2) * (N - 2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local).__enter__()   # started before the warm-up so it is polling by the timed region;
    for _ in range(max(args.warmup, 3)):       # it keeps polling through the config suite and is stopped at the end
        step()
    barrier()
    launches0 = _lib.launch_count()
    clocks.mark_start()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    kernel_name = _lib.last_kernel()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([float(n_win)], device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_win = float(tot.item())
    else:
        total_win = float(n_win)
    ms_per_step = ms / args.steps
    value = total_win / (ms_per_step * 1e-3) / 1e9

    # ---- parity spot check against the oracle (not timed) ---------------------------
    parity = None
    if world > 1:
        # every rank checks the first and last 64 output rows of its slab (they involve the exchanged halo rows)
        from oracle import kx_oracle as ko
        L, ok = slab.layout, True
        for out0 in (0, slab.out.shape[0] - 64):
            in0 = out0 + (1 if L.first else 0) + L.halo_top - 1   # ext row of the first window row of output row out0
            band = slab.ext[in0:in0 + 66].cpu().numpy()
            want = ko.affine_map(band, (3, 3), (1, 1), ((0, 0), (0, 0)), TAPS, COEFS)
            scale = ko.affine_map(np.abs(band), (3, 3), (1, 1), ((0, 0), (0, 0)), TAPS, np.abs(COEFS))
            err = np.abs(slab.out[out0:out0 + 64].cpu().numpy().astype(np.float64) - want)
            ok = ok and bool(np.all(err <= 1e-5 * scale + 1e-30))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = bool(flag.item())
    if world == 1:
        from oracle import kx_oracle as ko
        y = holder["y"]
        band = x[4000:4066].cpu().numpy()
        want = ko.affine_map(band, (3, 3), (1, 1), ((0, 0), (0, 0)), TAPS, COEFS)
        scale = ko.affine_map(np.abs(band), (3, 3), (1, 1), ((0, 0), (0, 0)), TAPS, np.abs(COEFS))
        err = np.abs(y[4000:4064].cpu().numpy().astype(np.float64) - want)
        parity = bool(np.all(err <= 1e-5 * scale + 1e-30))

    # ---- roofline of the dominant kernel ----------------------------------------------
    peak, peak_src = measured_peak()
    per_rank_ms = ms_per_step  # one stencil2d launch dominates the step (halo kernels are O(row))
    achieved = BYTES_PER_LUP * n_win / (per_rank_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "stencil2d_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "profiles/stencil2d_traffic.json (one `ncu --set full` capture of this "
                "kernel on this workload: dram__bytes_read.sum + dram__bytes_write.sum; static, not re-measured per run)",
                "kernel": "stencil2d_kernel", "last_kernel": kernel_name, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": BYTES_PER_LUP * n_win,
                "frac_of_8TBs_spec": achieved / 8000.0}

    # ---- the other BASELINE configs (device-resident), then the end-to-end arm ---------------------------------
    if world == 1:
        del holder["y"], x
    torch.cuda.empty_cache()
    configs = []
    if not args.no_configs:
        try:
            configs = run_config_suite(world, rank, dev, clocks, peak)
        except Exception as e:   # the headline line must survive a failure in the extra configs
            configs = [{"config": "suite aborted", "error": f"{type(e).__name__}: {e}"[:300]}]
    torch.cuda.empty_cache()

    # ---- end-to-end arm: HOST arrays in, HOST arrays out, H2D + kernel(s) + D2H inside the timed region --------
    e2e = None
    cpu = None
    e2e_steps = max(3, min(args.steps, 8))
    pinned = torch.empty((N, N), dtype=torch.float32, pin_memory=True)
    x_host = pinned.numpy()
    x_host[...] = host_grid(N, seed=rank)
    if world == 1:
        # one GPU: the reference-facing call on a NumPy array (hostpipe.py pipelines H2D / kernel / D2H over row slabs)
        for _ in range(2):
            y_host = lap(x_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            y_host = lap(x_host)  # returns a NumPy array backed by pinned memory, already synchronised
        dt = (time.perf_counter() - t0) / e2e_steps
        api = ("kernex_b200.kmap(kernel_size=(3,3), padding='valid', relative=True)(laplacian)(numpy_array), one "
               "(16384, 16384) pinned host grid; wall clock")
        out_bytes = int(y_host.nbytes)
    else:
        # several GPUs: the DISTRIBUTED path -- every rank holds its (16384, 16384) rows of the global host grid in
        # pinned memory; SlabStencil.step_host streams them H2D in chunks, pulls the halo rows from the neighbouring
        # GPUs over NVLink, launches the stencil chunk by chunk and streams the result rows D2H
        y_pinned = torch.empty(tuple(slab.out.shape), dtype=torch.float32, pin_memory=True)
        for _ in range(2):
            slab.step_host(pinned, y_pinned)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            slab.step_host(pinned, y_pinned)
        dt = (time.perf_counter() - t0) / e2e_steps
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # slowest rank
        dt = float(t.item())
        y_host = y_pinned.numpy()
        # the streamed result must equal the device-resident step on the same rows
        slab.set_local(pinned.cuda(non_blocking=True))
        same = bool(torch.equal(slab.step().cpu(), y_pinned))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = bool(parity and flag.item())
        api = (f"kernex_b200.dist.SlabStencil.step_host(pinned rows, pinned out) on every rank: global "
               f"({world * N}, {N}) host grid in row slabs, halo rows over NVLink ({halo_transport}); wall clock, max over ranks")
        out_bytes = int(total_win * 4 / world)
    e2e = {"value": total_win / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(x_host.nbytes) * world,
           "d2h_bytes_per_step": out_bytes * world, "steps": e2e_steps, "ms_per_step": dt * 1e3, "api": api}
    if world == 1:
        # ---- CPU baseline on the host cores (rank 0, N = 1 only) ---------------------------
        arms = CpuArms(x_host)
        ref_glups = arms.best_of(arms.reference_algorithm, 2)
        ref_out = arms.reference_algorithm().copy()
        stream = arms.best_of(arms.streaming_stencil, 3)
        cpu = {"value": ref_glups, "unit": UNIT, "cores": arms.threads, "kind": "port", "sample": CPU_SAMPLE,
               "streaming_stencil_glups": stream,
               "streaming_stencil_note": "hand-written OpenMP stencil for the same result (kxo_affine_f32): the best "
                                         "CPU code, not what the reference executes"}
        # the e2e result must agree with both CPU restatements (fp32: mul+add vs FMA chain)
        diff = float(np.max(np.abs(y_host[:2048] - ref_out[:2048])))
        diff2 = float(np.max(np.abs(y_host[:2048] - arms.streaming_stencil()[:2048])))
        parity = bool(parity and diff <= 1e-4 and diff2 <= 1e-4)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 2-D 5-pt relative Laplacian kmap, 16384x16384 fp32 per GPU, padding=valid",
                       "parallelism": "single GPU" if world == 1 else
                       f"axis-0 row slabs x{world}, 1-row halo exchange per step, transport={halo_transport}",
                       "cache": "input 1.07 GB + output 1.07 GB per GPU >> 126 MB L2; no flush needed",
                       "lattice_updates_per_step": total_win},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(), "parity_check": parity, "configs": configs,
        }
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)   # anything printed during teardown stays off stdout as well
    clocks.__exit__(None, None, None)
    if world > 1:
        dist.barrier()
        slab.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-configs", action="store_true", help="skip the extra BASELINE configs (`configs` in the line)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
