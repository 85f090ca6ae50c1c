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

    def summary(self):
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e30)]
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
    with ClockSampler(local) as clocks:   # started before the warm-up so it is polling by the timed region
        for _ in range(max(args.warmup, 3)):
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
                "traffic": traffic, "kernel": kernel_name, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": BYTES_PER_LUP * n_win,
                "frac_of_8TBs_spec": achieved / 8000.0}

    # ---- end-to-end arm: public API, pinned HOST arrays, H2D + kernel + D2H every step ----
    e2e = None
    cpu = None
    if world == 1:
        del holder["y"]
    else:
        slab = None
    torch.cuda.empty_cache()
    # every rank pushes its own (N, N) host grid through the public call: weak scaling, like `value`
    pinned = torch.empty((N, N), dtype=torch.float32, pin_memory=True)
    x_host = pinned.numpy()
    x_host[...] = host_grid(N, seed=rank)
    e2e_steps = max(3, min(args.steps, 8))
    for _ in range(2):
        y_host = lap(x_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        y_host = lap(x_host)  # returns a NumPy array backed by pinned memory, already synchronised
    dt = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # slowest rank
        dt = float(t.item())
    e2e = {"value": world * (N - 2) * (N - 2) / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(x_host.nbytes) * world,
           "d2h_bytes_per_step": int(y_host.nbytes) * world, "steps": e2e_steps, "ms_per_step": dt * 1e3,
           "api": "kernex_b200.kmap(kernel_size=(3,3), padding='valid', relative=True)(laplacian)(numpy_array), "
                  "one (16384, 16384) pinned host grid per rank; wall clock, max over ranks"}
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
            "clocks": clocks.summary(), "parity_check": parity,
        }
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)   # anything printed during teardown stays off stdout as well
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
