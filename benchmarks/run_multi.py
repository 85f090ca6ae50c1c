#!/usr/bin/env python
"""Multi-GPU measurements of the sharded paths (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29507 benchmarks/run_multi.py [--only c5_weak,c5_strong,c4] [--reps 10]

  c5_weak    3-D 7-pt Laplacian, 256 x 2048 x 2048 planes per GPU (global 256N x 2048^2), axis-0 slabs,
             one-plane NCCL halo exchange per application overlapped with the interior launch
  c5_strong  the same stencil on a fixed 2048^3 grid split over the N GPUs
  c4         kscan linear convection nt=4096, nx = 2^21 columns per GPU (global nx = N * 2^21), axis-1
             shards with a recomputed nt-column dependency cone instead of communication

Time = CUDA events around `reps` back-to-back applications, barrier + synchronize on both sides,
max over ranks; one JSON line per experiment on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lap3(x):
    return (x[1, 0, 0] + x[-1, 0, 0] + x[0, 1, 0] + x[0, -1, 0] + x[0, 0, 1] + x[0, 0, -1] - 6 * x[0, 0, 0])


def timed(step, reps, dev, world):
    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda c: not only or c in only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import kernex_b200 as kex
    from kernex_b200.dist import ColumnShardedScan, SlabStencil

    def emit(name, lups_total, bytes_per_lup, ms, extra=None):
        if rank == 0:
            gl = lups_total / (ms * 1e-3) / 1e9
            print(json.dumps({"experiment": name, "n_gpus": world, "ms": ms, "glups_total": gl,
                              "glups_per_gpu": gl / world, "gbs_per_gpu": gl / world * bytes_per_lup,
                              **(extra or {})}), flush=True)

    def total(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    for name, rows in (("c5_weak", 256), ("c5_strong", 2048 // world)):
        if not want(name):
            continue
        slab = SlabStencil(lap3, kernel_size=(3, 3, 3), relative=True, rows_per_rank=rows,
                           trailing_shape=(2048, 2048), device=dev, padding="valid")
        slab.fill_random(torch.Generator(device=dev).manual_seed(rank))
        ms = timed(slab.step, args.reps, dev, world)
        emit(name, total(slab.local_windows), 8, ms, {"planes_per_gpu": rows,
                                                      "halo_bytes_per_neighbour": 2048 * 2048 * 4})
        del slab
        torch.cuda.empty_cache()

    if want("c4"):
        nt, nx_per = 4096, 1 << 21
        nx = nx_per * world
        kappa = 0.5
        F = kex.kscan(kernel_size=(3, 3), padding=((1, 1), (1, 1)), named_axis={0: "n", 1: "i"}, relative=True)
        F[:, 0] = F[:, -1] = lambda u: 1
        F[:, : int((nx - 1) / 4) + 1] = F[:, int((nx - 1) / 2):] = lambda u: 1
        F[0:1, int((nx - 1) / 4) + 1: int((nx - 1) / 2)] = lambda u: 2
        F[1:, 1:-1] = lambda u: u["i", "n-1"] - kappa * (u["i", "n-1"] - u["i-1", "n-1"])
        sh = ColumnShardedScan(F, nt, nx, rank=rank, world=world, device=dev)
        ms = timed(sh.run, max(args.reps // 2, 3), dev, world)
        emit("c4", total(sh.local_windows), 4, ms, {"nt": nt, "nx_global": nx, "held_columns": sh.width,
                                                    "recomputed_cone_columns": sh.keep_lo})
        del sh
        torch.cuda.empty_cache()
        if world > 1:   # halo-EXCHANGE form: T*R ghost values per neighbour per 64-row block (NCCL send/recv)
            sh = ColumnShardedScan(F, nt, nx, rank=rank, world=world, device=dev, exchange=True)
            ms = timed(sh.run, max(args.reps // 2, 3), dev, world)
            emit("c4_exchange", total(sh.local_windows), 4, ms,
                 {"nt": nt, "nx_global": nx, "held_columns": sh.width, "ghost_columns": sh.ghost_l,
                  "exchanges": (nt - 1) // sh.t_rows, "bytes_per_exchange": sh.ghost_l * 4})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
