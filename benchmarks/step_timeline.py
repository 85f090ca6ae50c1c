#!/usr/bin/env python
"""Device-side timeline of SlabStencil.step (config 2 slabs, 16384 x 16384 per rank) on N GPUs, per rank, from CUDA
events on the two streams of a step (nsys is not installed in this image):

    t_halo   side stream: start of the step -> halo rows have arrived (NCCL send/recv done, or halo_pull_kernel done)
    t_strips side stream: halo rows arrived -> both strip launches finished
    t_inter  main stream: interior launch start -> end
    t_step   main stream: start of the step -> `main` has joined the side stream (= what bench.py times per step)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 \
        benchmarks/step_timeline.py [--transport p2p|nccl] [--steps 100]

Prints one JSON line (rank 0): per-rank medians in microseconds, plus the back-to-back step time of an untimed-phase
run (the instrumentation adds events, so the two are reported separately)."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = 16384


def laplacian(x):
    return x[1, 0] + x[0, -1] - 4 * x[0, 0] + x[0, 1] + x[-1, 0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transport", default=None)
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from kernex_b200 import dist as kd

    slab = kd.SlabStencil(laplacian, kernel_size=(3, 3), relative=True, rows_per_rank=N, width=N, device=dev,
                          transport=args.transport)
    slab.fill_random(torch.Generator(device=dev).manual_seed(rank))
    for _ in range(10):
        slab.step()
    dist.barrier()
    torch.cuda.synchronize()

    # plain back-to-back steps (what bench.py measures)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        slab.step()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    plain_us = e0.elapsed_time(e1) / args.steps * 1e3

    # instrumented steps: the body of SlabStencil.step with events between its phases
    main_s = torch.cuda.current_stream(dev)
    rows = []
    for _ in range(args.steps):
        ev = {k: torch.cuda.Event(enable_timing=True) for k in ("s0", "halo", "strips", "i0", "i1", "end")}
        ev["s0"].record(main_s)
        slab.side.wait_stream(main_s)
        reqs = []
        if slab._peers is not None:
            slab._pull_halos(slab.side)
        else:
            with torch.cuda.stream(slab.side):
                reqs = kd.exchange_halos(slab.ext, slab.layout, slab.group)
        ev["i0"].record(main_s)
        for p, pm in slab.maps:
            if not p.needs_halo:
                pm(slab.ext[p.in0:p.in1], out=slab.out[p.out0:p.out1])
        ev["i1"].record(main_s)
        with torch.cuda.stream(slab.side):
            for r in reqs:
                r.wait()
            ev["halo"].record(slab.side)
            for p, pm in slab.maps:
                if p.needs_halo:
                    pm(slab.ext[p.in0:p.in1], out=slab.out[p.out0:p.out1])
            ev["strips"].record(slab.side)
        main_s.wait_stream(slab.side)
        ev["end"].record(main_s)
        rows.append(ev)
    dist.barrier()
    torch.cuda.synchronize()
    med = lambda a, b: statistics.median(r[a].elapsed_time(r[b]) for r in rows) * 1e3
    mine = {"rank": rank, "t_halo_us": med("s0", "halo"), "t_strips_us": med("halo", "strips"), "t_inter_us": med("i0", "i1"),
            "t_step_us": med("s0", "end"), "halo_done_after_interior_end_us": med("i1", "halo"), "plain_step_us": plain_us}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        print(json.dumps({"experiment": "SlabStencil.step timeline, C2 slabs", "n_gpus": world, "transport": slab.transport,
                          "note": slab.transport_note, "ranks": allr,
                          "max_plain_step_us": max(r["plain_step_us"] for r in allr)}), flush=True)
    dist.barrier()
    slab.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
