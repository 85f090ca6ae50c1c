"""world_size-2/3 gloo tests (CPU) of the slab decomposition + halo exchange used by the
multi-GPU path.  Compute is done by the oracle here; on GPUs the same layout drives the
CUDA engine (tests/test_gpu_dist.py)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kernex_b200.dist import SlabLayout, exchange_halos
from oracle import kx_oracle as ko

TAPS = [(0, 1), (1, 0), (1, 1), (1, 2), (2, 1)]
COEFS = [1.0, 1.0, -4.0, 1.0, 1.0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rows, width, k0, border0, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        glob = rng.standard_normal((world * rows, width)).astype(np.float32)
        L = SlabLayout(rank, world, rows, k0, border0)
        ext = torch.zeros((L.ext_rows, width), dtype=torch.float32)
        ext[L.own] = torch.from_numpy(glob[rank * rows:(rank + 1) * rows])
        for r in exchange_halos(ext, L):
            r.wait()
        kx = 3
        taps = [(a, b) for a in range(k0) for b in range(kx)]
        coefs = list(np.linspace(-1, 1, len(taps)))
        outs = []
        for p in L.pieces:
            outs.append(ko.affine_map(ext[p.in0:p.in1].numpy(), (k0, kx), (1, 1), (p.border0, (1, 1)), taps, coefs))
            assert outs[-1].shape[0] == p.out1 - p.out0
        local = np.concatenate(outs, axis=0)
        assert local.shape[0] == L.out_rows
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        if rank == 0:
            want = ko.affine_map(glob, (k0, kx), (1, 1), (border0, (1, 1)), taps, coefs)
            got = np.concatenate(gathered, axis=0)
            q.put((got.shape == want.shape) and bool(np.array_equal(got, want)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k0,border0", [(2, 3, (0, 0)), (2, 3, (1, 1)), (3, 3, (1, 1)), (2, 2, (0, 1)),
                                              (3, 5, (2, 2)), (2, 1, (0, 0)), (2, 4, (1, 2))])
def test_slab_decomposition_matches_global_oracle(world, k0, border0):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 7, 12, k0, border0, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_layout_row_accounting():
    for world in (1, 2, 4, 8):
        for k0, b in ((3, (0, 0)), (3, (1, 1)), (5, (2, 2)), (2, (0, 1))):
            total = sum(SlabLayout(r, world, 16, k0, b).out_rows for r in range(world))
            assert total == world * 16 + b[0] + b[1] - k0 + 1
