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


# ---- adjoint of the halo exchange (multi-GPU VJP) -------------------------------------
def _adjoint_worker(rank, world, port, rows, width, k0, border0, q):
    from kernex_b200.dist import exchange_halo_grads
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(40 + rank)
        L = SlabLayout(rank, world, rows, k0, border0)
        # <E x, y> == <x, E^T y>: x lives on the own rows, y on the whole ext buffer
        x = torch.zeros((L.ext_rows, width), dtype=torch.float64)
        x[L.own] = torch.from_numpy(rng.standard_normal((rows, width)))
        x_own = x[L.own].clone()
        y = torch.from_numpy(rng.standard_normal((L.ext_rows, width)))
        for r in exchange_halos(x, L):
            r.wait()
        lhs = torch.tensor([float((x * y).sum())], dtype=torch.float64)
        yt = exchange_halo_grads(y.clone(), L)
        rhs = torch.tensor([float((x_own * yt[L.own]).sum())], dtype=torch.float64)
        dist.all_reduce(lhs)
        dist.all_reduce(rhs)
        # full VJP of a slab-decomposed stencil against the global NumPy adjoint
        glob = np.random.default_rng(5).standard_normal((world * rows, width))
        kx = 3
        taps = [(a, b) for a in range(k0) for b in range(kx)]
        coefs = np.linspace(-1, 1, len(taps))
        ext = torch.zeros((L.ext_rows, width), dtype=torch.float64)
        ext[L.own] = torch.from_numpy(glob[rank * rows:(rank + 1) * rows])
        for r in exchange_halos(ext, L):
            r.wait()
        n_out_glob = world * rows + border0[0] + border0[1] - k0 + 1
        g_glob = np.random.default_rng(6).standard_normal((n_out_glob, width))
        counts = [SlabLayout(r, world, rows, k0, border0).out_rows for r in range(world)]
        start = sum(counts[:rank])
        g_loc = g_glob[start:start + counts[rank]]
        dext = np.zeros((L.ext_rows, width))
        for p in L.pieces:                      # dx_ext += stencil^T (g piece), written with explicit loops
            gp = g_loc[p.out0:p.out1]
            for (a, b), c in zip(taps, coefs):
                for j in range(gp.shape[0]):
                    row = j - p.border0[0] + a
                    if 0 <= row < p.in1 - p.in0:
                        lo, hi = max(0, 1 - b), min(width, width + 1 - b)      # out col x reads in col x - 1 + b
                        dext[p.in0 + row, lo - 1 + b:hi - 1 + b] += c * gp[j, lo:hi]
        dl = exchange_halo_grads(torch.from_numpy(dext), L)[L.own].numpy()
        want = np.zeros_like(glob)
        for (a, b), c in zip(taps, coefs):
            for j in range(n_out_glob):
                row = j - border0[0] + a
                if 0 <= row < world * rows:
                    lo, hi = max(0, 1 - b), min(width, width + 1 - b)
                    want[row, lo - 1 + b:hi - 1 + b] += c * g_glob[j, lo:hi]
        ok = bool(np.allclose(dl, want[rank * rows:(rank + 1) * rows], rtol=1e-12, atol=1e-12))
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        if rank == 0:
            q.put(bool(abs(float(lhs) - float(rhs)) <= 1e-9 * max(1.0, abs(float(lhs)))) and all(flags))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k0,border0", [(2, 3, (0, 0)), (3, 3, (1, 1)), (2, 5, (2, 2)), (3, 2, (0, 1)), (2, 4, (1, 2))])
def test_halo_exchange_adjoint_and_sharded_vjp(world, k0, border0):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_adjoint_worker, args=(r, world, port, 7, 12, k0, border0, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_column_exchange_plan_partitions_the_grid():
    """kscan halo-exchange shards (axis 1): own columns tile [0, nx) exactly, ghost strips are one temporal
    block of the dependency cone (rounded to 4 columns) and absent at the global edges."""
    from kernex_b200.dist import column_exchange_plan, column_shard_plan
    for nx, world, t_rows, rl, rr in [(8192, 2, 64, 1, 0), (5000, 3, 64, 1, 1), (1 << 21, 8, 64, 1, 0), (4099, 4, 64, 2, 2)]:
        covered = 0
        for rank in range(world):
            base, width, k0, k1, gl, gr = column_exchange_plan(nx, world, rank, t_rows, rl, rr)
            assert base + k0 == covered                                   # own columns are contiguous over the ranks
            covered += k1 - k0
            assert gl == (0 if rank == 0 else -(-t_rows * rl // 4) * 4) and k0 == gl
            assert base + width <= nx and width == k1 + gr
            if rank == world - 1:
                assert gr == 0
            b2, w2, a0, a1 = column_shard_plan(nx, world, rank, 4096, rl, rr)  # recompute form holds the whole cone
            assert w2 >= width - gl - gr and b2 <= base + k0
        assert covered == nx
