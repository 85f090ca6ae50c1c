"""Run under torchrun on >= 2 GPUs: the slab-decomposed stencil must reproduce, bit for bit,
the single-GPU kmap of the gathered global array (same kernel, same accumulation order).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def lap(x):
    return x[1, 0] + x[0, -1] - 4 * x[0, 0] + x[0, 1] + x[-1, 0]


def lap3(x):
    return (x[1, 0, 0] + x[-1, 0, 0] + x[0, 1, 0] + x[0, -1, 0] + x[0, 0, 1] + x[0, 0, -1] - 6 * x[0, 0, 0])


def main():
    import kernex_b200 as kex
    from kernex_b200.dist import SlabStencil

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for (fn, ksize, rows, trailing, padding) in [
        (lap, (3, 3), 64, (256,), "valid"), (lap, (3, 3), 50, (130,), "same"),
        (lap3, (3, 3, 3), 16, (24, 64), "valid"), (lap3, (3, 3, 3), 12, (20, 36), "same"),
    ]:
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        slab = SlabStencil(fn, kernel_size=ksize, relative=True, rows_per_rank=rows, trailing_shape=trailing,
                           device=dev, padding=padding, transport="nccl")
        slab.fill_random(g)
        for _ in range(2):  # a second step must give the same answer (halos re-exchanged)
            out = slab.step()
        torch.cuda.synchronize()
        # the default transport on several GPUs: halo rows pulled out of the neighbours' peer-mapped slabs by one kernel
        p2p = SlabStencil(fn, kernel_size=ksize, relative=True, rows_per_rank=rows, trailing_shape=trailing,
                          device=dev, padding=padding)
        row_bytes = 4 * int(torch.tensor(trailing).prod())
        expect = "p2p" if row_bytes % 16 == 0 else "nccl"
        assert p2p.transport == expect, (p2p.transport, p2p.transport_note)
        p2p.set_local(slab.ext[slab.layout.own])
        for it in range(3):
            out2 = p2p.step()
            if it == 1:   # new rows between steps: the neighbours' DONE flags gate the overwrite, the next pull sees them
                p2p.set_local(slab.ext[slab.layout.own])
        torch.cuda.synchronize()
        same2 = torch.tensor([1 if torch.equal(out2, out) else 0], device=dev)
        dist.all_reduce(same2, op=dist.ReduceOp.MIN)
        ok = ok and bool(same2.item())
        if rank == 0:
            print(f"  {p2p.transport} transport vs nccl: {'bit-identical' if same2.item() else 'MISMATCH'}", flush=True)
        # end-to-end form: my rows in pinned host memory on both sides, streamed in small chunks (several per slab)
        xh = torch.empty(tuple(slab.ext[slab.layout.own].shape), dtype=torch.float32, pin_memory=True)
        xh.copy_(slab.ext[slab.layout.own])
        yh = torch.empty(tuple(out.shape), dtype=torch.float32, pin_memory=True)
        for it in range(2):
            yh.zero_()
            p2p.step_host(xh, yh, chunk_bytes=7 * xh[0].numel() * 4)
        same3 = torch.tensor([1 if torch.equal(yh, out.cpu()) else 0], device=dev)
        dist.all_reduce(same3, op=dist.ReduceOp.MIN)
        ok = ok and bool(same3.item())
        if rank == 0:
            print(f"  step_host (pinned host rows, chunked) vs device step: {'bit-identical' if same3.item() else 'MISMATCH'}",
                  flush=True)
        own = slab.ext[slab.layout.own].contiguous()
        parts = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(parts, own)
        glob = torch.cat(parts, dim=0)
        want = kex.kmap(kernel_size=ksize, padding=padding, relative=True)(fn)(glob)
        counts = torch.tensor([out.shape[0]], device=dev)
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        start = int(sum(int(c.item()) for c in allc[:rank]))
        mine = want[start:start + out.shape[0]]
        same = bool(torch.equal(mine, out)) and int(sum(int(c.item()) for c in allc)) == want.shape[0]
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        if rank == 0:
            print(f"{fn.__name__} {ksize} rows={rows} trailing={trailing} padding={padding}: "
                  f"{'bit-identical' if flag.item() else 'MISMATCH'}", flush=True)
    # ---- adjoint: dx through the transposed halo exchange, dw through one all-reduce ----
    import numpy as np

    def conv(x, w):
        return np.sum(x * w)

    for (fn, ksize, rows, trailing, padding, wshape) in [
        (lap, (3, 3), 48, (256,), "same", None), (lap, (3, 3), 40, (132,), "valid", None),
        (conv, (3, 3), 64, (260,), "same", (3, 3)), (conv, (3, 3, 3), 16, (24, 64), "valid", (3, 3, 3)),
    ]:
        gen = torch.Generator(device=dev).manual_seed(300 + rank)
        w = None
        if wshape is not None:
            w = torch.randn(wshape, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        slab = SlabStencil(fn, kernel_size=ksize, relative=wshape is None, rows_per_rank=rows, trailing_shape=trailing,
                           device=dev, padding=padding, args=() if w is None else (w,))
        slab.fill_random(gen)
        out = slab.step()
        g_local = torch.randn(out.shape, device=dev, generator=gen)
        dx, dw = slab.vjp(g_local, want_input=True, want_coef=w is not None)
        torch.cuda.synchronize()
        own = slab.ext[slab.layout.own].contiguous()
        parts = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(parts, own)
        glob = torch.cat(parts, dim=0).requires_grad_()
        counts = torch.tensor([out.shape[0]], device=dev)
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        gparts = []
        for r in range(world):
            buf = torch.empty((int(allc[r].item()),) + tuple(out.shape[1:]), device=dev)
            if r == rank:
                buf.copy_(g_local)
            dist.broadcast(buf, src=r)
            gparts.append(buf)
        g_glob = torch.cat(gparts, dim=0)
        wg = None if w is None else w.clone().requires_grad_()
        f1 = kex.kmap(kernel_size=ksize, padding=padding, relative=wshape is None)(fn)
        y = f1(glob) if w is None else f1(glob, wg)
        grads = torch.autograd.grad(y, [glob] if w is None else [glob, wg], g_glob)
        want_dx = grads[0][rank * rows:(rank + 1) * rows]
        scale = float(g_glob.abs().max()) * 30
        same = bool(torch.allclose(dx, want_dx, rtol=1e-5, atol=1e-5 * scale))
        if w is not None:
            tol = 1e-4 * float(np.sqrt(g_glob.numel())) * 4
            same = same and bool(torch.allclose(dw.reshape(wshape), grads[1], rtol=1e-4, atol=tol))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        if rank == 0:
            print(f"VJP {fn.__name__} {ksize} rows={rows} padding={padding}: {'ok' if flag.item() else 'MISMATCH'}",
                  flush=True)
    # ---- kscan, axis-1 shards: halo-EXCHANGE form vs the single-GPU scan (bit-identical) ----
    from kernex_b200.dist import ColumnShardedScan

    def convection_mesh(nt, nx, kappa, two_sided):
        F = kex.kscan(kernel_size=(3, 3), padding=((1, 1), (1, 1)), named_axis={0: "n", 1: "i"}, relative=True)
        q1, q2 = int((nx - 1) / 4) + 1, int((nx - 1) / 2)
        F[:, 0] = F[:, -1] = lambda u: 1
        F[:, :q1] = F[:, q2:] = lambda u: 1
        F[0:1, q1:q2] = lambda u: 2
        if two_sided:   # diffusion-like update: reads both neighbours of the row above (RL = RR = 1)
            F[1:, 1:-1] = lambda u: u["i", "n-1"] + 0.25 * (u["i-1", "n-1"] - 2 * u["i", "n-1"] + u["i+1", "n-1"])
        else:
            F[1:, 1:-1] = lambda u: u["i", "n-1"] - kappa * (u["i", "n-1"] - u["i-1", "n-1"])
        return F

    for (nt, nx, two_sided) in [(200, 8192, False), (130, 5000, False), (150, 8192, True)]:
        whole = convection_mesh(nt, nx, 0.5, two_sided)(torch.ones((nt, nx), device=dev))
        same = True
        for exchange in (True, False):
            sh = ColumnShardedScan(convection_mesh(nt, nx, 0.5, two_sided), nt, nx, rank=rank, world=world, device=dev,
                                   exchange=exchange)
            mine = sh.run()
            torch.cuda.synchronize()
            c0 = sh.col_base + sh.keep_lo
            same = same and bool(torch.equal(mine, whole[:, c0:c0 + mine.shape[1]]))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        if rank == 0:
            print(f"kscan nt={nt} nx={nx} two_sided={two_sided} exchange + recompute: "
                  f"{'bit-identical' if flag.item() else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
