"""GPU tests of the host pipeline and of the slab-decomposed multi-GPU path."""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import kx_oracle as ko

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kex():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kernex_b200
    return kernex_b200


def lap(x):
    return x[1, 0] + x[0, -1] - 4 * x[0, 0] + x[0, 1] + x[-1, 0]


@pytest.mark.parametrize("padding", ["valid", "same"])
def test_pinned_host_pipeline_matches_plain_path(kex, padding):
    """>= 64 MB pinned NumPy input goes through hostpipe.py (3-stream slab pipeline)."""
    from kernex_b200 import _lib
    rows, cols = 4200, 4096
    pinned = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True)
    x = pinned.numpy()
    x[...] = np.random.default_rng(0).standard_normal((rows, cols), dtype=np.float32)
    f = kex.kmap(kernel_size=(3, 3), padding=padding, relative=True)(lap)
    before = _lib.launch_count()
    got = f(x)
    assert _lib.launch_count() - before >= 2, "expected several slab launches"
    want = f(torch.from_numpy(x).cuda()).cpu().numpy()   # device-resident path, single launch
    np.testing.assert_array_equal(got, want)
    b = ((0, 0), (0, 0)) if padding == "valid" else ((1, 1), (1, 1))
    ref = ko.affine_map(x[:300], (3, 3), (1, 1), (b[0] if padding == "valid" else (1, 0), b[1]),
                        [(2, 1), (1, 0), (1, 1), (1, 2), (0, 1)], [1, 1, -4, 1, 1])
    n = ref.shape[0] - 2
    np.testing.assert_allclose(got[:n], ref[:n], rtol=0, atol=2e-5)


def test_slab_stencil_world1_equals_kmap(kex):
    from kernex_b200.dist import SlabStencil
    slab = SlabStencil(lap, kernel_size=(3, 3), relative=True, rows_per_rank=96, width=200, padding="same")
    slab.fill_random(torch.Generator(device="cuda").manual_seed(1))
    out = slab.step()
    want = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lap)(slab.ext)
    assert torch.equal(out, want)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_slab_stencil_multi_gpu_bit_identical():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "MISMATCH" not in res.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_tensor_on_a_non_current_device(kex):
    """Several GPUs in ONE process: the library switches to the device that owns the buffers for the call."""
    assert torch.cuda.current_device() == 0
    rng = np.random.default_rng(3)
    xh = rng.standard_normal((300, 516)).astype(np.float32)
    x1 = torch.from_numpy(xh).to("cuda:1")
    f = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lap)
    y1 = f(x1)
    assert y1.device == x1.device and torch.cuda.current_device() == 0
    y0 = f(torch.from_numpy(xh).to("cuda:0"))
    assert torch.equal(y1.cpu(), y0.cpu())
    s1 = kex.kscan(kernel_size=(3, 3), padding="same", relative=True)(lambda u: 0.25 * (u[0, -1] + u[-1, 0] + u[0, 1] + u[1, 0]))
    assert torch.equal(s1(x1).cpu(), s1(torch.from_numpy(xh).to("cuda:0")).cpu())


def lap3(x):
    return (x[1, 0, 0] + x[-1, 0, 0] + x[0, 1, 0] + x[0, -1, 0] + x[0, 0, 1] + x[0, 0, -1] - 6 * x[0, 0, 0])


@pytest.mark.parametrize("world,fn,ksize,rows,trailing,padding", [
    (2, lap, (3, 3), 64, (256,), "valid"), (3, lap, (3, 3), 50, (132,), "same"), (3, lap3, (3, 3, 3), 12, (20, 36), "same"),
    (3, lap3, (3, 3, 3), 16, (24, 64), "valid"), (2, lambda v: v[0, 0] + 2 * v[1, 1], (2, 2), 40, (64,), ((0, 1), (0, 1)))])
def test_peer_halo_pull_with_emulated_ranks_on_one_gpu(kex, world, fn, ksize, rows, trailing, padding):
    """The multi-GPU transport on a 1-GPU box: `world` emulated ranks in this process, every rank with its own streams
    and its own peer buffer (`PeerGroup.local`: the "peer" pointers are local).  Exercises halo_pull_kernel, the
    READY / DONE flag protocol over three steps with new rows in between, and the strip launches; the stitched output
    must equal the single-launch kmap of the whole array bit for bit.  (At most 3 emulated ranks: every rank needs two
    streams that make progress independently, and one process has 8 hardware queues by default -- with more streams than
    queues a spinning pull kernel can sit in front of the kernel it is waiting for.  Real ranks are separate processes.)"""
    from kernex_b200 import _lib, peer
    from kernex_b200.dist import SlabStencil
    dev = torch.device("cuda", 0)
    row_bytes = 4 * int(np.prod(trailing))
    rows_max = rows + (ksize[0] - 1) // 2 + ksize[0] // 2
    groups = peer.PeerGroup.local(world, _lib.KX_PEER_FLAG_BYTES + rows_max * row_bytes, dev)
    mains = [torch.cuda.Stream(device=dev) for _ in range(world)]
    slabs = [SlabStencil(fn, kernel_size=ksize, relative=True, rows_per_rank=rows, trailing_shape=trailing, device=dev,
                         padding=padding, transport="p2p", rank=r, world=world, peers=groups[r]) for r in range(world)]
    assert all(s.transport == "p2p" for s in slabs)
    gen = torch.Generator(device=dev).manual_seed(5)
    f1 = kex.kmap(kernel_size=ksize, padding=padding, relative=True)(fn)
    for trial in range(3):
        glob = torch.randn((world * rows,) + tuple(trailing), device=dev, generator=gen)
        torch.cuda.synchronize()
        for r, s in enumerate(slabs):
            with torch.cuda.stream(mains[r]):
                s.set_local(glob[r * rows:(r + 1) * rows])   # waits for the neighbours' DONE flags of the last step
        outs = []
        for r, s in enumerate(slabs):                        # launches only: rank r's pull spins until r+1's has started
            with torch.cuda.stream(mains[r]):
                outs.append(s.step())
        torch.cuda.synchronize()
        assert _lib.last_kernel() in ("stencil2d_kernel", "stencil3d_star_kernel", "stencil3d_kernel", "halo_pull_kernel")
        want = f1(glob)
        got = torch.cat(outs, dim=0)
        assert got.shape == want.shape and torch.equal(got, want), trial
    for g in groups:
        g.close()


@pytest.mark.parametrize("world,padding", [(2, "valid"), (3, "same")])
def test_peer_pull_add_vjp_with_emulated_ranks_on_one_gpu(kex, world, padding):
    """The adjoint of the halo exchange on a 1-GPU box: every emulated rank accumulates the flipped-tap stencils of its
    cotangent rows into a peer-mapped ext-shaped buffer, then ONE pull kernel in `accumulate` mode adds the rows that
    landed in the neighbours' halos to its edge rows.  Two rounds (the second waits for the neighbours' DONE flags before
    the buffer is rewritten); against autograd of the single-launch kmap on the whole array at the 1e-5 bound."""
    from kernex_b200 import _lib, peer
    from kernex_b200.dist import SlabStencil
    from vjp_ref import assert_within, dx_reference
    dev = torch.device("cuda", 0)
    rows, width, ksize = 40, 132, (3, 3)
    nbytes = _lib.KX_PEER_FLAG_BYTES + (rows + 2) * width * 4
    groups, ggroups = peer.PeerGroup.local(world, nbytes, dev), peer.PeerGroup.local(world, nbytes, dev)
    mains = [torch.cuda.Stream(device=dev) for _ in range(world)]
    slabs = [SlabStencil(lap, kernel_size=ksize, relative=True, rows_per_rank=rows, width=width, device=dev, padding=padding,
                         transport="p2p", rank=r, world=world, peers=groups[r]) for r in range(world)]
    gen = torch.Generator(device=dev).manual_seed(9)
    border = ko.resolve_padding(padding, ksize)
    taps, coefs = [(2, 1), (1, 0), (1, 1), (1, 2), (0, 1)], [1.0, 1.0, -4.0, 1.0, 1.0]
    for trial in range(2):
        glob = torch.randn((world * rows, width), device=dev, generator=gen)
        outs = []
        torch.cuda.synchronize()
        for r, s in enumerate(slabs):
            with torch.cuda.stream(mains[r]):
                s.set_local(glob[r * rows:(r + 1) * rows])
                outs.append(s.step())
        torch.cuda.synchronize()
        g_glob = torch.randn((sum(o.shape[0] for o in outs),) + tuple(outs[0].shape[1:]), device=dev, generator=gen)
        dxs, o0 = [], 0
        torch.cuda.synchronize()
        for r, s in enumerate(slabs):
            with torch.cuda.stream(mains[r]):
                dx, _ = s.vjp(g_glob[o0:o0 + outs[r].shape[0]], grad_peers=ggroups[r])
                dxs.append(dx)
            o0 += outs[r].shape[0]
        torch.cuda.synchronize()
        assert _lib.last_kernel() == "halo_pull_kernel"
        got = torch.cat(dxs, dim=0).cpu().numpy()
        want, mag = dx_reference(g_glob.cpu().numpy(), (world * rows, width), ksize, (1, 1), border, taps, coefs)
        assert_within(got, want, mag, f"sharded dx, trial {trial}")
    for g in groups + ggroups:
        g.close()
