"""GPU parity of the multi-function kmap mesh kernel (kx_stencil2d_mesh.cu: `F[slice] = fn` region
assignment with affine functions) against the literal oracle (region selection of
kernex/_src/utils.py:338-382) on small meshes, and against the generic kernel on large ones."""
from __future__ import annotations

import itertools

import numpy as np
import pytest

from oracle import kx_oracle as ko

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
INF = float("inf")


@pytest.fixture(scope="module")
def kex():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kernex_b200
    return kernex_b200


def _last_kernel():
    from kernex_b200 import _lib
    return _lib.last_kernel()


def _fn(taps, coefs, bias, integer):
    def fn(x):
        acc = int(bias) if integer else float(bias)
        for tp, c in zip(taps, coefs):
            acc = acc + (int(c) if integer else float(c)) * x[tp]
        return acc
    return fn


def _random_key_item(rng, n):
    kind = int(rng.integers(0, 5))
    if kind == 0:
        return int(rng.integers(0, n))
    if kind == 1:
        a = int(rng.integers(0, n))
        return (a, int(rng.integers(a, n + 1)))
    if kind == 2:
        a = int(rng.integers(0, n))
        return (a, int(rng.integers(a, n + 1)), int(rng.integers(1, 4)))
    if kind == 3:
        return (-INF, INF)
    return (int(rng.integers(0, n)), INF)


@pytest.mark.parametrize("seed", range(30))
def test_random_meshes_match_literal_oracle(kex, seed):
    rng = np.random.default_rng(12000 + seed)
    nd = 1 if seed % 5 == 4 else 2
    ksize = (3,) if nd == 1 else [(3, 3), (1, 3), (1, 1)][seed % 3]
    shape = (int(rng.integers(20, 200)),) if nd == 1 else (int(rng.integers(8, 40)), 4 * int(rng.integers(3, 40)) + (seed % 2))
    border = tuple((int(rng.integers(0, 3)), int(rng.integers(0, 3))) for _ in range(nd)) if seed % 3 else \
        tuple(((k - 1) // 2, k // 2) for k in ksize)
    if any(n + lo + hi < k for n, k, (lo, hi) in zip(shape, ksize, border)):
        pytest.skip("empty output")
    integer = seed % 2 == 0
    nf = 2 + seed % 5          # up to 6 functions
    all_taps = list(itertools.product(*[range(k) for k in ksize]))
    fmap = {}
    for q in range(nf):
        taps = [all_taps[i] for i in rng.choice(len(all_taps), min(len(all_taps), int(rng.integers(1, 5))), replace=False)]
        if integer:
            f = _fn(taps, rng.integers(-3, 4, len(taps)), int(rng.integers(-2, 3)), True)
        else:
            f = _fn(taps, rng.standard_normal(len(taps)), float(rng.standard_normal()), False)
        keys = [[]] if q == 0 else [[_random_key_item(rng, shape[d]) for d in range(int(rng.integers(1, nd + 1)))]
                                    for _ in range(int(rng.integers(1, 3)))]
        fmap[f] = keys
    x = rng.integers(-9, 10, shape).astype(np.int32) if integer else rng.standard_normal(shape).astype(np.float32)
    want = ko.kernel_map(fmap, shape, ksize, (1,) * nd, border, False)(x)
    got = kex.kernel_map(fmap, shape, ksize, (1,) * nd, border, False)(x)
    assert _last_kernel() == "stencil2d_mesh_kernel"
    assert got.shape == want.shape and got.dtype == want.dtype
    if integer:
        np.testing.assert_array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


def test_interface_mesh_large_matches_generic_kernel(kex, monkeypatch):
    """Interior Laplacian + identity edges + a source block on 2000 x 4100 (several bands and column
    blocks, region boundaries inside thread quads), and a batched mesh whose keys constrain the batch axis."""
    rng = np.random.default_rng(5)
    n0, n1 = 2000, 4100
    x = torch.from_numpy(rng.standard_normal((n0, n1)).astype(np.float32)).cuda()
    F = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)
    F[:, :] = lambda v: v[0, 0]
    F[1:-1, 1:-1] = lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0]
    F[501:777, 1001:2003] = lambda v: 0.5 * v[0, 0] + 1.0
    F[::7, 3::5] = lambda v: v[0, 1] - v[0, -1]
    got = F(x)
    assert _last_kernel() == "stencil2d_mesh_kernel"
    monkeypatch.setenv("KX_NO_MESH", "1")
    monkeypatch.setenv("KX_NO_DECOMPOSE", "1")      # ONE generic launch over everything
    want = F(x)
    assert _last_kernel() == "generic_map_kernel"
    monkeypatch.delenv("KX_NO_MESH")
    monkeypatch.delenv("KX_NO_DECOMPOSE")
    # same taps and coefficients; the generic kernel accumulates in expression order, this one in window order
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    # independent check of the three rectangular regions
    xh = x.cpu().numpy()
    g = got.cpu().numpy()
    np.testing.assert_array_equal(g[0, 1::5][:3], xh[0, 1::5][:3])                     # identity edge
    lap = xh[2:, 1:-1] + xh[1:-1, :-2] - 4 * xh[1:-1, 1:-1] + xh[1:-1, 2:] + xh[:-2, 1:-1]
    np.testing.assert_allclose(g[1:-1, 1:-1][3, 4], lap[3, 4], rtol=1e-5, atol=1e-5)  # row 4: not a ::7 row, col 5
    # batched: keys on the leading axis
    xb = torch.from_numpy(rng.integers(-5, 6, (3, 40, 64)).astype(np.int32)).cuda()
    f0, f1 = (lambda v: v[0, 1, 1]), (lambda v: 2 * v[0, 1, 1] - v[0, 1, 0])
    keys = {f0: [[]], f1: [[1, (5, 30), (8, 50, 3)]]}
    gb = kex.kernel_map(keys, (3, 40, 64), (1, 3, 3), (1, 1, 1), ((0, 0), (1, 1), (1, 1)), False)(xb)
    assert _last_kernel() == "stencil2d_mesh_kernel"
    wb = ko.kernel_map(keys, (3, 40, 64), (1, 3, 3), (1, 1, 1), ((0, 0), (1, 1), (1, 1)), False)(xb.cpu().numpy())
    np.testing.assert_array_equal(gb.cpu().numpy(), wb)


@pytest.mark.parametrize("dtype", ["f32", "i32"])
def test_two_pass_mesh_equals_one_pass_bit_for_bit(kex, dtype, monkeypatch):
    """stencil2d_mesh_kernel runs the plain stencil loop on tiles ONE function covers and the per-output selection loop on
    tiles a region boundary cuts.  Same accumulation order in both, so the result must equal the run with every tile on
    the general loop (KX_MESH_NO_UNIFORM=1) bit for bit -- with rectangular regions, regions narrower than a tile or
    a thread's 4 columns, and strided keys (every 64th row of tiles mixed)."""
    from kernex_b200 import _lib
    n = 8192
    rng = np.random.default_rng(77)
    if dtype == "f32":
        x = torch.from_numpy(rng.standard_normal((n, n)).astype(np.float32)).cuda()
        src = lambda v: 0.5 * v[0, 0] + 1.0            # noqa: E731
    else:
        x = torch.from_numpy(rng.integers(-1000, 1000, (n, n)).astype(np.int32)).cuda()
        src = lambda v: 3 * v[0, 0] + 1                # noqa: E731

    def mesh(strided):
        F = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)
        F[:, :] = lambda v: v[0, 0]
        F[1:-1, 1:-1] = lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0]
        F[n // 4: n // 2, n // 4: n // 2] = src
        F[4000:4003, 100:7000] = lambda v: v[0, 1] - v[0, -1]        # thinner than a band
        F[6000:6500, 5000:5003] = lambda v: 2 * v[-1, 0]             # thinner than a thread's 4 columns
        if strided:
            F[::64, ::2] = lambda v: v[1, 1]                         # cuts every 64th row of tiles
        return F

    for strided in (False, True):
        F = mesh(strided)
        got = F(x)
        assert _last_kernel() == "stencil2d_mesh_kernel"
        monkeypatch.setenv("KX_MESH_NO_UNIFORM", "1")
        want = mesh(strided)(x)
        monkeypatch.delenv("KX_MESH_NO_UNIFORM")
        assert torch.equal(got, want)
        # spot check against the literal oracle on a corner block and on the source-block corner
        for (r0, c0) in ((0, 0), (n // 4 - 8, n // 4 - 8), (3996, 96), (n - 24, n - 40)):
            blk = x[max(r0 - 1, 0):r0 + 17, max(c0 - 1, 0):c0 + 33].cpu().numpy()
            g = got[r0:r0 + 16, c0:c0 + 32].cpu().numpy()
            xh = np.zeros((18, 34), blk.dtype)
            xh[(1 if r0 == 0 else 0):(1 if r0 == 0 else 0) + blk.shape[0], (1 if c0 == 0 else 0):(1 if c0 == 0 else 0) + blk.shape[1]] = blk
            lap = xh[2:, 1:-1] + xh[1:-1, :-2] - 4 * xh[1:-1, 1:-1] + xh[1:-1, 2:] + xh[:-2, 1:-1]
            rr, cc = np.meshgrid(np.arange(r0, r0 + 16), np.arange(c0, c0 + 32), indexing="ij")
            interior = (rr >= 1) & (rr < n - 1) & (cc >= 1) & (cc < n - 1)
            plain = interior & ~((rr >= n // 4) & (rr < n // 2) & (cc >= n // 4) & (cc < n // 2)) \
                & ~((rr >= 4000) & (rr < 4003) & (cc >= 100) & (cc < 7000)) & ~((rr >= 6000) & (rr < 6500) & (cc >= 5000) & (cc < 5003))
            if strided:
                plain &= ~((rr % 64 == 0) & (cc % 2 == 0))
            if dtype == "i32":
                np.testing.assert_array_equal(g[plain], lap[plain])
            else:
                np.testing.assert_allclose(g[plain], lap[plain], rtol=1e-5, atol=1e-5)
            edge = ~interior
            if strided:
                edge &= ~((rr % 64 == 0) & (cc % 2 == 0))
            np.testing.assert_array_equal(g[edge], xh[1:-1, 1:-1][edge])


@pytest.mark.parametrize("seed", range(6))
def test_mesh_block_order_covers_every_tile(kex, seed, monkeypatch):
    """stencil2d_mesh_kernel dispatches the tile columns / row bands a region edge runs through FIRST (linear grid, the
    rest follows band by band).  Whatever the edges -- none, a few, more than the priority tables hold (17+ distinct
    tile columns: natural order) -- every tile must be computed exactly once: the result equals the generic kernel's
    and the run in natural order (KX_MESH_NO_PRIO=1) bit for bit."""
    rng = np.random.default_rng(900 + seed)
    n0, n1 = int(rng.integers(100, 700)), 4 * int(rng.integers(300, 1400))
    x = torch.from_numpy(rng.integers(-50, 50, (n0, n1)).astype(np.int32)).cuda()
    F = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)
    F[:, :] = lambda v: v[0, 0]
    F[1:-1, 1:-1] = lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0]
    fns = [lambda v: 2 * v[0, 0] + 1, lambda v: v[0, 1] - v[0, -1], lambda v: 3 * v[-1, 0], lambda v: v[1, 1] + 7]
    n_rect = [0, 1, 3, 9, 12, 2][seed]          # 9+ rectangles: more than 16 distinct edge columns are likely
    for i in range(n_rect):
        r0, c0 = int(rng.integers(0, n0 - 2)), int(rng.integers(0, n1 - 2))
        r1, c1 = int(rng.integers(r0 + 1, n0)), int(rng.integers(c0 + 1, n1))
        F[r0:r1, c0:c1] = fns[i % len(fns)]
    got = F(x)
    assert _last_kernel() == "stencil2d_mesh_kernel"
    monkeypatch.setenv("KX_MESH_NO_PRIO", "1")
    natural = F(x)
    monkeypatch.delenv("KX_MESH_NO_PRIO")
    monkeypatch.setenv("KX_NO_MESH", "1")
    monkeypatch.setenv("KX_NO_DECOMPOSE", "1")      # ONE generic launch over everything
    want = F(x)
    assert _last_kernel() == "generic_map_kernel"
    monkeypatch.delenv("KX_NO_MESH")
    monkeypatch.delenv("KX_NO_DECOMPOSE")
    assert torch.equal(got, natural)
    assert torch.equal(got, want)
