"""Multi-function meshes outside the 2-D mesh kernel (3-D windows, window reductions, strides, wide windows): the C
dispatcher runs the bulk function as a single-function map on the fast templates and recomputes the boxes of the other
functions' keys with the generic kernel (kx_abi.cu try_mesh_decompose).  Checked against the literal oracle (region
selection of kernex/_src/utils.py:338-382) and against the plain generic kernel (KX_NO_DECOMPOSE=1)."""
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


def _launches():
    from kernex_b200 import _lib
    return _lib.launch_count()


def _fn(taps, coefs, bias, integer):
    def fn(x):
        acc = int(bias) if integer else float(bias)
        for tp, c in zip(taps, coefs):
            acc = acc + (int(c) if integer else float(c)) * x[tp]
        return acc
    return fn


def _rect_item(rng, n):
    kind = int(rng.integers(0, 4))
    if kind == 0:
        return int(rng.integers(0, n))
    if kind == 1:
        a = int(rng.integers(0, n))
        return (a, int(rng.integers(a, n + 1)))
    if kind == 2:
        return (-INF, INF)
    return (int(rng.integers(0, n)), INF)


@pytest.mark.parametrize("seed", range(24))
def test_random_box_meshes_match_literal_oracle(kex, seed, monkeypatch):
    """Random rectangular regions (ints, ranges, wildcards, open ranges, partial keys), 3-D and strided / wide 2-D
    geometry, up to 6 functions, function 0 with or without a match-all key; int32 bit exact, fp32 at 1e-5."""
    rng = np.random.default_rng(31000 + seed)
    if seed % 3 == 0:
        ksize, shape = (3, 3, 3), (int(rng.integers(6, 14)), int(rng.integers(6, 20)), 4 * int(rng.integers(2, 12)))
        strides = (1, 1, 1)
    elif seed % 3 == 1:
        ksize, shape = (2, 3, 3), (int(rng.integers(5, 12)), int(rng.integers(6, 20)), 4 * int(rng.integers(2, 12)) + 1)
        strides = (1, 2, 1)
    else:
        ksize, shape = (5, 5), (int(rng.integers(12, 40)), 4 * int(rng.integers(4, 30)))
        strides = (1, 1) if seed % 2 else (2, 2)
    nd = len(shape)
    border = tuple(((k - 1) // 2, k // 2) for k in ksize) if seed % 4 else tuple((int(rng.integers(0, 2)), int(rng.integers(0, 3))) for _ in ksize)
    integer = seed % 2 == 0
    nf = 2 + seed % 5
    all_taps = list(itertools.product(*[range(k) for k in ksize]))
    fmap = {}
    for q in range(nf):
        taps = [all_taps[i] for i in rng.choice(len(all_taps), int(rng.integers(1, 6)), replace=False)]
        f = _fn(taps, rng.integers(-3, 4, len(taps)), int(rng.integers(-2, 3)), True) if integer else \
            _fn(taps, rng.standard_normal(len(taps)), float(rng.standard_normal()), False)
        if q == 0 and seed % 7 != 6:
            keys = [[]]                                      # match-all: the usual `F[:, :] = f0` first
        else:
            keys = [[_rect_item(rng, shape[d]) for d in range(int(rng.integers(1, nd + 1)))] for _ in range(int(rng.integers(1, 3)))]
        fmap[f] = keys
    x = rng.integers(-9, 10, shape).astype(np.int32) if integer else rng.standard_normal(shape).astype(np.float32)
    want = ko.kernel_map(fmap, shape, ksize, strides, border, False)(x)
    got = kex.kernel_map(fmap, shape, ksize, strides, border, False)(x)
    monkeypatch.setenv("KX_NO_DECOMPOSE", "1")
    plain = kex.kernel_map(fmap, shape, ksize, strides, border, False)(x)
    assert got.shape == want.shape and got.dtype == want.dtype
    if integer:
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(plain, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(plain, want, rtol=1e-5, atol=1e-5)


def test_heat_equation_3d_with_dirichlet_faces(kex, monkeypatch):
    """The README's 2-D mesh example in 3-D: interior 7-point update, six Dirichlet faces, a source block.  The interior
    runs on stencil3d_star_kernel, the faces and the block on the generic kernel: several launches, same values as one
    generic launch over everything (fp32: the star kernel adds in window order, the generic kernel in expression order)."""
    from kernex_b200 import _lib
    rng = np.random.default_rng(8)
    n0, n1, n2 = 96, 120, 256
    x = torch.from_numpy(rng.standard_normal((n0, n1, n2)).astype(np.float32)).cuda()

    def step(v):
        return v[0, 0, 0] + 0.1 * (v[1, 0, 0] + v[-1, 0, 0] + v[0, 1, 0] + v[0, -1, 0] + v[0, 0, 1] + v[0, 0, -1] - 6 * v[0, 0, 0])

    F = kex.kmap(kernel_size=(3, 3, 3), padding="same", relative=True)
    F[:, :, :] = lambda v: 1.0                       # Dirichlet faces
    F[1:-1, 1:-1, 1:-1] = step
    F[40:50, 30:60, 100:140] = lambda v: v[0, 0, 0] + 2.0
    before = _lib.launch_count()
    got = F(x)
    assert _lib.launch_count() - before >= 2         # bulk launch + boxes
    monkeypatch.setenv("KX_NO_DECOMPOSE", "1")
    before = _lib.launch_count()
    want = F(x)
    assert _lib.launch_count() - before == 1
    assert _lib.last_kernel() == "generic_map_kernel"
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    g, xh = got.cpu().numpy(), x.cpu().numpy()
    assert np.all(g[0] == 1.0) and np.all(g[-1] == 1.0) and np.all(g[:, 0] == 1.0) and np.all(g[:, :, -1] == 1.0)
    np.testing.assert_array_equal(g[40:50, 30:60, 100:140], xh[40:50, 30:60, 100:140] + np.float32(2.0))
    lap = (xh[2:, 1:-1, 1:-1] + xh[:-2, 1:-1, 1:-1] + xh[1:-1, 2:, 1:-1] + xh[1:-1, :-2, 1:-1] + xh[1:-1, 1:-1, 2:]
           + xh[1:-1, 1:-1, :-2] - 6 * xh[1:-1, 1:-1, 1:-1])
    ref = xh[1:-1, 1:-1, 1:-1] + np.float32(0.1) * lap
    np.testing.assert_allclose(g[1:-1, 1:-1, 1:-1][5, 7, 9:60], ref[5, 7, 9:60], rtol=1e-5, atol=1e-5)


def test_reduction_mesh_2d(kex, monkeypatch):
    """Window reductions are not affine, so the 2-D mesh kernel declines: max filter inside, identity on the frame."""
    rng = np.random.default_rng(9)
    x = torch.from_numpy(rng.standard_normal((300, 1028)).astype(np.float32)).cuda()
    F = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)
    F[:, :] = lambda v: v[0, 0]
    F[1:-1, 1:-1] = lambda v: np.max(v)
    got = F(x).cpu().numpy()
    xh = x.cpu().numpy()
    want = xh.copy()
    win = np.lib.stride_tricks.sliding_window_view(xh, (3, 3))
    want[1:-1, 1:-1] = win.max(axis=(2, 3))
    np.testing.assert_array_equal(got, want)
