"""GPU parity of the offset variants (smap / offset_kernel_map, kernex/_src/map.py:127-156):
the fused single-pass path (stride 1, affine: every cell written by the stencil kernel, cells
outside the interior keep their input) and the copy + strided write-back path, vs the oracle."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import kx_oracle as ko

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def kex():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kernex_b200
    return kernex_b200


def _last_kernel():
    from kernex_b200 import _lib
    return _lib.last_kernel()


@pytest.mark.parametrize("seed", range(16))
def test_offset_map_matches_oracle(kex, seed):
    rng = np.random.default_rng(8100 + seed)
    nd = 1 + seed % 2
    ksize = tuple(int(rng.integers(1, 4)) for _ in range(nd)) if seed % 4 else (3,) * nd
    shape = (int(rng.integers(20, 90)), 4 * int(rng.integers(8, 80)))[-nd:]
    offset = tuple((int(rng.integers(0, 3)), int(rng.integers(0, 3))) for _ in range(nd))
    strides = (1,) * nd if seed % 3 else tuple(int(rng.integers(1, 3)) for _ in range(nd))
    integer = seed % 2 == 0
    taps = [tuple(int(rng.integers(0, k)) for k in ksize) for _ in range(3)]
    taps = list(dict.fromkeys(taps))
    if integer:
        x = rng.integers(-50, 50, shape).astype(np.int32)
        coefs, bias = [int(c) or 2 for c in rng.integers(-4, 5, len(taps))], 3
    else:
        x = rng.standard_normal(shape).astype(np.float32)
        coefs, bias = [float(c) for c in rng.standard_normal(len(taps))], 0.25

    def fn(v):
        acc = bias
        for t, c in zip(taps, coefs):
            acc = acc + c * v[t]
        return acc

    want = ko.offset_kernel_map({fn: ()}, shape, ksize, strides, offset, False)(x)
    got = kex.offset_kernel_map({fn: ()}, shape, ksize, strides, offset, False)(x)
    if all(s == 1 for s in strides):
        assert _last_kernel() == "stencil2d_kernel"      # fused pass
    assert got.shape == want.shape and got.dtype == want.dtype
    if integer:
        np.testing.assert_array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


def test_smap_laplacian_large_and_device_resident(kex):
    rng = np.random.default_rng(8200)
    x = rng.standard_normal((300, 1028)).astype(np.float32)
    f = lambda v: v[0, 1] + v[0, -1] + v[1, 0] + v[-1, 0] - 4 * v[0, 0]
    got = kex.smap(kernel_size=(3, 3), offset=1, relative=True)(f)(torch.from_numpy(x).cuda())
    assert got.is_cuda and _last_kernel() == "stencil2d_kernel"
    want = x.copy()
    want[1:-1, 1:-1] = x[1:-1, 2:] + x[1:-1, :-2] + x[2:, 1:-1] + x[:-2, 1:-1] - 4 * x[1:-1, 1:-1]
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(got.cpu().numpy()[0], x[0])          # untouched border = input bits
    np.testing.assert_array_equal(got.cpu().numpy()[:, -1], x[:, -1])


def test_smap_max_and_mean_use_the_write_back_path(kex):
    x = np.random.default_rng(8300).integers(-9, 9, (12, 16)).astype(np.int32)
    f = lambda v: np.max(v)
    got = kex.smap(kernel_size=(3, 3), offset=1)(f)(x)
    want = ko.offset_kernel_map({f: ()}, x.shape, (3, 3), (1, 1), ((1, 1), (1, 1)), False)(x)
    np.testing.assert_array_equal(got, want)
    xf = x.astype(np.float32)
    g = lambda v: np.mean(v)
    got = kex.smap(kernel_size=(3, 3), offset=1)(g)(xf)
    want = ko.offset_kernel_map({g: ()}, x.shape, (3, 3), (1, 1), ((1, 1), (1, 1)), False)(xf)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", [
    ((40, 50, 64), (3, 3, 3), ((1, 1), (1, 1), (1, 1))),        # a Jacobi step: 7-point stencil, one-cell ring kept
    ((36, 44, 68), (3, 3, 3), ((1, 0), (0, 1), (1, 1))),        # one-sided offsets: zero-padded windows on the open sides
    ((70, 33, 36), (3, 3, 3), ((1, 1), (1, 1), (1, 1))),        # rows not 16-byte aligned
    ((300, 260), (5, 5), ((2, 2), (1, 2))),                     # 2-D, 5x5
    ((2, 20, 40, 48), (1, 3, 3, 3), ((0, 0), (1, 1), (1, 1), (1, 1))),   # leading batch axis
])
@pytest.mark.parametrize("integer", [False, True])
def test_fused_offset_map_beyond_2d(kex, case, integer, monkeypatch):
    """smap over 3-D windows (and anything else the 2-D fused pass declines): the map over EVERY cell on the fast
    templates + the ring outside the offset interior restored from the input (kx_map_offset) -- against the literal
    oracle and the copy + generic kernel + strided write-back path (KX_NO_FUSED_OFFSET=1)."""
    import itertools
    shape, ksize, offset = case
    rng = np.random.default_rng(sum(shape) + integer)
    taps = [t for t in itertools.product(*[range(k) for k in ksize]) if sum(abs(a - (k - 1) // 2) for a, k in zip(t, ksize)) <= 1]
    coefs = rng.integers(-3, 4, len(taps)) if integer else rng.standard_normal(len(taps))

    def fn(v):
        acc = 0 if integer else 0.0
        for t, c in zip(taps, coefs):
            acc = acc + (int(c) if integer else float(c)) * v[t]
        return acc

    x = rng.integers(-9, 10, shape).astype(np.int32) if integer else rng.standard_normal(shape).astype(np.float32)
    st = (1,) * len(shape)
    want = ko.offset_kernel_map({fn: ()}, shape, ksize, st, offset, False)(x)
    got = kex.offset_kernel_map({fn: ()}, shape, ksize, st, offset, False)(x)
    monkeypatch.setenv("KX_NO_FUSED_OFFSET", "1")
    plain = kex.offset_kernel_map({fn: ()}, shape, ksize, st, offset, False)(x)
    assert got.shape == want.shape and got.dtype == want.dtype
    if integer:
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(plain, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(plain, want, rtol=1e-5, atol=1e-5)


def test_fused_offset_map_reductions_and_meshes(kex, monkeypatch):
    """The general fused offset pass takes any scalar program: a window reduction (smap max filter) and a two-function
    `F[slice] = fn` mesh, against the literal oracle and the write-back path."""
    rng = np.random.default_rng(77)
    shape = (300, 512)
    x = rng.integers(-50, 50, shape).astype(np.int32)
    f = lambda v: np.max(v)                                          # noqa: E731
    want = ko.offset_kernel_map({f: ()}, shape, (3, 3), (1, 1), ((1, 1), (1, 1)), False)(x)
    got = kex.offset_kernel_map({f: ()}, shape, (3, 3), (1, 1), ((1, 1), (1, 1)), False)(x)
    np.testing.assert_array_equal(got, want)
    f0 = lambda v: 2 * v[1, 1] + 1                                   # noqa: E731
    f1 = lambda v: v[0, 1] + v[2, 1] - v[1, 0] - v[1, 2]             # noqa: E731
    fmap = {f0: [[]], f1: [[(50, 250), (100, 400)]]}
    want = ko.offset_kernel_map(fmap, shape, (3, 3), (1, 1), ((1, 1), (1, 1)), False)(x)
    got = kex.offset_kernel_map(fmap, shape, (3, 3), (1, 1), ((1, 1), (1, 1)), False)(x)
    monkeypatch.setenv("KX_NO_FUSED_OFFSET", "1")
    plain = kex.offset_kernel_map(fmap, shape, (3, 3), (1, 1), ((1, 1), (1, 1)), False)(x)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(plain, want)
