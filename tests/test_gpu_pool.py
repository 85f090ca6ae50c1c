"""GPU parity of the strided-window / window-reduction kernel (kx_pool2d.cu): average, sum,
max and min pooling, strided weighted windows and stride-1 max / min filters against the
oracle, over every window misalignment (left border 0..3), ragged heights / widths, several
bands, warps and column blocks, fp32 and int32, plus batched (vmap) and 1-D forms."""
from __future__ import annotations

import itertools

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


GEOMS = [((2, 2), (2, 2)), ((3, 3), (2, 2)), ((3, 3), (3, 3)), ((3, 3), (1, 1)), ((2, 2), (1, 1)),
         ((1, 3), (1, 2)), ((1, 2), (1, 2)), ((1, 3), (1, 1)), ((4, 4), (4, 4)), ((4, 4), (2, 2)), ((1, 3), (1, 3)),
         ((1, 4), (1, 4)), ((5, 5), (2, 2)), ((1, 5), (1, 2))]
FNS = {"max": lambda v: np.max(v), "min": lambda v: np.min(v), "mean": lambda v: np.mean(v),
       "sum": lambda v: np.sum(v)}


def _pool2d_cases():
    """(geometry, op, case, legacy): `legacy` = stride-1 max / min on the row-ring kernel (their home before stencil2d
    learnt them); it only exists for those combinations, so only those are generated (no skipped placeholders)."""
    out = []
    for gi, (ksize, strides) in enumerate(GEOMS):
        for op in ("max", "min", "mean", "sum"):
            for case in range(4):
                out.append((gi, op, case, False))
                if strides == (1, 1) and op in ("max", "min"):
                    out.append((gi, op, case, True))
    return out


@pytest.mark.parametrize("gi,op,case,legacy", _pool2d_cases())
def test_pool2d_matches_oracle(kex, gi, op, case, legacy, monkeypatch):
    ksize, strides = GEOMS[gi]
    if legacy:   # stride-1 max / min / mean on the row-ring kernel (their home before stencil2d learnt them)
        monkeypatch.setenv("KX_S2_LEGACY_SHAPES", "1")
    expect = "stencil2d_kernel" if strides == (1, 1) and not legacy else "pool2d_kernel"
    rng = np.random.default_rng(7000 + 100 * gi + case)
    h = [70, 33, 129, 200][case]
    w = [516, 132, 1028, 260][(case + gi) % 4]
    border = ((int(rng.integers(0, 3)), int(rng.integers(0, 3))), (case % 4, int(rng.integers(0, 3))))
    if ksize[0] == 1:
        border = ((0, 0), border[1])
    integer = (case % 2 == 1) and op != "mean"
    if integer:
        x = rng.integers(-1000, 1000, (h, w)).astype(np.int32)
    else:
        x = rng.standard_normal((h, w)).astype(np.float32)
    got = kex.kernel_map({FNS[op]: ()}, (h, w), ksize, strides, border, False)(x)
    assert _last_kernel() == expect
    want = ko.reduce_map(x, ksize, strides, border, op)
    assert got.shape == want.shape and got.dtype == want.dtype
    if integer or op in ("max", "min"):
        np.testing.assert_array_equal(got, want)
    else:
        tol = 1e-5 * float(np.abs(x).max()) * ksize[0] * ksize[1]
        np.testing.assert_allclose(got, want, rtol=0, atol=tol)


@pytest.mark.parametrize("seed", range(8))
def test_strided_weighted_window(kex, seed):
    rng = np.random.default_rng(7700 + seed)
    ksize, strides = [((3, 3), (2, 2)), ((2, 2), (2, 2)), ((3, 3), (3, 3)), ((5, 5), (2, 2))][seed % 4]
    h, w = int(rng.integers(40, 150)), 4 * int(rng.integers(40, 300))
    border = ((int(rng.integers(0, 2)), int(rng.integers(0, 2))), (int(rng.integers(0, 4)), 1))
    taps = list(itertools.product(range(ksize[0]), range(ksize[1])))
    integer = seed % 2 == 0
    if integer:
        x = rng.integers(-100, 100, (h, w)).astype(np.int32)
        coefs, bias = [int(c) or 1 for c in rng.integers(-5, 6, len(taps))], 7
    else:
        x = rng.standard_normal((h, w)).astype(np.float32)
        coefs, bias = [float(c) for c in rng.standard_normal(len(taps))], 0.5

    def fn(v):
        acc = bias
        for t, c in zip(taps, coefs):
            acc = acc + c * v[t]
        return acc

    got = kex.kernel_map({fn: ()}, (h, w), ksize, strides, border, False)(x)
    assert _last_kernel() == "pool2d_kernel"
    want = ko.affine_map(x, ksize, strides, border, taps, coefs, bias)
    if integer:
        np.testing.assert_array_equal(got, want)
    else:
        scale = ko.affine_map(np.abs(x), ksize, strides, border, taps, np.abs(coefs), abs(bias))
        assert np.all(np.abs(got.astype(np.float64) - want) <= 1e-5 * scale + 1e-30)


def test_batched_and_1d_pooling(kex):
    rng = np.random.default_rng(7800)
    x = rng.standard_normal((5, 64, 256)).astype(np.float32)
    pool = kex.vmap(kex.kmap(kernel_size=(3, 3), strides=(2, 2))(lambda v: np.mean(v)))   # README.md:331-336
    got = pool(torch.from_numpy(x).cuda())
    assert _last_kernel() == "pool2d_kernel"
    want = np.stack([ko.reduce_map(x[b], (3, 3), (2, 2), ((0, 0), (0, 0)), "mean") for b in range(5)])
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=1e-5 * 9 * float(np.abs(x).max()))
    x1 = rng.integers(-50, 50, (4096,)).astype(np.int32)
    got1 = kex.kmap(kernel_size=(2,), strides=(2,))(lambda v: np.max(v))(x1)
    assert _last_kernel() == "pool2d_kernel"
    np.testing.assert_array_equal(got1, x1.reshape(-1, 2).max(axis=1))


def test_unaligned_rows_fall_back_to_the_generic_kernel(kex):
    x = np.random.default_rng(1).standard_normal((40, 131)).astype(np.float32)   # rows not 16-byte aligned
    got = kex.kmap(kernel_size=(2, 2), strides=(2, 2))(lambda v: np.max(v))(x)
    assert _last_kernel() == "generic_map_kernel"
    np.testing.assert_array_equal(got, ko.reduce_map(x, (2, 2), (2, 2), ((0, 0), (0, 0)), "max"))


GEOMS3 = [(2, 2), (3, 2), (3, 3)]


@pytest.mark.parametrize("gi", range(len(GEOMS3)))
@pytest.mark.parametrize("op", ["max", "min", "mean", "sum"])
@pytest.mark.parametrize("case", range(5))
def test_pool3d_matches_oracle(kex, gi, op, case):
    """kx_pool3d: cubic K^3 windows with stride S^3 over a volume (3-D pooling), every window misalignment (left x border
    0..3), padded / cropped faces on every axis, ragged sizes, several column blocks and z chunks, fp32 and int32."""
    K, S = GEOMS3[gi]
    rng = np.random.default_rng(8000 + 100 * gi + case)
    d = [9, 17, 33, 6, 40][case]
    h = [14, 33, 20, 9, 50][case]
    w = [132, 516, 260, 1028, 64][(case + gi) % 5]
    border = tuple((int(rng.integers(0, K)), int(rng.integers(0, K))) for _ in range(2)) + ((case % 4, int(rng.integers(0, K))),)
    if case == 3:
        border = ((0, 0), (0, 0), (0, 0))
    if case == 4:
        border = ((-1, 1), (1, -1), (min(2, K - 1), 0))      # negative borders crop
    integer = (case % 2 == 1) and op != "mean"
    x = (rng.integers(-1000, 1000, (d, h, w)).astype(np.int32) if integer else rng.standard_normal((d, h, w)).astype(np.float32))
    got = kex.kernel_map({FNS[op]: ()}, (d, h, w), (K,) * 3, (S,) * 3, border, False)(x)
    assert _last_kernel() == "pool3d_kernel"
    want = ko.reduce_map(x, (K,) * 3, (S,) * 3, border, op)
    assert got.shape == want.shape and got.dtype == want.dtype
    if integer or op in ("max", "min"):
        np.testing.assert_array_equal(got, want)
    else:
        # 1e-5 * sum |x| over the window (the zero padding takes part in the mean, like the reference)
        scale = ko.reduce_map(np.abs(x), (K,) * 3, (S,) * 3, border, "sum") / (K ** 3 if op == "mean" else 1)
        assert np.all(np.abs(got.astype(np.float64) - want) <= 1e-5 * scale + 1e-30)


@pytest.mark.parametrize("seed", range(6))
def test_pool3d_weighted_windows_batched_and_chunked(kex, seed, monkeypatch):
    """Strided WEIGHTED 3-D windows (host weights) on pool3d_kernel: exact accumulation order is bias, then (kz, ky, kx)
    row-major; a leading batch axis; forced z chunks must give bit-identical results."""
    rng = np.random.default_rng(8800 + seed)
    K, S = GEOMS3[seed % 3]
    b, d, h, w = 2, int(rng.integers(10, 40)), int(rng.integers(8, 40)), 4 * int(rng.integers(20, 150))
    border = ((0, 0),) + tuple((int(rng.integers(0, K)), int(rng.integers(0, K))) for _ in range(3))
    taps = list(itertools.product(range(K), range(K), range(K)))
    integer = seed % 2 == 0
    if integer:
        x = rng.integers(-100, 100, (b, d, h, w)).astype(np.int32)
        wts = rng.integers(-5, 6, (K, K, K)).astype(np.int32)
        wts[wts == 0] = 2
        bias = 3
    else:
        x = rng.standard_normal((b, d, h, w)).astype(np.float32)
        wts = rng.standard_normal((K, K, K)).astype(np.float32)
        bias = 0.5
    fn = lambda v: bias + np.sum(v[0] * wts)          # noqa: E731
    f = kex.kernel_map({fn: ()}, (b, d, h, w), (1, K, K, K), (1, S, S, S), border, False)
    got = f(x)
    assert _last_kernel() == "pool3d_kernel"
    taps4 = [(0,) + t for t in taps]
    coefs = [wts[t].item() for t in taps]
    want = ko.affine_map(x, (1, K, K, K), (1, S, S, S), border, taps4, coefs, bias)
    assert got.shape == want.shape
    if integer:
        np.testing.assert_array_equal(got, want)
    else:
        scale = ko.affine_map(np.abs(x), (1, K, K, K), (1, S, S, S), border, taps4, np.abs(coefs), abs(bias))
        assert np.all(np.abs(got.astype(np.float64) - want) <= 1e-5 * scale + 1e-30)
    for nzc in ("1", "3", "7"):
        monkeypatch.setenv("KX_POOL3_NZC", nzc)
        np.testing.assert_array_equal(f(x), got)
    monkeypatch.delenv("KX_POOL3_NZC")
    # device weights and unaligned rows are not this kernel's: they must still be right (generic kernel)
    x2 = x[..., : w - 1].copy()
    got2 = kex.kernel_map({fn: ()}, x2.shape, (1, K, K, K), (1, S, S, S), border, False)(x2)
    assert _last_kernel() != "pool3d_kernel"
    want2 = ko.affine_map(x2, (1, K, K, K), (1, S, S, S), border, taps4, coefs, bias)
    if integer:
        np.testing.assert_array_equal(got2, want2)
    else:
        np.testing.assert_allclose(got2, want2, rtol=1e-4, atol=1e-4)


def test_pool3d_avgpool_readme_generalisation_under_vmap(kex):
    """README.md:327-337 average pooling, one dimension up: kmap(kernel (3,3,3), strides (2,2,2), 'valid')(mean) mapped
    over a leading channel axis with vmap; and its input gradient (generic strided VJP, carries the 1/27)."""
    rng = np.random.default_rng(8900)
    x = rng.standard_normal((3, 21, 30, 132)).astype(np.float32)
    pool = kex.vmap(kex.kmap(kernel_size=(3, 3, 3), strides=(2, 2, 2), padding="valid")(lambda v: np.mean(v)))
    got = pool(torch.from_numpy(x).cuda())
    assert _last_kernel() == "pool3d_kernel"
    want = np.stack([ko.reduce_map(x[c], (3, 3, 3), (2, 2, 2), ((0, 0),) * 3, "mean") for c in range(3)])
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
