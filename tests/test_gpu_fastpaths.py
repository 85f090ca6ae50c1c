"""GPU parity of the roofline-tuned kernel templates against the NumPy oracle on seeded
random geometry (aligned and ragged widths, positive / zero / negative borders), and the
size-independent properties used at BASELINE.json's full sizes."""
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
    from kernex_b200 import _lib
    _lib.load()
    return kernex_b200


def _last_kernel():
    from kernex_b200 import _lib
    return _lib.last_kernel()


def _check_affine(got, x, ksize, strides, border, taps, coefs, bias=0):
    want = ko.affine_map(x, ksize, strides, border, taps, coefs, bias)
    assert got.shape == want.shape
    if x.dtype == np.int32:
        np.testing.assert_array_equal(got, want)
    else:
        # tolerance of BASELINE.json: |gpu - oracle| <= 1e-5 * sum_t |c_t x_t| (fp32 FMA chain vs mul+add)
        scale = ko.affine_map(np.abs(x), ksize, strides, border, taps, np.abs(np.asarray(coefs)), abs(bias))
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        assert np.all(err <= 1e-5 * scale + 1e-30), float((err / (scale + 1e-30)).max())


@pytest.mark.parametrize("seed", range(40))
def test_stencil2d_random_geometry(kex, seed):
    rng = np.random.default_rng(1000 + seed)
    ky, kx = [(3, 3), (1, 3), (3, 1), (2, 2), (5, 5), (2, 3), (3, 2), (1, 5), (1, 1), (5, 1)][seed % 10]
    h = int(rng.integers(max(ky, 2), 70))
    w = int(rng.integers(max(kx, 2), 300))
    if seed % 3 == 0:
        w = (w // 4 + 1) * 4  # 16-byte aligned rows -> vector-load path
    border = tuple((int(rng.integers(-2, 4)), int(rng.integers(-2, 4))) for _ in range(2))
    if any(o <= 0 for o in ko.output_shape((h, w), (ky, kx), (1, 1), border)):
        border = ((0, 0), (0, 0))
    taps = [t for t in itertools.product(range(ky), range(kx)) if rng.random() < 0.7] or [(0, 0)]
    if seed % 2:
        x = rng.standard_normal((h, w)).astype(np.float32)
        coefs = [float(c) for c in rng.standard_normal(len(taps))]
        bias = 0.25
    else:
        x = rng.integers(-1000, 1000, (h, w)).astype(np.int32)
        coefs = [int(c) for c in rng.integers(-9, 10, len(taps))]
        bias = 3
    coefs = [c if c != 0 else 1 for c in coefs]

    def fn(v):
        acc = bias
        for t, c in zip(taps, coefs):
            acc = acc + c * v[t]
        return acc

    got = kex.kernel_map({fn: ()}, (h, w), (ky, kx), (1, 1), border, False)(x)
    assert _last_kernel() == "stencil2d_kernel"
    _check_affine(got, x, (ky, kx), (1, 1), border, taps, coefs, bias)


def test_stencil2d_nonfinite_inputs_do_not_leak_through_missing_taps(kex):
    x = np.ones((16, 64), np.float32)
    x[5, 7] = np.inf
    x[9, 20] = np.nan
    out = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lambda v: v[0, 1] + v[0, -1])(x)
    want = ko.affine_map(x, (3, 3), (1, 1), ((1, 1), (1, 1)), [(1, 2), (1, 0)], [1, 1])
    np.testing.assert_array_equal(np.isnan(out), np.isnan(want))
    np.testing.assert_array_equal(np.isinf(out), np.isinf(want))


def test_stencil1d_config1_bit_exact(kex):
    """BASELINE config 1: README kmap sum_all, kernel_size=(3,), int32 arange(1e6), valid."""
    n = 1_000_000
    x = np.arange(1, n + 1, dtype=np.int32)
    out = kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))(x)
    assert _last_kernel() == "stencil2d_kernel"
    assert out.dtype == np.int32 and out.shape == (n - 2,)
    np.testing.assert_array_equal(out, x[:-2] + x[1:-1] + x[2:])


def test_batched_leading_axis_folds_into_the_fast_path(kex):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 33, 64)).astype(np.float32)
    f = lambda v: v[0, 0, 0] - 0.5 * v[0, 1, 1] + 2 * v[0, -1, 0]
    got = kex.kmap(kernel_size=(1, 3, 3), padding="same", relative=True)(f)(x)
    assert _last_kernel() == "stencil2d_kernel"
    taps, coefs = [(0, 1, 1), (0, 2, 2), (0, 0, 1)], [1.0, -0.5, 2.0]
    _check_affine(got, x, (1, 3, 3), (1, 1, 1), ((0, 0), (1, 1), (1, 1)), taps, coefs)


def test_laplacian_full_size_properties(kex):
    """BASELINE config 2 at full size (16384^2 fp32): linearity, exact zero on constants,
    and oracle parity on row bands the oracle finishes in seconds."""
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, n), device="cuda", dtype=torch.float32, generator=g)
    lap = kex.kmap(kernel_size=(3, 3), padding="valid", relative=True)(
        lambda v: (0 * v[1, -1] + 1 * v[1, 0] + 0 * v[1, 1] + 1 * v[0, -1] + -4 * v[0, 0] + 1 * v[0, 1]
                   + 0 * v[-1, -1] + 1 * v[-1, 0] + 0 * v[-1, 1]))
    y = lap(x)
    assert _last_kernel() == "stencil2d_kernel"
    assert y.shape == (n - 2, n - 2)
    # constants are annihilated exactly (README.md:135-143)
    assert float(lap(torch.full((n, n), 3.0, device="cuda")).abs().max()) == 0.0
    # linearity: L(2x) == 2 L(x) bit for bit (power-of-two scaling is exact in fp32)
    assert torch.equal(lap(2 * x), 2 * y)
    # oracle on three bands
    taps, coefs = [(2, 1), (1, 0), (1, 1), (1, 2), (0, 1)], [1, 1, -4, 1, 1]
    for r0 in (0, 8000, n - 66):
        band = x[r0:r0 + 66].cpu().numpy()
        _check_affine(y[r0:r0 + 64].cpu().numpy(), band, (3, 3), (1, 1), ((0, 0), (0, 0)), taps, coefs)
    del x, y
    torch.cuda.empty_cache()
