"""int32 arrays whose window function returns floats (jnp.mean of integers, a Python-float coefficient): JAX converts
the window to float32 first (weak-type promotion), so large arrays are converted once by
kx_cast_i32_f32 and run on the float32 templates; small ones stay on the mixed-dtype generic kernel.  Both must give
the oracle's float32 values."""
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


@pytest.mark.parametrize("shape", [(40, 64), (300, 1024), (257, 1030)])
def test_int_array_float_result(kex, shape):
    rng = np.random.default_rng(shape[0])
    x = rng.integers(-1000, 1000, shape).astype(np.int32)
    big = x.size >= (1 << 16)
    cases = [
        ((3, 3), "same", False, lambda v: np.mean(v), [(a, b) for a in range(3) for b in range(3)], [1.0 / 9] * 9),
        ((3, 3), "valid", True, lambda v: 0.25 * (v[1, 0] + v[-1, 0] + v[0, 1] + v[0, -1]),
         [(2, 1), (0, 1), (1, 2), (1, 0)], [0.25] * 4),
    ]
    for ks, pad, rel, fn, taps, coefs in cases:
        got = kex.kmap(kernel_size=ks, padding=pad, relative=rel)(fn)(torch.from_numpy(x).cuda())
        assert got.dtype == torch.float32
        assert (_last_kernel() == "stencil2d_kernel") == big, _last_kernel()
        border = ko.resolve_padding(pad, ks)
        xf = x.astype(np.float64)
        want = ko.affine_map(xf, ks, (1, 1), border, taps, coefs)
        mag = ko.affine_map(np.abs(xf), ks, (1, 1), border, taps, np.abs(coefs))
        assert np.all(np.abs(got.cpu().numpy().astype(np.float64) - want) <= 1e-5 * mag + 1e-30)


def test_int_results_stay_int(kex):
    """Integer-valued functions on int32 arrays are NOT converted: bit-exact int32 (wraparound and all)."""
    x = np.arange(1, 300 * 1024 + 1, dtype=np.int32).reshape(300, 1024)
    got = kex.kmap(kernel_size=(3, 3), padding="same")(lambda v: np.sum(v))(torch.from_numpy(x).cuda())
    assert got.dtype == torch.int32
    xp = np.pad(x.astype(np.int64), 1)
    want = sum(xp[a:a + 300, b:b + 1024] for a in range(3) for b in range(3)).astype(np.int32)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("case", [
    ((40, 44, 64), (3, 3, 3), "same"), ((33, 50, 72), (3, 3, 3), "valid"), ((24, 40, 68), (2, 3, 5), ((1, 0), (0, 2), (2, 2))), ((20, 40, 68), (3, 5, 7), ((1, 0), (0, 2), (3, 3))),
    ((2, 30, 36, 64), (1, 3, 3, 3), ((0, 0), (1, 1), (1, 1), (1, 1))),
])
@pytest.mark.parametrize("op", ["max", "min"])
def test_separable_3d_reduction_is_exact(kex, case, op):
    """3-D max / min filters run as three passes of the 2-D row-march kernel (engine._separable_reduce_3d): bit for bit
    the literal oracle's zero-padded 3-D window reduction, int32 and float32."""
    shape, ksize, pad = case
    border = ko.resolve_padding(pad, ksize)
    rng = np.random.default_rng(sum(shape))
    fn = (lambda v: np.max(v)) if op == "max" else (lambda v: np.min(v))
    for integer in (False, True):
        x = rng.integers(-50, 50, shape).astype(np.int32) if integer else rng.standard_normal(shape).astype(np.float32)
        want = ko.kernel_map({fn: ()}, shape, ksize, (1,) * len(shape), border, False)(x)
        got = kex.kernel_map({fn: ()}, shape, ksize, (1,) * len(shape), border, False)(torch.from_numpy(x).cuda())
        if tuple(ksize[-3:]) == (3, 3, 3):
            assert _last_kernel() == "stencil2d_kernel", _last_kernel()
        assert tuple(got.shape) == want.shape
        np.testing.assert_array_equal(got.cpu().numpy(), want)
