"""Patch extraction (kmap(lambda x: x), README.md:157-186) with windows beyond 3x3 -- patch2d_direct_kernel -- bit exact
against the literal oracle, plain and `relative` (rolled) layout, every padding style, ragged widths, batches."""
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


SIZES = [(5, 5), (7, 7), (4, 4), (3, 5), (5, 3), (2, 3), (3, 2), (1, 2), (1, 4), (1, 7), (1, 9)]


@pytest.mark.parametrize("ksize", SIZES)
@pytest.mark.parametrize("relative", [False, True])
def test_wide_patches_match_oracle(kex, ksize, relative):
    rng = np.random.default_rng(sum(ksize) * 7 + relative)
    for case, (shape, border) in enumerate([
            ((37, 300), tuple(((k - 1) // 2, k // 2) for k in ksize)),          # 'same'
            ((23, 1029), ((0, 0), (0, 0))),                                     # 'valid', ragged width, > 4 blocks wide
            ((19, 64), ((1, 0), (2, 1))),                                       # mixed explicit borders
            ((300,), None)]):                                                  # 1-D
        if border is None:
            if ksize[0] != 1:
                continue
            ks, shape_, bd = (ksize[1],), shape, (((ksize[1] - 1) // 2, ksize[1] // 2),)
        else:
            ks, shape_, bd = ksize, shape, border
        if any(n + lo + hi < k for n, k, (lo, hi) in zip(shape_, ks, bd)):
            continue
        x = rng.integers(-1000, 1000, shape_).astype(np.int32) if case % 2 else rng.standard_normal(shape_).astype(np.float32)
        fmap = {(lambda v: v): ()}
        st = (1,) * len(ks)
        want = ko.kernel_map(fmap, shape_, ks, st, bd, relative)(x)
        got = kex.kernel_map(fmap, shape_, ks, st, bd, relative)(x)
        assert _last_kernel() == "patch2d_direct_kernel", (ksize, case, _last_kernel())
        assert got.shape == want.shape and got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)


def test_wide_patches_batched_and_large(kex):
    """A leading batch axis (1-wide window) and a grid of several bands / block columns: against NumPy's own window view."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 200, 1000)).astype(np.float32)
    f = kex.kmap(kernel_size=(1, 5, 5), padding=("valid", "valid", "valid"))(lambda v: v)
    got = f(torch.from_numpy(x).cuda()).cpu().numpy()
    assert _last_kernel() == "patch2d_direct_kernel"
    win = np.lib.stride_tricks.sliding_window_view(x, (1, 5, 5), axis=(0, 1, 2))
    np.testing.assert_array_equal(got, win)


@pytest.mark.parametrize("ksize", [(3, 3, 3), (2, 2, 2)])
@pytest.mark.parametrize("relative", [False, True])
def test_patches_3d_match_oracle(kex, ksize, relative):
    """3-D windows (patch3d_direct_kernel): 'same', 'valid' and mixed borders, ragged widths, int32 and float32."""
    rng = np.random.default_rng(sum(ksize) + relative)
    for case, (shape, border) in enumerate([
            ((9, 14, 300), tuple(((k - 1) // 2, k // 2) for k in ksize)),
            ((7, 11, 261), ((0, 0), (0, 0), (0, 0))),
            ((6, 20, 64), ((1, 0), (0, 1), (2, 1)))]):
        x = rng.integers(-1000, 1000, shape).astype(np.int32) if case % 2 else rng.standard_normal(shape).astype(np.float32)
        fmap = {(lambda v: v): ()}
        want = ko.kernel_map(fmap, shape, ksize, (1, 1, 1), border, relative)(x)
        got = kex.kernel_map(fmap, shape, ksize, (1, 1, 1), border, relative)(x)
        assert _last_kernel() == "patch3d_direct_kernel", (ksize, case, _last_kernel())
        assert got.shape == want.shape and got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)


def test_patches_3d_batched(kex):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, 20, 30, 260)).astype(np.float32)
    f = kex.kmap(kernel_size=(1, 3, 3, 3), padding=("valid", "valid", "valid", "valid"))(lambda v: v)
    got = f(torch.from_numpy(x).cuda()).cpu().numpy()
    assert _last_kernel() == "patch3d_direct_kernel"
    win = np.lib.stride_tricks.sliding_window_view(x, (1, 3, 3, 3), axis=(0, 1, 2, 3))
    np.testing.assert_array_equal(got, win)
