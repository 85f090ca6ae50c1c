"""CPU check of the 1-D -> 2-D fold (kernex_b200/fold1d.py) with the ORACLE standing in for the CUDA kernels: the folded
evaluation must equal the oracle's direct 1-D kmap, bit for bit on integers."""
from __future__ import annotations

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from kernex_b200.fold1d import folded_map_1d, plan_fold
from oracle import kx_oracle as ko


def _runner(k, border, taps, coefs, bias):
    def run(t2d):
        a = t2d.numpy()
        return torch.from_numpy(ko.affine_map(a, (1, k), (1, 1), ((0, 0), border), [(0, t) for t in taps], coefs, bias))
    return run


@pytest.mark.parametrize("k,lo,hi", [(3, 1, 1), (3, 0, 0), (5, 2, 2), (5, 0, 0), (4, 1, 2), (7, 0, 3), (7, 6, 0), (2, 0, 1),
                                     (1, 0, 0), (15, 7, 7), (6, 2, 1)])
@pytest.mark.parametrize("dtype", [np.int32, np.float32])
def test_folded_map_equals_the_direct_1d_map(k, lo, hi, dtype):
    rng = np.random.default_rng(100 * k + 10 * lo + hi)
    w, h = 1024, 7
    n = w * h
    x = (rng.integers(-1000, 1000, n).astype(np.int32) if dtype == np.int32 else rng.standard_normal(n).astype(np.float32))
    taps = [t for t in range(k) if rng.random() < 0.8] or [0]
    coefs = [int(c) or 2 for c in rng.integers(-9, 10, len(taps))]
    bias = 3
    want = ko.affine_map(x, (k,), (1,), ((lo, hi),), [(t,) for t in taps], coefs, bias)
    got = folded_map_1d(torch.from_numpy(x), k, lo, hi, w,
                        _runner(k, (lo, k - 1 - lo), taps, coefs, bias), _runner(k, (0, 0), taps, coefs, bias)).numpy()
    assert got.shape == want.shape
    if dtype == np.int32:
        np.testing.assert_array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-4)


def test_plan_fold_conditions():
    assert plan_fold(1 << 27, 15, 7, 7) == 8192 or plan_fold(1 << 27, 15, 7, 7) == 16384
    assert plan_fold(1 << 27, 3, 0, 0) in (16384, 8192)
    assert plan_fold(10 ** 6, 3, 0, 0) is None            # BASELINE config 1 is launch-latency bound: not folded
    assert plan_fold(1 << 27, 3, 2, 2) is None            # over-padded
    assert plan_fold(1 << 27, 3, -1, 0) is None           # cropping border
    assert plan_fold(1 << 27, 3, 1, 1, stride=2) is None
    assert plan_fold((1 << 22) + 4, 3, 1, 1) is None      # no row width divides it
