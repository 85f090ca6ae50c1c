"""float64 restatements of the kmap VJPs and north_star's 1e-5 bound (test infrastructure).

dx: scatter-add transpose of the (strided) AFFINE stencil; dw: window correlation -- what JAX derives by
transposing gather / pad (kernex tests/test_interface.py:345-348)."""
from __future__ import annotations

import numpy as np

from oracle import kx_oracle as ko

TOL = 1e-5


def dx_reference(g, shape, ksize, strides, border, taps, coefs, post_div=0.0):
    """Scatter-add transpose of the (strided) stencil in float64; returns (dx, sum |c||g|)."""
    dx = np.zeros(shape, dtype=np.float64)
    mag = np.zeros(shape, dtype=np.float64)
    g64 = np.asarray(g, dtype=np.float64)
    for t, c in zip(taps, coefs):
        # x[j*s - lo + t] receives c * g[j]
        src, dst = [], []
        for d in range(len(shape)):
            lo, s = border[d][0], strides[d]
            j0 = max(0, -((t[d] - lo) // s))                      # first j with j*s - lo + t >= 0
            j1 = min(g64.shape[d], (shape[d] - 1 + lo - t[d]) // s + 1)
            src.append(slice(j0, max(j1, j0)))
            dst.append(slice(j0 * s - lo + t[d], max(j1, j0) * s - lo + t[d], s) if j1 > j0 else slice(0, 0))
        if all(s_.stop > s_.start for s_ in src):
            dx[tuple(dst)] += c * g64[tuple(src)]
            mag[tuple(dst)] += abs(c) * np.abs(g64[tuple(src)])
    if post_div:
        dx, mag = dx / post_div, mag / abs(post_div)
    return dx, mag


def dw_reference(x, g, ksize, strides, border, taps):
    """dw[t] = sum_j g[j] x[origin(j)+t] and its magnitude sum_j |g||x| in float64, per listed tap."""
    work, out_shape = ko._padded_for_slices(np.asarray(x), ksize, strides, border)
    g64 = np.asarray(g, dtype=np.float64).reshape(out_shape)
    dw, mag = [], []
    for t in taps:
        sl = ko._tap_slice(work, t, out_shape, strides).astype(np.float64)
        dw.append(float(np.sum(g64 * sl)))
        mag.append(float(np.sum(np.abs(g64) * np.abs(sl))))
    return np.asarray(dw), np.asarray(mag)


def assert_within(got, want, mag, what):
    err = np.abs(np.asarray(got, dtype=np.float64) - want)
    bound = TOL * mag + 1e-30
    assert np.all(err <= bound), f"{what}: worst err / (sum|c g|) = {float((err / (mag + 1e-30)).max()):.3e} > {TOL}"
