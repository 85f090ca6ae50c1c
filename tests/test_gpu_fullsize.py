"""BASELINE.json configs at their FULL sizes on the GPU, checked through size-independent
properties (linearity, exact annihilation / reproduction of constants, bit-exact shifted views)
and against the oracle on bands it finishes in seconds.  Config 2 lives in
test_gpu_fastpaths.py::test_laplacian_full_size_properties, config 4 in
test_gpu_scan.py::test_rowmarch_config4_prefix_strip_full_depth."""
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


def test_config1_readme_sum_int32(kex):
    """configs[0]: README kmap sum_all, kernel_size=(3,), 1-D int32 arange(1e6), padding='valid': bit exact."""
    x = np.arange(1, 1_000_001, dtype=np.int32)
    got = kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))(x)
    assert got.dtype == np.int32 and got.shape == (999_998,)
    np.testing.assert_array_equal(got, x[:-2] + x[1:-1] + x[2:])
    np.testing.assert_array_equal(got[:3], [6, 9, 12])    # kernel_interface.py:379-391


def test_config3a_conv_full_size(kex):
    """conv-as-kmap (3,3,3) on (3, 8192, 8192), padding (valid, same, same), device weights."""
    n = 8192
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((3, n, n), device="cuda", generator=g)
    w = torch.randn((3, 3, 3), device="cuda", generator=g)
    conv = kex.kmap(kernel_size=(3, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    y = conv(x, w)
    assert _last_kernel() == "conv_fwd_rows_kernel" and y.shape == (1, n, n)
    assert torch.equal(conv(2 * x, w), 2 * y)                       # exact power-of-two linearity in x
    assert torch.equal(conv(x, 4 * w), 4 * y)                       # ... and in w
    # a one-hot weight reproduces a shifted input plane bit for bit (zero padding at the edge)
    for (c, a, b) in [(0, 0, 0), (1, 1, 1), (2, 2, 1), (1, 0, 2)]:
        e = torch.zeros((3, 3, 3), device="cuda")
        e[c, a, b] = 1.0
        s = conv(x, e)[0]
        want = torch.zeros((n, n), device="cuda")
        ys, xs = a - 1, b - 1
        want[max(-ys, 0):n - max(ys, 0), max(-xs, 0):n - max(xs, 0)] = \
            x[c, max(ys, 0):n - max(-ys, 0), max(xs, 0):n - max(-xs, 0)]
        assert torch.equal(s, want)
    # oracle on bands (top edge, middle, bottom edge)
    taps = list(itertools.product(range(3), range(3), range(3)))
    wh = w.cpu().numpy()
    coefs = [float(wh[t]) for t in taps]
    for r0 in (0, 4000, n - 20):
        lo, hi = max(r0 - 1, 0), min(r0 + 21, n)
        band = x[:, lo:hi].cpu().numpy()
        border = ((0, 0), (1 if r0 == 0 else 0, 1 if hi == n else 0), (1, 1))
        want = ko.affine_map(band, (3, 3, 3), (1, 1, 1), border, taps, coefs)
        scale = ko.affine_map(np.abs(band), (3, 3, 3), (1, 1, 1), border, taps, np.abs(coefs))
        rows = want.shape[1]
        got = y[:, r0:r0 + rows].cpu().numpy() if r0 == 0 else y[:, lo + 1:lo + 1 + rows].cpu().numpy()
        err = np.abs(got.astype(np.float64) - want)
        assert np.all(err <= 1e-5 * scale + 1e-30)
    del x, y
    torch.cuda.empty_cache()


def test_config2_backward_full_size(kex):
    """configs[1] "plus its custom_vjp backward" at 16384^2: dx = flipped-tap stencil of the (16382, 16382)
    cotangent.  float64 transpose on row bands (top edge, interior, bottom edge) at 1e-5 * sum |c g|, exact
    power-of-two linearity, exact annihilation of a constant cotangent in the interior, and the adjoint identity
    <L x, g> == <x, L^T g> in float64 over the whole grid."""
    from vjp_ref import assert_within, dx_reference
    n = 16384
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((n, n), device="cuda", generator=gen).requires_grad_()
    lap = kex.kmap(kernel_size=(3, 3), padding="valid", relative=True)(
        lambda v: (0 * v[1, -1] + 1 * v[1, 0] + 0 * v[1, 1] + 1 * v[0, -1] + -4 * v[0, 0] + 1 * v[0, 1]
                   + 0 * v[-1, -1] + 1 * v[-1, 0] + 0 * v[-1, 1]))
    y = lap(x)
    g = torch.randn((n - 2, n - 2), device="cuda", generator=gen)
    (dx,) = torch.autograd.grad(y, x, g, retain_graph=True)
    assert _last_kernel() == "stencil2d_kernel" and dx.shape == (n, n)
    taps, coefs = [(2, 1), (1, 0), (1, 1), (1, 2), (0, 1)], [1.0, 1.0, -4.0, 1.0, 1.0]
    border = ((0, 0), (0, 0))
    # dx rows [r0, r1) depend on cotangent rows [r0-2, r1): run the float64 transpose on that band of g as if it
    # were the whole cotangent of a (rows+2, n) grid and compare the rows whose every contribution is in the band
    for r0, r1 in ((0, 64), (8000, 8064), (n - 64, n)):
        j0, j1 = max(r0 - 2, 0), min(r1, n - 2)
        gb = g[j0:j1].cpu().numpy()
        want, mag = dx_reference(gb, (j1 - j0 + 2, n), (3, 3), (1, 1), border, taps, coefs)
        # band-local row i is global row j0 + i; rows r0..r1 are complete when j0 <= r - 2 or r < 2, and r <= j1 + 1 ...
        lo = r0 - j0
        hi = lo + (r1 - r0)
        assert_within(dx[r0:r1].cpu().numpy(), want[lo:hi], mag[lo:hi], f"dx rows [{r0},{r1})")
    (dx2,) = torch.autograd.grad(y, x, 2 * g, retain_graph=True)
    assert torch.equal(dx2, 2 * dx)
    del dx2
    (dxc,) = torch.autograd.grad(y, x, torch.full_like(g, 3.0), retain_graph=True)
    assert float(dxc[2:n - 2, 2:n - 2].abs().max()) == 0.0          # coefficients sum to zero
    del dxc
    lhs = float((y.detach().double() * g.double()).sum())
    rhs = float((x.detach().double() * dx.double()).sum())
    mag = float((y.detach().abs().double() * g.abs().double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * mag, (lhs, rhs, mag)
    del x, y, g, dx
    torch.cuda.empty_cache()


def test_config3a_weight_gradient_full_size(kex):
    """configs[2] weight gradient at (3, 8192, 8192): dW[t] = sum_j g[j] x[origin(j)+t] (27 numbers out of 2 x 10^8
    products each).  (1) cotangent supported on three row bands -> the NumPy float64 oracle over those bands, at
    1e-5 * sum |g x|; (2) dense cotangent -> the same correlation restated in float64 with torch slices on the device
    (the oracle's formula, evaluated where the data is; cross-checked against the NumPy oracle by (1)); (3) exact
    power-of-two linearity; (4) determinism (fixed-order reduction: two runs are bit-identical)."""
    from vjp_ref import TOL, assert_within, dw_reference
    n = 8192
    gen = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((3, n, n), device="cuda", generator=gen)
    w = torch.randn((3, 3, 3), device="cuda", generator=gen).requires_grad_()
    conv = kex.kmap(kernel_size=(3, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    y = conv(x, w)
    assert y.shape == (1, n, n)
    taps = list(itertools.product(range(3), range(3), range(3)))
    border = ((0, 0), (1, 1), (1, 1))

    def dw_f64(g):
        """float64 window correlation with zero padding, torch slices (restates oracle/kx_oracle.py weight_grad)."""
        xp = torch.nn.functional.pad(x, (1, 1, 1, 1)).double()
        g64, out, mag = g[0].double(), [], []
        for (c, a, b) in taps:
            sl = xp[c, a:a + n, b:b + n]
            out.append(float((g64 * sl).sum()))
            mag.append(float((g64.abs() * sl.abs()).sum()))
        del xp
        return np.asarray(out), np.asarray(mag)

    # (1) banded cotangent against the NumPy oracle
    bands = [(0, 48), (4000, 4064), (n - 48, n)]
    g = torch.zeros((1, n, n), device="cuda")
    for r0, r1 in bands:
        g[0, r0:r1].normal_(generator=gen)
    (dw,) = torch.autograd.grad(y, w, g, retain_graph=True)
    assert _last_kernel() in ("conv_wgrad_finish", "conv_wgrad_rows_kernel", "conv_wgrad_kernel"), _last_kernel()
    want = np.zeros(27)
    mag = np.zeros(27)
    for r0, r1 in bands:
        lo, hi = max(r0 - 1, 0), min(r1 + 1, n)
        bx = x[:, lo:hi].cpu().numpy()
        bb = ((0, 0), (1 if r0 == 0 else 0, 1 if r1 == n else 0), (1, 1))
        d, m = dw_reference(bx, g[:, r0:r1].cpu().numpy(), (3, 3, 3), (1, 1, 1), bb, taps)
        want += d
        mag += m
    assert_within(dw.cpu().numpy().reshape(-1), want, mag, "dW (banded cotangent, NumPy oracle)")
    t_want, t_mag = dw_f64(g)
    np.testing.assert_allclose(t_want, want, rtol=0, atol=1e-9 * float(mag.max()))   # the torch restatement == the oracle
    # (2) dense cotangent
    g.normal_(generator=gen)
    (dw,) = torch.autograd.grad(y, w, g, retain_graph=True)
    want, mag = dw_f64(g)
    assert_within(dw.cpu().numpy().reshape(-1), want, mag, "dW (dense cotangent)")
    # (3) + (4)
    (dw2,) = torch.autograd.grad(y, w, 2 * g, retain_graph=True)
    assert torch.equal(dw2, 2 * dw)
    (dw3,) = torch.autograd.grad(y, w, g, retain_graph=True)
    assert torch.equal(dw3, dw)
    assert TOL == 1e-5
    del x, y, g
    torch.cuda.empty_cache()


def test_config3b_patches_full_size(kex):
    """3x3 patch extraction 'same' on 8192^2: every patch slot is a shifted copy of the input, bit exact."""
    n = 8192
    x = torch.randn((n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    p = kex.kmap(kernel_size=(3, 3), padding="same")(lambda v: v)(x)
    assert _last_kernel() == "patch2d_kernel" and p.shape == (n, n, 3, 3)
    for a in range(3):
        for b in range(3):
            ys, xs = a - 1, b - 1
            want = torch.zeros((n, n), device="cuda")
            want[max(-ys, 0):n - max(ys, 0), max(-xs, 0):n - max(xs, 0)] = \
                x[max(ys, 0):n - max(-ys, 0), max(xs, 0):n - max(-xs, 0)]
            assert torch.equal(p[:, :, a, b], want)
    del x, p
    torch.cuda.empty_cache()


def test_config5_laplacian3d_full_size(kex):
    """3-D 7-point Laplacian on 2048^3 (34 GB in, 34 GB out; > 2^32 elements: 64-bit offsets everywhere)."""
    n = 2048
    lap3 = kex.kmap(kernel_size=(3, 3, 3), padding="valid", relative=True)(
        lambda v: v[1, 0, 0] + v[-1, 0, 0] + v[0, 1, 0] + v[0, -1, 0] + v[0, 0, 1] + v[0, 0, -1] - 6 * v[0, 0, 0])
    x = torch.empty((n, n, n), device="cuda")
    x.normal_(generator=torch.Generator(device="cuda").manual_seed(2))
    y = lap3(x)
    assert _last_kernel() == "stencil3d_star_kernel" and y.shape == (n - 2,) * 3
    taps = [(2, 1, 1), (0, 1, 1), (1, 2, 1), (1, 0, 1), (1, 1, 2), (1, 1, 0), (1, 1, 1)]
    coefs = [1, 1, 1, 1, 1, 1, -6]
    for z0 in (0, 1023, n - 6):                                     # slabs of 4 output planes, incl. beyond 2^32 elements
        slab = x[z0:z0 + 6].cpu().numpy()
        want = ko.affine_map(slab, (3, 3, 3), (1, 1, 1), ((0, 0),) * 3, taps, coefs)
        scale = ko.affine_map(np.abs(slab), (3, 3, 3), (1, 1, 1), ((0, 0),) * 3, taps, np.abs(coefs))
        err = np.abs(y[z0:z0 + 4].cpu().numpy().astype(np.float64) - want)
        assert np.all(err <= 1e-5 * scale + 1e-30)
    # linearity, in place to stay inside HBM: L(2x) == 2 L(x) bit for bit
    x.mul_(2.0)
    y2 = lap3(x)
    y.mul_(2.0)
    assert torch.equal(y2, y)
    del y2
    # constants are annihilated exactly
    x.fill_(3.0)
    y = lap3(x)
    assert float(y.abs().max()) == 0.0
    del x, y
    torch.cuda.empty_cache()


def test_empty_and_tiny_inputs(kex):
    """Ragged / degenerate sizes on every fast kernel family: outputs smaller than one tile, windows
    larger than the array (empty output), single rows and single columns."""
    rng = np.random.default_rng(3)
    for shape, ksize, strides, padding, fn, op in [
        ((3, 4), (3, 3), (1, 1), "valid", lambda v: np.sum(v), "sum"),          # one output row
        ((1, 8), (1, 3), (1, 2), "valid", lambda v: np.max(v), "max"),          # pooling on a single row
        ((5, 4), (2, 2), (2, 2), "valid", lambda v: np.max(v), "max"),          # odd height: last row dropped
        ((4, 4), (3, 3), (2, 2), "same", lambda v: np.mean(v), "mean"),
        ((7, 12), (3, 3), (3, 3), "valid", lambda v: np.min(v), "min"),
    ]:
        x = rng.standard_normal(shape).astype(np.float32)
        border = ko.resolve_padding(padding, ksize)
        want = ko.reduce_map(x, ksize, strides, border, op)
        got = kex.kmap(kernel_size=ksize, strides=strides, padding=padding)(fn)(x)
        assert got.shape == want.shape, (shape, ksize, strides)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
    # a window larger than the array: the interface refuses it like the reference (resolve_utils.py:176-184),
    # the engine factory returns the empty result
    with pytest.raises(AssertionError):
        kex.kmap(kernel_size=(3, 3))(lambda v: np.sum(v))(np.ones((2, 4), np.float32))
    e = kex.kernel_map({(lambda v: np.sum(v)): ()}, (2, 4), (3, 3), (1, 1), ((0, 0), (0, 0)))(np.ones((2, 4), np.float32))
    assert e.shape == (0, 2)
    # conv-as-kmap on an image narrower than one warp and shorter than one band
    x = rng.standard_normal((3, 3, 4)).astype(np.float32)
    w = rng.standard_normal((3, 3, 3)).astype(np.float32)
    got = kex.kmap(kernel_size=(3, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))(
        torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda())
    taps = list(itertools.product(range(3), range(3), range(3)))
    want = ko.affine_map(x, (3, 3, 3), (1, 1, 1), ((0, 0), (1, 1), (1, 1)), taps, [float(w[t]) for t in taps])
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    # zero-size arrays
    # zero-size arrays: rejected by the interface's kernel_size check as in the reference; the engine
    # factories return the empty result without launching anything
    z = kex.kernel_map({(lambda v: np.sum(v)): ()}, (0, 8), (1, 3), (1, 1), ((0, 0), (1, 1)))(np.zeros((0, 8), np.float32))
    assert z.shape == (0, 8)
