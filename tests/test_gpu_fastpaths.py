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


_WIDE = [(1, 4), (4, 4), (3, 5), (5, 3), (2, 5), (4, 1), (1, 7), (7, 1), (7, 7), (5, 2), (4, 3), (2, 4),
         (1, 6), (1, 8), (1, 9), (1, 11), (1, 13), (1, 15)]


@pytest.mark.parametrize("ki", range(len(_WIDE)))
@pytest.mark.parametrize("variant", ["f32-aligned", "i32-aligned", "f32-ragged", "f32-even"])
def test_stencil2d_every_window_shape_up_to_5x5_and_7x7(kex, ki, variant):
    ky, kx = _WIDE[ki]
    rng = np.random.default_rng(9100 + 10 * ki + len(variant))
    h = int(rng.integers(ky + 3, 90))
    w = {"f32-aligned": 260, "i32-aligned": 132, "f32-ragged": 131, "f32-even": 150}[variant]
    border = ((int(rng.integers(-1, ky)), int(rng.integers(-1, ky))), (int(rng.integers(-1, kx)), int(rng.integers(-1, kx))))
    taps = [t for t in itertools.product(range(ky), range(kx)) if rng.random() < 0.8] or [(0, 0)]
    if variant.startswith("i32"):
        x = rng.integers(-1000, 1000, (h, w)).astype(np.int32)
        coefs, bias = [int(c) or 2 for c in rng.integers(-9, 10, len(taps))], 3
    else:
        x = rng.standard_normal((h, w)).astype(np.float32)
        coefs, bias = [float(c) for c in rng.standard_normal(len(taps))], -0.25

    def fn(v):
        acc = bias
        for t, c in zip(taps, coefs):
            acc = acc + c * v[t]
        return acc

    got = kex.kernel_map({fn: ()}, (h, w), (ky, kx), (1, 1), border, False)(x)
    assert _last_kernel() == "stencil2d_kernel"
    _check_affine(got, x, (ky, kx), (1, 1), border, taps, coefs, bias)


@pytest.mark.parametrize("ksize", [(2, 2), (3, 3), (1, 3), (3, 1), (5, 5), (1, 7), (2, 4), (1, 1)])
@pytest.mark.parametrize("op", ["max", "min", "mean"])
@pytest.mark.parametrize("dtype", ["f32", "i32"])
def test_stencil2d_window_reductions_stride1(kex, ksize, op, dtype):
    """np.max / np.min / np.mean over stride-1 windows (max / min filters, box blurs): zero padding takes part."""
    int_mean = op == "mean" and dtype == "i32"   # the mean of an int32 array is float32 (JAX promotion): generic kernel
    rng = np.random.default_rng(9300 + 7 * ksize[0] + ksize[1])
    for (h, w), pad in [((45, 260), "same"), ((33, 131), "valid"), ((20, 64), ((1, 0), (2, 1)))]:
        if ksize[0] > h or ksize[1] > w:
            continue
        border = ko.resolve_padding(pad, ksize) if isinstance(pad, str) else tuple(
            (min(l, k - 1), min(r, k - 1)) for (l, r), k in zip(pad, ksize))
        x = (rng.standard_normal((h, w)).astype(np.float32) if dtype == "f32"
             else rng.integers(-1000, 1000, (h, w)).astype(np.int32))
        fn = {"max": lambda v: np.max(v), "min": lambda v: np.min(v), "mean": lambda v: np.mean(v)}[op]
        got = kex.kernel_map({fn: ()}, (h, w), ksize, (1, 1), border, False)(x)
        if not int_mean:
            assert _last_kernel() == "stencil2d_kernel"
        want = ko.reduce_map(x, ksize, (1, 1), border, op)
        assert got.shape == want.shape and got.dtype == want.dtype == (np.float32 if op == "mean" else x.dtype)
        if op == "mean":
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-5 * float(np.abs(x).max()) * ksize[0] * ksize[1])
        else:
            np.testing.assert_array_equal(got, want)


def _gauss(ky, kx, rng):
    gy = np.exp(-0.5 * np.linspace(-(ky - 1) / 2, (ky - 1) / 2, ky) ** 2 / 1.7).astype(np.float32)
    gx = np.exp(-0.5 * np.linspace(-(kx - 1) / 2, (kx - 1) / 2, kx) ** 2 / 0.9).astype(np.float32)
    w = np.outer(gy, gx).astype(np.float32)
    return (w / w.sum()).astype(np.float32)


@pytest.mark.parametrize("ksize", [(4, 4), (5, 5), (3, 5), (5, 3), (7, 7), (9, 9), (11, 11), (13, 13), (15, 15)])
@pytest.mark.parametrize("kind", ["gauss", "signed", "mean", "max", "sum-i32"])
def test_stencil2d_separable_windows_up_to_15x15(kex, ksize, kind, monkeypatch):
    """README.md:277-294 Gaussian blur (w = outer(g, g) / sum, any kernel size) and other outer-product tables run as
    KY + KX FMAs per output; box sums / means and max / min as row-then-column folds; vs the dense oracle."""
    ky, kx = ksize
    rng = np.random.default_rng(9500 + 16 * ky + kx)
    taps = list(itertools.product(range(ky), range(kx)))
    for (h, w_), pad in [((40, 260), "same"), ((37, 131), "valid"), ((48, 72), tuple(((k - 1) // 2, 0) for k in ksize))]:
        border = ko.resolve_padding(pad, ksize) if isinstance(pad, str) else pad
        if kind == "sum-i32":
            x = rng.integers(-1000, 1000, (h, w_)).astype(np.int32)
            got = kex.kernel_map({(lambda v: np.sum(v)): ()}, (h, w_), ksize, (1, 1), border, False)(x)
            assert _last_kernel() == ("stencil2d_kernel")
            np.testing.assert_array_equal(got, ko.reduce_map(x, ksize, (1, 1), border, "sum"))
            continue
        x = rng.standard_normal((h, w_)).astype(np.float32)
        if kind in ("mean", "max"):
            fn = (lambda v: np.mean(v)) if kind == "mean" else (lambda v: np.max(v))
            got = kex.kernel_map({fn: ()}, (h, w_), ksize, (1, 1), border, False)(x)
            assert _last_kernel() == "stencil2d_kernel"
            want = ko.reduce_map(x, ksize, (1, 1), border, kind)
            if kind == "max":
                np.testing.assert_array_equal(got, want)
            else:
                np.testing.assert_allclose(got, want, rtol=0, atol=1e-5 * float(np.abs(x).max()))
            continue
        wt = _gauss(ky, kx, rng)
        if kind == "signed":   # derivative-of-Gaussian style: sign changes in one factor (zero entries would be
            wt = np.outer(wt[:, kx // 2], np.linspace(-1, 1.37, kx)).astype(np.float32)   # dropped by the tracer)
        got = kex.kernel_map({(lambda v: np.sum(v * wt)): ()}, (h, w_), ksize, (1, 1), border, False)(x)
        # square odd windows >= 7x7 over 16-byte aligned rows take the TMA row-ring kernel (kx_sep2d.cu)
        ring = ky == kx and ky >= 7 and ky % 2 == 1 and w_ % 4 == 0
        assert _last_kernel() == ("sep2d_kernel" if ring else "stencil2d_kernel")
        _check_affine(got, x, ksize, (1, 1), border, taps, [float(wt[t]) for t in taps])
        if ring:   # ... and must agree bit for bit with stencil2d_kernel's separable path (same accumulation order)
            monkeypatch.setenv("KX_NO_SEP2D", "1")
            got2 = kex.kernel_map({(lambda v: np.sum(v * wt)): ()}, (h, w_), ksize, (1, 1), border, False)(x)
            monkeypatch.delenv("KX_NO_SEP2D")
            assert _last_kernel() == "stencil2d_kernel"
            np.testing.assert_array_equal(got, got2)


def test_stencil2d_dense_tables_are_not_mistaken_for_outer_products(kex):
    """A table one entry away from an outer product must take the dense kernel (<= 7x7) or the generic one."""
    rng = np.random.default_rng(77)
    x = rng.standard_normal((40, 132)).astype(np.float32)
    for k in (5, 7, 9):
        wt = _gauss(k, k, rng)
        wt[1, 2] *= np.float32(1.001)
        taps = list(itertools.product(range(k), range(k)))
        got = kex.kernel_map({(lambda v: np.sum(v * wt)): ()}, x.shape, (k, k), (1, 1), ((1, 1), (2, 2)), False)(x)
        assert _last_kernel() == ("stencil2d_kernel" if k <= 7 else "generic_map_kernel")
        _check_affine(got, x, (k, k), (1, 1), ((1, 1), (2, 2)), taps, [float(wt[t]) for t in taps])


@pytest.mark.parametrize("k", [6, 9, 11, 15])
def test_fir_filter_on_a_1d_array(kex, k):
    """1-D arrays are one-row windows: FIR filters / moving averages of up to 15 taps on the row-march kernel."""
    rng = np.random.default_rng(9700 + k)
    x = rng.standard_normal(100003).astype(np.float32)
    w = rng.standard_normal(k).astype(np.float32)
    taps = [(t,) for t in range(k)]
    got = kex.kmap(kernel_size=(k,), padding="same")(lambda v: np.sum(v * w))(x)
    assert _last_kernel() == "stencil2d_kernel"
    _check_affine(got, x, (k,), (1,), ko.resolve_padding("same", (k,)), taps, [float(c) for c in w])
    xi = rng.integers(-1000, 1000, 4096).astype(np.int32)
    goti = kex.kmap(kernel_size=(k,), padding="valid")(lambda v: np.sum(v))(xi)
    assert _last_kernel() == "stencil2d_kernel"
    np.testing.assert_array_equal(goti, ko.reduce_map(xi, (k,), (1,), ((0, 0),), "sum"))
    gotm = kex.kmap(kernel_size=(k,), padding="same")(lambda v: np.max(v))(xi)
    assert _last_kernel() == "stencil2d_kernel"
    np.testing.assert_array_equal(gotm, ko.reduce_map(xi, (k,), (1,), ko.resolve_padding("same", (k,)), "max"))


def test_window_max_propagates_nan_like_numpy(kex):
    """np.max / jnp.max return NaN when any window element is NaN, whichever position it has in the window."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((40, 132)).astype(np.float32)
    x[7, 9] = x[20, 100] = x[39, 131] = np.nan
    for ksize, strides in [((3, 3), (1, 1)), ((2, 2), (2, 2)), ((3, 5), (1, 1)), ((3, 3), (1, 3))]:
        for op, fn in [("max", lambda v: np.max(v)), ("min", lambda v: np.min(v))]:
            got = kex.kernel_map({fn: ()}, x.shape, ksize, strides, ((1, 1), (1, 1)), False)(x)
            want = ko.reduce_map(x, ksize, strides, ((1, 1), (1, 1)), op)
            assert np.isnan(want).sum() > 0
            np.testing.assert_array_equal(got, want)   # NaNs compare equal position-wise here


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


@pytest.mark.parametrize("seed", range(24))
def test_stencil3d_random_geometry(kex, seed):
    rng = np.random.default_rng(3000 + seed)
    ks = [(3, 3, 3), (2, 2, 2), (3, 3, 3), (5, 5, 5)][seed % 4]
    star = seed % 4 in (2, 3)
    d = int(rng.integers(ks[0], 14))
    h = int(rng.integers(ks[1], 40))
    w = int(rng.integers(ks[2], 150))
    if seed % 3 == 0:
        w = (w // 4 + 1) * 4
    border = tuple((int(rng.integers(-1, 3)), int(rng.integers(-1, 3))) for _ in range(3))
    if any(o <= 0 for o in ko.output_shape((d, h, w), ks, (1, 1, 1), border)):
        border = ((0, 0),) * 3
    c = tuple((k - 1) // 2 for k in ks)
    all_taps = list(itertools.product(*[range(k) for k in ks]))
    if star:
        taps = [t for t in all_taps if sum(int(t[i] == c[i]) for i in range(3)) >= 2]
    else:
        taps = [t for t in all_taps if rng.random() < 0.8] or [c]
    if seed % 2:
        x = rng.standard_normal((d, h, w)).astype(np.float32)
        coefs = [float(v) for v in rng.standard_normal(len(taps))]
        bias = -0.5
    else:
        x = rng.integers(-1000, 1000, (d, h, w)).astype(np.int32)
        coefs = [int(v) or 2 for v in rng.integers(-9, 10, len(taps))]
        bias = 0

    def fn(v):
        acc = bias
        for t, cf in zip(taps, coefs):
            acc = acc + cf * v[t]
        return acc

    got = kex.kernel_map({fn: ()}, (d, h, w), ks, (1, 1, 1), border, False)(x)
    assert _last_kernel() in ("stencil3d_kernel", "stencil3d_star_kernel", "conv_fwd_rows_kernel")
    _check_affine(got, x, ks, (1, 1, 1), border, taps, coefs, bias)


_DENSE3D = [((9, 37, 132), "same"), ((3, 20, 64), "same"), ((5, 18, 260), "valid"), ((4, 70, 1028), "same"),
            ((7, 33, 136), ((2, 0), (0, 2), (1, 1))), ((6, 21, 72), ((1, 2), (2, 1), (0, 2))),
            ((8, 19, 68), ((-1, 1), (1, -1), (2, 0))), ((1, 16, 32), "same"), ((2, 16, 32), ((2, 2), (1, 1), (1, 1))),
            ((12, 40, 516), ((0, 1), (1, 1), (1, 1)))]


@pytest.mark.parametrize("case", range(len(_DENSE3D)))
@pytest.mark.parametrize("dtype", ["f32", "i32"])
def test_dense3d_runs_on_the_conv_row_ring(kex, case, dtype, monkeypatch):
    """Dense 27-point windows over a volume on the conv row-ring kernel with the three z planes as channels (the path
    of device-resident weights; host weights reach it with KX_NO_DENSE3D_PLANES=1); output planes at a padded z face
    use the 1- and 2-plane instantiations (kx_conv2d.cu run_dense3d)."""
    monkeypatch.setenv("KX_NO_DENSE3D_PLANES", "1")
    shape, pad = _DENSE3D[case]
    rng = np.random.default_rng(7000 + case)
    ks = (3, 3, 3)
    border = ko.resolve_padding(pad, ks)
    taps = list(itertools.product(range(3), range(3), range(3)))
    if dtype == "f32":
        x = rng.standard_normal(shape).astype(np.float32)
        w = rng.standard_normal(ks).astype(np.float32)
        bias = 0.25
    else:
        x = rng.integers(-1000, 1000, shape).astype(np.int32)
        w = rng.integers(-9, 10, ks).astype(np.int32)
        w[w == 0] = 3
        bias = 0
    coefs = [w[t].item() for t in taps]

    def fn(v):
        acc = bias
        for t, cf in zip(taps, coefs):
            acc = acc + cf * v[t]
        return acc

    got = kex.kernel_map({fn: ()}, shape, ks, (1, 1, 1), border, False)(x)
    assert _last_kernel() == "conv_fwd_rows_kernel"
    _check_affine(got, x, ks, (1, 1, 1), border, taps, coefs, bias)
    if dtype == "f32" and shape[0] >= 3:   # runtime weights on the device (conv3d-as-kmap) and a leading batch axis
        conv = kex.kmap(kernel_size=ks, padding=pad)(lambda v, w: np.sum(v * w))
        got = conv(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda())
        assert _last_kernel() == "conv_fwd_rows_kernel"
        _check_affine(got.cpu().numpy(), x, ks, (1, 1, 1), border, taps, coefs)
        xb = np.stack([x, x[::-1].copy()])
        gotb = kex.kmap(kernel_size=(1,) + ks, padding=((0, 0),) + tuple(border))(lambda v: np.sum(v[0] * w))(xb)
        assert _last_kernel() == "conv_fwd_rows_kernel"
        for b in range(2):
            _check_affine(gotb[b], xb[b], ks, (1, 1, 1), border, taps, coefs)


def test_dense3d_input_vjp_runs_on_the_conv_row_ring(kex):
    """dx of a dense 3-D stencil with device weights = the same kernel on the reversed coefficient table."""
    rng = np.random.default_rng(17)
    shape, ks = (6, 24, 68), (3, 3, 3)
    taps = list(itertools.product(range(3), range(3), range(3)))
    w = rng.standard_normal(ks).astype(np.float32)
    x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda().requires_grad_()
    wd = torch.from_numpy(w).cuda()
    y = kex.kmap(kernel_size=ks, padding="same")(lambda v, w: np.sum(v * w))(x, wd)
    g = rng.standard_normal(shape).astype(np.float32)
    (dx,) = torch.autograd.grad(y, x, torch.from_numpy(g).cuda())
    assert _last_kernel() == "conv_fwd_rows_kernel"
    # transpose of a 'same' 3x3x3 stencil = the stencil with flipped taps
    flipped = [tuple(2 - i for i in t) for t in taps]
    _check_affine(dx.cpu().numpy(), g, ks, (1, 1, 1), ((1, 1),) * 3, flipped, [float(w[t]) for t in taps])


def test_conv_as_kmap_config3a_shape(kex):
    """README conv example / BASELINE config 3a at reduced size: kernel (C,3,3), padding
    (valid, same, same), runtime weights on the device, vs the oracle (fp32 tolerance)."""
    rng = np.random.default_rng(11)
    C, H, W = 3, 70, 132
    x = rng.standard_normal((C, H, W)).astype(np.float32)
    w = rng.standard_normal((C, 3, 3)).astype(np.float32)
    conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    taps = list(itertools.product(range(C), range(3), range(3)))
    b = ((0, 0), (1, 1), (1, 1))
    got_host_w = conv(x, w)                                   # weights known on the host
    assert _last_kernel() in ("conv_fwd_kernel", "conv_fwd_rows_kernel") and got_host_w.shape == (1, H, W)
    _check_affine(got_host_w, x, (C, 3, 3), (1, 1, 1), b, taps, [float(w[t]) for t in taps])
    got_dev_w = conv(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda())   # device weights
    assert _last_kernel() in ("conv_fwd_kernel", "conv_fwd_rows_kernel")
    _check_affine(got_dev_w.cpu().numpy(), x, (C, 3, 3), (1, 1, 1), b, taps, [float(w[t]) for t in taps])


def test_patch_extraction_fast_path(kex):
    """README get_3x3_patches / BASELINE config 3b semantics, bit exact."""
    rng = np.random.default_rng(5)
    for shape, ks, pad in [((37, 53), (3, 3), "same"), ((64, 64), (3, 3), "valid"), ((1000,), (3,), "same"),
                           ((9, 10, 11), (2, 3, 2), "same")]:
        x = rng.standard_normal(shape).astype(np.float32)
        for relative in (False, True):
            got = kex.kmap(kernel_size=ks, padding=pad, relative=relative)(lambda v: v)(x)
            assert _last_kernel() in ("patch_kernel", "patch2d_kernel")
            border = ko.resolve_padding(pad, ks)
            want = ko.patch_map(x, ks, (1,) * len(ks), border, relative)
            np.testing.assert_array_equal(got, want)
    xi = np.arange(1, 26, dtype=np.int32).reshape(5, 5)
    p = kex.kmap(kernel_size=(3, 3), padding="same")(lambda v: v)(xi)
    np.testing.assert_array_equal(p[0, 0], [[0, 0, 0], [0, 1, 2], [0, 6, 7]])   # README.md:183-186


def test_weight_gradient_matches_oracle(kex):
    """tests/test_interface.py:345-348: d/dw sum(conv(x, w)) and d/dx, fp32, atol 1e-3 there."""
    rng = np.random.default_rng(9)
    C, H = 3, 16
    x = rng.standard_normal((C, H, H)).astype(np.float32)
    w = rng.standard_normal((C, 3, 3)).astype(np.float32)
    xt = torch.from_numpy(x).cuda().requires_grad_()
    wt = torch.from_numpy(w).cuda().requires_grad_()
    conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    g = torch.from_numpy(rng.standard_normal((1, H, H)).astype(np.float32)).cuda()
    y = conv(xt, wt)
    (y * g).sum().backward()
    b = ((0, 0), (1, 1), (1, 1))
    from vjp_ref import assert_within, dw_reference, dx_reference
    taps = list(itertools.product(range(C), range(3), range(3)))
    dw, dw_mag = dw_reference(x, g.cpu().numpy(), (C, 3, 3), (1, 1, 1), b, taps)
    assert_within(wt.grad.cpu().numpy().reshape(-1), dw, dw_mag, "dW")      # 1e-5 * sum |g x|
    # dx = flipped-tap stencil of g: compare against finite structure via float64 numpy
    dx = np.zeros_like(x, dtype=np.float64)
    gp = np.pad(g.cpu().numpy()[0].astype(np.float64), 1)
    for c in range(C):
        for a in range(3):
            for bb in range(3):
                # x[c, i+a-1, j+bb-1] contributes w[c,a,bb] to out[i,j]
                dx[c] += w[c, a, bb] * gp[2 - a:2 - a + H, 2 - bb:2 - bb + H]
    dx2, dx_mag = dx_reference(g.cpu().numpy(), (C, H, H), (C, 3, 3), (1, 1, 1), b, taps, [float(w[t]) for t in taps])
    np.testing.assert_allclose(dx, dx2, rtol=1e-12, atol=1e-12)               # the two float64 transposes agree
    assert_within(xt.grad.cpu().numpy(), dx2, dx_mag, "dx")                  # 1e-5 * sum |c g|


@pytest.mark.parametrize("case", range(12))
def test_conv_rows_kernels_geometry(kex, case):
    """The per-warp cp.async ring kernels (aligned rows): every x border class (lx = 0, 1, 2),
    y borders, several warps / column blocks / bands, warp-edge halo lanes, fp32 and int32,
    forward against the oracle and the weight gradient against the oracle's correlation."""
    rng = np.random.default_rng(4200 + case)
    C = [3, 2, 4, 3][case % 4]
    H = [200, 97, 150, 64][case % 4]
    W = [1028, 516, 132, 260][(case // 2) % 4]
    bx = [(1, 1), (0, 0), (2, 0), (0, 2), (2, 2), (1, 0)][case % 6]
    by = [(1, 1), (0, 0), (2, 1), (0, 2)][(case // 3) % 4]
    border = ((0, 0), by, bx)
    integer = case % 5 == 4
    taps = list(itertools.product(range(C), range(3), range(3)))
    if integer:
        x = rng.integers(-50, 50, (C, H, W)).astype(np.int32)
        w = rng.integers(-5, 6, (C, 3, 3)).astype(np.int32)
    else:
        x = rng.standard_normal((C, H, W)).astype(np.float32)
        w = rng.standard_normal((C, 3, 3)).astype(np.float32)
    op = kex.kernel_map({(lambda v, w: np.sum(v * w)): ()}, (C, H, W), (C, 3, 3), (1, 1, 1), border, False)
    xt = torch.from_numpy(x).cuda()
    wt = torch.from_numpy(w).cuda()
    if not integer:
        wt.requires_grad_()
    got = op(xt, wt)
    assert _last_kernel() == "conv_fwd_rows_kernel"
    coefs = [w[t].item() for t in taps]
    _check_affine(got.detach().cpu().numpy(), x, (C, 3, 3), (1, 1, 1), border, taps, coefs)
    if not integer:
        g = rng.standard_normal(tuple(got.shape)).astype(np.float32)
        (dw,) = torch.autograd.grad(got, wt, torch.from_numpy(g).cuda())
        from vjp_ref import assert_within, dw_reference
        want, mag = dw_reference(x, g, (C, 3, 3), (1, 1, 1), border, taps)
        assert_within(dw.cpu().numpy().reshape(-1), want, mag, "dW")          # 1e-5 * sum |g x|


def test_conv_rows_batched_and_host_weights(kex):
    rng = np.random.default_rng(4300)
    B, C, H, W = 3, 3, 80, 384
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = rng.standard_normal((C, 3, 3)).astype(np.float32)
    border = ((0, 0), (0, 0), (1, 1), (1, 1))
    op = kex.kernel_map({(lambda v, w: np.sum(v[0] * w)): ()}, (B, C, H, W), (1, C, 3, 3), (1, 1, 1, 1), border, False)
    got = op(torch.from_numpy(x).cuda(), w)     # host weights -> constant-bank coefficients
    assert _last_kernel() == "conv_fwd_rows_kernel"
    taps = list(itertools.product(range(C), range(3), range(3)))
    for b in range(B):
        _check_affine(got[b].cpu().numpy(), x[b], (C, 3, 3), (1, 1, 1), border[1:], taps, [float(w[t]) for t in taps])


@pytest.mark.parametrize("case", range(len(_DENSE3D)))
@pytest.mark.parametrize("dtype", ["f32", "i32"])
def test_plane_stationary_dense3d_kernel(kex, case, dtype):
    """Dense 27-point windows with host weights: stencil3d_dense_kernel (kx_stencil3d_dense.cu) against the oracle."""
    shape, pad = _DENSE3D[case]
    rng = np.random.default_rng(7100 + case)
    ks = (3, 3, 3)
    border = ko.resolve_padding(pad, ks)
    taps = list(itertools.product(range(3), range(3), range(3)))
    if dtype == "f32":
        x = rng.standard_normal(shape).astype(np.float32)
        coefs, bias = [float(c) for c in rng.standard_normal(27)], 0.25
    else:
        x = rng.integers(-1000, 1000, shape).astype(np.int32)
        coefs, bias = [int(c) or 3 for c in rng.integers(-9, 10, 27)], 2

    def fn(v):
        acc = bias
        for t, cf in zip(taps, coefs):
            acc = acc + cf * v[t]
        return acc

    got = kex.kernel_map({fn: ()}, shape, ks, (1, 1, 1), border, False)(x)
    assert _last_kernel() == "stencil3d_dense_kernel"
    _check_affine(got, x, ks, (1, 1, 1), border, taps, coefs, bias)


@pytest.mark.parametrize("k,pad", [(3, "same"), (3, "valid"), (15, "same"), (6, ((2, 1),)), (9, "valid")])
def test_folded_1d_map_equals_the_one_row_launch(kex, k, pad, monkeypatch):
    """kernex_b200/fold1d.py: a large 1-D array as rows + seam fix-up must equal the plain one-row launch bit for bit
    (KX_NO_FOLD_1D=1), for weighted sums, integer max and int -> float promotion."""
    rng = np.random.default_rng(9900 + k)
    x = torch.from_numpy(rng.standard_normal(1 << 22).astype(np.float32)).cuda()
    xi = torch.from_numpy(rng.integers(-1000, 1000, 1 << 22).astype(np.int32)).cuda()
    w = rng.standard_normal(k).astype(np.float32)
    fs = [kex.kmap(kernel_size=(k,), padding=pad)(lambda v: np.sum(v * w)),
          kex.kmap(kernel_size=(k,), padding=pad)(lambda v: np.max(v)),
          kex.kmap(kernel_size=(k,), padding=pad)(lambda v: 0.5 * v[0] + v[k - 1])]
    monkeypatch.setenv("KX_NO_FOLD_1D", "1")
    want = [fs[0](x), fs[1](xi), fs[2](xi)]
    monkeypatch.delenv("KX_NO_FOLD_1D")
    got = [kex.kmap(kernel_size=(k,), padding=pad)(lambda v: np.sum(v * w))(x),
           kex.kmap(kernel_size=(k,), padding=pad)(lambda v: np.max(v))(xi),
           kex.kmap(kernel_size=(k,), padding=pad)(lambda v: 0.5 * v[0] + v[k - 1])(xi)]
    for g_, w_ in zip(got, want):
        assert g_.shape == w_.shape and g_.dtype == w_.dtype and torch.equal(g_, w_)
    # ... and against the oracle on the first and last 4096 outputs (true array ends = zero padding)
    border = ko.resolve_padding(pad, (k,))
    taps = [(i,) for i in range(k)]
    n_out = got[0].shape[0]
    head = ko.affine_map(x[:4096 + k].cpu().numpy(), (k,), (1,), (border[0][0:1] + (0,),), taps, [float(c) for c in w])
    m = min(4096, head.shape[0])
    scale = ko.affine_map(np.abs(x[:4096 + k].cpu().numpy()), (k,), (1,), (border[0][0:1] + (0,),), taps, np.abs(w))
    assert np.all(np.abs(got[0][:m].cpu().numpy() - head[:m]) <= 1e-5 * scale[:m] + 1e-30)
    tail = ko.affine_map(x[-(4096 + k):].cpu().numpy(), (k,), (1,), ((0,) + border[0][1:2],), taps, [float(c) for c in w])
    scale = ko.affine_map(np.abs(x[-(4096 + k):].cpu().numpy()), (k,), (1,), ((0,) + border[0][1:2],), taps, np.abs(w))
    m = min(4096, tail.shape[0])
    assert np.all(np.abs(got[0][n_out - m:].cpu().numpy() - tail[-m:]) <= 1e-5 * scale[-m:] + 1e-30)


def test_vector_valued_1d_functions_do_not_take_the_fold(kex):
    """A GATHER function (patch extraction) on a large 1-D device array: the fold only serves scalar functions, the
    patches must still be bit exact (BASELINE test_benchmark_patch shape: k=3, 'same')."""
    n = 1 << 22
    x = torch.arange(n, dtype=torch.int32, device="cuda")
    p = kex.kmap(kernel_size=(3,), padding="same")(lambda v: v)(x)
    assert p.shape == (n, 3)
    want = torch.stack([torch.cat([torch.zeros(1, dtype=torch.int32, device="cuda"), x[:-1]]), x,
                        torch.cat([x[1:], torch.zeros(1, dtype=torch.int32, device="cuda")])], dim=1)
    assert torch.equal(p, want)


@pytest.mark.parametrize("k", [7, 9, 11, 13, 15])
def test_sep2d_row_ring_geometry(kex, k, monkeypatch):
    """kx_sep2d.cu on the shapes that stress its ring: several bands (zero-filled stages above / below the array), several
    warps and column blocks with a ragged last block, every window misalignment (left x border 0..3), a bias, a batch axis,
    the README Gaussian with a mean-style final division -- against the oracle and bit for bit against stencil2d_kernel."""
    rng = np.random.default_rng(9700 + k)
    g1 = np.exp(-0.5 * np.linspace(-(k - 1) / 2, (k - 1) / 2, k) ** 2 / 2.0)
    wt = np.outer(g1, g1)
    wt = (wt / wt.sum()).astype(np.float32)
    taps = list(itertools.product(range(k), range(k)))
    for (h, w_), border in [((300, 644), ((k // 2, k // 2), (0, k - 1))), ((70, 1028), ((0, 0), (1, 2))),
                            ((1100, 516), ((k - 1, 0), (2, k // 2))), ((40, 260), ((3, 3), (3, 3)))]:
        x = rng.standard_normal((h, w_)).astype(np.float32)
        fn = lambda v: np.sum(v * wt) + 0.25      # noqa: E731
        got = kex.kernel_map({fn: ()}, (h, w_), (k, k), (1, 1), border, False)(x)
        assert _last_kernel() == "sep2d_kernel"
        _check_affine(got, x, (k, k), (1, 1), border, taps, [float(wt[t]) for t in taps], 0.25)
        monkeypatch.setenv("KX_NO_SEP2D", "1")
        got2 = kex.kernel_map({(lambda v: np.sum(v * wt) + 0.25): ()}, (h, w_), (k, k), (1, 1), border, False)(x)
        monkeypatch.delenv("KX_NO_SEP2D")
        np.testing.assert_array_equal(got, got2)
    xb = rng.standard_normal((3, 50, 132)).astype(np.float32)
    f = kex.vmap(kex.kmap(kernel_size=(k, k), padding="same")(lambda v: np.sum(v * wt)))
    gotb = f(torch.from_numpy(xb).cuda()).cpu().numpy()
    assert _last_kernel() == "sep2d_kernel"
    b = ko.resolve_padding("same", (k, k))
    for c in range(3):
        _check_affine(gotb[c], xb[c], (k, k), (1, 1), b, taps, [float(wt[t]) for t in taps])
