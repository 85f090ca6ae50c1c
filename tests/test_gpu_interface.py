"""GPU parity at the interface level: the scenarios of the reference's tests/test_interface.py
(re-stated with NumPy window functions), its error conventions, and seeded random
engine-vs-oracle property checks over (shape, kernel, stride, border, function class)."""
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


# ---- tests/test_interface.py:203-239 (test_mesh) ------------------------------------------
@pytest.mark.parametrize("map_kind", ["vmap", "lmap", "pmap"])
def test_mesh(kex, map_kind):
    Mesh = kex.kmap(kernel_size=(1,), relative=True, padding="same", map_kind=map_kind)
    Mesh[0] = lambda x: x[0] * 10
    Mesh[1] = lambda x: x[0] * -10
    Mesh[2:] = lambda x: x[0] * 100
    np.testing.assert_array_equal(Mesh(np.arange(1, 11)), [10, -20, 300, 400, 500, 600, 700, 800, 900, 1000])

    Mesh = kex.kmap(kernel_size=(3, 3), relative=True, padding="same")
    Mesh[0, 0] = lambda x: x[0, 0] * 10
    Mesh[1, 1:] = lambda x: x[0, 0] * -10
    Mesh[2:] = lambda x: x[0, 0] * 100
    np.testing.assert_array_equal(
        Mesh(np.arange(1, 26).reshape(5, 5)),
        [[10, 20, 30, 40, 50], [60, -70, -80, -90, -100], [1100, 1200, 1300, 1400, 1500],
         [1600, 1700, 1800, 1900, 2000], [2100, 2200, 2300, 2400, 2500]])


# ---- tests/test_interface.py:242-312 (composition with an outer scan = repeated application) ----
def test_repeated_application_of_kscan_and_kmap(kex):
    A = np.ones([10], dtype=np.float32)
    F = kex.kscan(kernel_size=(1,), relative=True, padding="same")
    F[1:-1] = lambda x: x[0] + 1
    F[0] = F[-1] = lambda x: x[0]
    np.testing.assert_allclose(F(A), [1, 2, 2, 2, 2, 2, 2, 2, 2, 1])
    u = A
    for _ in range(9):
        u = F(u)
    np.testing.assert_allclose(u, [1, 10, 10, 10, 10, 10, 10, 10, 10, 1])

    F = kex.kscan(kernel_size=(3,), relative=True, padding="same")
    F[1:-1] = lambda x: x[0] + x[-1] + x[1]
    F[0] = F[-1] = lambda x: x[0]
    np.testing.assert_allclose(F(A), [1.0, 3.0, 5.0, 7.0, 9.0, 11.0, 13.0, 15.0, 17.0, 1.0])
    u = A
    for _ in range(2):
        u = F(u)
    np.testing.assert_allclose(u, [1.0, 9.0, 21.0, 37.0, 57.0, 81.0, 109.0, 141.0, 159.0, 1.0])

    F = kex.kmap(kernel_size=(3,), relative=True, padding="same")
    F[1:-1] = lambda x: x[0] + x[-1] + x[1]
    F[0] = F[-1] = lambda x: x[0]
    np.testing.assert_allclose(F(A), [1.0, 3.0, 3.0, 3.0, 3.0, 3.0, 3.0, 3.0, 3.0, 1.0])
    u = A
    for _ in range(2):
        u = F(u)
    np.testing.assert_allclose(u, [1.0, 7.0, 9, 9, 9, 9, 9, 9, 7, 1.0])


# ---- tests/test_interface.py:106-200 (linear convection: sscan and kscan vs the NumPy FDM) ----
def _fdm_convection(nt, nx, tmax, xmax, c):
    dt, dx = tmax / (nt - 1), xmax / (nx - 1)
    u = np.zeros((nx, nt))
    u[0, :] = u[nx - 1, :] = 1
    for i in range(1, nx - 1):
        u[i, 0] = 2 if (nx - 1) / 4 < i < (nx - 1) / 2 else 1
    for n in range(0, nt - 1):
        for i in range(1, nx - 1):
            u[i, n + 1] = u[i, n] - c * (dt / dx) * (u[i, n] - u[i - 1, n])
    return u


def test_linear_convection_sscan_and_kscan(kex):
    tmax, xmax, nt, nx, c = 0.5, 2.0, 151, 51, 0.5
    dt, dx = tmax / (nt - 1), xmax / (nx - 1)
    fdm = _fdm_convection(nt, nx, tmax, xmax, c)
    u = np.ones([nx, nt])
    for i in range(1, nx - 1):
        u[i, 0] = 2 if (nx - 1) / 4 < i < (nx - 1) / 2 else 1
    u = u.T.astype(np.float32)

    F = kex.sscan(kernel_size=(3, 3), named_axis={0: "n", 1: "i"}, relative=True)
    F[:] = lambda u: u["i", "n"]
    F[1:, 1:-1] = lambda u: u["i", "n-1"] - (c * dt / dx) * (u["i", "n-1"] - u["i-1", "n-1"])
    np.testing.assert_allclose(F(u), fdm.T, atol=1e-3)

    F = kex.kscan(kernel_size=(3, 3), padding=((1, 1), (1, 1)), named_axis={0: "n", 1: "i"}, relative=True)
    F[:, 0] = F[:, -1] = lambda u: 1
    F[0:1, int((nx - 1) / 4) + 1: int((nx - 1) / 2)] = lambda u: 2
    F[:, : int((nx - 1) / 4) + 1] = F[:, int((nx - 1) / 2):] = lambda u: 1
    F[1:, 1:-1] = lambda u: u["i", "n-1"] - (c * dt / dx) * (u["i", "n-1"] - u["i-1", "n-1"])
    assert "kscan(kernel_size=(3, 3)" in repr(F)
    np.testing.assert_allclose(F(u), fdm.T, atol=1e-3)


# ---- docstring examples of the four interface classes ------------------------------------------
def test_docstring_examples(kex):
    x = np.array([1, 2, 3, 4, 5], dtype=np.int32)
    np.testing.assert_array_equal(kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))(x), [6, 9, 12])       # :443-455
    np.testing.assert_array_equal(kex.kscan(kernel_size=(3,))(lambda v: np.sum(v))(x), [6, 13, 22])     # :379-391
    f = lambda v: v[0] + v[1] + v[-1]
    np.testing.assert_array_equal(kex.sscan(kernel_size=(3,), offset=((2, 1),), relative=True)(f)(x), [1, 2, 9, 18, 5])
    np.testing.assert_array_equal(kex.smap(kernel_size=(3,), offset=((1, 1),), relative=True)(f)(x),
                                  ko.offset_kernel_map({f: ()}, (5,), (3,), (1,), ((1, 1),), True)(x))
    # an int kernel_size works here (the reference crashes on it, resolve_utils.py:188-189)
    np.testing.assert_array_equal(kex.kmap(kernel_size=1)(lambda v: v[0] * 2)(x), x * 2)


# ---- error conventions ----------------------------------------------------------------
def test_error_conventions(kex):
    with pytest.raises(ValueError):
        kex.kmap(kernel_size=(3,), padding="bogus")
    with pytest.raises(TypeError):
        kex.kmap(kernel_size=(3,), padding=1.5)
    with pytest.raises(TypeError):
        F = kex.kmap(kernel_size=(3,))
        F[0] = 3
    with pytest.raises(AssertionError):
        kex.kmap(kernel_size=(3, 3))(lambda v: v[0, 0])(np.arange(5.0))          # rank mismatch
    with pytest.raises(AssertionError):
        kex.kmap(kernel_size=(9,))(lambda v: v[0])(np.arange(5.0))               # kernel larger than array
    with pytest.raises(ValueError):
        kex.smap(kernel_size=(3,), offset=((-1, 0),))(lambda v: v[0])(np.arange(5.0))
    with pytest.raises(ValueError):
        kex.smap(kernel_size=(3,), offset=((1, 1),))(lambda v: v)(np.arange(5.0))  # non-scalar function
    with pytest.raises(ValueError):
        kex.kscan(kernel_size=(3,))(lambda v: v)(np.arange(5.0))
    with pytest.raises(ValueError):
        kex.kmap(kernel_size=(3,))("not an array or callable", 1)
    with pytest.raises(kex.UnsupportedStencilError):
        kex.kmap(kernel_size=(3,), map_kwargs={"axis_name": "i"})(lambda v: v[0])(np.arange(5.0))


# ---- property checks: engine == oracle over random geometry and function classes -------------------
@pytest.mark.parametrize("seed", range(60))
def test_random_geometry_engine_equals_oracle(kex, seed):
    import catalog
    rng = np.random.default_rng(7000 + seed)
    nd = 1 + seed % 4
    hi = {1: 40, 2: 14, 3: 8, 4: 5}[nd]
    shape = tuple(int(v) for v in rng.integers(3, hi, size=nd))
    ksize = tuple(int(rng.integers(1, min(4, n) + 1)) for n in shape)
    strides = tuple(int(rng.integers(1, 4)) for _ in shape)
    border = tuple((int(rng.integers(-1, 4)), int(rng.integers(-1, 4))) for _ in shape)
    if any(o <= 0 for o in ko.output_shape(shape, ksize, strides, border)):
        border = tuple((0, 0) for _ in shape)
    relative = bool(rng.integers(0, 2))
    names = catalog.for_kernel(ksize, relative)
    fname = names[int(rng.integers(0, len(names)))]
    fn, extra = catalog.build(fname, ksize, relative, rng)
    dtype = np.int32 if seed % 2 == 0 else np.float32
    x = rng.integers(-50, 50, shape).astype(dtype) if dtype == np.int32 else rng.standard_normal(shape).astype(dtype)
    want = ko.kernel_map({fn: ()}, shape, ksize, strides, border, relative)(x, *extra)
    got = kex.kernel_map({fn: ()}, shape, ksize, strides, border, relative)(x, *extra)
    assert got.shape == want.shape and got.dtype == want.dtype, (fname, got.dtype, want.dtype)
    if want.dtype.kind in "iu":
        np.testing.assert_array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(x).max())) * 30)
    if catalog.is_scalar(fname) and nd <= 3 and dtype == np.int32:   # float recurrences amplify rounding
        want = ko.kernel_scan({fn: ()}, shape, ksize, strides, border, relative)(x, *extra)
        got = kex.kernel_scan({fn: ()}, shape, ksize, strides, border, relative)(x, *extra)
        assert got.shape == want.shape and got.dtype == want.dtype
        if want.dtype.kind in "iu":
            np.testing.assert_array_equal(got, want)
        else:
            scale = max(1.0, float(np.abs(want[np.isfinite(want)]).max(initial=1.0)))
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5 * scale)


def test_reverse_scan_matches_oracle(kex):
    rng = np.random.default_rng(1)
    x = rng.integers(-5, 6, (7, 9)).astype(np.int32)
    f = lambda v: v[0, 0] + v[0, 1] - v[1, 0]
    for kw in ({"reverse": True}, {"reverse": False}):
        got = kex.kscan(kernel_size=(3, 3), padding="same", relative=True, scan_kwargs=kw)(f)(x)
        want = ko.kernel_scan({f: ()}, x.shape, (3, 3), (1, 1), ((1, 1), (1, 1)), True, "scan", kw)(x)
        np.testing.assert_array_equal(got, want)


# ---- README.md:270-297 (example 5, Gaussian blur): host-computed weights, 'same' padding ----------------
@pytest.mark.parametrize("ksize", [3, 5, 7])
def test_readme_gaussian_blur(kex, ksize):
    rng = np.random.default_rng(50 + ksize)
    image = rng.standard_normal((61, 84)).astype(np.float32)
    sigma = 1.5
    xx, yy = np.meshgrid(np.linspace(-1, 1, ksize), np.linspace(-1, 1, ksize))
    w = np.exp(-0.5 * (xx ** 2 + yy ** 2) / sigma ** 2)
    w = (w / w.sum()).astype(np.float32)

    @kex.kmap(kernel_size=(ksize, ksize), padding="same")
    def conv(x):
        return np.sum(x * w)

    got = conv(image)
    pad = ksize // 2
    padded = np.pad(image.astype(np.float64), pad)
    want = np.zeros_like(image, dtype=np.float64)
    for a in range(ksize):
        for b in range(ksize):
            want += float(w[a, b]) * padded[a:a + image.shape[0], b:b + image.shape[1]]
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    taps = [(a, b) for a in range(ksize) for b in range(ksize)]
    oracle = ko.affine_map(image, (ksize, ksize), (1, 1), ((pad, pad), (pad, pad)), taps, [float(w[t]) for t in taps])
    np.testing.assert_allclose(got, oracle, rtol=1e-5, atol=1e-5)


def test_rank5_array_with_batch_axes(kex):
    """(N, C, D, H, W) with kernel (1, 1, 3, 3, 3): the two leading 1-wide axes are merged into one batch axis."""
    rng = np.random.default_rng(77)
    x = rng.standard_normal((2, 3, 6, 10, 16)).astype(np.float32)
    f = lambda v: (v[0, 0, 1, 0, 0] + v[0, 0, -1, 0, 0] + v[0, 0, 0, 1, 0] + v[0, 0, 0, -1, 0] + v[0, 0, 0, 0, 1]
                   + v[0, 0, 0, 0, -1] - 6 * v[0, 0, 0, 0, 0])
    got = kex.kmap(kernel_size=(1, 1, 3, 3, 3), padding=(0, 0, "same", "same", "same"), relative=True)(f)(x)
    assert got.shape == x.shape
    taps = [(2, 1, 1), (0, 1, 1), (1, 2, 1), (1, 0, 1), (1, 1, 2), (1, 1, 0), (1, 1, 1)]
    for n in range(2):
        for c in range(3):
            want = ko.affine_map(x[n, c], (3, 3, 3), (1, 1, 1), ((1, 1),) * 3, taps, [1, 1, 1, 1, 1, 1, -6])
            np.testing.assert_allclose(got[n, c], want, rtol=1e-5, atol=1e-5)
    p = kex.kmap(kernel_size=(1, 1, 1, 2, 2), strides=(1, 1, 1, 2, 2))(lambda v: np.max(v))(x)   # pooling over (H, W)
    np.testing.assert_array_equal(p, x.reshape(2, 3, 6, 5, 2, 8, 2).max(axis=(4, 6)))
    with pytest.raises(kex.UnsupportedStencilError):
        kex.kmap(kernel_size=(2, 1, 3, 3, 3), padding="same")(lambda v: np.sum(v))(x)              # 4 windowed axes + 1
