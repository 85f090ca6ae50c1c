"""GPU parity of the hyperplane-wavefront kscan kernel (kx_scan_wavefront.cu): scans whose
functions read cells of the row being written (Gauss-Seidel sweeps, 1-D recurrences, README
RK4) against the literal sequential oracle (kernex/_src/scan.py:84-113 restated), and against
the library's own one-thread sweep (KX_SCAN_SEQUENTIAL=1) where the oracle loop is too slow."""
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


WAVE = ("scan_wavefront_kernel", "scan_wavefront_rows_kernel")


def _tap_fn(taps, coefs, bias, integer):
    def fn(x):
        acc = int(bias) if integer else float(bias)
        for tp, c in zip(taps, coefs):
            acc = acc + (int(c) if integer else float(c)) * x[tp]
        return acc
    return fn


def _random_case(rng, nd, integer):
    ksize = tuple(int(rng.integers(1, 4)) for _ in range(nd))
    shape = tuple(int(rng.integers(3, 12 if nd == 3 else 24)) for _ in range(nd))
    strides = tuple(int(rng.integers(1, 3)) for _ in range(nd))
    border = tuple((int(rng.integers(0, 3)), int(rng.integers(0, 3))) for _ in range(nd))
    all_taps = list(itertools.product(*[range(k) for k in ksize]))
    n = int(rng.integers(1, min(len(all_taps), 6) + 1))
    taps = [all_taps[i] for i in rng.choice(len(all_taps), n, replace=False)]
    if integer:
        coefs, bias = rng.integers(-3, 4, n), int(rng.integers(-2, 3))
    else:
        coefs, bias = rng.uniform(-0.4, 0.4, n), float(rng.uniform(-1, 1))  # contraction: no blow-up
    return shape, ksize, strides, border, taps, coefs, bias


@pytest.mark.parametrize("seed", range(40))
def test_random_scans_match_literal_oracle(kex, seed):
    rng = np.random.default_rng(9000 + seed)
    nd = 1 + seed % 3
    integer = seed % 2 == 0
    reverse = seed % 5 == 3
    shape, ksize, strides, border, taps, coefs, bias = _random_case(rng, nd, integer)
    if any(n + lo + hi < k for n, k, (lo, hi) in zip(shape, ksize, border)):
        pytest.skip("empty output")
    fn = _tap_fn(taps, coefs, bias, integer)
    x = (rng.integers(-5, 6, shape).astype(np.int32) if integer
         else rng.standard_normal(shape).astype(np.float32))
    kw = {"reverse": True} if reverse else None
    want = ko.kernel_scan({fn: ()}, shape, ksize, strides, border, False, "scan", kw)(x)
    got = kex.kernel_scan({fn: ()}, shape, ksize, strides, border, False, "scan", kw)(x)
    assert _last_kernel() in WAVE + ("scan_rowmarch_kernel",)
    assert got.shape == want.shape and got.dtype == want.dtype
    if integer:
        np.testing.assert_array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("relative", [False, True])
@pytest.mark.parametrize("seed", range(6))
def test_random_multi_function_scans(kex, seed, relative):
    rng = np.random.default_rng(9500 + seed)
    shape, ksize = (int(rng.integers(6, 14)), int(rng.integers(6, 20))), (3, 3)
    border = ((1, 1), (1, 1))
    fns = []
    for q in range(3):
        idx = rng.choice(9, 3, replace=False)
        taps = [(int(i) // 3 - (1 if relative else 0), int(i) % 3 - (1 if relative else 0)) for i in idx]
        fns.append(_tap_fn(taps, rng.integers(-2, 3, 3), int(rng.integers(-2, 3)), True))
    keys = {fns[0]: [[]], fns[1]: [[(2, shape[0] - 1, 1), (1, shape[1], 2)]], fns[2]: [[3], [(0, 2), (0, 4)]]}
    x = rng.integers(-4, 5, shape).astype(np.int32)
    want = ko.kernel_scan(keys, shape, ksize, (1, 1), border, relative)(x)
    got = kex.kernel_scan(keys, shape, ksize, (1, 1), border, relative)(x)
    assert _last_kernel() in WAVE
    np.testing.assert_array_equal(got, want)


def test_gauss_seidel_sweep(kex):
    # one in-place Gauss-Seidel sweep of the 5-point Laplace equation: reads the updated west and
    # north neighbours and the old east and south ones
    rng = np.random.default_rng(11)
    x = rng.standard_normal((37, 53)).astype(np.float32)
    fn = lambda u: 0.25 * (u[0, -1] + u[-1, 0] + u[0, 1] + u[1, 0])
    F = kex.kscan(kernel_size=(3, 3), padding="same", relative=True)(fn)
    got = F(x)
    assert _last_kernel() == "scan_wavefront_rows_kernel"     # in-row dependency, 'same' borders: warp-skewed kernel
    want = ko.kernel_scan({fn: ()}, x.shape, (3, 3), (1, 1), ((1, 1), (1, 1)), True)(x)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    # independent statement of the same sweep
    ref = np.pad(x.astype(np.float64), 1)
    for i in range(1, 38):
        for j in range(1, 54):
            ref[i, j] = 0.25 * (ref[i, j - 1] + ref[i - 1, j] + ref[i, j + 1] + ref[i + 1, j])
    np.testing.assert_allclose(got, ref[1:-1, 1:-1], rtol=1e-4, atol=1e-5)


def test_readme_rk4(kex):
    # README.md:340-392: dy/dt = y by RK4 as a kscan over a (y, t) table
    t = np.linspace(0, 1, 5, dtype=np.float32)
    x = np.stack([np.zeros(5, np.float32), t], axis=0)
    dt = float(t[1] - t[0])
    f = lambda tn, yn: yn

    def ic(x):
        return 1.0

    def rk4(x):
        t0, y0 = x[1, -1], x[0, -1]
        k1 = dt * f(t0, y0)
        k2 = dt * f(t0 + dt / 2, y0 + 1 / 2 * k1)
        k3 = dt * f(t0 + dt / 2, y0 + 1 / 2 * k2)
        k4 = dt * f(t0 + dt, y0 + k3)
        return y0 + 1 / 6 * (k1 + 2 * k2 + 2 * k3 + k4)

    F = kex.kscan(kernel_size=(2, 3), relative=True, padding=((0, 1)))
    F[0:1, 1:] = rk4
    F[0, 0] = ic
    y = F(x)[0, :]
    np.testing.assert_allclose(y, np.exp(t), rtol=2e-4)
    step = 1 + dt + dt ** 2 / 2 + dt ** 3 / 6 + dt ** 4 / 24
    np.testing.assert_allclose(y, step ** np.arange(5), rtol=1e-5)


@pytest.mark.parametrize("shape,fn_taps,reverse", [
    ((40, 20000), [(0, 1), (0, 2), (1, 1)], False),     # row-above only, int32: a = (1, 0), cooperative grid path
    ((300, 300), [(1, 0), (0, 1), (0, 2), (2, 1)], False),  # a = (2, 1): x[-1,+1] tap
    ((300, 300), [(1, 0), (0, 1), (0, 2), (2, 1)], True),
    ((12, 70, 90), [(0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 2)], False),
])
def test_wavefront_equals_one_thread_sweep(kex, monkeypatch, shape, fn_taps, reverse):
    rng = np.random.default_rng(77)
    ksize = (3,) * len(shape)
    fn = _tap_fn(fn_taps, rng.integers(-2, 3, len(fn_taps)), 1, True)
    x = torch.from_numpy(rng.integers(-3, 4, shape).astype(np.int32)).cuda()
    kw = {"reverse": True} if reverse else None
    op = kex.kernel_scan({fn: ()}, shape, ksize, (1,) * len(shape), ((1, 1),) * len(shape), False, "scan", kw)
    got = op(x)
    assert _last_kernel() in WAVE
    monkeypatch.setenv("KX_SCAN_NO_ROWS", "1")                # hyperplane kernel only
    got2 = op(x)
    assert _last_kernel() == "scan_wavefront_kernel"
    monkeypatch.setenv("KX_SCAN_SEQUENTIAL", "1")
    want = op(x)
    assert _last_kernel() not in WAVE
    assert torch.equal(got, want) and torch.equal(got2, want)


@pytest.mark.parametrize("case", range(8))
def test_warp_skewed_kernel_geometries(kex, monkeypatch, case):
    """scan_wavefront_rows_kernel: several 32-row bands (inter-warp hand-off through global progress
    flags), reverse sweeps, 5x5 windows (skew 3, two rows up), 1-D recurrences, multi-function meshes;
    bit-identical to the one-thread sweep."""
    rng = np.random.default_rng(9900 + case)
    shape, ksize = [((200, 150), (3, 3)), ((70, 300), (3, 3)), ((97, 64), (5, 5)), ((1000,), (3,)),
                    ((33, 40), (3, 5)), ((260, 90), (3, 3)), ((64, 64), (5, 3)), ((5000,), (5,))][case]
    nd = len(shape)
    reverse = case in (1, 5, 7)
    all_taps = list(itertools.product(*[range(k) for k in ksize]))
    taps = [all_taps[i] for i in rng.choice(len(all_taps), min(6, len(all_taps)), replace=False)]
    integer = case % 2 == 0
    if integer:
        fns = [_tap_fn(taps, rng.integers(-2, 3, len(taps)), 1, True), _tap_fn(taps[:2], [1, -1], 0, True)]
        x = torch.from_numpy(rng.integers(-3, 4, shape).astype(np.int32)).cuda()
    else:
        fns = [_tap_fn(taps, rng.uniform(-0.15, 0.15, len(taps)), 0.5, False), _tap_fn(taps[:2], [0.5, -0.25], 0.0, False)]
        x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda()
    border = tuple(((k - 1) // 2, k // 2) for k in ksize)
    keys = {fns[0]: [[]], fns[1]: [[(3, shape[0] - 2, 1)] + ([(0, shape[1], 2)] if nd == 2 else [])]}
    kw = {"reverse": True} if reverse else None
    op = kex.kernel_scan(keys, shape, ksize, (1,) * nd, border, False, "scan", kw)
    got = op(x)
    monkeypatch.setenv("KX_SCAN_SEQUENTIAL", "1")
    want = op(x)
    monkeypatch.delenv("KX_SCAN_SEQUENTIAL")
    assert torch.equal(got, want)
