"""The steady-state launcher of an engine closure (engine._FastMap): repeat calls on device-resident inputs skip the
plan lookup, so every way the general path could have answered differently must send the call back there -- captured
values that changed (the un-jitted reference re-evaluates the function per call, kernel_interface.py:130-157), other
dtypes / devices / layouts, gradients, cleared caches.  All results against the oracle."""
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


def _oracle_map(f, x, ksize, border, relative=False, strides=None):
    strides = strides or (1,) * x.ndim
    return ko.kernel_map({f: ()}, x.shape, ksize, strides, border, relative)(x)


def test_repeat_calls_use_the_launcher_and_agree_with_the_oracle(kex):
    from kernex_b200 import _lib, engine
    rng = np.random.default_rng(0)
    x = rng.standard_normal((37, 64)).astype(np.float32)
    f = lambda v: v[1, 0] + v[0, -1] - 4 * v[0, 0] + v[0, 1] + v[-1, 0]   # noqa: E731
    lap = kex.kmap(kernel_size=(3, 3), padding="valid", relative=True)(f)
    want = _oracle_map(f, x, (3, 3), ((0, 0), (0, 0)), True)
    xd = torch.from_numpy(x).cuda()
    n0 = _lib.launch_count()
    outs = [lap(xd) for _ in range(4)]
    assert _lib.launch_count() - n0 == 4
    for o in outs:
        np.testing.assert_allclose(o.cpu().numpy(), want, rtol=0, atol=1e-5 * 8 * np.abs(x).max())
    # the same closure on other inputs: int32 (its own launcher), NumPy (general path), non-contiguous (general path)
    xi = rng.integers(-50, 50, (37, 64)).astype(np.int32)
    for _ in range(3):
        np.testing.assert_array_equal(lap(torch.from_numpy(xi).cuda()).cpu().numpy(), _oracle_map(f, xi, (3, 3), ((0, 0), (0, 0)), True))
    np.testing.assert_allclose(lap(x), want, rtol=0, atol=1e-5 * 8 * np.abs(x).max())
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).cuda().T          # (37, 64) view with transposed strides
    assert not xt.is_contiguous()
    np.testing.assert_allclose(lap(xt).cpu().numpy(), want, rtol=0, atol=1e-5 * 8 * np.abs(x).max())
    # a second device-resident call after all that still takes the launcher and still agrees
    np.testing.assert_allclose(lap(xd).cpu().numpy(), want, rtol=0, atol=1e-5 * 8 * np.abs(x).max())
    assert engine._PLAN_GEN[0] >= 0


def test_launcher_follows_captured_values(kex):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((48,)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    dt = 0.5
    f = lambda v: v[0] + dt * (v[1] - v[0])            # noqa: E731
    step = kex.kmap(kernel_size=(2,), padding="valid")(f)
    for _ in range(3):
        got = step(xd).cpu().numpy()
    np.testing.assert_allclose(got, x[:-1] + 0.5 * (x[1:] - x[:-1]), rtol=0, atol=1e-6 * 4 * np.abs(x).max())
    dt = 0.125                                          # rebinding the closure variable: the next call re-traces
    for _ in range(3):
        got = step(xd).cpu().numpy()
    np.testing.assert_allclose(got, x[:-1] + 0.125 * (x[1:] - x[:-1]), rtol=0, atol=1e-6 * 4 * np.abs(x).max())
    w = np.array([1.0, -2.0, 1.0], np.float32)
    g = lambda v: np.sum(v * w)                          # noqa: E731
    fir = kex.kmap(kernel_size=(3,), padding="valid")(g)
    for _ in range(3):
        got = fir(xd).cpu().numpy()
    np.testing.assert_allclose(got, x[:-2] - 2 * x[1:-1] + x[2:], rtol=0, atol=1e-6 * 8 * np.abs(x).max())
    w[1] = 3.0                                           # in-place mutation of a captured host table
    for _ in range(3):
        got = fir(xd).cpu().numpy()
    np.testing.assert_allclose(got, x[:-2] + 3 * x[1:-1] + x[2:], rtol=0, atol=1e-6 * 8 * np.abs(x).max())


def test_launcher_with_a_device_weight_argument(kex):
    """sum(x * w) with w a CUDA tensor: the weight tensor itself is the coefficient table (zero copy), so the launcher
    must read the CURRENT tensor's pointer and contents on every call, and step aside for autograd."""
    rng = np.random.default_rng(2)
    for C in (3, 8):
        x = rng.standard_normal((C, 20, 32)).astype(np.float32)
        conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
        xd = torch.from_numpy(x).cuda()
        for trial in range(4):
            w = rng.standard_normal((C, 3, 3)).astype(np.float32)
            wd = torch.from_numpy(w).cuda()
            got = conv(xd, wd).cpu().numpy()
            want = ko.kernel_map({(lambda v, w=w: np.sum(v * w)): ()}, x.shape, (C, 3, 3), (1, 1, 1),
                                 ((0, 0), (1, 1), (1, 1)), False)(x)
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-5 * np.abs(w).sum() * np.abs(x).max())
            wd.mul_(2.0)                                 # same storage, new values
            np.testing.assert_allclose(conv(xd, wd).cpu().numpy(), 2 * want, rtol=0,
                                       atol=2e-5 * np.abs(w).sum() * np.abs(x).max())
        wg = torch.from_numpy(w).cuda().requires_grad_()
        y = conv(xd, wg)                                 # gradients: the general (autograd) path
        assert y.requires_grad
        y.sum().backward()
        assert wg.grad is not None and wg.grad.shape == wg.shape
        with torch.no_grad():
            np.testing.assert_allclose(conv(xd, wg).cpu().numpy(), want, rtol=0, atol=1e-5 * np.abs(w).sum() * np.abs(x).max())
        # a host weight of the same shape: baked into the program, a different plan
        np.testing.assert_allclose(conv(xd, w).cpu().numpy(), want, rtol=0, atol=1e-5 * np.abs(w).sum() * np.abs(x).max())
        # a weight of another dtype is converted by the general path
        np.testing.assert_allclose(conv(xd, torch.from_numpy(w.astype(np.float64)).cuda()).cpu().numpy(), want, rtol=0,
                                   atol=1e-5 * np.abs(w).sum() * np.abs(x).max())


def test_launcher_retires_when_the_cache_is_cleared_and_for_patches(kex):
    x = np.arange(1, 41, dtype=np.int32)
    xd = torch.from_numpy(x).cuda()
    patches = kex.kmap(kernel_size=(3,), padding="same")(lambda v: v)
    xp = np.pad(x, 1)
    want = np.stack([xp[:-2], xp[1:-1], xp[2:]], axis=1)
    for _ in range(3):
        np.testing.assert_array_equal(patches(xd).cpu().numpy(), want)
    kex.clear_plan_cache()
    for _ in range(3):
        np.testing.assert_array_equal(patches(xd).cpu().numpy(), want)
    # input gradients requested: not the launcher's business
    f = lambda v: 2.0 * v[0] + v[1]                      # noqa: E731
    m = kex.kmap(kernel_size=(2,), padding="valid")(f)
    xf = torch.arange(8, dtype=torch.float32, device="cuda")
    for _ in range(2):
        m(xf)
    xg = xf.clone().requires_grad_()
    y = m(xg)
    y.sum().backward()
    np.testing.assert_allclose(xg.grad.cpu().numpy(), np.array([2, 3, 3, 3, 3, 3, 3, 1], np.float32))


def test_kscan_in_a_time_loop_traces_once(kex):
    """A kscan called per step hits the scan plan cache (engine._scan_plan); results equal the oracle each step."""
    from kernex_b200 import engine
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 24)).astype(np.float32)
    c = 0.3
    f = lambda u: u[0, 0] - c * (u[0, 0] - u[0, -1])    # noqa: E731
    scan = kex.kscan(kernel_size=(1, 3), padding=((0, 0), (1, 1)), relative=True)(f)
    engine.clear_plan_cache()
    xd = torch.from_numpy(x).cuda()
    want = x
    n_before = len(engine._PLAN_CACHE)
    for _ in range(3):
        xd = scan(xd)
        want = ko.kernel_scan({f: ()}, x.shape, (1, 3), (1, 1), ((0, 0), (1, 1)), True)(want)
        np.testing.assert_allclose(xd.cpu().numpy(), want, rtol=0, atol=1e-5 * 4 * np.abs(x).max())
    assert len(engine._PLAN_CACHE) == n_before + 1
