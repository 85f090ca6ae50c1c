"""GPU parity of conv-as-kmap with MANY channels (kx_convc.cu): the shapes of the reference's own benchmark
tests_and_benchmarks/test_benchmark_conv2d.py:21-44 -- kernel (C, 3, 3), padding ("valid", "same", "same"), C in
{4, 8, 16, 32, 64} -- forward against the oracle, and both cotangents at the 1e-5 bound (dx = C batched flipped-tap
2-D stencils of the one cotangent plane, dW = the fused window correlation per group of 4 channels)."""
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


def _check(got, x, C, border, w, bias=0):
    taps = list(itertools.product(range(C), range(3), range(3)))
    coefs = [w[t].item() for t in taps]
    want = ko.affine_map(x, (C, 3, 3), (1, 1, 1), border, taps, coefs, bias)
    assert got.shape == want.shape
    if x.dtype == np.int32:
        np.testing.assert_array_equal(got, want)
    else:
        scale = ko.affine_map(np.abs(x), (C, 3, 3), (1, 1, 1), border, taps, np.abs(coefs), abs(bias))
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        assert np.all(err <= 1e-5 * scale + 1e-30), float((err / (scale + 1e-30)).max())


@pytest.mark.parametrize("C", [5, 8, 16, 32, 64])
@pytest.mark.parametrize("case", range(4))
def test_many_channel_conv_forward(kex, C, case):
    rng = np.random.default_rng(6100 + 10 * C + case)
    H = [16, 33, 64, 70][case]
    W = [16, 132, 64, 516][case]
    border = [((0, 0), (1, 1), (1, 1)), ((0, 0), (0, 0), (0, 0)), ((0, 0), (2, 0), (3, 1)), ((0, 0), (1, 2), (2, 2))][case]
    integer = case == 2
    if integer:
        x = rng.integers(-50, 50, (C, H, W)).astype(np.int32)
        w = rng.integers(-4, 5, (C, 3, 3)).astype(np.int32)
        w[w == 0] = 1
    else:
        x = rng.standard_normal((C, H, W)).astype(np.float32)
        w = rng.standard_normal((C, 3, 3)).astype(np.float32)
    op = kex.kernel_map({(lambda v, w: np.sum(v * w)): ()}, (C, H, W), (C, 3, 3), (1, 1, 1), border, False)
    got = op(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda())      # device weights
    assert _last_kernel() == "convc_kernel"
    _check(got.cpu().numpy(), x, C, border, w)
    got_h = op(x, w)                                                       # host weights
    assert _last_kernel() == "convc_kernel"
    _check(got_h, x, C, border, w)


def test_reference_benchmark_shape_c64(kex, monkeypatch):
    """(64, 128, 128): 576 taps (needs KX_MAX_TAPS >= 576); also against the generic kernel (KX_NO_CONVC=1)."""
    rng = np.random.default_rng(6400)
    C, H = 64, 128
    x = rng.standard_normal((C, H, H)).astype(np.float32)
    w = rng.standard_normal((C, 3, 3)).astype(np.float32)
    conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"), relative=False)(lambda v, w: np.sum(v * w))
    got = conv(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda())
    assert _last_kernel() == "convc_kernel" and got.shape == (1, H, H)
    _check(got.cpu().numpy(), x, C, ((0, 0), (1, 1), (1, 1)), w)
    monkeypatch.setenv("KX_NO_CONVC", "1")
    want = conv(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda())
    assert _last_kernel() == "generic_map_kernel"
    monkeypatch.delenv("KX_NO_CONVC")
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=1e-3)   # test_benchmark_conv2d.py:50


def test_partial_windows_and_bias_take_the_masked_variant(kex):
    rng = np.random.default_rng(6500)
    C, H, W = 6, 30, 68
    x = rng.standard_normal((C, H, W)).astype(np.float32)
    taps = [t for t in itertools.product(range(C), range(3), range(3)) if rng.random() < 0.7]
    coefs = [float(c) for c in rng.standard_normal(len(taps))]

    def fn(v):
        acc = 0.25
        for t, c in zip(taps, coefs):
            acc = acc + c * v[t]
        return acc

    border = ((0, 0), (1, 1), (1, 1))
    got = kex.kernel_map({fn: ()}, (C, H, W), (C, 3, 3), (1, 1, 1), border, False)(x)
    assert _last_kernel() == "convc_kernel"
    want = ko.affine_map(x, (C, 3, 3), (1, 1, 1), border, taps, coefs, 0.25)
    scale = ko.affine_map(np.abs(x), (C, 3, 3), (1, 1, 1), border, taps, np.abs(coefs), 0.25)
    assert np.all(np.abs(got.astype(np.float64) - want) <= 1e-5 * scale + 1e-30)


@pytest.mark.parametrize("C,H,W", [(8, 40, 132), (16, 64, 64), (5, 33, 260)])
@pytest.mark.parametrize("dev_weights", [True, False])
def test_many_channel_conv_gradients(kex, C, H, W, dev_weights):
    from vjp_ref import assert_within, dw_reference, dx_reference
    rng = np.random.default_rng(6600 + C)
    x = rng.standard_normal((C, H, W)).astype(np.float32)
    w = rng.standard_normal((C, 3, 3)).astype(np.float32)
    border = ((0, 0), (1, 1), (1, 1))
    taps = list(itertools.product(range(C), range(3), range(3)))
    xt = torch.from_numpy(x).cuda().requires_grad_()
    conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    g = rng.standard_normal((1, H, W)).astype(np.float32)
    if dev_weights:
        wt = torch.from_numpy(w).cuda().requires_grad_()
        y = conv(xt, wt)
        dx, dw = torch.autograd.grad(y, [xt, wt], torch.from_numpy(g).cuda())
        want_dw, mag_dw = dw_reference(x, g, (C, 3, 3), (1, 1, 1), border, taps)
        assert_within(dw.cpu().numpy().reshape(-1), want_dw, mag_dw, "dW")
    else:
        y = conv(xt, w)
        (dx,) = torch.autograd.grad(y, xt, torch.from_numpy(g).cuda())
    want_dx, mag_dx = dx_reference(g, (C, H, W), (C, 3, 3), (1, 1, 1), border, taps, [float(w[t]) for t in taps])
    assert_within(dx.cpu().numpy(), want_dx, mag_dx, "dx")


def test_gradient_kernels_of_the_many_channel_conv(kex):
    """dx = ONE batched stencil2d launch (device weights), dW = conv_wgrad per group of 4 channels -- not the generic kernels."""
    from kernex_b200.engine import PreparedMap
    rng = np.random.default_rng(6700)
    C, H, W = 12, 48, 132
    x = torch.from_numpy(rng.standard_normal((C, H, W)).astype(np.float32)).cuda()
    w = torch.from_numpy(rng.standard_normal((C, 3, 3)).astype(np.float32)).cuda()
    pm = PreparedMap({(lambda v, w: np.sum(v * w)): ()}, (C, H, W), (C, 3, 3), (1, 1, 1), ((0, 0), (1, 1), (1, 1)), False, args=(w,))
    y = pm(x)
    assert _last_kernel() == "convc_kernel"
    g = torch.randn_like(y)
    pm.vjp_input(g)
    assert _last_kernel() == "stencil2d_kernel"
    pm.vjp_coef(x, g)
    assert _last_kernel() == "conv_wgrad_finish"


@pytest.mark.parametrize("C,K,H,W,pad", [(3, 5, 40, 260, "same"), (2, 7, 33, 132, "same"), (8, 5, 21, 516, "valid"), (6, 7, 50, 64, "same"),
                                         (3, 5, 64, 64, ((0, 0), (1, 2), (3, 0)))])
@pytest.mark.parametrize("device_weights", [False, True])
def test_wide_window_multichannel_conv(kex, C, K, H, W, pad, device_weights):
    """conv-as-kmap with (C, 5, 5) / (C, 7, 7) windows (convc_kernel<K>): forward against the oracle at the 1e-5 bound, int32
    bit exact, and both gradients (dx: per-channel flipped-tap 2-D stencils, dW: wgrad_rows_kernel) against float64."""
    import itertools
    from kernex_b200 import _lib
    from vjp_ref import assert_within, dw_reference, dx_reference
    rng = np.random.default_rng(C * 100 + K + H)
    padding = ("valid", pad, pad) if isinstance(pad, str) else pad
    border = ko.resolve_padding(padding, (C, K, K))
    taps = list(itertools.product(range(C), range(K), range(K)))
    x = rng.standard_normal((C, H, W)).astype(np.float32)
    w = rng.standard_normal((C, K, K)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    if device_weights:
        wt = torch.from_numpy(w).cuda().requires_grad_()
        conv = kex.kmap(kernel_size=(C, K, K), padding=padding)(lambda v, w: np.sum(v * w))
        xt.requires_grad_()
        y = conv(xt, wt)
    else:
        conv = kex.kmap(kernel_size=(C, K, K), padding=padding)(lambda v: np.sum(v * w))
        y = conv(xt)
    assert _lib.last_kernel() == "convc_kernel", _lib.last_kernel()
    coefs = [float(w[t]) for t in taps]
    want = ko.affine_map(x, (C, K, K), (1, 1, 1), border, taps, coefs)
    scale = ko.affine_map(np.abs(x), (C, K, K), (1, 1, 1), border, taps, np.abs(coefs))
    assert np.all(np.abs(y.detach().cpu().numpy().astype(np.float64) - want) <= 1e-5 * scale + 1e-30)
    if device_weights:
        g = rng.standard_normal(tuple(y.shape)).astype(np.float32)
        dx, dw = torch.autograd.grad(y, (xt, wt), torch.from_numpy(g).cuda())
        dw_want, mag = dw_reference(x, g, (C, K, K), (1, 1, 1), border, taps)
        assert_within(dw.cpu().numpy().reshape(-1), dw_want, mag, "dW")
        dx_want, magx = dx_reference(g, (C, H, W), (C, K, K), (1, 1, 1), border, taps, [float(np.float32(c)) for c in coefs])
        assert_within(dx.cpu().numpy(), dx_want, magx, "dx")
    else:
        xi = rng.integers(-9, 10, (C, H, W)).astype(np.int32)
        wi = rng.integers(-3, 4, (C, K, K)).astype(np.int32)
        yi = kex.kmap(kernel_size=(C, K, K), padding=padding)(lambda v: np.sum(v * wi))(torch.from_numpy(xi).cuda())
        wanti = ko.affine_map(xi.astype(np.int64), (C, K, K), (1, 1, 1), border, taps, [int(wi[t]) for t in taps])
        np.testing.assert_array_equal(yi.cpu().numpy(), wanti.astype(np.int32))
