"""GPU parity of the explicit VJPs (the reference gets them from JAX autodiff,
tests/test_interface.py:345-348): dx = flipped-tap stencil of the cotangent, dw = window
correlation.  Both are checked against float64 NumPy restatements of the transpose (no library
comparator) at north_star's bound:

    |dx_gpu - dx_f64| <= 1e-5 * sum_t |c_t| |g_j(t)|        (element-wise)
    |dw_gpu - dw_f64| <= 1e-5 * sum_j |g_j| |x_{origin(j)+t}|  (per tap)

Accumulation order on the GPU: dx = one FMA per tap in row-major window order of the transposed
program; dw = per-lane serial FMA chain over the rows of a band, warp shuffle tree, per-block
shared-memory sum, fixed-order second stage (deterministic, no atomics)."""
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


from vjp_ref import TOL, assert_within, dw_reference, dx_reference  # noqa: E402


@pytest.mark.parametrize("case", [((40, 64), (3, 3), "valid"), ((33, 50), (3, 3), "same"), ((20, 36), (2, 3), ((1, 0), (0, 2))),
                                  ((9, 12, 16), (3, 3, 3), "same"), ((300,), (5,), "same"), ((200, 1028), (3, 3), "valid"),
                                  ((64, 516), (5, 5), "same")])
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
    coefs32 = [float(np.float32(c)) for c in coefs]
    want, mag = dx_reference(g.cpu().numpy(), shape, ksize, (1,) * len(shape), border, taps, coefs32)
    assert_within(dx.cpu().numpy(), want, mag, "dx")


@pytest.mark.parametrize("fn_name", ["sum", "mean", "weighted"])
@pytest.mark.parametrize("geom", [((17, 20), (3, 3), (2, 2), "same"), ((31, 44), (2, 2), (2, 2), "valid"),
                                  ((12, 9, 14), (2, 3, 2), (2, 1, 2), "same"), ((101,), (4,), (3,), ((2, 1),)),
                                  ((64, 260), (3, 3), (2, 2), ((1, 1), (1, 1))), ((40, 128), (4, 4), (4, 4), "valid"),
                                  ((33, 64), (3, 3), (3, 3), ((-1, 2), (0, -2))), ((9, 20, 36), (3, 3, 3), (2, 2, 2), "same"),
                                  ((5, 24, 40), (1, 3, 3), (1, 2, 2), ((0, 0), (1, 0), (0, 1)))])
def test_strided_input_vjp_generic_kernel(kex, fn_name, geom, monkeypatch):
    """Strided dense windows take strided_vjp_input_kernel (a thread walks only the windows that contain its input
    elements); `mean` carries the forward's final division (KxFunc.post_div) into the cotangent -- the average-pool
    gradient of README.md:327-337.  Same accumulation order as generic_vjp_input_kernel: bit-identical to it."""
    from kernex_b200 import _lib
    shape, ksize, strides, pad = geom
    rng = np.random.default_rng(5)
    border = ko.resolve_padding(pad, ksize)
    taps = list(itertools.product(*[range(k) for k in ksize]))
    n = len(taps)
    wts = rng.standard_normal(ksize).astype(np.float32)
    if fn_name == "sum":
        fn, coefs, div = (lambda v: np.sum(v)), [1.0] * n, 0.0
    elif fn_name == "mean":
        fn, coefs, div = (lambda v: np.mean(v)), [1.0] * n, float(n)
    else:
        fn, coefs, div = (lambda v: np.sum(v * wts)), [float(wts[t]) for t in taps], 0.0
    x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda().requires_grad_()
    y = kex.kmap(kernel_size=ksize, strides=strides, padding=pad)(fn)(x)
    g = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32)).cuda()
    (dx,) = torch.autograd.grad(y, x, g, retain_graph=True)
    assert _lib.last_kernel() == "strided_vjp_input_kernel"
    want, mag = dx_reference(g.cpu().numpy(), shape, ksize, strides, border, taps, coefs, div)
    assert_within(dx.cpu().numpy(), want, mag, f"dx[{fn_name}]")
    monkeypatch.setenv("KX_NO_STRIDED_VJP", "1")
    (dx2,) = torch.autograd.grad(y, x, g)
    assert _lib.last_kernel() == "generic_vjp_input_kernel"
    assert torch.equal(dx, dx2)


@pytest.mark.parametrize("strides", [(1, 1), (2, 2)])
def test_post_div_weight_gradient(kex, strides):
    """d/dw of sum(x*w)/c (a post_div program): only the generic coefficient kernel takes it, and it divides."""
    rng = np.random.default_rng(8)
    shape, ksize = (40, 52), (3, 3)
    border = ko.resolve_padding("same", ksize)
    x = rng.standard_normal(shape).astype(np.float32)
    w = torch.from_numpy(rng.standard_normal(ksize).astype(np.float32)).cuda().requires_grad_()
    f = kex.kmap(kernel_size=ksize, strides=strides, padding="same")(lambda v, w: np.mean(v * w))
    y = f(torch.from_numpy(x).cuda(), w)
    g = rng.standard_normal(tuple(y.shape)).astype(np.float32)
    (dw,) = torch.autograd.grad(y, w, torch.from_numpy(g).cuda())
    taps = list(itertools.product(range(3), range(3)))
    want, mag = dw_reference(x, g, ksize, strides, border, taps)
    assert_within(dw.cpu().numpy().reshape(-1), want / 9.0, mag / 9.0, "dw of mean(x*w)")
    # forward value for completeness
    w_np = w.detach().cpu().numpy()
    fwd = ko.affine_map(x, ksize, strides, border, taps, [float(w_np[t]) / 1.0 for t in taps]) / np.float32(9.0)
    np.testing.assert_allclose(y.detach().cpu().numpy(), fwd, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C,H,W", [(3, 40, 64), (3, 33, 50), (2, 17, 128), (4, 70, 36), (1, 25, 44), (3, 200, 1028)])
def test_conv_weight_gradient_matches_oracle(kex, C, H, W):
    """conv-as-kmap forward + dW against the oracle (tests/test_interface.py:336-348)."""
    from kernex_b200 import _lib
    rng = np.random.default_rng(6)
    x = rng.standard_normal((C, H, W)).astype(np.float32)
    w = rng.standard_normal((C, 3, 3)).astype(np.float32)
    border = ((0, 0), (1, 1), (1, 1))
    taps = list(itertools.product(range(C), range(3), range(3)))
    xt = torch.from_numpy(x).cuda()
    wt = torch.from_numpy(w).cuda().requires_grad_()
    conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
    y = conv(xt, wt)
    if C >= 2 and W % 4 == 0:
        assert _lib.last_kernel() in ("conv_fwd_kernel", "conv_fwd_rows_kernel")
    coefs = [float(w[t]) for t in taps]
    want = ko.affine_map(x, (C, 3, 3), (1, 1, 1), border, taps, coefs)
    scale = ko.affine_map(np.abs(x), (C, 3, 3), (1, 1, 1), border, taps, np.abs(coefs))
    assert np.all(np.abs(y.detach().cpu().numpy().astype(np.float64) - want) <= TOL * scale + 1e-30)
    g = rng.standard_normal(tuple(y.shape)).astype(np.float32)
    (dw,) = torch.autograd.grad(y, wt, torch.from_numpy(g).cuda())
    dw_want, mag = dw_reference(x, g, (C, 3, 3), (1, 1, 1), border, taps)
    assert_within(dw.cpu().numpy().reshape(-1), dw_want, mag, "dW")
    np.testing.assert_allclose(ko.weight_grad(x, g, (C, 3, 3), (1, 1, 1), border).reshape(-1), dw_want, rtol=1e-12)


@pytest.mark.parametrize("case", [
    ((70, 200), (5, 5), "same"), ((65, 129), (5, 5), "valid"), ((130, 260), (7, 7), "same"), ((40, 300), (2, 2), "same"),
    ((50, 64), (4, 4), ((1, 2), (0, 3))), ((20, 40, 68), (3, 3, 3), "same"), ((9, 70, 130), (3, 3, 3), "valid"),
    ((2, 30, 36, 64), (1, 3, 3, 3), ((0, 0), (1, 1), (1, 1), (1, 1))), ((3, 60, 132), (1, 5, 5), ((0, 0), (2, 2), (2, 2))),
])
@pytest.mark.parametrize("mean", [False, True])
def test_wgrad_rows_any_window(kex, case, mean, monkeypatch):
    """Weight gradient of stride-1 windows beyond 3x3 (5x5, 7x7, 2x2, 4x4, 3x3x3, with batch axes): wgrad_rows_kernel, at the
    1e-5 bound against the float64 window correlation, deterministic, and equal within that bound to the generic kernel."""
    from kernex_b200 import _lib
    shape, ksize, pad = case
    rng = np.random.default_rng(sum(shape) + sum(ksize) + mean)
    border = ko.resolve_padding(pad, ksize)
    x = rng.standard_normal(shape).astype(np.float32)
    w = torch.from_numpy(rng.standard_normal(ksize).astype(np.float32)).cuda().requires_grad_()
    f = kex.kmap(kernel_size=ksize, padding=pad)((lambda v, w: np.mean(v * w)) if mean else (lambda v, w: np.sum(v * w)))
    y = f(torch.from_numpy(x).cuda(), w)
    g = rng.standard_normal(tuple(y.shape)).astype(np.float32)
    gt = torch.from_numpy(g).cuda()
    (dw,) = torch.autograd.grad(y, w, gt, retain_graph=True)
    assert _lib.last_kernel() == "wgrad_stage2_kernel", _lib.last_kernel()
    (dw_again,) = torch.autograd.grad(y, w, gt, retain_graph=True)
    assert torch.equal(dw, dw_again)
    taps = list(itertools.product(*[range(k) for k in ksize]))
    want, mag = dw_reference(x, g, ksize, (1,) * len(shape), border, taps)
    div = float(len(taps)) if mean else 1.0
    assert_within(dw.cpu().numpy().reshape(-1), want / div, mag / div, "dW")
    monkeypatch.setenv("KX_NO_WGRAD_ROWS", "1")
    (dw_gen,) = torch.autograd.grad(y, w, gt)
    assert _lib.last_kernel() != "wgrad_stage2_kernel"
    assert_within(dw_gen.cpu().numpy().reshape(-1), want / div, mag / div, "dW (generic)")
