"""CPU tests of the window-function tracer / classifier (kernex_b200/tracer.py): the classified program,
evaluated by hand, must equal the function itself applied to a concrete (rolled) window; everything
outside the three classes must raise UnsupportedStencilError, never compute something else."""
from __future__ import annotations

import itertools

import numpy as np
import pytest

from kernex_b200.errors import UnsupportedStencilError
from kernex_b200.named_axis import named_axis_wrapper
from kernex_b200.tracer import trace_function
from oracle import kx_oracle as ko

W33 = np.arange(1, 10, dtype=np.float32).reshape(3, 3) / 7


def _eval(spec, patch, device=None):
    """Evaluate a FuncSpec on a concrete patch given in ORIGIN coordinates (not rolled)."""
    if spec.kind == "affine":
        def cval(c):
            v = float(c.const)
            for (arg, flat), scale in c.w.items():
                v += float(scale) * float(np.asarray(device[arg]).reshape(-1)[flat])
            return v
        acc = cval(spec.bias) + sum(cval(c) * float(patch[t]) for t, c in zip(spec.taps, spec.coefs))
        return acc / spec.post_div if spec.post_div else acc
    if spec.kind in ("max", "min"):
        vals = [float(patch[t]) for t in spec.taps]
        return max(vals) if spec.kind == "max" else min(vals)
    return np.asarray([patch[t] for t in spec.taps]).reshape(spec.out_shape)


AFFINE_CASES = {
    "readme_laplacian": (lambda x: 0 * x[1, -1] + 1 * x[1, 0] + 0 * x[1, 1] + 1 * x[0, -1] + -4 * x[0, 0] + 1 * x[0, 1]
                         + 0 * x[-1, -1] + 1 * x[-1, 0] + 0 * x[-1, 1], True),
    "sum": (lambda x: np.sum(x), False),
    "method_sum": (lambda x: x.sum(), False),
    "mean": (lambda x: np.mean(x), False),
    "average_weights": (lambda x: np.average(x, weights=W33), False),
    "conv": (lambda x: np.sum(x * W33), False),
    "dot": (lambda x: np.dot(x.ravel(), W33.ravel()), False),
    "tensordot": (lambda x: np.tensordot(x, W33, axes=2), False),
    "row_minus_corner": (lambda x: x[1].sum() - x[0, 0], False),
    "axis_sum_pick": (lambda x: np.sum(x, axis=0)[1], False),
    "col_mean": (lambda x: np.mean(x[:, 1]), False),
    "scaled": (lambda x: (2 * x - 1)[1, 1] / 4 + 3, False),
    "neg_div": (lambda x: -x[0, 0] + x[2, 2] / 4, False),
    "constant": (lambda x: 1.5, False),
    "fdm_relative": (lambda u: u[0, -1] - 0.5 * (u[0, -1] - u[-1, -1]), True),
}


@pytest.mark.parametrize("name", list(AFFINE_CASES))
def test_affine_programs_reproduce_the_function(name):
    fn, relative = AFFINE_CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % 1000)
    spec, dev = trace_function(fn, (3, 3), relative)
    assert spec.kind == "affine" and not dev
    for _ in range(3):
        patch = rng.standard_normal((3, 3))
        seen = ko.roll_view(patch) if relative else patch
        np.testing.assert_allclose(_eval(spec, patch), float(fn(seen)), rtol=1e-6, atol=1e-6)
    assert all(0 <= t[d] < 3 for t in spec.taps for d in range(2))


def test_zero_coefficient_terms_are_dropped():
    spec, _ = trace_function(AFFINE_CASES["readme_laplacian"][0], (3, 3), True)
    assert len(spec.taps) == 5 and spec.integral


@pytest.mark.parametrize("fn,kind", [(lambda x: np.max(x), "max"), (lambda x: x.min(), "min"),
                                     (lambda x: np.maximum.reduce(x.ravel()), "max"), (lambda x: np.max(x[1]), "max")])
def test_reduce_programs(fn, kind):
    spec, _ = trace_function(fn, (3, 3), False)
    assert spec.kind == kind
    patch = np.random.default_rng(1).standard_normal((3, 3))
    assert _eval(spec, patch) == float(fn(patch))


@pytest.mark.parametrize("fn", [lambda x: x, lambda x: x.T, lambda x: x.reshape(-1), lambda x: x[::2, ::2], lambda x: x[::-1],
                                lambda x: np.stack([x[0], x[2]])])
@pytest.mark.parametrize("relative", [False, True])
def test_gather_programs(fn, relative):
    spec, _ = trace_function(fn, (3, 3), relative)
    assert spec.kind == "gather"
    patch = np.arange(9.0).reshape(3, 3)
    seen = ko.roll_view(patch) if relative else patch
    np.testing.assert_array_equal(_eval(spec, patch), np.asarray(fn(seen)))


def test_device_weights_stay_symbolic():
    torch = pytest.importorskip("torch")
    w = torch.arange(1.0, 10.0).reshape(3, 3).requires_grad_()     # a tensor that must not be read at trace time
    spec, dev = trace_function(lambda x, w: np.sum(x * w) + 2 * x[0, 0] * w[1, 1], (3, 3), False, (w,))
    assert spec.kind == "affine" and spec.device_coefs and list(dev) == [0] and spec.weak_float
    patch = np.random.default_rng(2).standard_normal((3, 3))
    wn = w.detach().numpy()
    np.testing.assert_allclose(_eval(spec, patch, {0: wn}), float(np.sum(patch * wn) + 2 * patch[0, 0] * wn[1, 1]), rtol=1e-6)
    wi = torch.arange(9, dtype=torch.int32).reshape(3, 3).requires_grad_(False)
    spec_i, _ = trace_function(lambda x, w: np.sum(x * w), (3, 3), False, (wi.numpy(),))
    assert spec_i.integral and not spec_i.weak_float           # integer host weights keep an integer stencil integer
    with pytest.raises(UnsupportedStencilError):
        trace_function(lambda x, w, v: np.sum(x * w * v), (3, 3), False, (w, w))   # product of two device arguments


def test_named_axis_wrapper_feeds_the_tracer():
    fn = named_axis_wrapper((3, 3), {0: "n", 1: "i"})(lambda u: u["i", "n-1"] - 0.5 * (u["i", "n-1"] - u["i-1", "n-1"]))
    spec, _ = trace_function(fn, (3, 3), True)
    got = {t: float(c.const) for t, c in zip(spec.taps, spec.coefs)}
    assert got == {(0, 1): 0.5, (0, 0): 0.5}                    # window-origin taps of u[n-1, i] and u[n-1, i-1]


UNSUPPORTED = {
    "product": lambda x: x[0, 0] * x[1, 1],
    "exp": lambda x: np.exp(x[0, 0]),
    "abs": lambda x: np.abs(x).sum(),
    "power": lambda x: x[1, 1] ** 2,
    "compare": lambda x: x[1, 1] if x[0, 0] > 0 else x[0, 0],
    "where": lambda x: np.where(x[0, 0] > 0, x[1, 1], x[0, 0]),
    "median": lambda x: np.median(x),
    "std": lambda x: np.std(x),
    "prod": lambda x: np.prod(x),
    "float()": lambda x: float(x[1, 1]),
    "division_by_window": lambda x: 1.0 / x[1, 1],
    "sum_where": lambda x: np.sum(x, where=np.eye(3, dtype=bool)),
    "max_initial": lambda x: np.max(x, initial=0.0),
    "mean_keepdims": lambda x: np.mean(x, keepdims=True),
    "ravel_F": lambda x: np.ravel(x, order="F"),
    "max_of_sums": lambda x: np.max(x + x),
    "vector_of_sums": lambda x: x.sum(axis=0),
    "matmul": lambda x: (x @ W33)[0, 0],
}


@pytest.mark.parametrize("name", list(UNSUPPORTED))
def test_functions_outside_the_three_classes_raise(name):
    with pytest.raises(UnsupportedStencilError):
        trace_function(UNSUPPORTED[name], (3, 3), False)


def test_other_ranks():
    spec, _ = trace_function(lambda v: np.sum(v), (3,), False)
    assert spec.kind == "affine" and spec.taps == [(0,), (1,), (2,)]
    f3 = lambda v: v[1, 0, 0] + v[-1, 0, 0] + v[0, 1, 0] + v[0, -1, 0] + v[0, 0, 1] + v[0, 0, -1] - 6 * v[0, 0, 0]
    spec, _ = trace_function(f3, (3, 3, 3), True)
    assert sorted(spec.taps) == sorted([(2, 1, 1), (0, 1, 1), (1, 2, 1), (1, 0, 1), (1, 1, 2), (1, 1, 0), (1, 1, 1)])
    patch = np.random.default_rng(3).standard_normal((3, 3, 3))
    np.testing.assert_allclose(_eval(spec, patch), float(f3(ko.roll_view(patch))), rtol=1e-6)
