"""Pins oracle/kx_oracle.py: reference golden vectors + reference-on-shim outputs."""
from __future__ import annotations

import itertools
import os
import sys

import numpy as np
import pytest

import catalog
import golden_io
from oracle import kx_oracle as ko
from oracle_iface import OracleIface as _OracleIface

OPS = {
    "kernel_map": ko.kernel_map,
    "kernel_scan": ko.kernel_scan,
    "offset_kernel_map": ko.offset_kernel_map,
    "offset_kernel_scan": ko.offset_kernel_scan,
}


@pytest.mark.parametrize("case", golden_io.engine_vectors(), ids=lambda c: c["id"])
def test_reference_golden_vectors(case):
    """All 101 live ``assert_array_equal`` of the reference's tests/test_scan_and_map.py."""
    fmap = golden_io.eval_func_map(case["func_map"], np)
    array = golden_io.eval_array(case["array"])
    fn = OPS[case["op"]](fmap, tuple(case["shape"]), tuple(case["kernel_size"]), tuple(case["strides"]),
                         golden_io.to_border(case["border"]), case["relative"])
    got = fn(array)
    want = np.asarray(case["expect"]["data"]).reshape(case["expect"]["shape"])
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)


def test_roll_view_vectors():
    # reference tests/test_utils.py:14-30
    np.testing.assert_array_equal(ko.roll_view(np.array([])), np.array([]))
    np.testing.assert_array_equal(ko.roll_view(np.array([1])), np.array([1]))
    np.testing.assert_array_equal(ko.roll_view(np.array([1, 2])), np.array([1, 2]))
    np.testing.assert_array_equal(ko.roll_view(np.array([1, 2, 3])), np.array([2, 3, 1]))
    np.testing.assert_array_equal(ko.roll_view(np.array([1, 2, 3, 4])), np.array([2, 3, 4, 1]))
    np.testing.assert_array_equal(ko.roll_view(np.arange(1, 10).reshape(3, 3)),
                                  np.array([[5, 6, 4], [8, 9, 7], [2, 3, 1]]))
    # docstring example _src/utils.py:166-175
    np.testing.assert_array_equal(
        ko.roll_view(np.arange(1, 26).reshape(5, 5)),
        np.array([[13, 14, 15, 11, 12], [18, 19, 20, 16, 17], [23, 24, 25, 21, 22],
                  [3, 4, 5, 1, 2], [8, 9, 10, 6, 7]]))


def test_docstring_and_readme_examples():
    x = np.array([1, 2, 3, 4, 5], dtype=np.int32)
    one = ((0, 0),)
    # kernel_interface.py:443-455 / :379-391, README.md:43-69
    np.testing.assert_array_equal(ko.kernel_map({lambda v: np.sum(v): ()}, (5,), (3,), (1,), one)(x), [6, 9, 12])
    np.testing.assert_array_equal(ko.kernel_scan({lambda v: np.sum(v): ()}, (5,), (3,), (1,), one)(x), [6, 13, 22])
    # kernel_interface.py:214-230 (sscan offset=(2,1)?) and :297-313 (smap)
    f = lambda v: v[0] + v[1] + v[-1]
    np.testing.assert_array_equal(
        ko.offset_kernel_scan({f: ()}, (5,), (3,), (1,), ((2, 1),), True)(x), [1, 2, 9, 18, 5])
    # README.md:183-186 patches, :135-143 laplacian of ones
    m = golden_io.mat(5, 5)
    p = ko.kernel_map({lambda v: v: ()}, (5, 5), (3, 3), (1, 1), ((1, 1), (1, 1)))(m)
    assert p.shape == (5, 5, 3, 3)
    np.testing.assert_array_equal(p[0, 0], [[0, 0, 0], [0, 1, 2], [0, 6, 7]])
    lap = lambda v: (0 * v[1, -1] + 1 * v[1, 0] + 0 * v[1, 1] + 1 * v[0, -1] + -4 * v[0, 0] + 1 * v[0, 1]
                     + 0 * v[-1, -1] + 1 * v[-1, 0] + 0 * v[-1, 1])
    out = ko.kernel_map({lap: ()}, (10, 10), (3, 3), (1, 1), ((0, 0), (0, 0)), True)(np.ones((10, 10), np.float32))
    np.testing.assert_array_equal(out, np.zeros((8, 8), np.float32))


def _shim_ids():
    meta, _ = golden_io.shim_cases()
    return meta


@pytest.mark.parametrize("kind", ["kernel_map", "kernel_scan", "offset_kernel_map", "offset_kernel_scan"])
def test_reference_on_shim_cases(kind):
    """Outputs of the reference's own engine code (run on oracle/jax_shim)."""
    meta, arrays = golden_io.shim_cases()
    rng = np.random.default_rng(0)
    n = 0
    for spec in meta:
        if spec["kind"] != kind:
            continue
        i = spec["id"]
        x, extra = golden_io.shim_inputs(arrays, i)
        ksize = tuple(spec["kernel_size"])
        fn, _ = catalog.build(spec["fn"], ksize, spec["relative"], rng)
        geom = golden_io.to_border(spec["offset" if "offset" in spec else "border"])
        call = OPS[kind]({fn: ()}, tuple(spec["shape"]), ksize, tuple(spec["strides"]), geom, spec["relative"])
        got = call(x, *extra)
        want = arrays[f"out_{i}"]
        assert got.shape == want.shape, spec
        assert got.dtype == want.dtype, spec
        np.testing.assert_array_equal(got, want, err_msg=str(spec))
        n += 1
    assert n > 100


@pytest.mark.parametrize("name", list(catalog.MESH_SCENARIOS))
@pytest.mark.parametrize("inplace", [False, True])
def test_reference_on_shim_mesh(name, inplace):
    meta, arrays = golden_io.shim_cases()
    spec = next(s for s in meta if s["kind"] == "mesh" and s["scenario"] == name and s["inplace"] == inplace)
    rng = np.random.default_rng(20261019)
    cls = lambda **kw: _OracleIface(inplace, **kw)
    F, _ = catalog.MESH_SCENARIOS[name](cls, rng)
    x = arrays[f"x_{spec['id']}"]
    got = F(x)
    want = arrays[f"out_{spec['id']}"]
    assert got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)


# --------------------------------------------------------------------------- #
# fast evaluators == literal evaluator                                         #
# --------------------------------------------------------------------------- #


@pytest.mark.parametrize("seed", range(12))
def test_fast_evaluators_match_literal(seed):
    rng = np.random.default_rng(seed)
    nd = 1 + seed % 3
    shape = tuple(int(v) for v in rng.integers(4, 8, size=nd))
    ksize = tuple(int(rng.integers(1, 4)) for _ in shape)
    strides = tuple(int(rng.integers(1, 3)) for _ in shape)
    border = tuple((int(rng.integers(-1, 3)), int(rng.integers(-1, 3))) for _ in shape)
    if any(o <= 0 for o in ko.output_shape(shape, ksize, strides, border)):
        pytest.skip("empty")
    x = rng.standard_normal(shape).astype(np.float32) if seed % 2 else rng.integers(-9, 9, shape).astype(np.int32)
    taps = list(itertools.product(*[range(k) for k in ksize]))
    w = rng.integers(-3, 4, size=ksize).astype(x.dtype)
    lit = ko.kernel_map({lambda v, w: np.sum(v * w): ()}, shape, ksize, strides, border)(x, w)
    fast = ko.affine_map(x, ksize, strides, border, taps, [w[t] for t in taps])
    if x.dtype == np.int32:
        np.testing.assert_array_equal(lit, fast)
    else:
        np.testing.assert_allclose(lit, fast, rtol=1e-5, atol=1e-5)
    for op, f in (("sum", np.sum), ("max", np.max), ("min", np.min), ("mean", np.mean)):
        lit = ko.kernel_map({lambda v: f(v): ()}, shape, ksize, strides, border)(x)
        fast = ko.reduce_map(x, ksize, strides, border, op)
        np.testing.assert_allclose(lit, fast, rtol=1e-5, atol=1e-5)
    for rel in (False, True):
        lit = ko.kernel_map({lambda v: v: ()}, shape, ksize, strides, border, rel)(x)
        np.testing.assert_array_equal(lit, ko.patch_map(x, ksize, strides, border, rel))


@pytest.mark.skipif(not os.path.isdir("/root/reference/kernex"), reason="reference tree not present on this box")
def test_live_reference_on_shim_matches_oracle():
    """When the reference tree is mounted (build container), run it live."""
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "jax_shim")
    saved = list(sys.path)
    sys.path[:0] = [shim, "/root/reference"]
    try:
        import jax.numpy as jnp
        import kernex as kex
        rng = np.random.default_rng(7)
        x = rng.integers(-9, 9, (6, 7)).astype(np.int32)
        f = lambda v: v[0, 1] - 3 * v[-1, 0] + v[1, -1]
        for border in (((1, 1), (1, 1)), ((0, 2), (-1, 1)), ((-1, 0), (2, 0))):
            for op in ("kernel_map", "kernel_scan"):
                ref = getattr(kex, op)({f: ()}, x.shape, (3, 3), (1, 2), border, True)(jnp.array(x))
                got = OPS[op]({f: ()}, x.shape, (3, 3), (1, 2), border, True)(x)
                np.testing.assert_array_equal(np.asarray(ref), got)
    finally:
        sys.path[:] = saved
        for m in [m for m in sys.modules if m == "jax" or m.startswith("jax.") or m == "kernex" or m.startswith("kernex.")]:
            del sys.modules[m]
