"""GPU parity: the CUDA engine (through the public kernex-style API -> C ABI) against the
reference's golden vectors, the reference-on-shim fixtures and the NumPy oracle.
Integer / index results must be bit exact; float32 within 1e-5 * sum_t|c_t x_t|."""
from __future__ import annotations

import numpy as np
import pytest

import catalog
import golden_io

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def kex():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kernex_b200
    from kernex_b200 import _lib
    _lib.load()  # fails loudly if the extension is missing
    return kernex_b200


def _ops(kex):
    return {"kernel_map": kex.kernel_map, "kernel_scan": kex.kernel_scan,
            "offset_kernel_map": kex.offset_kernel_map, "offset_kernel_scan": kex.offset_kernel_scan}


def test_launch_counter_moves(kex):
    from kernex_b200 import _lib
    before = _lib.launch_count()
    out = kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))(np.array([1, 2, 3, 4, 5], dtype=np.int32))
    np.testing.assert_array_equal(out, [6, 9, 12])
    assert _lib.launch_count() > before


@pytest.mark.parametrize("case", golden_io.engine_vectors(), ids=lambda c: c["id"])
def test_reference_golden_vectors(kex, case):
    """The 101 live assert_array_equal of the reference's tests/test_scan_and_map.py."""
    fmap = golden_io.eval_func_map(case["func_map"], np)
    array = golden_io.eval_array(case["array"])
    fn = _ops(kex)[case["op"]](fmap, tuple(case["shape"]), tuple(case["kernel_size"]), tuple(case["strides"]),
                               golden_io.to_border(case["border"]), case["relative"])
    got = fn(array)
    want = np.asarray(case["expect"]["data"]).reshape(case["expect"]["shape"])
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("kind", ["kernel_map", "kernel_scan", "offset_kernel_map", "offset_kernel_scan"])
def test_reference_on_shim_cases(kex, kind):
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
        call = _ops(kex)[kind]({fn: ()}, tuple(spec["shape"]), ksize, tuple(spec["strides"]), geom, spec["relative"])
        got = call(x, *extra)
        want = arrays[f"out_{i}"]
        assert got.shape == want.shape, spec
        assert got.dtype == want.dtype, (spec, got.dtype, want.dtype)
        if want.dtype.kind in "iu":
            np.testing.assert_array_equal(got, want, err_msg=str(spec))
        else:
            # fixtures were computed by NumPy in float64-then-cast or float32 order; CUDA uses an fp32 FMA
            # chain in tap order: |diff| <= 1e-5 * sum|c_t x_t| <= 1e-5 * (k-volume * max|x| * max|c|)
            tol = 1e-5 * max(1.0, float(np.abs(x).max())) * 10 * int(np.prod(ksize))
            np.testing.assert_allclose(got, want, rtol=0, atol=tol, err_msg=str(spec))
        n += 1
    assert n > 100


@pytest.mark.parametrize("name", list(catalog.MESH_SCENARIOS))
@pytest.mark.parametrize("inplace", [False, True])
def test_reference_on_shim_mesh(kex, name, inplace):
    meta, arrays = golden_io.shim_cases()
    spec = next(s for s in meta if s["kind"] == "mesh" and s["scenario"] == name and s["inplace"] == inplace)
    rng = np.random.default_rng(20261019)
    F, _ = catalog.MESH_SCENARIOS[name](kex.kscan if inplace else kex.kmap, rng)
    x = arrays[f"x_{spec['id']}"]
    want = arrays[f"out_{spec['id']}"]
    for attempt in range(2):  # a mesh object can be called repeatedly
        got = F(x)
        assert got.dtype == want.dtype
        if want.dtype.kind in "iu":
            np.testing.assert_array_equal(got, want)
        else:
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_device_tensors_stay_on_device(kex):
    x = torch.arange(1, 26, dtype=torch.int32, device="cuda").reshape(5, 5)
    out = kex.kmap(kernel_size=(3, 3), padding="same")(lambda v: v)(x)
    assert out.is_cuda and out.shape == (5, 5, 3, 3)
    np.testing.assert_array_equal(out[0, 0].cpu().numpy(), [[0, 0, 0], [0, 1, 2], [0, 6, 7]])  # README.md:183-186


def test_unsupported_function_raises(kex):
    with pytest.raises(kex.UnsupportedStencilError):
        kex.kmap(kernel_size=(3,))(lambda v: v[0] * v[1])(np.arange(5, dtype=np.float32))
    with pytest.raises(kex.UnsupportedStencilError):
        kex.kmap(kernel_size=(3,))(lambda v: np.exp(v[0]))(np.arange(5, dtype=np.float32))
