"""CPU tests of the host-side mirror of kernex's interface layer (kernex_b200/resolve.py,
named_axis.py, interface.py): fixed expectations taken from kernex/interface/resolve_utils.py and
named_axis.py, and -- when /root/reference is mounted (build container) -- a LIVE comparison
with the reference's own modules running on oracle/jax_shim."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

from kernex_b200 import resolve as R
from kernex_b200.named_axis import named_axis_wrapper

HAVE_REF = os.path.isdir("/root/reference/kernex")


def test_resolve_padding():
    assert R.resolve_padding("same", (3, 4)) == ((1, 1), (1, 2))          # resolve_utils.py:60-66: ((k-1)//2, k//2)
    assert R.resolve_padding("VALID", (3, 3)) == ((0, 0), (0, 0))
    assert R.resolve_padding(2, (3, 3)) == ((2, 2), (2, 2))
    assert R.resolve_padding(("same", (0, 2), 1), (5, 3, 3)) == ((2, 2), (0, 2), (1, 1))
    assert R.resolve_padding(((1, 1), "valid"), (3, 3)) == ((1, 1), (0, 0))
    with pytest.raises(ValueError):
        R.resolve_padding("full", (3,))                                    # :57,73
    with pytest.raises(ValueError):
        R.resolve_padding(("same", "nope"), (3, 3))
    with pytest.raises(TypeError):
        R.resolve_padding(1.5, (3,))                                       # :75
    with pytest.raises(AssertionError):
        R.resolve_padding(((1, 1),), (3, 3))                               # :37 dimension mismatch


def test_resolve_offset_kernel_size_strides():
    assert R.resolve_offset(1, (3, 3)) == [(1, 1), (1, 1)]
    assert R.resolve_offset(((1, 0), 2), (3, 3)) == [(1, 0), (2, 2)]
    with pytest.raises(NotImplementedError):
        R.resolve_offset("same", (3,))                                     # :117
    assert R.resolve_kernel_size((3, -1), (10, 7)) == (3, 7)               # -1 = the whole axis
    assert R.resolve_kernel_size(3, (10, 7)) == (3, 3)                     # int works here (reference bug :188-189)
    with pytest.raises(AssertionError):
        R.resolve_kernel_size((3, 3), (10,))
    with pytest.raises(AssertionError):
        R.resolve_kernel_size((11,), (10,))
    with pytest.raises(ValueError):
        R.resolve_kernel_size("3", (10,))
    assert R.resolve_strides(2, (10, 7)) == (2, 2)
    assert R.resolve_strides((1, 2), (10, 7)) == (1, 2)
    with pytest.raises(AssertionError):
        R.resolve_strides((1,), (10, 7))


def test_region_key_normalisation():
    shape = (10, 8)
    assert R.resolve_index((slice(None), 0), shape) == [(0, 10, 1), 0]
    assert R.resolve_index((slice(1, -1), slice(None, None, 2)), shape) == [(1, 9, 1), (0, 8, 2)]
    assert R.resolve_index(-1, shape) == [9]                               # partial key: trailing dims unconstrained (:151)
    assert R.resolve_index((slice(2, 0),), shape) == [(2, 10, 1)]          # stop 0 -> n, the reference's `or` (:128-139)
    f, g = (lambda x: x), (lambda x: 2 * x)
    cont = {f: [(slice(None), 0), (slice(None), -1)], g: [(slice(1, None), slice(1, -1))]}
    norm = R.normalize_slices(cont, shape)
    assert norm[f] == [[(0, 10, 1), 0], [(0, 10, 1), 7]] and norm[g] == [[(1, 10, 1), (1, 7, 1)]]
    assert cont[f][0] == (slice(None), 0)                                  # not mutated (the reference mutates, :157-165)
    with pytest.raises(NotImplementedError):
        R.resolve_index(("a",), shape)


def test_named_axis_wrapper():
    x = np.arange(9.0).reshape(3, 3)        # a rolled 3x3 patch: index 0 is the centre, -1 the element before it
    f = named_axis_wrapper((3, 3), {0: "n", 1: "i"})(lambda u: u["i+1", "n-1"] * 10 + u["n", "i"])
    assert f(x) == x[-1, 1] * 10 + x[0, 0]                                 # fully named: key order does not matter
    g = named_axis_wrapper((3, 3), {0: "n"})(lambda u: u["n+1", 1])
    assert g(x) == x[1, 1]                                                 # partially named: positions are kept
    h = named_axis_wrapper((3,), {0: "i"})(lambda u, c: c * u["i-1"])
    assert h(np.array([5.0, 6.0, 7.0]), 2.0) == 14.0                       # extra arguments pass through
    with pytest.raises(KeyError):
        f2 = named_axis_wrapper((3, 3), {0: "n", 1: "i"})(lambda u: u["i+2", "n"])
        f2(x)


def test_interface_error_conventions():
    import kernex_b200 as kex
    with pytest.raises(TypeError):
        F = kex.kmap(kernel_size=(3,))
        F[0] = 3                                                           # kernel_interface.py:87 non-callable
    with pytest.raises(ValueError):
        kex.kmap(kernel_size=(3,))(1, 2)                                   # :169 neither array nor callable
    with pytest.raises(ValueError):
        kex.kmap(kernel_size=(3,), padding="nope")
    with pytest.raises(KeyError):
        kex.kernel_map({(lambda v: v[0]): ()}, (5,), (3,), (1,), ((0, 0),), False, "zmap")
    F = kex.kscan(kernel_size=(3, 3), padding="same", named_axis={0: "n", 1: "i"}, relative=True)
    fn = lambda u: 1
    F[:, 0] = F[:, -1] = fn
    assert list(F.container) == [fn] and len(F.container[fn]) == 2          # one function, two keys, insertion order
    assert "kscan" in repr(F)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present on this box")
def test_live_comparison_with_the_reference_modules():
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "jax_shim")
    saved = list(sys.path)
    sys.path[:0] = [shim, "/root/reference"]
    try:
        from kernex.interface import resolve_utils as ref
        from kernex.interface.named_axis import named_axis_wrapper as ref_named
        for ks in [(3,), (3, 4), (5, 3, 2)]:
            for pad in ["same", "valid", "SAME", 1, 0, tuple(["same"] * len(ks)), tuple([(1, 2)] * len(ks)),
                        tuple(("valid" if i % 2 else (0, 1)) for i in range(len(ks)))]:
                assert tuple(map(tuple, ref._resolve_padding_argument(pad, ks))) == tuple(map(tuple, R.resolve_padding(pad, ks)))
            for off in [0, 1, tuple([(1, 0)] * len(ks)), tuple(range(len(ks)))]:
                assert [tuple(o) for o in ref._resolve_offset_argument(off, ks)] == [tuple(o) for o in R.resolve_offset(off, ks)]
        for shape, ks in [((10, 7), (3, -1)), ((10, 7, 4), (1, 3, 3)), ((9,), (-1,))]:
            assert tuple(ref._resolve_kernel_size(ks, shape)) == R.resolve_kernel_size(ks, shape)
        for shape, st in [((10, 7), 2), ((10, 7), (1, 3)), ((9,), 1)]:
            assert tuple(ref._resolve_strides(st, shape)) == R.resolve_strides(st, shape)
        shape = (12, 9)
        for idx in [(slice(None), 0), (slice(1, -1), slice(None, None, 2)), -1, (3,), (slice(-4, None), slice(2, 7, 3)),
                    (slice(2, 0),), (0, slice(None))]:
            want = ref._resolve_index(idx, shape)
            got = R.resolve_index(idx, shape)
            assert [tuple(w) if isinstance(w, (tuple, list)) else w for w in want] == \
                   [tuple(g) if isinstance(g, (tuple, list)) else g for g in got], idx
        x = np.arange(9.0).reshape(3, 3)
        for na, fn in [({0: "n", 1: "i"}, lambda u: u["i+1", "n-1"] * 10 + u["n", "i"]), ({0: "n"}, lambda u: u["n+1", 1]),
                       ({1: "j"}, lambda u: u[0, "j-1"] - u[-1, "j"])]:
            assert float(ref_named(kernel_size=(3, 3), named_axis=na)(fn)(x)) == float(named_axis_wrapper((3, 3), na)(fn)(x))
    finally:
        sys.path[:] = saved
        for m in [m for m in sys.modules if m == "jax" or m.startswith("jax.") or m == "kernex" or m.startswith("kernex.")]:
            del sys.modules[m]


@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.uint8, np.uint16, np.bool_])
def test_narrow_integer_inputs_are_refused_not_widened(dtype):
    """JAX keeps int8 / int16 / uint8 / bool arrays narrow (wraparound, `sum` promotion, logical bool arithmetic); the
    engine computes in int32 / float32 only, so these inputs raise instead of silently returning int32 results.
    The check runs before any device work, so it is testable without a GPU."""
    import kernex_b200 as kex
    from kernex_b200.errors import UnsupportedStencilError
    x = np.arange(8).astype(dtype)
    with pytest.raises(UnsupportedStencilError, match="narrow"):
        kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))(x)
    with pytest.raises(UnsupportedStencilError, match="narrow"):
        kex.kscan(kernel_size=(3,))(lambda v: np.sum(v))(x)
