"""The plan cache keys a traced window function on the VALUES it captured (closure cells, referenced globals,
defaults), because the tracer bakes them into the coefficients while the un-jitted reference re-evaluates the
function on every call (kernex/interface/kernel_interface.py:130-157).  Pure host logic: no GPU needed."""
from __future__ import annotations

import numpy as np
import pytest

from kernex_b200 import engine as E
from kernex_b200 import _lib
from kernex_b200.named_axis import named_axis_wrapper

ALPHA = 0.25
TABLE = np.array([1.0, 2.0, 3.0])


def _key(f, *a, **k):
    return E._plan_key("map", {f: ()}, (8, 8), (3, 3), (1, 1), ((0, 0), (0, 0)), True, _lib.KX_F32, a, k)


def test_rebinding_a_closure_variable_changes_the_key():
    dt = 0.1
    f = lambda x: x[0, 0] + dt * x[1, 0]       # noqa: E731
    k1 = _key(f)
    assert k1 is not None and _key(f) == k1     # stable while nothing changes
    dt = 0.2
    assert _key(f) not in (None, k1)


def test_mutating_a_captured_array_in_place_changes_the_key():
    w = np.ones((3, 3), np.float32)
    f = lambda x: np.sum(x * w)                 # noqa: E731
    k1 = _key(f)
    w[1, 1] = -4.0
    k2 = _key(f)
    assert k1 is not None and k2 is not None and k1 != k2


def test_globals_and_defaults_are_part_of_the_key():
    global ALPHA
    f = lambda x: ALPHA * x[0, 0] + TABLE[1] * x[0, 1]   # noqa: E731
    k1 = _key(f)
    ALPHA = 0.5
    k2 = _key(f)
    ALPHA = 0.25
    assert k1 is not None and k1 != k2 and _key(f) == k1
    TABLE[1] = 7.0
    assert _key(f) != k1
    TABLE[1] = 2.0

    def g(x, c=1.5):
        return c * x[0, 0]
    kg = _key(g)
    g.__defaults__ = (2.5,)
    assert kg is not None and _key(g) != kg


def test_wrapped_functions_are_fingerprinted_through_the_wrapper():
    kappa = 0.5
    f = lambda u: u["i", "n-1"] - kappa * (u["i", "n-1"] - u["i-1", "n-1"])   # noqa: E731
    wrapped = named_axis_wrapper(kernel_size=(3, 3), named_axis={0: "n", 1: "i"})(f)
    k1 = _key(wrapped)
    assert k1 is not None
    kappa = 0.75
    assert _key(wrapped) not in (None, k1)


def test_unfingerprintable_captures_are_not_cached():
    big = np.zeros(100_000, np.float32)
    f = lambda x: x[0, 0] * big[3]              # noqa: E731
    assert _key(f) is None

    class Opaque:
        v = 2.0
    o = Opaque()
    g = lambda x: x[0, 0] * o.v                 # noqa: E731
    assert _key(g) is None


def test_clear_plan_cache_is_public():
    import kernex_b200
    E._PLAN_CACHE[("dummy",)] = ((), ())
    kernex_b200.clear_plan_cache()
    assert not E._PLAN_CACHE


def test_fingerprint_is_cheap():
    import time
    f = lambda v: (0 * v[1, -1] + 1 * v[1, 0] + 0 * v[1, 1] + 1 * v[0, -1] + -4 * v[0, 0] + 1 * v[0, 1])  # noqa: E731
    _key(f)
    t0 = time.perf_counter()
    for _ in range(2000):
        _key(f)
    assert (time.perf_counter() - t0) / 2000 < 50e-6


def test_scan_plans_are_cached_per_captured_value():
    """kernel_scan keeps its traced program per (function, geometry, captured values) like kernel_map does: the second
    lookup returns the SAME program object, a changed closure value a new one (scan.py:84-113 re-traces per call)."""
    E.clear_plan_cache()
    c = 0.5
    f = lambda u: u[0, 0] - c * (u[0, 0] - u[0, -1])      # noqa: E731
    args = ({f: ()}, (6, 16), (1, 3), (1, 1), ((0, 0), (1, 1)), True, _lib.KX_F32, (), {})
    p1 = E._scan_plan(*args)
    p2 = E._scan_plan(*args)
    assert p1[2] is p2[2] and p1[3] is p2[3]
    c = 0.25
    p3 = E._scan_plan(*args)
    assert p3[2] is not p1[2]
    assert abs(p3[2].c.coef[0] - p1[2].c.coef[0]) > 0 or any(
        p3[2].c.coef[i] != p1[2].c.coef[i] for i in range(p1[2].c.n_taps))
    # a map of the same function and geometry is a different entry (tag), not a collision
    m = E._map_plan(*args)
    assert m[3] is not p3[2]
