"""Loaders for the committed golden fixtures (tests/golden/*)."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mat(*args):
    # the reference tests' input helper (tests/test_scan_and_map.py:12-16)
    return np.arange(1, int(np.prod(args)) + 1, dtype=np.int32).reshape(*args)


def engine_vectors():
    with open(os.path.join(GOLDEN, "engine_vectors.json")) as f:
        return json.load(f)


def eval_func_map(src, jnp):
    """Rebuild the ``{lambda: keys}`` dict of a vector; ``jnp`` is the namespace the
    lambdas see as ``jnp`` (NumPy for the oracle and for the tracer)."""
    return eval(src, {"jnp": jnp, "np": np})


def eval_array(src):
    return eval(src, {"mat": mat, "np": np})


def to_border(b):
    return tuple(tuple(int(v) for v in p) for p in b)


_shim = None


def shim_cases():
    global _shim
    if _shim is None:
        with open(os.path.join(GOLDEN, "shim_cases.json")) as f:
            meta = json.load(f)
        arrays = np.load(os.path.join(GOLDEN, "shim_cases.npz"))
        _shim = (meta, arrays)
    return _shim


def shim_inputs(arrays, i):
    x = arrays[f"x_{i}"]
    extra = []
    j = 0
    while f"arg{j}_{i}" in arrays:
        extra.append(arrays[f"arg{j}_{i}"])
        j += 1
    return x, extra
