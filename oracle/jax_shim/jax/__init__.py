"""NumPy-backed stand-in for the handful of JAX entry points kernex calls.

TEST INFRASTRUCTURE ONLY (lives under ``oracle/``).  Real ``jax``/``jaxlib`` are
not installed in this image and cannot be fetched (no network), so the
reference package at ``/root/reference`` cannot be imported as-is.  Putting this
directory on ``sys.path`` lets the reference's *own, unmodified* Python
(``kernex/_src/*.py``, ``kernex/interface/*.py``) run eagerly on NumPy: every
control decision (view generation, roll, key search, switch clamp, scan order,
write-back) is then the reference's code, only the array primitives are NumPy.
It is used by ``tests/golden/make_golden.py`` to generate fixtures and by
``tests/test_oracle_golden.py`` (when ``/root/reference`` is present) to
cross-check ``oracle/kx_oracle.py``.  It is NOT a JAX re-implementation: no
tracing, no jit, no autodiff; ``vmap``/``scan``/``map`` are Python loops.
"""

from __future__ import annotations

import functools as _ft

import numpy as _np

from . import lax, numpy, profiler, tree_util  # noqa: F401
from ._core import ShimArray as Array
from ._core import tree_map_leaves as _tree_map_leaves
from ._core import stack_results as _stack_results

__version__ = "0.0.shim"


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def _slice_tree(tree, i, axis):
    if axis is None:
        return tree
    return _tree_map_leaves(lambda a: _np.take(_np.asarray(a), i, axis=axis), tree)


def _leading(tree, axis):
    if isinstance(tree, (tuple, list)):
        for t in tree:
            n = _leading(t, axis)
            if n is not None:
                return n
        return None
    return _np.asarray(tree).shape[axis]


def vmap(fun, in_axes=0, out_axes=0, **_kw):
    assert out_axes == 0

    @_ft.wraps(fun)
    def mapped(*args):
        axes = list(in_axes) if isinstance(in_axes, (list, tuple)) else [in_axes] * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _leading(a, ax)
                break
        outs = [fun(*[_slice_tree(a, i, ax) for a, ax in zip(args, axes)]) for i in range(n)]
        return _stack_results(outs)

    return mapped


pmap = vmap
