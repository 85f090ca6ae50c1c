"""``jax.lax`` stand-in (eager Python loops)."""
from __future__ import annotations

import numpy as _np

from ._core import as_shim as _as_shim
from ._core import stack_results as _stack_results
from ._core import tree_map_leaves as _tree_map_leaves


def broadcasted_iota(dtype, shape, dimension):
    idx = _np.arange(shape[dimension], dtype=dtype)
    view = [1] * len(shape)
    view[dimension] = shape[dimension]
    return _as_shim(_np.broadcast_to(idx.reshape(view), shape).copy())


def slice_in_dim(a, start, limit, stride=1, axis=0):
    sl = [slice(None)] * _np.ndim(a)
    sl[axis] = slice(start, limit, stride)
    return _as_shim(_np.asarray(a)[tuple(sl)])


def concatenate(operands, dimension):
    return _as_shim(_np.concatenate([_np.asarray(o) for o in operands], axis=dimension))


def broadcast_in_dim(a, shape, broadcast_dimensions):
    a = _np.asarray(a)
    view = [1] * len(shape)
    for src, dst in enumerate(broadcast_dimensions):
        view[dst] = a.shape[src]
    return _as_shim(_np.broadcast_to(a.reshape(view), shape).copy())


def switch(index, branches, *operands):
    i = int(_np.clip(int(index), 0, len(branches) - 1))
    return branches[i](*operands)


def _take(tree, i):
    return _tree_map_leaves(lambda a: _np.asarray(a)[i], tree)


def _length(tree):
    while isinstance(tree, (tuple, list)):
        tree = tree[0]
    return _np.asarray(tree).shape[0]


def scan(f, init, xs, length=None, reverse=False, unroll=1):
    n = _length(xs) if xs is not None else length
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        carry, y = f(carry, _take(xs, i) if xs is not None else None)
        ys[i] = y
    if n == 0 or ys[0] is None:
        return carry, None
    return carry, _stack_results(ys)


def map(f, xs, **_kw):  # noqa: A001
    n = _length(xs)
    return _stack_results([f(_take(xs, i)) for i in range(n)])
