from __future__ import annotations

import numpy as np


class _At:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _index(self):
        idx = self.idx
        if isinstance(idx, tuple):
            return tuple(np.asarray(i) if isinstance(i, np.ndarray) else i for i in idx)
        return idx

    def get(self, **_kw):
        return as_shim(np.asarray(self.arr)[self._index()])

    def set(self, value, **_kw):
        out = np.array(self.arr, copy=True)
        out[self._index()] = np.asarray(value)
        return as_shim(out)


class _AtHelper:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _At(self.arr, idx)


class ShimArray(np.ndarray):
    @property
    def at(self):
        return _AtHelper(self)


def _demote(dt):
    dt = np.dtype(dt)
    if dt == np.float64:
        return np.dtype(np.float32)
    if dt == np.int64:
        return np.dtype(np.int32)
    return dt


def as_shim(a, dtype=None):
    a = np.asarray(a, dtype=dtype)
    if dtype is None:
        a = a.astype(_demote(a.dtype), copy=False)
    return a.view(ShimArray)


def tree_map_leaves(f, tree):
    if isinstance(tree, tuple):
        return tuple(tree_map_leaves(f, t) for t in tree)
    if isinstance(tree, list):
        return [tree_map_leaves(f, t) for t in tree]
    return f(tree)


def _weak(v):
    return isinstance(v, (bool, int, float)) and not isinstance(v, np.generic)


def _stack_leaf(values):
    strong = [np.asarray(v).dtype for v in values if not _weak(v)]
    if strong:
        dt = _demote(np.result_type(*strong))
        if dt.kind in "iub" and any(isinstance(v, float) for v in values if _weak(v)):
            dt = np.dtype(np.float32)
    elif any(isinstance(v, float) for v in values):
        dt = np.dtype(np.float32)
    else:
        dt = np.dtype(np.int32)
    return as_shim(np.stack([np.asarray(v, dtype=dt) for v in values]), dtype=dt)


def stack_results(outs):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(stack_results([o[i] for o in outs]) for i in range(len(first)))
    if first is None:
        return None
    return _stack_leaf(outs)
