"""Named-axis indexing of a window: ``x['i+1', 'n-1']``.

Host-side mirror of the reference's ``kernex/interface/named_axis.py``
(``generate_named_axis`` :42-160, ``named_axis_wrapper`` :163-179).  Behaviour
kept: keys always address *relative* offsets; when every
axis is named the key is order-insensitive (the reference sorts it, :33-39);
when only some axes are named the key is positional and unnamed axes take ints.
Unlike the reference, the window is indexed lazily on ``__getitem__`` instead of
eagerly for all prod(k) offsets per window.
"""

from __future__ import annotations

import functools
import itertools
from typing import Callable


def _axis_range(k: int, relative: bool):
    return tuple(range(-((k - 1) // 2), k // 2 + 1)) if relative else tuple(range(k))


def _label(name: str, off: int) -> str:
    return f"{name}+{off}" if off > 0 else (f"{name}{off}" if off < 0 else name)


def generate_named_axis(kernel_size, named_axis, relative: bool = True):
    """Return ``(table, order_insensitive)``: ``table`` maps a key tuple to the
    integer offset tuple it addresses."""
    ndim = len(kernel_size)
    names = {d: d for d in range(ndim)}
    names.update(named_axis)
    partial = len(named_axis) != ndim
    per_dim_keys, per_dim_vals = [], []
    for d in range(ndim):
        rng = _axis_range(kernel_size[d], relative)
        nm = names[d]
        if isinstance(nm, str):
            per_dim_keys.append([_label(nm, o) for o in rng])
        elif isinstance(nm, int):
            partial = True
            per_dim_keys.append(list(rng))
        else:
            raise ValueError("Wrong format for named_axis.")
        per_dim_vals.append(rng)
    table = {}
    for key, val in zip(itertools.product(*per_dim_keys), itertools.product(*per_dim_vals)):
        table[key if partial else tuple(sorted(key))] = val
    return table, not partial


class NamedWindow:
    """What a window function sees when ``named_axis`` is given."""

    __slots__ = ("_window", "_table", "_sorted")

    def __init__(self, window, table, order_insensitive):
        self._window, self._table, self._sorted = window, table, order_insensitive

    def _norm(self, key):
        if isinstance(key, str):
            return (key,)
        key = tuple(key) if isinstance(key, (tuple, list)) else (key,)
        return tuple(sorted(key)) if self._sorted else key

    def __getitem__(self, key):
        return self._window[self._table[self._norm(key)]]

    def __contains__(self, key):
        return self._norm(key) in self._table

    def keys(self):
        return self._table.keys()


def named_axis_wrapper(kernel_size, named_axis) -> Callable:
    table, order_insensitive = generate_named_axis(tuple(kernel_size), named_axis)

    def wrap(func: Callable):
        @functools.wraps(func)
        def inner(window, *args, **kwargs):
            return func(NamedWindow(window, table, order_insensitive), *args, **kwargs)

        return inner

    return wrap
