"""Argument resolution for the kmap/kscan front-end.

Host-side mirror of ``kernex/interface/resolve_utils.py``.  Results are
identical to the reference for every input the reference accepts; inputs on
which the reference only crashes are accepted: an int
``kernel_size`` is broadcast instead of hitting the ``arg*ndim`` bug
(``resolve_utils.py:188-189``) and slice normalisation is idempotent instead of
mutating the container (``:157-165``).
"""

from __future__ import annotations

from typing import Any, Sequence


def _same(k: int):
    return ((k - 1) // 2, k // 2)


def _pad_item(item: Any, k):
    if isinstance(item, int) and not isinstance(item, bool):
        return (item, item)
    if isinstance(item, tuple):
        return item
    if isinstance(item, str):
        low = item.lower()
        if item in ("same", "SAME") or low == "same":
            return _same(k) if k is not None else "same"
        if item in ("valid", "VALID") or low == "valid":
            return (0, 0)
        raise ValueError(f'string argument must be in ["same","SAME","VALID","valid"].Found {item}')
    raise TypeError("input_argument must be tuple or int or str")


def resolve_padding(arg, kernel_size):
    """``_resolve_padding_argument`` (``resolve_utils.py:26-75``).

    ``kernel_size`` may still contain ints only; when it is an int (ndim not
    known yet) resolution of ``'same'`` is deferred by returning the raw
    argument -- callers re-resolve once the array rank is known."""
    if not isinstance(kernel_size, tuple):
        return arg  # deferred
    if isinstance(arg, tuple):
        msg = (f"kernel_size dimension != padding dimension."
               f"Found length(kernel_size)={len(kernel_size)} length(padding)={len(arg)}")
        assert len(arg) == len(kernel_size), msg
        return tuple(_pad_item(item, k) for item, k in zip(arg, kernel_size))
    if isinstance(arg, int) and not isinstance(arg, bool):
        return ((arg, arg),) * len(kernel_size)
    if isinstance(arg, str):
        if arg.lower() not in ("same", "valid"):
            raise ValueError(f'string argument must be in ["same","SAME","VALID","valid"].Found {arg}')
        return tuple(_pad_item(arg, k) for k in kernel_size)
    raise TypeError("input_argument must be tuple or int or str")


def resolve_offset(arg, kernel_size):
    """``_resolve_offset_argument`` (``resolve_utils.py:103-117``)."""
    if not isinstance(kernel_size, tuple):
        return arg
    if isinstance(arg, (tuple, list)):
        return [(item, item) if isinstance(item, int) else tuple(item) for item in arg]
    if isinstance(arg, int):
        return [(arg, arg)] * len(kernel_size)
    raise NotImplementedError(f"input_argument type={type(arg)} is not implemented")


def resolve_kernel_size(arg, shape: Sequence[int]):
    """``_resolve_kernel_size`` (``resolve_utils.py:168-191``); ``-1`` = whole axis."""
    kw = "kernel_size"
    if isinstance(arg, int) and not isinstance(arg, bool):
        arg = (arg,) * len(shape)
    if isinstance(arg, tuple):
        assert all(isinstance(k, int) for k in arg), f"{kw} input must be a tuple of int."
        assert len(arg) == len(shape), (f"{kw} dimension must be equal to array dimension."
                                        f"Found len({arg}) != len{tuple(shape)}")
        assert all(k <= n for k, n in zip(arg, shape)), (f"{kw} shape must be less than array shape.\n"
                                                          f"Found {kw} = {arg} array shape = {tuple(shape)}")
        return tuple(n if k == -1 else k for n, k in zip(shape, arg))
    raise ValueError(f"{kw} must be instance of int or tuple. Found {type(arg)}")


def resolve_strides(arg, shape: Sequence[int]):
    """``_resolve_strides`` (``resolve_utils.py:194-220``)."""
    kw = "strides"
    if isinstance(arg, tuple):
        assert all(isinstance(s, int) for s in arg), f"{kw} input must be a tuple of int."
        assert len(arg) == len(shape), f"{kw} dimension must be equal to array dimension."
        assert all(s <= n for s, n in zip(arg, shape)), f"{kw} shape must be less than array shape."
        return arg
    if isinstance(arg, int) and not isinstance(arg, bool):
        return (arg,) * len(shape)
    raise ValueError(f"{kw} must be instance of int or tuple. Found {type(arg)}")


def _resolve_single_index(index, n: int):
    if isinstance(index, int) and not isinstance(index, bool):
        return index + n if index < 0 else index
    if isinstance(index, slice):
        # ``x or default`` on purpose: an explicit stop of 0 becomes n (:128-139)
        start = index.start or 0
        start += n if start < 0 else 0
        end = index.stop or n
        end += n if end < 0 else 0
        return (start, end, index.step or 1)
    if isinstance(index, (list, tuple)):
        flat = []
        stack = [index]
        while stack:
            cur = stack.pop()
            for v in cur:
                stack.append(v) if isinstance(v, (list, tuple)) else flat.append(v)
        assert all(isinstance(v, int) for v in flat), "All items in tuple must be int"
        return tuple(index)
    raise NotImplementedError(f"index type={type(index)} is not implemented")


def resolve_index(index, shape: Sequence[int]):
    """``_resolve_index`` (``resolve_utils.py:120-154``): one region key ->
    per-dim ``int | (start, end, step)``; a partial tuple leaves the trailing
    dims unconstrained (``zip`` truncation, :151)."""
    index = index if isinstance(index, tuple) else (index,)
    return [_resolve_single_index(item, n) for item, n in zip(index, shape)]


def normalize_slices(container: dict, shape: Sequence[int]) -> dict:
    """``_normalize_slices`` (``resolve_utils.py:157-165``) without the in-place
    mutation: returns a fresh ``{func: [key, ...]}``."""
    return {func: [resolve_index(index, shape) for index in keys] for func, keys in container.items()}
