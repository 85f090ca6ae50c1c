"""User-facing ``kmap`` / ``kscan`` / ``smap`` / ``sscan``.

Host-side mirror of ``kernex/interface/kernel_interface.py`` (``KernelInterface``
:46-172 and its four subclasses :175-467): identical constructor arguments,
usable as a decorator (``kmap(...)(fn)(array, *args)``) or as a mesh with region
assignment (``F[slice] = fn`` then ``F(array)``).  The engine behind it is the
CUDA library, selected exactly as the reference selects its four factories from
``(inplace, use_offset)`` (:112-116).

Deliberate differences (crash-only items): an int ``kernel_size``
works; a mesh object can be called repeatedly and on NumPy arrays or CUDA
tensors; the object's constructor arguments are not overwritten by the resolved
values of the first call.
"""

from __future__ import annotations

import functools
from typing import Any, Callable

import numpy as np

from . import engine
from .named_axis import named_axis_wrapper
from .resolve import (normalize_slices, resolve_kernel_size, resolve_offset, resolve_padding,
                      resolve_strides)

try:
    import torch
    _ARRAY_TYPES = (np.ndarray, torch.Tensor)
except Exception:  # pragma: no cover
    _ARRAY_TYPES = (np.ndarray,)

__all__ = ["KernelInterface", "kmap", "kscan", "smap", "sscan"]


def _is_array(x):
    return isinstance(x, _ARRAY_TYPES) or (hasattr(x, "__array__") and hasattr(x, "shape") and not callable(x))


class KernelInterface:
    def __init__(self, kernel_size, strides=1, border=0, relative=False, inplace=False, use_offset=False,
                 named_axis=None, container=None, transform_kind=None, transform_kwargs=None):
        self.kernel_size = kernel_size
        self.strides = strides
        self.relative = relative
        self.inplace = inplace
        self.use_offset = use_offset
        self.named_axis = named_axis
        self.container = container or dict()
        self._raw_border = border
        # eager resolution as the reference does (:67-71) so bad arguments fail at construction
        self.border = (resolve_offset if use_offset else resolve_padding)(border, kernel_size)
        self.transform_kind = transform_kind
        self.transform_kwargs = transform_kwargs

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(kernel_size={self.kernel_size}, strides={self.strides}, "
                f"{'offset' if self.use_offset else 'padding'}={self.border}, relative={self.relative}, "
                f"named_axis={self.named_axis})")

    def __setitem__(self, index, func):
        if not callable(func):
            raise TypeError(f"Input not callable. Found {type(func)}")
        # dict order = order of each function object's FIRST assignment (:90)
        self.container[func] = [*self.container.get(func, []), index]
        self.__dict__.pop("_mesh_ops", None)   # resolved engine closures are per region table

    # -- dispatch ---------------------------------------------------------- #

    def _resolved(self, shape):
        ksize = resolve_kernel_size(self.kernel_size, shape)
        strides = resolve_strides(self.strides, shape)
        border = (resolve_offset if self.use_offset else resolve_padding)(self._raw_border, ksize)
        return ksize, strides, border

    def _engine(self):
        if self.use_offset:
            return engine.offset_kernel_scan if self.inplace else engine.offset_kernel_map
        return engine.kernel_scan if self.inplace else engine.kernel_map

    def _named(self, func, ksize):
        if self.named_axis is None:
            return func
        # one wrapper object per (function, kernel size): keeps the engine's plan cache warm
        cache = self.__dict__.setdefault("_named_cache", {})
        key = (func, tuple(ksize))
        if key not in cache:
            cache[key] = named_axis_wrapper(kernel_size=ksize, named_axis=self.named_axis)(func)
        return cache[key]

    def _wrap_mesh(self, array, *a, **k):
        shape = tuple(array.shape)
        ops = self.__dict__.setdefault("_mesh_ops", {})   # shape -> engine closure (dropped by __setitem__)
        op = ops.get(shape)
        if op is None:
            ksize, strides, border = self._resolved(shape)
            normalized = normalize_slices(self.container, shape)
            resolved = {self._named(func, ksize): keys for func, keys in normalized.items()}
            op = self._engine()(resolved, shape, ksize, strides, border, self.relative,
                                self.transform_kind, self.transform_kwargs)
            if len(ops) < 64:
                ops[shape] = op
        return op(array, *a, **k)

    def _wrap_decorator(self, func):
        ops: dict = {}   # shape -> engine closure: argument resolution and validation happen once per shape

        def call(array, *args, **kwargs):
            shape = tuple(array.shape)
            op = ops.get(shape)
            if op is None:
                ksize, strides, border = self._resolved(shape)
                resolved = {self._named(func, ksize): ()}
                op = self._engine()(resolved, shape, ksize, strides, border, self.relative,
                                    self.transform_kind, self.transform_kwargs)
                if len(ops) < 64:
                    ops[shape] = op
            return op(array, *args, **kwargs)

        call._kx_interface, call._kx_func = self, func  # lets batching.vmap fold a batch axis into the launch
        return call

    def __call__(self, *args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not _is_array(args[0]) and len(kwargs) == 0:
            return functools.wraps(args[0])(self._wrap_decorator(args[0]))
        if len(args) > 0 and _is_array(args[0]):
            return self._wrap_mesh(*args, **kwargs)
        raise ValueError(f"Expected an array or `Callable` for the first argument. "
                         f"Found {tuple(type(a).__name__ for a in args)}")


class sscan(KernelInterface):
    """Scalar function over a sliding window, sequentially and in place, writing into a
    copy of the input at the offset interior (``kernel_interface.py:175-255``)."""

    def __init__(self, kernel_size=1, strides=1, offset=0, relative=False, named_axis=None,
                 scan_kind="scan", scan_kwargs=None):
        super().__init__(kernel_size=kernel_size, strides=strides, border=offset, relative=relative,
                         inplace=True, use_offset=True, named_axis=named_axis,
                         transform_kind=scan_kind, transform_kwargs=scan_kwargs)


class smap(KernelInterface):
    """Scalar function over a sliding window, in parallel, writing into a copy of the
    input at the offset interior (``kernel_interface.py:258-340``)."""

    def __init__(self, kernel_size=1, strides=1, offset=0, relative=False, named_axis=None,
                 map_kind="vmap", map_kwargs=None):
        super().__init__(kernel_size=kernel_size, strides=strides, border=offset, relative=relative,
                         inplace=False, use_offset=True, named_axis=named_axis,
                         transform_kind=map_kind, transform_kwargs=map_kwargs)


class kscan(KernelInterface):
    """Function over a sliding window applied sequentially in row-major order with
    in-place updates (``kernel_interface.py:343-403``)."""

    def __init__(self, kernel_size=1, strides=1, padding=0, relative=False, named_axis=None,
                 scan_kind="scan", scan_kwargs=None):
        super().__init__(kernel_size=kernel_size, strides=strides, border=padding, relative=relative,
                         inplace=True, use_offset=False, named_axis=named_axis,
                         transform_kind=scan_kind, transform_kwargs=scan_kwargs)


class kmap(KernelInterface):
    """Function over a sliding window applied to all windows in parallel
    (``kernel_interface.py:406-467``)."""

    def __init__(self, kernel_size=1, strides=1, padding=0, relative=False, named_axis=None,
                 map_kind="vmap", map_kwargs=None):
        super().__init__(kernel_size=kernel_size, strides=strides, border=padding, relative=relative,
                         inplace=False, use_offset=False, named_axis=named_axis,
                         transform_kind=map_kind, transform_kwargs=map_kwargs)
