"""``vmap``: batch a kmap over a leading axis in ONE launch.

The reference composes with ``jax.vmap`` (README.md:310-316 depthwise convolution,
:331-336 average pooling).  Here the batch axis is folded into the launch instead:

* arguments that are NOT mapped (``in_axes`` entry ``None``): the batched call is the same
  stencil with a 1-wide window on a new leading axis, which the fast kernels fold into
  their batch dimension;
* mapped arguments (e.g. one weight tensor per channel): the window function is traced
  once on a symbolic per-item weight tensor and the ``[batch, n_taps]`` coefficient table
  is handed to ``kx_map`` with a per-batch stride.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _Arr, _KX2TORCH, _map_plan, _ptr, _stream_ptr
from .errors import UnsupportedStencilError
from .interface import KernelInterface
from .named_axis import named_axis_wrapper
from .resolve import normalize_slices

__all__ = ["vmap"]


def _unwrap(fn):
    if isinstance(fn, KernelInterface):
        return fn, None
    iface = getattr(fn, "_kx_interface", None)
    if iface is None:
        raise TypeError("vmap expects a function decorated with kmap/smap (or a kmap mesh object)")
    return iface, fn._kx_func


def vmap(fn, in_axes=0):
    """Map a ``kmap``-decorated function (or a kmap mesh) over axis 0 of its array argument and
    of every extra argument whose ``in_axes`` entry is 0 (``None`` = broadcast)."""
    iface, func = _unwrap(fn)
    if iface.inplace or iface.use_offset:
        raise UnsupportedStencilError("vmap is implemented for kmap (parallel, non-offset) functions")

    def call(array, *args):
        axes = list(in_axes) if isinstance(in_axes, (tuple, list)) else [in_axes] * (1 + len(args))
        if len(axes) != 1 + len(args) or axes[0] != 0 or any(a not in (0, None) for a in axes):
            raise ValueError("in_axes must be 0 for the array and 0 / None for every extra argument")
        shape = tuple(array.shape)
        item_shape = shape[1:]
        ksize, strides, border = iface._resolved(item_shape)
        if func is None:
            containers = normalize_slices(iface.container, item_shape)
        else:
            containers = {func: ()}
        fmap_item = {iface._named(f, ksize): keys for f, keys in containers.items()}
        mapped = [i for i, a in enumerate(axes[1:]) if a == 0]
        if not mapped:
            # leading-axis trick: window (1, *k), stride (1, *s), border ((0,0), *b)
            def lift(f):
                return lambda w, *a: f(w[0], *a)
            fmap = {}
            for f, keys in fmap_item.items():
                fmap[lift(f)] = keys if keys == () else [[(0, shape[0], 1)] + list(k) for k in keys]
            from .engine import kernel_map
            op = kernel_map(fmap, shape, (1,) + tuple(ksize), (1,) + tuple(strides), ((0, 0),) + tuple(border),
                            iface.relative, iface.transform_kind, iface.transform_kwargs)
            return op(array, *args)
        # ---- mapped extra arguments: per-batch coefficient rows ---------------------------------
        if len(fmap_item) != 1:
            raise UnsupportedStencilError("vmap over extra arguments needs a single-function kmap")
        arr = _Arr(array)
        dev = arr.t.device
        full_args = []
        for i, a in enumerate(args):
            if axes[1 + i] == 0:
                t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
                if t.shape[0] != shape[0]:
                    raise ValueError("mapped arguments must share the leading axis of the array")
                full_args.append(t.to(dev))
            else:
                full_args.append(a)
        item_args = tuple(a[0] if axes[1 + i] == 0 else a for i, a in enumerate(full_args))
        specs, device_args, out_kx, prog, geom, full_shape = _map_plan(
            fmap_item, item_shape, tuple(ksize), tuple(strides), tuple(border), iface.relative, arr.kx_dtype,
            item_args, {})
        spec = specs[0]
        if spec.kind != "affine" or not spec.device_coefs:
            raise UnsupportedStencilError("vmap over extra arguments needs an affine function of device weights")
        comp = _KX2TORCH[out_kx]
        n, B = prog.c.n_taps, shape[0]
        by_arg: dict = {}
        for slot, c in prog.device_refs:
            for (ai, flat), scale in c.w.items():
                by_arg.setdefault(ai, []).append((slot, flat, scale))
        base = [prog.c.coef[i] for i in range(n)]
        identity = None
        if len(by_arg) == 1 and all(v == 0.0 for v in base):
            (ai, refs), = by_arg.items()
            if axes[1 + ai] == 0 and [r[0] for r in refs] == list(range(n)) and [r[1] for r in refs] == list(range(n)) \
                    and all(r[2] == 1 for r in refs) and full_args[ai][0].numel() == n:
                identity = ai
        if identity is not None:
            # the usual `sum(x * w)`: the per-item weight tensor IS the coefficient row, zero extra launches
            coef = full_args[identity].to(device=dev, dtype=comp).reshape(B, n).contiguous()
        else:
            cache = prog.__dict__.setdefault("_vmap_coef_cache", {})
            ck = (str(dev), comp, B)
            if ck not in cache:   # index / scale tensors are uploaded once per (plan, device, batch)
                cache[ck] = (torch.tensor(base, dtype=comp, device=dev).repeat(B, 1),
                             {ai: (torch.tensor([r[0] for r in refs], dtype=torch.long, device=dev),
                                   torch.tensor([r[1] for r in refs], dtype=torch.long, device=dev),
                                   torch.tensor([r[2] for r in refs], dtype=comp, device=dev))
                              for ai, refs in by_arg.items()})
            coef, plans = cache[ck]
            for ai, (slots, flats, scales) in plans.items():
                w = full_args[ai].to(device=dev, dtype=comp)
                if axes[1 + ai] == 0:
                    contrib = w.reshape(B, -1)[:, flats] * scales
                else:
                    contrib = (w.reshape(-1)[flats] * scales).expand(B, -1)
                coef = coef.index_add(1, slots, contrib)
            coef = coef.contiguous()
        out = torch.empty((B,) + tuple(full_shape), dtype=comp, device=dev)
        if out.numel():
            lib = _lib.load()
            _lib.check(lib.kx_map(C.byref(geom), C.byref(prog.c), _ptr(arr.t), _ptr(coef), n, _ptr(out), B,
                                  _stream_ptr(dev)))
        return arr.give_back(out)

    return call
