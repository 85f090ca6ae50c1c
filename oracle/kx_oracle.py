"""CPU oracle for the kernex kmap/kscan hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in plain NumPy, the algorithm of the reference engine
(ASEM000/kernex v0.2.1, files ``kernex/_src/{utils,map,scan}.py``).  It exists to
*check* the CUDA engine in ``kernex_b200``; it is never imported by the product
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this file against
(1) every golden vector of the reference's own ``tests/test_scan_and_map.py``,
    ``tests/test_utils.py`` and the docstring / README examples, extracted by
    ``tests/golden/make_golden.py`` into ``tests/golden/*.json``, and
(2) outputs of the reference's *own source files* executed here on top of a
    NumPy-backed stand-in for the few JAX primitives it calls
    (``oracle/jax_shim``; real JAX is not installable in this image), stored in
    ``tests/golden/shim_cases.npz``.

Two evaluators live here:

* the *literal* evaluator (``kernel_map`` ... ``offset_kernel_scan``): a window
  loop that accepts arbitrary Python window functions, one window at a time,
  exactly in the reference's row-major order.  Slow, obviously correct.
* the *fast* evaluators (``affine_map``, ``reduce_map``, ``patch_map``,
  ``rowmarch_scan``): slice-based NumPy for the three classified stencil
  classes, used at sizes where the literal loop would take hours.  They are
  validated against the literal evaluator in ``tests/test_oracle_golden.py``.

All ``file:line`` citations are relative to ``/root/reference/kernex``.
"""

from __future__ import annotations

import itertools
import math
from typing import Any, Callable, Sequence

import numpy as np

__all__ = [
    "pad_width",
    "output_shape",
    "window_starts",
    "centre_offset",
    "offset_to_border",
    "set_indices",
    "roll_view",
    "key_matches",
    "select_function",
    "kernel_map",
    "kernel_scan",
    "offset_kernel_map",
    "offset_kernel_scan",
    "affine_map",
    "reduce_map",
    "patch_map",
    "rowmarch_scan",
    "weight_grad",
    "resolve_padding",
    "normalize_index",
]


# --------------------------------------------------------------------------- #
# geometry                                                                      #
# --------------------------------------------------------------------------- #


def pad_width(border):
    """High-side-only zero pad, ``_src/utils.py:43-53``.

    The reference pads ``max(0,l)+max(0,r)`` cells at the END of each axis and
    reaches low-side pad cells through negative-index wraparound.
    """
    return tuple((0, max(0, lo) + max(0, hi)) for lo, hi in border)


def output_shape(shape, kernel_size, strides, border):
    """``(n + l + r - k)//s + 1`` per axis, ``_src/utils.py:93-112``."""
    out = []
    for n, k, s, (lo, hi) in zip(shape, kernel_size, strides, border, strict=True):
        out.append((n + lo + hi - k) // s + 1)
    return tuple(out)


def window_starts(n, k, s, lo, hi):
    """First input coordinate of every window along one axis.

    ``general_arange`` (``_src/utils.py:205-242``): stride-1 windows start at
    ``-lo`` and there are ``n+lo+hi-k+1`` of them; every ``s``-th is kept.
    """
    count = n + lo + hi - k + 1
    if count <= 0:
        return np.zeros((0,), dtype=np.int64)
    return np.arange(-lo, -lo + count, 1, dtype=np.int64)[::s]


def centre_offset(k):
    """Position of the window centre inside a window, ``_src/utils.py:56-71``:
    ``k//2`` for odd k, ``(k-1)//2`` for even k -- i.e. always ``(k-1)//2``."""
    return (k - 1) // 2


def offset_to_border(offset, kernel_size):
    """``_offset_to_padding``, ``_src/utils.py:115-144``."""
    border = []
    for item, k in zip(offset, kernel_size):
        same = ((k - 1) // 2, k // 2)
        if isinstance(item, int):
            item = (item, item)
        item = tuple(item)
        if item < (0, 0):
            raise ValueError(f"Negative value found, offset={item}")
        border.append((same[0] - item[0], same[1] - item[1]))
    return tuple(border)


def set_indices(shape, strides, offset):
    """Write-back coordinates of the offset variants, ``_src/utils.py:195-202``."""
    out = []
    for n, s, item in zip(shape, strides, offset, strict=True):
        o0, o1 = (item, item) if isinstance(item, int) else item
        out.append(np.arange(o0, n - o1, s))
    return tuple(out)


def roll_view(patch):
    """Rotate every axis by ``-(k-1)//2`` so index 0 is the centre,
    ``_src/utils.py:147-181``.  Golden vectors: ``tests/test_utils.py:14-30``."""
    patch = np.asarray(patch)
    for ax, k in enumerate(patch.shape):
        if k == 0:
            continue
        patch = np.roll(patch, -((k - 1) // 2), axis=ax)
    return patch


# --------------------------------------------------------------------------- #
# region keys                                                                   #
# --------------------------------------------------------------------------- #


def _item_matches(c, item):
    """One dimension of one key, ``_src/utils.py:312-333``."""
    arr = np.atleast_1d(np.asarray(item, dtype=np.float64))
    if arr.size == 3:
        return bool(arr[0] <= c < arr[1]) and (c % arr[2] == 0)
    if arr.size == 2:
        return bool(arr[0] <= c < arr[1])
    if arr.size == 1:
        return True if np.isinf(arr[0]) else bool(c == arr[0])
    raise ValueError(f"unsupported key item {item!r}")


def key_matches(centre, key):
    """``_compare_key`` (``_src/utils.py:302-335``): all dims *present in the key*
    must match; ``zip`` truncates so a short key leaves trailing dims free and an
    empty key matches everything."""
    return all(_item_matches(c, item) for c, item in zip(centre, key))


def select_function(centre, key_groups):
    """``_key_search`` (``_src/utils.py:338-382``) followed by the clamp of
    ``lax.switch``: returns the index (in *insertion* order) of the function to
    run.  Groups are tried last-inserted first; a group matches if any of its
    keys does; no match falls through to the first-inserted function."""
    n = len(key_groups)
    for rev in range(n):
        group = key_groups[n - 1 - rev]
        if any(key_matches(centre, key) for key in group):
            return n - 1 - rev
    return 0


# --------------------------------------------------------------------------- #
# dtype rules (JAX with x64 disabled, weak Python scalars)                      #
# --------------------------------------------------------------------------- #


def _demote(dt):
    dt = np.dtype(dt)
    if dt == np.float64:
        return np.dtype(np.float32)
    if dt == np.int64:
        return np.dtype(np.int32)
    if dt == np.uint64:
        return np.dtype(np.uint32)
    return dt


def _is_weak(v):
    return isinstance(v, (bool, int, float)) and not isinstance(v, np.generic)


def _result_dtype(values):
    """Common dtype of per-window results the way ``vmap``/``lax.switch`` would
    produce it: strong (array-derived) dtypes win over weak Python scalars."""
    strong = [np.asarray(v).dtype for v in values if not _is_weak(v)]
    if strong:
        dt = _demote(np.result_type(*strong))
        if any(isinstance(v, float) for v in values if _is_weak(v)) and dt.kind in "iub":
            dt = np.dtype(np.float32)
        return dt
    if any(isinstance(v, float) for v in values):
        return np.dtype(np.float32)
    if all(isinstance(v, bool) for v in values):
        return np.dtype(bool)
    return np.dtype(np.int32)


# --------------------------------------------------------------------------- #
# literal evaluator                                                             #
# --------------------------------------------------------------------------- #


def _as_input(array):
    array = np.asarray(array)
    return array.astype(_demote(array.dtype), copy=False)


class _Geometry:
    def __init__(self, shape, kernel_size, strides, border):
        self.shape = tuple(int(v) for v in shape)
        self.kernel_size = tuple(int(v) for v in kernel_size)
        self.strides = tuple(int(v) for v in strides)
        self.border = tuple((int(lo), int(hi)) for lo, hi in border)
        self.pad = pad_width(self.border)
        self.out_shape = output_shape(self.shape, self.kernel_size, self.strides, self.border)
        self.starts = [
            window_starts(n, k, s, lo, hi)
            for n, k, s, (lo, hi) in zip(self.shape, self.kernel_size, self.strides, self.border)
        ]
        self.k_ranges = [np.arange(k) for k in self.kernel_size]
        self.coff = tuple(centre_offset(k) for k in self.kernel_size)

    def windows(self):
        """Row-major enumeration of window origins (``_src/utils.py:245-280``)."""
        return itertools.product(*[list(map(int, s)) for s in self.starts])

    def view(self, origin):
        return tuple(o + r for o, r in zip(origin, self.k_ranges))

    def centre(self, origin):
        return tuple(o + c for o, c in zip(origin, self.coff))


def _gather(state, view):
    # open-mesh gather with Python negative-index wraparound into the high-side
    # pad, ``_src/map.py:46`` / ``_src/scan.py:49``
    return state[np.ix_(*view)]


def _key_groups(func_map):
    return [tuple(v) for v in func_map.values()]


def kernel_map(func_map, shape, kernel_size, strides, border, relative=False,
               map_kind="vmap", map_kwargs=None):
    """Literal restatement of ``kernel_map`` (``_src/map.py:61-124``)."""
    geo = _Geometry(shape, kernel_size, strides, border)
    funcs = list(func_map.keys())
    groups = _key_groups(func_map)
    single = len(funcs) == 1

    def call(array, *a, **k):
        array = _as_input(array)
        padded = np.pad(array, geo.pad)
        results = []
        for origin in geo.windows():
            fi = 0 if single else select_function(geo.centre(origin), groups)
            patch = _gather(padded, geo.view(origin))
            if relative:
                patch = roll_view(patch)
            results.append(funcs[fi](patch, *a, **k))
        if not results:
            return np.zeros(geo.out_shape, dtype=array.dtype)
        dt = _result_dtype(results)
        stacked = np.stack([np.asarray(v, dtype=dt) for v in results])
        return stacked.reshape(*geo.out_shape, *stacked.shape[1:])

    return call


def kernel_scan(func_map, shape, kernel_size, strides, border, relative=False,
                scan_kind="scan", scan_kwargs=None):
    """Literal restatement of ``kernel_scan`` (``_src/scan.py:65-115``): the
    carried state is the padded array, each window's value is written at the
    window centre before the next window is read, and the emitted values (not
    the final state) are returned."""
    geo = _Geometry(shape, kernel_size, strides, border)
    funcs = list(func_map.keys())
    groups = _key_groups(func_map)
    single = len(funcs) == 1
    reverse = bool((scan_kwargs or {}).get("reverse", False))

    def call(array, *a, **k):
        array = _as_input(array)
        state = np.pad(array, geo.pad)
        origins = list(geo.windows())
        order = range(len(origins) - 1, -1, -1) if reverse else range(len(origins))
        emitted = np.zeros((len(origins),), dtype=state.dtype)
        for w in order:
            origin = origins[w]
            centre = geo.centre(origin)
            fi = 0 if single else select_function(centre, groups)
            patch = _gather(state, geo.view(origin))
            if relative:
                patch = roll_view(patch)
            value = funcs[fi](patch, *a, **k)
            if np.ndim(value) != 0:
                value = np.asarray(value).reshape(())
            state[centre] = value  # cast to the state dtype, ``_src/scan.py:50``
            emitted[w] = state[centre]
        return emitted.reshape(geo.out_shape)

    return call


def _offset_writeback(inner, shape, strides, offset, what):
    idx = set_indices(shape, strides, offset)

    def call(array, *a, **k):
        array = _as_input(array)
        result = inner(array, *a, **k)
        if tuple(result.shape) > tuple(array.shape):
            raise ValueError(f"{what} output must be scalar. Found {result.shape=}")
        out = array.copy()
        out[np.ix_(*idx)] = result
        return out

    return call


def offset_kernel_map(func_map, shape, kernel_size, strides, offset, relative=False,
                      map_kind="vmap", map_kwargs=None):
    """``offset_kernel_map`` (``_src/map.py:127-156``)."""
    offset = [(o, o) if isinstance(o, int) else tuple(o) for o in offset]
    inner = kernel_map(func_map, shape, kernel_size, strides,
                       offset_to_border(offset, kernel_size), relative, map_kind, map_kwargs)
    return _offset_writeback(inner, shape, strides, offset, "map")


def offset_kernel_scan(func_map, shape, kernel_size, strides, offset, relative=False,
                       scan_kind="scan", scan_kwargs=None):
    """``offset_kernel_scan`` (``_src/scan.py:118-147``)."""
    offset = [(o, o) if isinstance(o, int) else tuple(o) for o in offset]
    inner = kernel_scan(func_map, shape, kernel_size, strides,
                        offset_to_border(offset, kernel_size), relative, scan_kind, scan_kwargs)
    return _offset_writeback(inner, shape, strides, offset, "scan")


# --------------------------------------------------------------------------- #
# host-side argument resolution (interface/resolve_utils.py)                   #
# --------------------------------------------------------------------------- #


def resolve_padding(arg, kernel_size):
    """``_resolve_padding_argument`` (``interface/resolve_utils.py:26-75``)."""
    same = lambda k: ((k - 1) // 2, k // 2)

    def one(item, k):
        if isinstance(item, bool):
            raise TypeError("padding item must be int, tuple or str")
        if isinstance(item, int):
            return (item, item)
        if isinstance(item, tuple):
            return item
        if isinstance(item, str):
            if item.lower() == "same":
                return same(k)
            if item.lower() == "valid":
                return (0, 0)
            raise ValueError(f"bad padding string {item!r}")
        raise TypeError("padding item must be int, tuple or str")

    if isinstance(arg, tuple):
        assert len(arg) == len(kernel_size)
        return tuple(one(item, k) for item, k in zip(arg, kernel_size))
    if isinstance(arg, int):
        return ((arg, arg),) * len(kernel_size)
    if isinstance(arg, str):
        return tuple(one(arg, k) for k in kernel_size)
    raise TypeError("input_argument must be tuple or int or str")


def normalize_index(index, shape):
    """``_resolve_index`` (``interface/resolve_utils.py:120-154``): slice ->
    ``(start,end,step)`` with ``x or default`` semantics, int -> non-negative."""
    index = index if isinstance(index, tuple) else (index,)
    out = []
    for item, n in zip(index, shape):
        if isinstance(item, int):
            out.append(item + n if item < 0 else item)
        elif isinstance(item, slice):
            start = item.start or 0
            start += n if start < 0 else 0
            end = item.stop or n
            end += n if end < 0 else 0
            out.append((start, end, item.step or 1))
        elif isinstance(item, (list, tuple)):
            out.append(item)
        else:
            raise NotImplementedError(type(item))
    return out


# --------------------------------------------------------------------------- #
# fast evaluators for the three classified stencil classes                      #
# --------------------------------------------------------------------------- #


def _padded_for_slices(array, kernel_size, strides, border):
    """True two-sided zero pad / crop so that window j starts at ``j*s`` in the
    returned array (equivalent to the reference's high-side pad + wraparound)."""
    array = np.asarray(array)
    out_shape = output_shape(array.shape, kernel_size, strides, border)
    pads, crops = [], []
    for n, (lo, hi) in zip(array.shape, border):
        pads.append((max(lo, 0), max(hi, 0)))
        crops.append(slice(max(-lo, 0), n - max(-hi, 0)))
    work = np.pad(array[tuple(crops)], pads)
    return work, out_shape


def _tap_slice(work, tap, out_shape, strides):
    sl = tuple(slice(t, t + (o - 1) * s + 1, s) for t, o, s in zip(tap, out_shape, strides))
    return work[sl]


def affine_map(array, kernel_size, strides, border, taps, coefs, bias=0, out_dtype=None):
    """``out[j] = bias + sum_t coefs[t] * x[origin(j) + taps[t]]`` accumulated in
    tap order in the array's dtype (fp32 FMA chain order = order of ``taps``).
    ``taps`` are offsets from the window ORIGIN (0 <= tap < k)."""
    array = _as_input(array)
    dt = np.dtype(out_dtype or array.dtype)
    work, out_shape = _padded_for_slices(array, kernel_size, strides, border)
    if any(o <= 0 for o in out_shape):
        return np.zeros([max(o, 0) for o in out_shape], dtype=dt)
    acc = np.full(out_shape, bias, dtype=dt)
    with np.errstate(over="ignore"):
        for tap, c in zip(taps, coefs):
            acc += np.asarray(c, dtype=dt) * _tap_slice(work, tap, out_shape, strides).astype(dt, copy=False)
    return acc


def reduce_map(array, kernel_size, strides, border, op):
    """Window reduction over the full window (zero padding participates, as in
    the reference where the pad cells are real zeros)."""
    array = _as_input(array)
    work, out_shape = _padded_for_slices(array, kernel_size, strides, border)
    taps = list(itertools.product(*[range(k) for k in kernel_size]))
    if op == "mean":
        acc = np.zeros(out_shape, dtype=np.float32)
        for tap in taps:
            acc += _tap_slice(work, tap, out_shape, strides).astype(np.float32)
        return (acc / np.float32(len(taps))).astype(np.float32)
    acc = _tap_slice(work, taps[0], out_shape, strides).copy()
    with np.errstate(over="ignore"):
        for tap in taps[1:]:
            v = _tap_slice(work, tap, out_shape, strides)
            if op == "sum":
                acc += v
            elif op == "max":
                np.maximum(acc, v, out=acc)
            elif op == "min":
                np.minimum(acc, v, out=acc)
            else:
                raise ValueError(op)
    return acc


def patch_map(array, kernel_size, strides, border, relative=False):
    """``kmap(lambda x: x)``: output ``out_shape + kernel_size``; with
    ``relative`` each patch is rolled so its centre sits at index 0."""
    array = _as_input(array)
    work, out_shape = _padded_for_slices(array, kernel_size, strides, border)
    out = np.zeros(tuple(out_shape) + tuple(kernel_size), dtype=array.dtype)
    for tap in itertools.product(*[range(k) for k in kernel_size]):
        dst = tuple((t - centre_offset(k)) % k for t, k in zip(tap, kernel_size)) if relative else tap
        out[(Ellipsis,) + dst] = _tap_slice(work, tap, out_shape, strides)
    return out


def weight_grad(array, cotangent, kernel_size, strides, border):
    """d/dw of ``sum_j g[j] * sum_t w[t] x[origin(j)+t]``: ``dw[t] = sum_j g[j]
    x[origin(j)+t]`` (JAX transposes the gather), in float64."""
    array = _as_input(array)
    work, out_shape = _padded_for_slices(array, kernel_size, strides, border)
    g = np.asarray(cotangent, dtype=np.float64)
    dw = np.zeros(kernel_size, dtype=np.float64)
    for tap in itertools.product(*[range(k) for k in kernel_size]):
        dw[tap] = np.sum(g * _tap_slice(work, tap, out_shape, strides))
    return dw


def rowmarch_scan(array, regions, nrows=None):
    """Row-marching evaluator for 2-D kscans (kernel (3,3), border ((1,1),(1,1)),
    stride 1) whose functions only read the previous row and constants.

    ``regions`` is the *resolved* per-row rule: a callable
    ``rule(n) -> list[(col_lo, col_hi, taps, coefs, bias)]`` in increasing
    priority (later entries overwrite earlier ones); ``taps`` are column offsets
    into row ``n-1`` of the *state* (row -1 is the zero pad).  Vectorised per row:
    equivalent to the sequential sweep because no function reads the row being
    written.
    """
    array = _as_input(array)
    nt, nx = array.shape
    nt = nrows or nt
    out = np.zeros((nt, nx), dtype=array.dtype)
    prev = np.zeros((nx + 2,), dtype=array.dtype)  # state row n-1 with pad cols
    for n in range(nt):
        row = np.zeros((nx,), dtype=array.dtype)
        for lo, hi, taps, coefs, bias in regions(n):
            acc = np.full((hi - lo,), bias, dtype=array.dtype)
            for d, c in zip(taps, coefs):
                acc += np.asarray(c, dtype=array.dtype) * prev[1 + lo + d: 1 + hi + d]
            row[lo:hi] = acc
        out[n] = row
        prev[1:-1] = row
    return out
