"""The four engine factories of kernex, re-implemented on the B200 CUDA library.

Drop-in for ``kernel_map`` / ``offset_kernel_map`` (``kernex/_src/map.py:61-124,
127-156``) and ``kernel_scan`` / ``offset_kernel_scan`` (``kernex/_src/scan.py:
65-115, 118-147``): same positional signature ``(func_map, shape, kernel_size,
strides, border|offset, relative, kind, kwargs)`` and same ``func_map`` key format
(``()`` single function, ``[[]]`` match-all, lists of per-dim ``int |
(start,end[,step]) | inf`` keys), returning ``fn(array, *args, **kwargs)``.

Nothing is computed in Python: the window functions are traced + classified
(``tracer.py``), described as a ``KxProgram`` and executed by ``libkx_b200.so``
through the C ABI of ``include/kx_b200.h``.  Arrays may be NumPy arrays (copied
to the device and back, the end-to-end path) or CUDA ``torch`` tensors (zero
copy, result stays on the device).  PyTorch is used for device memory, streams
and autograd plumbing only.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Any, Callable, Sequence

import numpy as np

from . import _lib
from .errors import EngineUnavailableError, UnsupportedStencilError
from .tracer import FuncSpec, trace_function

try:
    import torch
except Exception as e:  # pragma: no cover
    raise EngineUnavailableError(f"kernex_b200 needs PyTorch for device memory and streams: {e}")

__all__ = ["kernel_map", "kernel_scan", "offset_kernel_map", "offset_kernel_scan"]

_NP2KX = {np.dtype(np.float32): _lib.KX_F32, np.dtype(np.int32): _lib.KX_I32}
_KX2TORCH = {_lib.KX_F32: torch.float32, _lib.KX_I32: torch.int32}
_TORCH2KX = {torch.float32: _lib.KX_F32, torch.int32: _lib.KX_I32}
# JAX keeps narrow integer / bool arrays narrow (results wrap at 8 / 16 bits, `sum` promotes, bool arithmetic is
# logical); widening them silently to int32 would return different values AND a different dtype than the reference,
# so they are refused instead (float64 / int64 are demoted exactly like JAX without x64).
_NARROW_NP = (np.dtype(np.int16), np.dtype(np.int8), np.dtype(np.uint8), np.dtype(np.uint16), np.dtype(np.bool_))
_NARROW_TORCH = (torch.int16, torch.int8, torch.uint8, torch.bool)
_NARROW_MSG = ("dtype {} is not supported: the engine computes in float32 / int32 and will not widen a narrow integer "
               "or bool array silently (the reference keeps the narrow dtype and its wraparound); cast the input "
               "explicitly, e.g. x.astype(np.int32)")


# --------------------------------------------------------------------------- #
# geometry helpers (kernex/_src/utils.py:93-144,195-202)                        #
# --------------------------------------------------------------------------- #


def calculate_output_shape(shape, kernel_size, strides, border):
    return tuple((n + lo + hi - k) // s + 1
                 for n, k, s, (lo, hi) in zip(shape, kernel_size, strides, border, strict=True))


def offset_to_border(offset, kernel_size):
    out = []
    for item, k in zip(offset, kernel_size):
        item = (item, item) if isinstance(item, int) else tuple(item)
        if item < (0, 0):
            raise ValueError(f"Negative value found, offset={item}")
        out.append(((k - 1) // 2 - item[0], k // 2 - item[1]))
    return tuple(out)


# --------------------------------------------------------------------------- #
# array plumbing                                                                #
# --------------------------------------------------------------------------- #


def _require_cuda():
    if not torch.cuda.is_available():
        raise EngineUnavailableError(
            "no CUDA device: kernex_b200 executes on sm_100a only and has no CPU fallback")


class _Arr:
    """An input array moved to the device, remembering how to hand results back."""

    def __init__(self, array):
        self.was_numpy = not isinstance(array, torch.Tensor)
        if self.was_numpy:
            host = np.asarray(array)
            if host.dtype == np.float64:       # JAX without x64 demotes too
                host = host.astype(np.float32)
            elif host.dtype == np.int64:         # ... and int64 -> int32
                host = host.astype(np.int32)
            elif host.dtype in _NARROW_NP:
                raise UnsupportedStencilError(_NARROW_MSG.format(host.dtype))
            if host.dtype not in _NP2KX:
                raise UnsupportedStencilError(f"dtype {host.dtype} is not supported (float32 / int32 only)")
            _require_cuda()
            self.t = torch.from_numpy(np.ascontiguousarray(host)).cuda(non_blocking=True)
        else:
            t = array
            if t.dtype == torch.float64:
                t = t.float()
            elif t.dtype == torch.int64:
                t = t.int()
            elif t.dtype in _NARROW_TORCH:
                raise UnsupportedStencilError(_NARROW_MSG.format(t.dtype))
            if t.dtype not in _TORCH2KX:
                raise UnsupportedStencilError(f"dtype {t.dtype} is not supported (float32 / int32 only)")
            if not t.is_cuda:
                _require_cuda()
                t = t.cuda()
            self.t = t.contiguous()
        self.kx_dtype = _TORCH2KX[self.t.dtype]

    def give_back(self, t):
        if not self.was_numpy:
            return t
        # device -> pinned host (torch's caching host allocator recycles the staging
        # block once the returned array is dropped) -> NumPy view, one stream sync
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return host.numpy()


def _stream_ptr(device):
    # raw handle of torch's current stream (torch.cuda.current_stream() costs ~7 us of Python per call)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(idx))
    except AttributeError:  # pragma: no cover
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# --------------------------------------------------------------------------- #
# program construction                                                          #
# --------------------------------------------------------------------------- #


def _key_item(item):
    it = _lib.KxKeyItem()
    if isinstance(item, (tuple, list, np.ndarray)):
        vals = list(item)
    else:
        vals = [item]
    inf = [isinstance(v, float) and math.isinf(v) for v in vals]
    big = (1 << 62)
    if len(vals) == 1:
        if inf[0]:
            it.kind = _lib.KX_KEY_ANY
        else:
            it.kind, it.a = _lib.KX_KEY_EQ, int(vals[0])
    elif len(vals) in (2, 3):
        it.kind = _lib.KX_KEY_RANGE if len(vals) == 2 else _lib.KX_KEY_RANGE_STEP
        it.a = -big if (inf[0] and vals[0] < 0) else (big if inf[0] else int(vals[0]))
        it.b = big if (inf[1] and vals[1] > 0) else (-big if inf[1] else int(vals[1]))
        if len(vals) == 3:
            if inf[2]:
                raise ValueError("infinite step in a region key")
            it.step = int(vals[2])
            if it.step == 0:
                raise ZeroDivisionError("region key with step 0")
    else:
        raise ValueError(f"region key item {item!r} must have 1, 2 or 3 entries")
    return it


class _Program:
    """Host description of the traced functions -> ``KxProgram``."""

    def __init__(self, specs: Sequence[FuncSpec], key_groups, ndim):
        if len(specs) > _lib.KX_MAX_FUNCS:
            raise UnsupportedStencilError(f"at most {_lib.KX_MAX_FUNCS} functions per mesh")
        self.specs = list(specs)
        p = _lib.KxProgram()
        p.n_funcs = len(specs)
        t = 0
        self.device_refs = []  # (tap slot, Coef) with weight references
        for fi, s in enumerate(specs):
            if t + len(s.taps) > _lib.KX_MAX_TAPS:
                raise UnsupportedStencilError(f"more than {_lib.KX_MAX_TAPS} taps in one program")
            f = p.func[fi]
            f.kind = {"affine": _lib.KX_AFFINE, "max": _lib.KX_REDUCE_MAX, "min": _lib.KX_REDUCE_MIN,
                      "gather": _lib.KX_GATHER}[s.kind]
            f.tap_begin = t
            for ti, tap in enumerate(s.taps):
                for d in range(ndim):
                    p.tap_off[t][d] = int(tap[d])
                if s.kind == "affine":
                    c = s.coefs[ti]
                    p.coef[t] = float(c.const)
                    if not c.is_const:
                        self.device_refs.append((t, c))
                else:
                    p.coef[t] = 1.0
                t += 1
            f.tap_end = t
            if s.kind == "affine":
                if not s.bias.is_const:
                    raise UnsupportedStencilError("a device-resident additive term (bias) is not supported")
                f.bias = float(s.bias.const)
                f.post_div = float(s.post_div)
                f.coef_is_int = int(s.integral or not s.weak_float)  # integer device weights stay integer
            else:
                f.coef_is_int = 1
        p.n_taps = t
        nk = 0
        if len(specs) > 1:
            for fi, group in enumerate(key_groups):
                for key in group:
                    if nk >= _lib.KX_MAX_KEYS:
                        raise UnsupportedStencilError(f"more than {_lib.KX_MAX_KEYS} region keys")
                    k = p.key[nk]
                    k.func = fi
                    items = list(key)[:ndim]
                    k.n_items = len(items)
                    for d, item in enumerate(items):
                        k.item[d] = _key_item(item)
                    nk += 1
        p.n_keys = nk
        self.c = p

    def coef_tensor(self, device_args, dtype, device):
        """Device coefficient table ``const + sum scale * W.flatten()[idx]`` built with
        torch ops so that autograd reaches the weights.  The index / scale tensors are
        uploaded once per (program, device); when the table is exactly ``W.flatten()`` (the
        usual ``sum(x*w)``) the weight tensor itself is the table: zero extra launches."""
        if not self.device_refs:
            return None
        n = self.c.n_taps
        cache = self.__dict__.setdefault("_coef_cache", {})
        ck = (str(device), dtype)
        if ck not in cache:
            by_arg: dict[Any, list] = {}
            for slot, c in self.device_refs:
                for (arg, flat), scale in c.w.items():
                    by_arg.setdefault(arg, []).append((slot, flat, scale))
            base = [self.c.coef[i] for i in range(n)]
            identity = None
            if len(by_arg) == 1 and all(v == 0.0 for v in base):
                (arg, refs), = by_arg.items()
                if [r[0] for r in refs] == list(range(n)) and [r[1] for r in refs] == list(range(n)) \
                        and all(r[2] == 1 for r in refs):
                    identity = arg
            plan = []
            if identity is None:
                for arg, refs in by_arg.items():
                    plan.append((arg,
                                 torch.tensor([r[0] for r in refs], dtype=torch.long, device=device),
                                 torch.tensor([r[1] for r in refs], dtype=torch.long, device=device),
                                 torch.tensor([r[2] for r in refs], dtype=dtype, device=device)))
            cache[ck] = (identity, torch.tensor(base, dtype=dtype, device=device), plan)
        identity, base, plan = cache[ck]
        if identity is not None:
            w = device_args[identity]
            if w.numel() == n:
                return w.to(device=device, dtype=dtype).reshape(-1).contiguous()
        coef = base
        for arg, slots, flats, scales in plan if identity is None else []:
            w = device_args[arg]
            coef = coef.index_add(0, slots, w.to(device=device, dtype=dtype).reshape(-1)[flats] * scales)
        if identity is not None:  # numel mismatch (broadcast weights): rebuild the general way
            w = device_args[identity]
            idx = torch.arange(n, device=device)
            coef = base.index_add(0, idx, w.to(device=device, dtype=dtype).reshape(-1)[idx])
        return coef.contiguous()


def _geom(in_shape, out_shape, kernel_size, strides, border, in_dtype, out_dtype):
    nd = len(in_shape)
    if nd > _lib.KX_MAX_DIMS:
        raise UnsupportedStencilError(f"arrays of rank > {_lib.KX_MAX_DIMS} are not supported")
    g = _lib.KxGeom()
    g.ndim, g.in_dtype, g.out_dtype = nd, in_dtype, out_dtype
    for d in range(nd):
        g.in_shape[d], g.out_shape[d] = int(in_shape[d]), int(out_shape[d])
        g.ksize[d], g.stride[d] = int(kernel_size[d]), int(strides[d])
        g.lo[d], g.hi[d] = int(border[d][0]), int(border[d][1])
    return g


def _result_dtype(specs, in_dtype):
    """JAX's weak-type promotion for the per-window results: taps
    carry the array dtype, Python numbers are weak; int array + float factor -> f32."""
    strong = any(s.kind != "affine" or s.taps for s in specs)
    any_float = any(s.kind == "affine" and s.weak_float for s in specs)
    if strong:
        if in_dtype == _lib.KX_I32 and any_float:
            return _lib.KX_F32
        return in_dtype
    return _lib.KX_F32 if any_float else _lib.KX_I32


def _trace_all(func_map, kernel_size, relative, a, k):
    funcs = list(func_map.keys())
    specs, device_args = [], {}
    for f in funcs:
        spec, dev = trace_function(f, kernel_size, relative, a, k)
        specs.append(spec)
        device_args.update(dev)
    return specs, device_args


def _validate(shape, kernel_size, strides, border):
    if not (len(shape) == len(kernel_size) == len(strides) == len(border)):
        raise ValueError("shape, kernel_size, strides and border must have the same length")
    if any(s <= 0 for s in strides):
        raise ValueError(f"strides must be positive, found {strides}")
    if any(k <= 0 for k in kernel_size):
        raise ValueError(f"kernel_size must be positive, found {kernel_size}")


# --------------------------------------------------------------------------- #
# autograd (the reference gets this from JAX's transpose rules)     #
# --------------------------------------------------------------------------- #


class _AffineMapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coef, geom, prog, out_shape, out_dtype):
        lib = _lib.load()
        out = torch.empty(out_shape, dtype=out_dtype, device=x.device)
        if out.numel():
            _lib.check(lib.kx_map(C.byref(geom), C.byref(prog.c), _ptr(x), _ptr(coef), 0, _ptr(out), 1,
                                  _stream_ptr(x.device)))
        ctx.geom, ctx.prog = geom, prog
        ctx.save_for_backward(x, coef)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, coef = ctx.saved_tensors
        geom, prog = ctx.geom, ctx.prog
        g = g.contiguous()
        dx = dcoef = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.check(lib.kx_map_vjp_input(C.byref(geom), C.byref(prog.c), _ptr(g), _ptr(coef), _ptr(dx),
                                            _stream_ptr(x.device)))
        if coef is not None and ctx.needs_input_grad[1]:
            nbytes = lib.kx_vjp_coef_scratch_bytes(C.byref(geom), C.byref(prog.c))
            scratch = torch.empty(max(int(nbytes), 4), dtype=torch.uint8, device=x.device)
            dcoef = torch.empty_like(coef)
            _lib.check(lib.kx_map_vjp_coef(C.byref(geom), C.byref(prog.c), _ptr(x), _ptr(g), _ptr(scratch),
                                           _ptr(dcoef), _stream_ptr(x.device)))
        return dx, dcoef, None, None, None, None


# --------------------------------------------------------------------------- #
# kmap                                                                          #
# --------------------------------------------------------------------------- #


_PLAN_CACHE: "dict[tuple, tuple]" = {}
_PLAN_CACHE_MAX = 128
_PROMOTE_MIN = 1 << 16      # int32 arrays with a float32 result: convert first from this many elements on
_PLAN_GEN = [0]


def _arg_signature(v):
    """Hashable description of an extra argument as far as tracing is concerned: device
    tensors are symbolic (only shape / dtype matter), host values are baked into the taps."""
    if isinstance(v, torch.Tensor):
        if v.is_cuda or v.requires_grad:
            return ("dev", tuple(v.shape), str(v.dtype), bool(v.requires_grad))
        v = v.detach().numpy()
    if isinstance(v, np.ndarray):
        if v.nbytes > 8192:
            raise TypeError("large host array: not cached")
        return ("arr", v.shape, str(v.dtype), v.tobytes())
    if isinstance(v, (bool, int, float, str, type(None), np.generic)):
        return ("val", type(v).__name__, v)
    raise TypeError("unhashable argument kind")


def _captured_signature(v, depth):
    """Signature of a value a window function captured (closure cell / referenced global)."""
    import types
    if isinstance(v, (types.ModuleType, type, types.BuiltinFunctionType)) or isinstance(v, np.ufunc):
        return ("id", id(v))
    if isinstance(v, (types.FunctionType, types.MethodType)) or hasattr(v, "__wrapped__"):
        return _func_fingerprint(v, depth + 1)
    if isinstance(v, (list, tuple)) and len(v) <= 64:
        return (type(v).__name__,) + tuple(_captured_signature(e, depth) for e in v)
    if isinstance(v, dict) and len(v) <= 64:
        return ("dict",) + tuple((repr(n), _captured_signature(e, depth)) for n, e in v.items())
    if isinstance(v, torch.Tensor) and (v.is_cuda or v.requires_grad):
        raise TypeError("captured device tensor: its values are baked at trace time, not cached")
    return _arg_signature(v)


class _CapturePlan:
    """What a function can capture, resolved once per function object: its closure cells, the global names its code
    (and nested code objects) reads that are not modules / types / builtins, its defaults.  Evaluating the fingerprint
    is then a walk over these few references -- or nothing at all for a plain lambda."""

    __slots__ = ("code_id", "cells", "glb", "names", "fn", "trivial")

    def __init__(self, f):
        import types
        fn = getattr(f, "__func__", f)
        code = getattr(fn, "__code__", None)
        if code is None:
            fn = getattr(f, "__wrapped__", None) or getattr(type(f), "__call__", None)
            code = getattr(fn, "__code__", None)
            if code is None:
                raise TypeError("callable without code object")
        self.fn, self.code_id = fn, id(code)
        self.cells = tuple(zip(code.co_freevars, fn.__closure__ or ()))
        self.glb = fn.__globals__
        names = set(code.co_names)
        for const in code.co_consts:      # nested lambdas / comprehensions read globals too
            if hasattr(const, "co_names"):
                names.update(const.co_names)
        skip = (types.ModuleType, type, types.BuiltinFunctionType, np.ufunc)
        self.names = tuple(n for n in sorted(names) if n in self.glb and not isinstance(self.glb[n], skip))
        self.trivial = not self.cells and not self.names


_CAPTURE_PLANS: "dict[int, tuple]" = {}


def _capture_plan(f):
    ent = _CAPTURE_PLANS.get(id(f))
    if ent is not None and ent[0] is f:
        return ent[1]
    plan = _CapturePlan(f)
    if len(_CAPTURE_PLANS) > 4096:
        _CAPTURE_PLANS.clear()
    _CAPTURE_PLANS[id(f)] = (f, plan)   # keeps f alive: the id stays unique
    return plan


def _func_fingerprint(f, depth=0):
    """The tracer bakes the values a window function reads from its closure and globals (dt, alpha, host weight
    tables) into the coefficients, while the un-jitted reference re-evaluates the function on every call
    (kernel_interface.py:130-157).  The plan cache therefore keys on those VALUES too: a captured variable that
    changes -- rebinding or in-place mutation of a small array -- gives a new key and a fresh trace.  Anything the
    fingerprint cannot describe (large arrays, device tensors, arbitrary objects) raises -> the call is not cached."""
    if depth > 3:
        raise TypeError("captured functions nest too deeply")
    plan = _capture_plan(f)
    fn = plan.fn
    if plan.trivial and not fn.__defaults__ and not fn.__kwdefaults__:
        return plan.code_id
    parts = [plan.code_id]
    for name, cell in plan.cells:
        try:
            val = cell.cell_contents
        except ValueError:            # empty cell
            continue
        parts.append((name, _captured_signature(val, depth)))
    glb = plan.glb
    for name in plan.names:
        if name in glb:
            parts.append((name, _captured_signature(glb[name], depth)))
    if fn.__defaults__:
        parts.append(("defaults", tuple(_captured_signature(v, depth) for v in fn.__defaults__)))
    if fn.__kwdefaults__:
        parts.append(("kwdefaults", tuple((n, _captured_signature(v, depth)) for n, v in sorted(fn.__kwdefaults__.items()))))
    return tuple(parts)


def clear_plan_cache():
    """Drop every cached trace (public: ``kernex_b200.clear_plan_cache()``)."""
    _PLAN_CACHE.clear()
    _PLAN_GEN[0] += 1       # steady-state launchers of earlier generations retire themselves


def _static_key(tag, func_map, shape, kernel_size, strides, border, relative):
    """The part of a plan key that cannot change between two calls of one engine closure."""
    try:
        return (tag, tuple(id(f) for f in func_map), repr([v for v in func_map.values()]), shape, kernel_size, strides,
                border, bool(relative))
    except Exception:
        return None


def _plan_key(tag, func_map, shape, kernel_size, strides, border, relative, in_kx, a=(), k=None, static=None):
    if static is None:
        static = _static_key(tag, func_map, shape, kernel_size, strides, border, relative)
        if static is None:
            return None
    try:
        sig = tuple(_arg_signature(v) for v in a) if a else ()
        if k:
            sig += tuple((n, _arg_signature(v)) for n, v in sorted(k.items()))
        captured = tuple(_func_fingerprint(f) for f in func_map)
    except Exception:
        return None
    return (static, in_kx, sig, captured)


def _fresh_device_args(a, k):
    """Device tensors are looked up afresh on a cache hit: the cached plan only remembers their positions."""
    fresh = {i: v for i, v in enumerate(a) if isinstance(v, torch.Tensor) and (v.is_cuda or v.requires_grad)}
    fresh.update({("kw", n): v for n, v in (k or {}).items()
                  if isinstance(v, torch.Tensor) and (v.is_cuda or v.requires_grad)})
    return fresh


def _cache_plan(key, func_map, plan):
    if key is not None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        _PLAN_CACHE[key] = (tuple(func_map), plan)  # keep the functions alive: ids stay unique


def _map_plan(func_map, shape, kernel_size, strides, border, relative, in_kx, a, k, static=None):
    """Trace + classify + build the KxGeom / KxProgram.  Plans of argument-free window
    functions are cached on the function objects' identity (the reference re-traces under
    ``jit`` at most once per shape too), so steady-state calls cost one ctypes launch."""
    key = _plan_key("map", func_map, shape, kernel_size, strides, border, relative, in_kx, a, k, static)
    if key is not None and key in _PLAN_CACHE:
        specs, _stale_args, out_kx, prog, geom, full_shape = _PLAN_CACHE[key][1]
        return specs, _fresh_device_args(a, k), out_kx, prog, geom, full_shape
    specs, device_args = _trace_all(func_map, kernel_size, relative, a, k)
    groups = [tuple(v) for v in func_map.values()]
    out_shape = calculate_output_shape(shape, kernel_size, strides, border)
    out_shape = tuple(max(o, 0) for o in out_shape)
    gather = [s for s in specs if s.kind == "gather"]
    if gather and len(specs) > 1:
        raise UnsupportedStencilError("vector-valued functions cannot be mixed in a multi-function mesh")
    out_kx = _result_dtype(specs, in_kx)
    prog = _Program(specs, groups, len(shape))
    geom = _geom(shape, out_shape, kernel_size, strides, border, in_kx, out_kx)
    full_shape = out_shape + (gather[0].out_shape if gather else ())
    plan = (specs, device_args, out_kx, prog, geom, full_shape)
    _cache_plan(key, func_map, plan)
    return plan


def _win_max(v):
    return np.max(v)


def _win_min(v):
    return np.min(v)


_SEP3D_OPS: dict = {}


def _separable_reduce_3d(specs, shape, kernel_size, strides, border, arr):
    """Whole-window max / min over a 3-D window with stride 1 (a 3-D max filter; the reference's zero padding takes part
    in the reduction): max over the window = max along x of max along y of max along z, each with its own zero
    padding, and a reduction has no rounding -- so three passes of the 2-D row-march kernel (1 x kx, ky x 1 and kz x 1
    windows, 24 B per cell) give bit for bit what one 3-D pass gives; the only single-pass kernel for it is the generic
    one (0.04 of the HBM roofline).  Device arrays of >= 2^16 elements; anything else returns None."""
    nd = len(shape)
    if nd < 3 or len(specs) != 1 or specs[0].kind not in ("max", "min") or arr.t.numel() < _PROMOTE_MIN:
        return None
    if any(s != 1 for s in strides) or kernel_size[nd - 3] < 2 or any(k != 1 for k in kernel_size[:nd - 3]):
        return None
    if any(tuple(b) != (0, 0) for b in border[:nd - 3]):
        return None
    if len(specs[0].taps) != int(np.prod(kernel_size)) or any(k not in (1, 3, 5, 7) for k in kernel_size[nd - 3:]):   # the 2-D kernel's reduction windows
        return None
    fn = _win_max if specs[0].kind == "max" else _win_min
    lead = int(np.prod(shape[:nd - 3])) if nd > 3 else 1
    D, H, W = shape[nd - 3:]
    kz, ky, kx = kernel_size[nd - 3:]
    bz, by, bx = [tuple(b) for b in border[nd - 3:]]
    Wo, Ho, Do = W + sum(bx) - kx + 1, H + sum(by) - ky + 1, D + sum(bz) - kz + 1
    if min(Wo, Ho, Do) < 1:
        return None
    key = (specs[0].kind, lead, D, H, W, kz, ky, kx, bz, by, bx)
    ops = _SEP3D_OPS.get(key)
    if ops is None:
        ops = (kernel_map({fn: ()}, (lead * D * H, W), (1, kx), (1, 1), ((0, 0), bx)),
               kernel_map({fn: ()}, (lead * D, H, Wo), (1, ky, 1), (1, 1, 1), ((0, 0), by, (0, 0))),
               kernel_map({fn: ()}, (lead, D, Ho * Wo), (1, kz, 1), (1, 1, 1), ((0, 0), bz, (0, 0))))
        if len(_SEP3D_OPS) < 64:
            _SEP3D_OPS[key] = ops
    t = arr.t.reshape(lead * D * H, W)
    t = ops[0](t).reshape(lead * D, H, Wo)
    t = ops[1](t).reshape(lead, D, Ho * Wo)
    t = ops[2](t)
    return t.reshape(tuple(shape[:nd - 3]) + (Do, Ho, Wo))


class _FastMap:
    """Steady-state launcher of one engine closure for device-resident inputs: everything the general path derives
    per call (plan key, program, geometry, output shape) is fixed once that path has run, so a repeat call is the
    checks below, one ``torch.empty`` and the ``kx_map`` call.  BASELINE config 1 and the reference's own benchmark
    shapes (tests_and_benchmarks/test_benchmark_{patch,conv2d}.py) are launch-latency bound: this is their hot path
    (benchmarks/call_layers.py).  Returns None whenever a precondition fails -> the general path decides."""

    __slots__ = ("funcs", "captured", "shape", "in_dtype", "device", "dev_index", "full_shape", "out_dtype", "geom", "prog",
                 "geom_ref", "prog_ref", "kx_map", "weight", "raw_stream", "gen")

    def __init__(self, func_map, key, shape, t, w, full_shape, out_dtype, geom, prog):
        self.funcs = tuple(func_map)
        plans = [_capture_plan(f) for f in self.funcs]
        trivial = all(pl.trivial and not pl.fn.__defaults__ and not pl.fn.__kwdefaults__ for pl in plans)
        self.captured = None if trivial else key[3]     # captured VALUES are re-checked per call unless there are none
        self.shape, self.in_dtype, self.device = tuple(shape), t.dtype, t.device
        self.dev_index = t.device.index if t.device.index is not None else torch.cuda.current_device()
        self.full_shape, self.out_dtype = tuple(full_shape), out_dtype
        self.geom, self.prog = geom, prog               # keep the ctypes structures alive
        self.geom_ref, self.prog_ref = C.byref(geom), C.byref(prog.c)
        self.kx_map = _lib.load().kx_map
        self.weight = None if w is None else (tuple(w.shape), w.dtype)
        self.raw_stream = torch._C._cuda_getCurrentRawStream
        self.gen = _PLAN_GEN[0]

    def __call__(self, t, w):
        if self.gen != _PLAN_GEN[0] or t.shape != self.shape or t.dtype is not self.in_dtype or t.device != self.device or not t.is_contiguous():
            return None
        coef = 0
        if w is not None:
            if self.weight is None or type(w) is not torch.Tensor or (tuple(w.shape), w.dtype) != self.weight \
                    or w.device != self.device or not w.is_contiguous() or w.requires_grad:
                return None
            coef = w.data_ptr()
        elif self.weight is not None:
            return None
        if t.requires_grad and torch.is_grad_enabled():
            return None
        if self.captured is not None:
            try:
                if tuple(_func_fingerprint(f) for f in self.funcs) != self.captured:
                    return None
            except Exception:
                return None
        out = torch.empty(self.full_shape, dtype=self.out_dtype, device=self.device)
        rc = self.kx_map(self.geom_ref, self.prog_ref, t.data_ptr(), coef, 0, out.data_ptr(), 1, self.raw_stream(self.dev_index))
        if rc:
            _lib.check(rc)
        return out


def _run_map(func_map, shape, kernel_size, strides, border, relative, array, a, k, static=None, fast_out=None):
    arr = _Arr(array)
    if tuple(arr.t.shape) != tuple(shape):
        raise ValueError(f"array shape {tuple(arr.t.shape)} != declared shape {tuple(shape)}")
    specs, device_args, out_kx, prog, geom, full_shape = _map_plan(
        func_map, shape, kernel_size, strides, border, relative, arr.kx_dtype, a, k, static)
    sep = _separable_reduce_3d(specs, shape, kernel_size, strides, border, arr)
    if sep is not None:
        return arr, sep
    if arr.kx_dtype == _lib.KX_I32 and out_kx == _lib.KX_F32 and arr.t.numel() >= _PROMOTE_MIN \
            and all(s.kind == "affine" for s in specs):
        # int32 array, float result (jnp.mean of integers, a Python-float coefficient): JAX converts the window to
        # float32 before the arithmetic, so convert the ARRAY once (kx_cast_i32_f32) and run the float32 templates; the
        # mixed-dtype form exists only in the generic kernel (0.04-0.06 of the HBM roofline on 8192^2)
        t32 = torch.empty(arr.t.shape, dtype=torch.float32, device=arr.t.device)
        _lib.check(_lib.load().kx_cast_i32_f32(_ptr(arr.t), _ptr(t32), arr.t.numel(), _stream_ptr(arr.t.device)))
        arr.t, arr.kx_dtype = t32, _lib.KX_F32
        specs, device_args, out_kx, prog, geom, full_shape = _map_plan(
            func_map, shape, kernel_size, strides, border, relative, _lib.KX_F32, a, k, static)
    comp_dtype = _KX2TORCH[out_kx]
    coef = prog.coef_tensor(device_args, comp_dtype, arr.t.device)
    needs_grad = torch.is_grad_enabled() and (arr.t.requires_grad or (coef is not None and coef.requires_grad))
    if fast_out is not None and not needs_grad and arr.t is array and not k and len(a) <= 1 and hasattr(torch._C, "_cuda_getCurrentRawStream"):
        # the call qualifies for the steady-state launcher if its plan is cacheable and the coefficient table is
        # either absent or the single device argument itself (sum(x * w): zero-copy identity layout)
        w = a[0] if a else None
        ok = (w is None and coef is None) or (isinstance(w, torch.Tensor) and coef is not None and w.is_cuda
                                              and w.is_contiguous() and coef.data_ptr() == w.data_ptr() and coef.dtype == w.dtype)
        key = _plan_key("map", func_map, shape, kernel_size, strides, border, relative, arr.kx_dtype, a, k, static) if ok else None
        if key is not None and key in _PLAN_CACHE and int(np.prod(full_shape)) > 0:
            fast_out.append(_FastMap(func_map, key, shape, arr.t, w, full_shape, comp_dtype, geom, prog))
    if needs_grad:
        if len(specs) != 1 or specs[0].kind != "affine" or out_kx != _lib.KX_F32 or arr.kx_dtype != _lib.KX_F32:
            raise UnsupportedStencilError("gradients are implemented for single-function float32 affine stencils")
        out = _AffineMapFn.apply(arr.t, coef, geom, prog, full_shape, comp_dtype)
    else:
        lib = _lib.load()
        out = torch.empty(full_shape, dtype=comp_dtype, device=arr.t.device)
        if out.numel():
            _lib.check(lib.kx_map(C.byref(geom), C.byref(prog.c), _ptr(arr.t), _ptr(coef), 0, _ptr(out), 1,
                                  _stream_ptr(arr.t.device)))
    return arr, out


class PreparedMap:
    """A kmap traced once for a fixed geometry, launched on caller-owned CUDA tensors
    (optionally into a caller-owned output).  This is what the slab decomposition
    (``dist.py``) and the pipelined host path use to issue many launches of sub-slabs
    without re-tracing; it is the same ``kx_map`` call as :func:`kernel_map`."""

    def __init__(self, func_map, shape, kernel_size, strides, border, relative, dtype=None, args=(), kwargs=None):
        dtype = torch.float32 if dtype is None else dtype
        self.shape = tuple(int(v) for v in shape)
        self.in_dtype = dtype
        plan = _map_plan(func_map, self.shape, tuple(kernel_size), tuple(strides),
                         tuple(tuple(b) for b in border), relative, _TORCH2KX[dtype], tuple(args), kwargs or {})
        self.specs, self.device_args, self.out_kx, self.prog, self.geom, self.out_shape = plan
        self.out_dtype = _KX2TORCH[self.out_kx]
        self._lib = _lib.load()

    def __call__(self, x, out=None, coef=None):
        if tuple(x.shape) != self.shape or x.dtype != self.in_dtype or not x.is_cuda or not x.is_contiguous():
            raise ValueError(f"PreparedMap expects a contiguous CUDA tensor {self.shape} {self.in_dtype}")
        if out is None:
            out = torch.empty(self.out_shape, dtype=self.out_dtype, device=x.device)
        elif tuple(out.shape) != tuple(self.out_shape) or out.dtype != self.out_dtype or not out.is_contiguous():
            raise ValueError(f"PreparedMap output must be contiguous {self.out_shape} {self.out_dtype}")
        if coef is None:
            coef = self.prog.coef_tensor(self.device_args, self.out_dtype, x.device)
        if out.numel():
            _lib.check(self._lib.kx_map(C.byref(self.geom), C.byref(self.prog.c), _ptr(x), _ptr(coef), 0, _ptr(out),
                                        1, _stream_ptr(x.device)))
        return out

    def _check_vjp(self, g):
        if len(self.specs) != 1 or self.specs[0].kind != "affine" or self.out_dtype != torch.float32 \
                or self.in_dtype != torch.float32:
            raise UnsupportedStencilError("gradients are implemented for single-function float32 affine stencils")
        if tuple(g.shape) != tuple(self.out_shape) or g.dtype != torch.float32 or not g.is_contiguous():
            raise ValueError(f"cotangent must be contiguous float32 {tuple(self.out_shape)}")

    def vjp_input(self, g, out=None, coef=None):
        """Cotangent of the input: the flipped-tap stencil of ``g`` (``kx_map_vjp_input``)."""
        self._check_vjp(g)
        if out is None:
            out = torch.empty(self.shape, dtype=torch.float32, device=g.device)
        if coef is None:
            coef = self.prog.coef_tensor(self.device_args, torch.float32, g.device)
        if out.numel():
            _lib.check(self._lib.kx_map_vjp_input(C.byref(self.geom), C.byref(self.prog.c), _ptr(g), _ptr(coef), _ptr(out),
                                                  _stream_ptr(g.device)))
        return out

    def vjp_coef(self, x, g):
        """Cotangent of the coefficient table ``[n_taps]`` (``kx_map_vjp_coef``: fused window correlation,
        deterministic two-stage reduction)."""
        self._check_vjp(g)
        scratch = self.__dict__.get("_wgrad_scratch")
        if scratch is None or scratch.device != x.device:   # per-block partials: sized once per plan, reused per call
            nbytes = int(self._lib.kx_vjp_coef_scratch_bytes(C.byref(self.geom), C.byref(self.prog.c)))
            scratch = self._wgrad_scratch = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=x.device)
        dcoef = torch.empty(self.prog.c.n_taps, dtype=torch.float32, device=x.device)
        if not g.numel():
            dcoef.zero_()
        if g.numel():
            _lib.check(self._lib.kx_map_vjp_coef(C.byref(self.geom), C.byref(self.prog.c), _ptr(x), _ptr(g), _ptr(scratch),
                                                 _ptr(dcoef), _stream_ptr(x.device)))
        return dcoef


def kernel_map(func_map: dict, shape, kernel_size, strides, border, relative: bool = False,
               map_kind: str = "vmap", map_kwargs: dict | None = None) -> Callable:
    """``kernex.kernel_map`` (``_src/map.py:61-124``).  ``map_kind`` selects a JAX
    transform in the reference (``vmap``/``lmap``/``pmap``, ``_src/utils.py:25-30``);
    all three give the same values, so it is validated and otherwise ignored."""
    if map_kind not in (None, "vmap", "lmap", "pmap"):
        raise KeyError(map_kind)
    if map_kwargs:
        raise UnsupportedStencilError(f"map_kwargs={map_kwargs!r} are JAX transform options with no CUDA meaning")
    shape, kernel_size, strides = tuple(shape), tuple(kernel_size), tuple(strides)
    border = tuple(tuple(b) for b in border)
    _validate(shape, kernel_size, strides, border)

    folded = _fold_leading_axes(func_map, shape, kernel_size, strides, border, relative)
    fold1d = _fold_1d(func_map, shape, kernel_size, strides, border, relative)

    fast: dict = {}      # (input dtype, has a weight argument) -> _FastMap, installed by the first general call

    def call(array, *a, **k):
        if type(array) is torch.Tensor and not k and len(a) <= 1:     # steady state, device-resident input
            ent = fast.get((array.dtype, len(a)))
            if ent is not None:
                out = ent(array, a[0] if a else None)
                if out is not None:
                    return out
        if folded is not None:
            return folded(array, *a, **k)
        if fold1d is not None and fold1d.applies(array, a, k):
            return fold1d(array, *a, **k)
        piped = _try_host_pipeline(func_map, shape, kernel_size, strides, border, relative, array, a, k)
        if piped is not None:
            return piped
        installed: list = []
        arr, out = _run_map(func_map, shape, kernel_size, strides, border, relative, array, a, k, static, installed)
        if installed and len(fast) < 8:
            fast[(arr.t.dtype, len(a))] = installed[0]
        return arr.give_back(out)

    static = _static_key("map", func_map, shape, kernel_size, strides, border, relative)
    return call


def _fold_leading_axes(func_map, shape, kernel_size, strides, border, relative):
    """Arrays of rank > KX_MAX_DIMS: leading axes with a 1-wide window (pure batch axes, e.g. (N, C, D, H, W)
    with kernel (1, 1, 3, 3, 3)) are merged into ONE batch axis on the host; the kernels fold that axis into
    their batch grid dimension.  Only single-function maps (region keys could constrain the merged axes)."""
    nd = len(shape)
    if nd <= _lib.KX_MAX_DIMS:
        return None
    n_fold = nd - _lib.KX_MAX_DIMS + 1
    pure = all(kernel_size[d] == 1 and strides[d] == 1 and tuple(border[d]) == (0, 0) for d in range(n_fold))
    if not pure or any(v != () for v in func_map.values()):
        raise UnsupportedStencilError(
            f"arrays of rank > {_lib.KX_MAX_DIMS} need leading axes with a 1-wide window (pure batch axes) and a single function")
    lead = int(np.prod(shape[:n_fold]))
    shape2 = (lead,) + tuple(shape[n_fold:])
    fmap2 = {}
    for f in func_map:
        def lifted(w, *a, _f=f, _n=n_fold, **k):
            return _f(w.reshape((1,) * _n + tuple(w.shape[1:])), *a, **k)
        fmap2[lifted] = ()
    inner = kernel_map(fmap2, shape2, (1,) + tuple(kernel_size[n_fold:]), (1,) + tuple(strides[n_fold:]),
                       ((0, 0),) + tuple(border[n_fold:]), relative)
    return _FoldedCall(inner, shape, shape2, n_fold)


def _fold_1d(func_map, shape, kernel_size, strides, border, relative):
    """Large 1-D arrays run as 2-D row problems plus a seam fix-up (``fold1d.py``): measured on B200 at 2^27 elements,
    15-tap FIR 0.53 -> 0.72 of the HBM roofline, moving average 0.64 -> 0.85, running max 0.65 -> 0.84
    (profiles/README.md, r2).  ``KX_NO_FOLD_1D=1`` keeps the one-row launch (cross-check)."""
    import os

    from .fold1d import plan_fold
    if os.environ.get("KX_NO_FOLD_1D") or len(shape) != 1 or len(func_map) != 1 or any(v != () for v in func_map.values()):
        return None
    (n,), (k,), (st,), ((lo, hi),) = shape, kernel_size, strides, border
    w = plan_fold(int(n), int(k), int(lo), int(hi), int(st))
    if w is None:
        return None
    (f,) = func_map

    def lifted(win, *a, _f=f, **kw):        # the (1, k) window of the folded array is the user's 1-D window
        return _f(win.reshape((k,)), *a, **kw)

    main_geom = ({lifted: ()}, (n // w, w), (1, k), (1, 1), ((0, 0), (lo, k - 1 - lo)), relative)
    main = kernel_map(*main_geom)
    seam = kernel_map({lifted: ()}, (n // w - 1, 2 * (k - 1)), (1, k), (1, 1), ((0, 0), (0, 0)), relative) if k > 1 else None
    return _Fold1dCall(main, seam, int(n), int(k), int(lo), int(hi), w, main_geom)


class _Fold1dCall:
    def __init__(self, main, seam, n, k, lo, hi, w, main_geom):
        self.main, self.seam, self.n, self.k, self.lo, self.hi, self.w = main, seam, n, k, lo, hi, w
        self.main_geom = main_geom

    def applies(self, array, a, kw):
        """Device-resident contiguous tensors outside autograd (the seam write-back is an in-place strided copy) and
        SCALAR window functions only: the fold flattens the (H, W) main result and overwrites (H-1, k-1) seam cells,
        which has no meaning for a vector-valued (GATHER / patch) function."""
        if not isinstance(array, torch.Tensor) or not array.is_cuda or not array.is_contiguous():
            return False
        if tuple(array.shape) != (self.n,) or array.dtype not in _TORCH2KX:
            return False
        if torch.is_grad_enabled() and (array.requires_grad or any(
                isinstance(v, torch.Tensor) and v.requires_grad for v in list(a) + list(kw.values()))):
            return False
        specs = _map_plan(*self.main_geom, _TORCH2KX[array.dtype], a, kw)[0]
        return all(s.kind != "gather" for s in specs)

    def __call__(self, array, *a, **kw):
        from .fold1d import folded_map_1d
        return folded_map_1d(array, self.k, self.lo, self.hi, self.w,
                             lambda t: self.main(t, *a, **kw), lambda t: self.seam(t, *a, **kw))


class _FoldedCall:
    def __init__(self, inner, shape, shape2, n_fold):
        self.inner, self.shape, self.shape2, self.n_fold = inner, tuple(shape), tuple(shape2), n_fold

    def __call__(self, array, *a, **k):
        if tuple(array.shape) != self.shape:
            raise ValueError(f"array shape {tuple(array.shape)} != declared shape {self.shape}")
        out = self.inner(array.reshape(self.shape2), *a, **k)
        return out.reshape(self.shape[:self.n_fold] + tuple(out.shape[1:]))


def _try_host_pipeline(func_map, shape, kernel_size, strides, border, relative, array, a, k):
    """Large pinned host arrays: overlap H2D / kernel / D2H over row slabs (hostpipe.py)."""
    if isinstance(array, torch.Tensor) or not isinstance(array, np.ndarray) or array.nbytes < (64 << 20):
        return None
    from . import hostpipe
    if tuple(array.shape) != tuple(shape) or not array.flags.c_contiguous:
        return None
    t_host = torch.from_numpy(array)
    if not hostpipe.eligible(array, t_host, strides, a, k, func_map):
        return None
    _require_cuda()
    out = hostpipe.pipelined_map(func_map, shape, kernel_size, strides, border, relative, t_host)
    return None if out is None else out.numpy()


def offset_kernel_map(func_map: dict, shape, kernel_size, strides, offset, relative: bool = False,
                      map_kind: str = "vmap", map_kwargs: dict | None = None) -> Callable:
    """``kernex.offset_kernel_map`` (``_src/map.py:127-156``): run the map with the
    border implied by ``offset`` and write the result into a copy of the input at
    ``[o0 : n-o1 : s]`` per axis."""
    shape, kernel_size, strides = tuple(shape), tuple(kernel_size), tuple(strides)
    offset = [(o, o) if isinstance(o, int) else tuple(o) for o in offset]
    border = offset_to_border(offset, kernel_size)
    inner = kernel_map(func_map, shape, kernel_size, strides, border, relative, map_kind, map_kwargs)

    def call(array, *a, **k):
        arr = _Arr(array)
        if tuple(arr.t.shape) != tuple(shape):
            raise ValueError(f"array shape {tuple(arr.t.shape)} != declared shape {tuple(shape)}")
        specs, device_args, out_kx, prog, geom, full_shape = _map_plan(
            func_map, shape, kernel_size, strides, border, relative, arr.kx_dtype, a, k)
        scalar = len(full_shape) == len(shape)
        needs_grad = torch.is_grad_enabled() and (arr.t.requires_grad or any(
            isinstance(v, torch.Tensor) and v.requires_grad for v in device_args.values()))
        if scalar and out_kx == arr.kx_dtype and not needs_grad and arr.t.numel():
            # one library call writes the whole result (fused pass, or copy + strided write-back)
            coef = prog.coef_tensor(device_args, _KX2TORCH[out_kx], arr.t.device)
            out = torch.empty_like(arr.t)
            off0 = (C.c_int32 * len(shape))(*[int(o[0]) for o in offset])
            _lib.check(_lib.load().kx_map_offset(C.byref(geom), C.byref(prog.c), _ptr(arr.t), _ptr(coef), off0,
                                                 _ptr(out), _stream_ptr(arr.t.device)))
            return arr.give_back(out)
        _, res = _run_map(func_map, shape, kernel_size, strides, border, relative, arr.t, a, k)
        return arr.give_back(_write_back(arr.t, res, shape, strides, offset, "map"))

    del inner
    return call


def _write_back(x, res, shape, strides, offset, what):
    if tuple(res.shape) > tuple(x.shape) or res.dim() != x.dim():  # lexicographic, as the reference (map.py:152)
        raise ValueError(f"{what} operation output must be scalar. Found result.shape={tuple(res.shape)}")
    out = x.clone()
    sl = tuple(slice(o0, n - o1, s) for n, s, (o0, o1) in zip(shape, strides, offset))
    if res.numel():
        out[sl] = res.to(out.dtype)  # strided device-to-device copy (cast as .at[].set does)
    return out


# --------------------------------------------------------------------------- #
# kscan                                                                         #
# --------------------------------------------------------------------------- #


def _scan_plan(func_map, shape, kernel_size, strides, border, relative, in_kx, a, k, static=None):
    """Trace + KxGeom / KxProgram of a scan, cached like the map plans (same key, tag "scan"): a time-stepping loop
    that calls one kscan per step re-traces nothing."""
    key = _plan_key("scan", func_map, shape, kernel_size, strides, border, relative, in_kx, a, k, static)
    if key is not None and key in _PLAN_CACHE:
        specs, _stale_args, prog, geom, out_shape = _PLAN_CACHE[key][1]
        return specs, _fresh_device_args(a, k), prog, geom, out_shape
    specs, device_args = _trace_all(func_map, kernel_size, relative, a, k)
    for s in specs:
        if s.kind == "gather" and s.out_shape != ():
            raise ValueError("scan operation output must be scalar.")
    groups = [tuple(v) for v in func_map.values()]
    out_shape = tuple(max(o, 0) for o in calculate_output_shape(shape, kernel_size, strides, border))
    prog = _Program(specs, groups, len(shape))
    # the carried state has the array's dtype; values are cast on write (scan.py:50)
    geom = _geom(shape, out_shape, kernel_size, strides, border, in_kx, in_kx)
    plan = (specs, device_args, prog, geom, out_shape)
    _cache_plan(key, func_map, plan)
    return plan


def _run_scan(func_map, shape, kernel_size, strides, border, relative, reverse, array, a, k, static=None):
    arr = _Arr(array)
    if tuple(arr.t.shape) != tuple(shape):
        raise ValueError(f"array shape {tuple(arr.t.shape)} != declared shape {tuple(shape)}")
    specs, device_args, prog, geom, out_shape = _scan_plan(
        func_map, shape, kernel_size, strides, border, relative, arr.kx_dtype, a, k, static)
    comp = torch.float32 if any(s.kind == "affine" and s.weak_float for s in specs) else arr.t.dtype
    coef = prog.coef_tensor(device_args, comp, arr.t.device)
    lib = _lib.load()
    out = torch.empty(out_shape, dtype=arr.t.dtype, device=arr.t.device)
    if out.numel():
        nbytes = int(lib.kx_scan_state_bytes(C.byref(geom)))
        state = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=arr.t.device)
        _lib.check(lib.kx_scan(C.byref(geom), C.byref(prog.c), _ptr(arr.t), _ptr(coef), _ptr(state), _ptr(out),
                               int(bool(reverse)), _stream_ptr(arr.t.device)))
    return arr, out


def _try_fused_offset_scan(func_map, shape, kernel_size, strides, border, offset, relative, arr, a, k, static):
    """sscan in ONE C call (``kx_scan_offset``: 3-D time-marching scans, a fused 2-D pass per plane that writes straight
    into the result).  Returns None when the library has no fused template for the call: the caller then composes
    ``kx_scan`` with the strided write-back."""
    if len(shape) != 3 or len(func_map) != 1 or any(s != 1 for s in strides) or tuple(arr.t.shape) != tuple(shape):
        return None
    specs, device_args, prog, geom, out_shape = _scan_plan(
        func_map, shape, kernel_size, strides, border, relative, arr.kx_dtype, a, k, static)
    if device_args or specs[0].kind != "affine" or not all(out_shape):
        return None
    if arr.kx_dtype == _lib.KX_I32 and specs[0].weak_float:
        return None
    lib = _lib.load()
    out = torch.empty_like(arr.t)
    off0 = (C.c_int32 * len(shape))(*[int(o[0]) for o in offset])
    rc = lib.kx_scan_offset(C.byref(geom), C.byref(prog.c), _ptr(arr.t), 0, off0, _ptr(out), _stream_ptr(arr.t.device))
    if rc == 2:          # KX_ERR_UNSUPPORTED: nothing was launched
        return None
    _lib.check(rc)
    return out


def _scan_kwargs(scan_kind, scan_kwargs):
    if scan_kind not in (None, "scan"):
        raise KeyError(scan_kind)
    kw = dict(scan_kwargs or {})
    reverse = bool(kw.pop("reverse", False))
    kw.pop("unroll", None)  # scheduling hint only
    if kw:
        raise UnsupportedStencilError(f"scan_kwargs={kw!r} are JAX transform options with no CUDA meaning")
    return reverse


def kernel_scan(func_map: dict, shape, kernel_size, strides, border, relative: bool = False,
                scan_kind: str = "scan", scan_kwargs: dict | None = None) -> Callable:
    """``kernex.kernel_scan`` (``_src/scan.py:65-115``)."""
    reverse = _scan_kwargs(scan_kind, scan_kwargs)
    shape, kernel_size, strides = tuple(shape), tuple(kernel_size), tuple(strides)
    border = tuple(tuple(b) for b in border)
    _validate(shape, kernel_size, strides, border)

    def call(array, *a, **k):
        arr, out = _run_scan(func_map, shape, kernel_size, strides, border, relative, reverse, array, a, k, static)
        return arr.give_back(out)

    static = _static_key("scan", func_map, shape, kernel_size, strides, border, relative)
    return call


def offset_kernel_scan(func_map: dict, shape, kernel_size, strides, offset, relative: bool = False,
                       scan_kind: str = "scan", scan_kwargs: dict | None = None) -> Callable:
    """``kernex.offset_kernel_scan`` (``_src/scan.py:118-147``)."""
    reverse = _scan_kwargs(scan_kind, scan_kwargs)
    shape, kernel_size, strides = tuple(shape), tuple(kernel_size), tuple(strides)
    offset = [(o, o) if isinstance(o, int) else tuple(o) for o in offset]
    border = offset_to_border(offset, kernel_size)
    _validate(shape, kernel_size, strides, border)

    def call(array, *a, **k):
        arr0 = _Arr(array)      # uploaded once, whichever path runs
        fused = None if reverse else _try_fused_offset_scan(func_map, shape, kernel_size, strides, border, offset, relative, arr0, a, k, static)
        if fused is not None:
            return arr0.give_back(fused)
        arr, res = _run_scan(func_map, shape, kernel_size, strides, border, relative, reverse, arr0.t, a, k, static)
        return arr0.give_back(_write_back(arr.t, res, shape, strides, offset, "scan"))

    static = _static_key("scan", func_map, shape, kernel_size, strides, border, relative)
    return call
