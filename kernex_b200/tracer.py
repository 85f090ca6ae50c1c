"""Symbolic tracer + classifier for per-window functions.

The reference hands the user function a gathered, optionally rolled patch and lets
XLA compile whatever it does (``kernex/_src/map.py:40-58``).  This engine has a
fixed set of CUDA templates instead, so the function is evaluated ONCE on a
symbolic window and the result is classified (BASELINE.json north_star; the
``jax.make_jaxpr`` front-end it names needs JAX, which this image lacks):

* ``AFFINE``  ``bias + sum_t coef_t * x[tap_t]`` -- explicit tap formulas
  (README.md:130-133), ``sum(x*w)`` with host or device weights (README.md:109-111),
  FDM updates with Python-float coefficients (README.md:251-252,
  tests/test_interface.py:36-45), constants (``bc``/``ic``), ``sum``/``mean``.
* ``REDUCE``  ``max`` / ``min`` over window elements.
* ``GATHER``  an array of plain window elements (``lambda x: x``, README.md:157-171).

Anything else raises :class:`UnsupportedStencilError` naming the operation.

The window is an object array of :class:`Affine` scalars laid out in *patch*
coordinates and, for ``relative=True``, rotated exactly like ``roll_view``
(``kernex/_src/utils.py:163-181``), so NumPy's own indexing rules (negative
indices, slices, tuples) give the reference's semantics for free.
"""

from __future__ import annotations

import numbers
from dataclasses import dataclass, field
from typing import Any, Callable, Sequence

import numpy as np

from .errors import UnsupportedStencilError

try:  # torch tensors may be passed as extra args (device weights)
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_number(v):
    return isinstance(v, (numbers.Real, np.generic)) and not isinstance(v, (bool, np.bool_))


def _is_int(v):
    return isinstance(v, (numbers.Integral, np.integer)) and not isinstance(v, (bool, np.bool_))


def _py(v):
    """numpy scalar -> exact Python number (ints stay ints)."""
    if isinstance(v, np.generic):
        return v.item()
    return v


# --------------------------------------------------------------------------- #
# coefficients: host constant + linear combination of device-weight elements   #
# --------------------------------------------------------------------------- #


class Coef:
    """``const + sum_r scale_r * W[arg_r][flat_r]`` where ``W`` are tensor args that
    live on the device (their values are not read at trace time)."""

    __slots__ = ("const", "w")

    def __init__(self, const=0, w=None):
        self.const = const
        self.w = w or {}

    @property
    def is_const(self):
        return not self.w

    def is_zero(self):
        return self.is_const and self.const == 0

    def add(self, other: "Coef", sign=1):
        w = dict(self.w)
        for k, s in other.w.items():
            w[k] = w.get(k, 0) + sign * s
        return Coef(self.const + sign * other.const, {k: s for k, s in w.items() if s != 0})

    def scale(self, s):
        if s == 0:
            return Coef(0 * self.const if isinstance(self.const, float) or isinstance(s, float) else 0)
        return Coef(self.const * s, {k: v * s for k, v in self.w.items()})

    def mul(self, other: "Coef"):
        if other.is_const:
            return self.scale(other.const)
        if self.is_const:
            return other.scale(self.const)
        raise UnsupportedStencilError("product of two device-resident arguments is not an affine stencil")

    def is_integral(self):
        return self.is_const and isinstance(self.const, int)


class Affine:
    """Scalar symbolic value ``bias + sum_t coef_t * x[tap_t]`` (taps in patch coords)."""

    __slots__ = ("terms", "bias", "post_div")
    __array_priority__ = 1000

    def __init__(self, terms=None, bias=None, post_div=None):
        self.terms: dict[tuple, Coef] = terms or {}
        self.bias: Coef = bias if bias is not None else Coef(0)
        # ``value / d`` kept as a true division while it is the last operation (window
        # mean: sum then divide, like jnp.mean); folded into the coefficients otherwise
        self.post_div = post_div

    def _folded(self):
        if self.post_div is None:
            return self
        c = Coef(1.0 / self.post_div)
        return Affine({k: v.mul(c) for k, v in self.terms.items()}, self.bias.mul(c))

    # -- construction helpers -- #
    @staticmethod
    def tap(offset):
        return Affine({tuple(offset): Coef(1)}, Coef(0))

    @staticmethod
    def const(v):
        return Affine({}, Coef(_py(v)))

    @staticmethod
    def lift(v):
        if isinstance(v, Affine):
            return v
        if _is_number(v):
            return Affine.const(v)
        if isinstance(v, np.ndarray) and v.ndim == 0:
            return Affine.lift(v[()])
        if isinstance(v, SymArray) and v.data.ndim == 0:
            return v.data[()]
        raise UnsupportedStencilError(f"cannot combine a window value with {type(v).__name__}")

    @property
    def is_constant(self):
        return not self.terms

    def is_single_tap(self):
        if len(self.terms) != 1 or not self.bias.is_zero() or self.post_div is not None:
            return False
        (c,) = self.terms.values()
        return c.is_const and c.const == 1

    # -- arithmetic -- #
    def _add(self, other, sign):
        other = Affine.lift(other)._folded()
        self = self._folded()
        terms = dict(self.terms)
        for k, c in other.terms.items():
            terms[k] = terms[k].add(c, sign) if k in terms else (c if sign == 1 else c.scale(-1))
        return Affine(terms, self.bias.add(other.bias, sign))

    def __add__(self, o):
        return self._add(o, 1)

    __radd__ = __add__

    def __sub__(self, o):
        return self._add(o, -1)

    def __rsub__(self, o):
        return Affine.lift(o)._add(self, -1)

    def __neg__(self):
        return self._scale(Coef(-1))

    def __pos__(self):
        return self

    def _scale(self, c: Coef):
        self = self._folded()
        return Affine({k: v.mul(c) for k, v in self.terms.items()}, self.bias.mul(c))

    def __mul__(self, o):
        o = Affine.lift(o)._folded()
        if o.is_constant:
            return self._scale(o.bias)
        if self.is_constant:
            return o._scale(self.bias)
        raise UnsupportedStencilError("product of two window elements (x[a]*x[b]) is not an affine stencil")

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Affine.lift(o)
        if not (o.is_constant and o.bias.is_const):
            raise UnsupportedStencilError("division by a window element / device argument is not an affine stencil")
        d = o.bias.const
        if d == 0:
            raise ZeroDivisionError("window function divides by zero")
        if self.post_div is None and self.terms:
            return Affine(dict(self.terms), self.bias, post_div=float(d))
        return self._scale(Coef(1.0 / d))

    def __rtruediv__(self, o):
        raise UnsupportedStencilError("division by a window element is not an affine stencil")

    def _unsupported(name):  # noqa: N805
        def op(self, *a, **k):
            raise UnsupportedStencilError(
                f"operation '{name}' on window elements is outside the affine / reduce / gather classes")
        return op

    __pow__ = _unsupported("**")
    __rpow__ = _unsupported("**")
    __floordiv__ = _unsupported("//")
    __rfloordiv__ = _unsupported("//")
    __mod__ = _unsupported("%")
    __abs__ = _unsupported("abs")
    __lt__ = _unsupported("<")
    __le__ = _unsupported("<=")
    __gt__ = _unsupported(">")
    __ge__ = _unsupported(">=")
    __bool__ = _unsupported("bool() / data-dependent control flow")
    __float__ = _unsupported("float()")
    __int__ = _unsupported("int()")
    __matmul__ = _unsupported("@")

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        return _dispatch_ufunc(ufunc, method, inputs, kwargs)

    def __repr__(self):
        parts = [f"{c.const if c.is_const else '<w>'}*x{list(k)}" for k, c in self.terms.items()]
        return "Affine(" + " + ".join(parts + [str(self.bias.const if self.bias.is_const else "<w>")]) + ")"


@dataclass
class Reduce:
    """``max``/``min`` over plain window elements."""

    op: str
    taps: list

    def _no(self, *a, **k):
        raise UnsupportedStencilError(
            f"arithmetic on the result of a window {self.op} is not supported; return it directly")

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __neg__ = _no


# --------------------------------------------------------------------------- #
# symbolic array                                                                #
# --------------------------------------------------------------------------- #


def _obj(a):
    """Any array-like of numbers / Affines -> object ndarray of Affine."""
    if isinstance(a, SymArray):
        return a.data
    if isinstance(a, Affine):
        out = np.empty((), dtype=object)
        out[()] = a
        return out
    arr = np.asarray(a)
    if arr.dtype == object:
        return arr
    out = np.empty(arr.shape, dtype=object)
    flat = out.reshape(-1)
    for i, v in enumerate(arr.reshape(-1)):
        flat[i] = Affine.const(v)
    return out


_BINARY = {
    np.add: lambda a, b: a + b,
    np.subtract: lambda a, b: a - b,
    np.multiply: lambda a, b: a * b,
    np.true_divide: lambda a, b: a / b,
}
_UNARY = {np.negative: lambda a: -a, np.positive: lambda a: a}


def _dispatch_ufunc(ufunc, method, inputs, kwargs):
    if kwargs.get("out") is not None:
        raise UnsupportedStencilError("in-place NumPy ops on the window are not supported")
    if method == "__call__" and ufunc in _BINARY and len(inputs) == 2:
        a, b = np.broadcast_arrays(_obj(inputs[0]), _obj(inputs[1]))
        out = np.empty(a.shape, dtype=object)
        f = _BINARY[ufunc]
        for idx in np.ndindex(a.shape):
            out[idx] = f(a[idx], b[idx])
        return SymArray(out)._unwrap()
    if method == "__call__" and ufunc in _UNARY and len(inputs) == 1:
        a = _obj(inputs[0])
        out = np.empty(a.shape, dtype=object)
        for idx in np.ndindex(a.shape):
            out[idx] = _UNARY[ufunc](a[idx])
        return SymArray(out)._unwrap()
    if method == "reduce" and ufunc is np.add:
        return SymArray(_obj(inputs[0])).sum(axis=kwargs.get("axis", 0))
    if method == "reduce" and ufunc is np.maximum:
        return SymArray(_obj(inputs[0])).max(axis=kwargs.get("axis", 0))
    if method == "reduce" and ufunc is np.minimum:
        return SymArray(_obj(inputs[0])).min(axis=kwargs.get("axis", 0))
    raise UnsupportedStencilError(
        f"NumPy ufunc '{ufunc.__name__}' ({method}) on the window is outside the affine / reduce / gather classes")


_ARRAY_FUNCTIONS = {}


def _implements(*funcs):
    def deco(f):
        for fn in funcs:
            _ARRAY_FUNCTIONS[fn] = f
        return f
    return deco


class SymArray:
    """N-d array of :class:`Affine` values with the slice of NumPy a stencil needs."""

    __array_priority__ = 1000

    def __init__(self, data: np.ndarray):
        self.data = data

    def _unwrap(self):
        return self.data[()] if self.data.ndim == 0 else self

    # -- array protocol -- #
    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def size(self):
        return self.data.size

    @property
    def T(self):
        return SymArray(self.data.T)

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        for i in range(len(self.data)):
            yield self[i]

    def __getitem__(self, idx):
        if isinstance(idx, SymArray) or isinstance(idx, Affine):
            raise UnsupportedStencilError("data-dependent indexing of the window is not supported")
        out = self.data[idx]
        return SymArray(out) if isinstance(out, np.ndarray) else out

    def reshape(self, *shape):
        return SymArray(self.data.reshape(*shape))

    def ravel(self):
        return SymArray(self.data.reshape(-1))

    flatten = ravel

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        return _dispatch_ufunc(ufunc, method, inputs, kwargs)

    def __array_function__(self, func, types, args, kwargs):
        if func not in _ARRAY_FUNCTIONS:
            raise UnsupportedStencilError(
                f"NumPy function '{func.__name__}' on the window is outside the affine / reduce / gather classes")
        return _ARRAY_FUNCTIONS[func](*args, **kwargs)

    # -- arithmetic (elementwise) -- #
    def __add__(self, o):
        return _dispatch_ufunc(np.add, "__call__", (self, o), {})

    def __radd__(self, o):
        return _dispatch_ufunc(np.add, "__call__", (o, self), {})

    def __sub__(self, o):
        return _dispatch_ufunc(np.subtract, "__call__", (self, o), {})

    def __rsub__(self, o):
        return _dispatch_ufunc(np.subtract, "__call__", (o, self), {})

    def __mul__(self, o):
        return _dispatch_ufunc(np.multiply, "__call__", (self, o), {})

    def __rmul__(self, o):
        return _dispatch_ufunc(np.multiply, "__call__", (o, self), {})

    def __truediv__(self, o):
        return _dispatch_ufunc(np.true_divide, "__call__", (self, o), {})

    def __neg__(self):
        return _dispatch_ufunc(np.negative, "__call__", (self,), {})

    def _unsupported(name):  # noqa: N805
        def op(self, *a, **k):
            raise UnsupportedStencilError(
                f"operation '{name}' on the window is outside the affine / reduce / gather classes")
        return op

    __pow__ = _unsupported("**")
    __matmul__ = _unsupported("@")
    __rmatmul__ = _unsupported("@")
    __abs__ = _unsupported("abs")
    __lt__ = _unsupported("<")
    __gt__ = _unsupported(">")
    __le__ = _unsupported("<=")
    __ge__ = _unsupported(">=")
    __bool__ = _unsupported("bool()")

    # -- reductions -- #
    @staticmethod
    def _only_defaults(what, kw, ignore=()):
        """Keyword arguments of a NumPy reduction that would change the VALUE (``where``, ``initial``,
        ``keepdims``, ``weights`` ...) are refused unless they hold their neutral default; ``dtype`` / ``out``
        = None are harmless.  Swallowing them silently would compute something else than the user wrote."""
        for k, v in kw.items():
            if k in ignore or v is None or (k == "keepdims" and v is False) or (k == "where" and v is True) \
                    or v is getattr(np, "_NoValue", object()):
                continue
            raise UnsupportedStencilError(f"{what}(..., {k}={v!r}) on the window is not supported")

    def _reduce_axes(self, axis):
        if axis is None:
            return None
        axes = (axis,) if isinstance(axis, int) else tuple(axis)
        return tuple(a % self.data.ndim for a in axes)

    def sum(self, axis=None, **kw):
        self._only_defaults("sum", kw)
        axes = self._reduce_axes(axis)
        if axes is None or len(axes) == self.data.ndim:
            acc = Affine.const(0)
            for v in self.data.reshape(-1):  # row-major: the documented accumulation order
                acc = acc + v
            return acc
        moved = np.moveaxis(self.data, axes, tuple(range(len(axes))))
        lead = int(np.prod(moved.shape[: len(axes)]))
        moved = moved.reshape((lead,) + moved.shape[len(axes):])
        out = np.empty(moved.shape[1:], dtype=object)
        for idx in np.ndindex(out.shape):
            acc = Affine.const(0)
            for r in range(lead):
                acc = acc + moved[(r,) + idx]
            out[idx] = acc
        return SymArray(out)._unwrap()

    def mean(self, axis=None, **kw):
        self._only_defaults("mean", kw)
        axes = self._reduce_axes(axis)
        n = self.data.size if axes is None else int(np.prod([self.data.shape[a] for a in axes]))
        return self.sum(axis=axis) / float(n)

    def _minmax(self, op, axis):
        if axis is not None and self._reduce_axes(axis) != tuple(range(self.data.ndim)):
            raise UnsupportedStencilError(f"window {op} over a subset of axes is not supported")
        flat = list(self.data.reshape(-1))
        if not flat:
            raise ValueError(f"{op} of an empty window")
        if not all(isinstance(v, Affine) and v.is_single_tap() for v in flat):
            raise UnsupportedStencilError(f"{op} is only supported over plain window elements")
        return Reduce(op, [next(iter(v.terms)) for v in flat])

    def max(self, axis=None, **kw):
        self._only_defaults("max", kw)
        return self._minmax("max", axis)

    def min(self, axis=None, **kw):
        self._only_defaults("min", kw)
        return self._minmax("min", axis)


@_implements(np.sum)
def _np_sum(a, axis=None, **kw):
    return SymArray(_obj(a)).sum(axis=axis, **kw)


@_implements(np.mean)
def _np_mean(a, axis=None, **kw):
    return SymArray(_obj(a)).mean(axis=axis, **kw)


@_implements(np.average)
def _np_average(a, axis=None, weights=None, **kw):
    arr = SymArray(_obj(a))
    if weights is None:
        return arr.mean(axis=axis, **kw)
    SymArray._only_defaults("average", kw)
    w = np.asarray(weights)
    if w.dtype == object or w.shape != arr.data.shape:
        raise UnsupportedStencilError("np.average needs host weights of the window's shape")
    # sum(x * w) / sum(w), as NumPy computes it
    return _dispatch_ufunc(np.multiply, "__call__", (arr, w), {}).sum(axis=axis) / w.sum(axis=axis)


@_implements(np.max, np.amax)
def _np_max(a, axis=None, **kw):
    return SymArray(_obj(a)).max(axis=axis, **kw)


@_implements(np.min, np.amin)
def _np_min(a, axis=None, **kw):
    return SymArray(_obj(a)).min(axis=axis, **kw)


@_implements(np.reshape)
def _np_reshape(a, shape, **kw):
    SymArray._only_defaults("reshape", kw, ignore=("order",) if kw.get("order") in (None, "C") else ())
    return SymArray(_obj(a)).reshape(shape)


@_implements(np.ravel)
def _np_ravel(a, **kw):
    SymArray._only_defaults("ravel", kw, ignore=("order",) if kw.get("order") in (None, "C") else ())
    return SymArray(_obj(a)).ravel()


@_implements(np.shape)
def _np_shape(a):
    return a.shape


@_implements(np.ndim)
def _np_ndim(a):
    return a.ndim


@_implements(np.size)
def _np_size(a, axis=None):
    return a.size if axis is None else a.shape[axis]


@_implements(np.transpose)
def _np_transpose(a, axes=None):
    return SymArray(np.transpose(_obj(a), axes))


@_implements(np.stack)
def _np_stack(arrays, axis=0, **kw):
    SymArray._only_defaults("stack", kw)
    return SymArray(np.stack([_obj(a) for a in arrays], axis=axis))


@_implements(np.concatenate)
def _np_concatenate(arrays, axis=0, **kw):
    SymArray._only_defaults("concatenate", kw)
    return SymArray(np.concatenate([np.atleast_1d(_obj(a)) for a in arrays], axis=axis))


@_implements(np.dot, np.vdot, np.inner)
def _np_dot(a, b, **kw):
    SymArray._only_defaults("dot", kw)
    a, b = _obj(a), _obj(b)
    if a.ndim != 1 or b.ndim != 1:
        raise UnsupportedStencilError("np.dot on the window is only supported for 1-D operands")
    return _dispatch_ufunc(np.multiply, "__call__", (SymArray(a), SymArray(b)), {}).sum()


@_implements(np.tensordot)
def _np_tensordot(a, b, axes=2):
    a, b = _obj(a), _obj(b)
    if a.shape == b.shape and (axes == a.ndim or axes == 2 and a.ndim == 2):
        return _dispatch_ufunc(np.multiply, "__call__", (SymArray(a), SymArray(b)), {}).sum()
    raise UnsupportedStencilError("np.tensordot on the window is only supported as a full contraction")


# --------------------------------------------------------------------------- #
# tracing                                                                       #
# --------------------------------------------------------------------------- #


def make_window(kernel_size: Sequence[int], relative: bool) -> SymArray:
    data = np.empty(tuple(kernel_size), dtype=object)
    for idx in np.ndindex(*kernel_size):
        data[idx] = Affine.tap(idx)
    if relative:  # roll_view, kernex/_src/utils.py:163-181
        for ax, k in enumerate(kernel_size):
            if k:
                data = np.roll(data, -((k - 1) // 2), axis=ax)
    return SymArray(data)


class DeviceArg:
    """Marker for a tensor argument that stays on the device."""

    def __init__(self, index, tensor):
        self.index, self.tensor = index, tensor


def _symbolic_arg(index, tensor):
    shape = tuple(tensor.shape)
    data = np.empty(shape, dtype=object)
    flat = data.reshape(-1)
    for i in range(flat.size):
        flat[i] = Affine({}, Coef(0, {(index, i): 1}))
    return SymArray(data)._unwrap() if shape == () else SymArray(data)


@dataclass
class FuncSpec:
    """One classified window function."""

    kind: str                       # "affine" | "max" | "min" | "gather"
    taps: list                      # offsets from the window origin, tuple per tap
    coefs: list = field(default_factory=list)   # Coef per tap (affine)
    bias: Coef = field(default_factory=lambda: Coef(0))
    out_shape: tuple = ()           # trailing result axes (gather)
    weak_float: bool = False        # a Python/NumPy float took part (int input -> f32 result)
    post_div: float = 0.0           # result = (bias + sum) / post_div when non-zero

    @property
    def device_coefs(self):
        return any(not c.is_const for c in self.coefs) or not self.bias.is_const

    @property
    def integral(self):
        return all(c.is_integral() for c in self.coefs) and self.bias.is_integral() and not self.post_div


def trace_function(func: Callable, kernel_size, relative: bool, args=(), kwargs=None):
    """Evaluate ``func`` on the symbolic window and classify the result.

    Returns ``(FuncSpec, device_args)`` where ``device_args`` lists the tensor
    arguments referenced by symbolic coefficients (index -> tensor)."""
    kwargs = kwargs or {}
    device_args: dict[int, Any] = {}

    def prep(i, a):
        if torch is not None and isinstance(a, torch.Tensor):
            if a.is_cuda or a.requires_grad:
                device_args[i] = a
                return _symbolic_arg(i, a)
            return a.detach().numpy()
        return a

    sym_args = [prep(i, a) for i, a in enumerate(args)]
    sym_kwargs = {k: prep(("kw", k), v) for k, v in kwargs.items()}
    window = make_window(kernel_size, relative)
    try:
        result = func(window, *sym_args, **sym_kwargs)
    except UnsupportedStencilError:
        raise
    except TypeError as e:  # e.g. math.sqrt(Affine)
        raise UnsupportedStencilError(f"window function is outside the affine / reduce / gather classes: {e}") from e
    int_args = {i for i, t in device_args.items() if not (t.dtype.is_floating_point or t.dtype.is_complex)}
    return classify(result, kernel_size, int_args), device_args


def classify(result, kernel_size, int_device_args=frozenset()) -> FuncSpec:
    """``int_device_args``: indices of device-resident arguments of integer dtype (they keep an
    integer stencil integer, like host integer weights do)."""
    if isinstance(result, Reduce):
        return FuncSpec(kind=result.op, taps=list(result.taps))
    if isinstance(result, SymArray) or (isinstance(result, np.ndarray) and result.dtype == object):
        data = _obj(result)
        flat = list(data.reshape(-1))
        if all(isinstance(v, Affine) and v.is_single_tap() for v in flat):
            return FuncSpec(kind="gather", taps=[next(iter(v.terms)) for v in flat], out_shape=tuple(data.shape))
        if data.ndim == 0:
            return classify(flat[0], kernel_size, int_device_args)
        raise UnsupportedStencilError(
            "vector-valued window functions are only supported when every element is a plain window element "
            "(patch extraction)")
    if _is_number(result) or (isinstance(result, np.ndarray) and result.ndim == 0):
        result = Affine.const(result[()] if isinstance(result, np.ndarray) else result)
    if isinstance(result, Affine):
        taps, coefs = [], []
        for t, c in result.terms.items():
            if c.is_zero():
                continue  # 0*x[...] terms of README.md:130-133 (differs from XLA only for non-finite x)
            taps.append(t)
            coefs.append(c)
        weak_float = any(isinstance(c.const, float) for c in coefs + [result.bias]) or \
            any(isinstance(sc, float) or arg not in int_device_args
                for c in coefs + [result.bias] for (arg, _), sc in c.w.items()) or result.post_div is not None
        return FuncSpec(kind="affine", taps=taps, coefs=coefs, bias=result.bias, weak_float=weak_float,
                        post_div=float(result.post_div or 0.0))
    raise UnsupportedStencilError(
        f"window function returned {type(result).__name__}; expected a scalar / array built from window elements")
