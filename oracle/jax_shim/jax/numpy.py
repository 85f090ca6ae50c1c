"""``jax.numpy`` stand-in: NumPy with int32/float32 defaults and ``.at``."""
from __future__ import annotations

import numpy as _np

from ._core import ShimArray as ndarray  # noqa: F401
from ._core import as_shim as _as_shim

inf = _np.inf
pi = _np.pi
int32 = _np.int32
float32 = _np.float32
int = _np.int32  # noqa: A001  (old alias the reference's type hints mention)


def array(obj, dtype=None):
    return _as_shim(obj, dtype=dtype)


asarray = array


def arange(*a, dtype=None, **k):
    return _as_shim(_np.arange(*a, **k), dtype=dtype)


def _wrap(name):
    fn = getattr(_np, name)

    def call(*a, **k):
        out = fn(*a, **k)
        return _as_shim(out) if isinstance(out, (_np.ndarray, _np.generic)) else out

    call.__name__ = name
    return call


for _n in ("pad", "sum", "mean", "max", "min", "where", "all", "any", "ones", "zeros", "stack",
           "linspace", "exp", "square", "outer", "abs", "maximum", "minimum", "prod", "reshape",
           "concatenate", "ones_like", "zeros_like", "sqrt", "roll", "full"):
    globals()[_n] = _wrap(_n)


def ix_(*args):
    return _np.ix_(*args)
