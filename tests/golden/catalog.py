"""Window functions and mesh scenarios shared by the fixture generator
(``make_golden.py``, where they run inside the real reference on the shim), the
oracle tests and the CUDA parity tests.  They only use operations all three
understand: integer / named indexing, ``+ - * /`` and ``np.sum/max/min/mean``.
"""

from __future__ import annotations

import numpy as np


def _first(ksize, relative):
    return tuple(-((k - 1) // 2) for k in ksize) if relative else (0,) * len(ksize)


def _last(ksize, relative):
    return tuple(k // 2 for k in ksize) if relative else tuple(k - 1 for k in ksize)


def _centre(ksize, relative):
    return (0,) * len(ksize) if relative else tuple((k - 1) // 2 for k in ksize)


SCALAR = ("sum", "max", "min", "mean", "centre10", "first_last", "wsum", "const", "affine_bias", "star")
VECTOR = ("patch",)


def is_scalar(name):
    return name in SCALAR


def for_kernel(ksize, relative):
    names = ["sum", "max", "min", "mean", "centre10", "first_last", "wsum", "const", "affine_bias", "patch"]
    if relative and all(k == 3 for k in ksize):
        names.append("star")
    return names


def build(name, ksize, relative, rng):
    """Return ``(fn, extra_args)``; ``extra_args`` are NumPy arrays / scalars
    passed positionally after the window."""
    nd = len(ksize)
    if name == "sum":
        return (lambda x: np.sum(x)), []
    if name == "max":
        return (lambda x: np.max(x)), []
    if name == "min":
        return (lambda x: np.min(x)), []
    if name == "mean":
        return (lambda x: np.mean(x)), []
    if name == "centre10":
        c = _centre(ksize, relative)
        return (lambda x: x[c] * 10), []
    if name == "first_last":
        a, b = _first(ksize, relative), _last(ksize, relative)
        return (lambda x: x[a] - 2 * x[b]), []
    if name == "wsum":
        w = rng.integers(-3, 4, size=ksize).astype(np.int32)
        return (lambda x, w: np.sum(x * w)), [w]
    if name == "const":
        return (lambda x: 7), []
    if name == "affine_bias":
        a = _first(ksize, relative)
        return (lambda x: 0.5 * x[a] + 3), []
    if name == "patch":
        return (lambda x: x), []
    if name == "star":
        def star(x):
            acc = -2 * nd * x[(0,) * nd]
            for d in range(nd):
                for s in (-1, 1):
                    idx = tuple(s if e == d else 0 for e in range(nd))
                    acc = acc + x[idx]
            return acc
        return star, []
    raise KeyError(name)


# --------------------------------------------------------------------------- #
# region-assignment scenarios: ``build(cls, rng) -> (F, x)``                   #
# --------------------------------------------------------------------------- #


def _mesh_1d(cls, rng):
    F = cls(kernel_size=(1,), relative=True, padding="same")
    F[0] = lambda x: x[0] * 10
    F[1] = lambda x: x[0] * -10
    F[2:] = lambda x: x[0] * 100
    return F, np.arange(1, 11, dtype=np.int32)


def _mesh_2d(cls, rng):
    F = cls(kernel_size=(3, 3), relative=True, padding="same")
    F[0, 0] = lambda x: x[0, 0] * 10
    F[1, 1:] = lambda x: x[0, 0] * -10
    F[2:] = lambda x: x[0, 0] * 100
    return F, np.arange(1, 26, dtype=np.int32).reshape(5, 5)


def _mesh_neighbors(cls, rng):
    F = cls(kernel_size=(3, 3), relative=True, padding="same")
    F[:, :] = lambda x: x[0, -1] + x[-1, 0] + 2 * x[0, 0]
    F[0, :] = lambda x: x[0, 0] - x[1, 1]
    F[-1] = F[:, -2:] = lambda x: 5
    F[1:-1, 0] = lambda x: x[1, 0] * 3
    return F, rng.integers(-5, 6, size=(6, 7)).astype(np.int32)


def _mesh_strided(cls, rng):
    # the step of a slice key tests ``centre % step == 0`` (NOT (centre-start) % step)
    F = cls(kernel_size=(3,), relative=True, padding="same")
    F[:] = lambda x: x[0]
    F[1::2] = lambda x: x[-1] + x[1]
    F[2:9:3] = lambda x: x[0] * 100
    return F, np.arange(1, 13, dtype=np.int32)


def _mesh_convection(cls, rng):
    # README linear convection, scaled down (README.md:231-266)
    tmax, xmax, nt, nx, c = 0.5, 2.0, 14, 19, 0.5
    dt, dx = tmax / (nt - 1), xmax / (nx - 1)
    F = cls(kernel_size=(3, 3), padding=((1, 1), (1, 1)), named_axis={0: "n", 1: "i"}, relative=True)

    def bc(u):
        return 1

    def ic1(u):
        return 1

    def ic2(u):
        return 2

    def linear_convection(u):
        return u["i", "n-1"] - (c * dt / dx) * (u["i", "n-1"] - u["i-1", "n-1"])

    F[:, 0] = F[:, -1] = bc
    F[:, : int((nx - 1) / 4) + 1] = F[:, int((nx - 1) / 2):] = ic1
    F[0:1, int((nx - 1) / 4) + 1: int((nx - 1) / 2)] = ic2
    F[1:, 1:-1] = linear_convection
    return F, np.ones((nt, nx), dtype=np.float32)


def _mesh_3d_partial(cls, rng):
    F = cls(kernel_size=(1, 3, 3), relative=True, padding="same")
    F[:] = lambda x: x[0, 0, 1] - x[0, 0, -1]
    F[0] = lambda x: -1
    F[1, 2:4] = lambda x: x[0, 1, 0] + x[0, -1, 0]
    return F, rng.integers(-4, 5, size=(3, 5, 6)).astype(np.int32)


MESH_SCENARIOS = {
    "mesh_1d": _mesh_1d,
    "mesh_2d": _mesh_2d,
    "mesh_neighbors": _mesh_neighbors,
    "mesh_strided": _mesh_strided,
    "mesh_convection": _mesh_convection,
    "mesh_3d_partial": _mesh_3d_partial,
}
