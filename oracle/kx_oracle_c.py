"""ctypes loader for ``oracle/kx_oracle_c.c`` (OpenMP C restatement, CPU baseline).

TEST / BENCH INFRASTRUCTURE ONLY -- see the header of ``kx_oracle.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libkx_oracle.so")
MAX_DIMS = 4


class KxoGeom(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("in_shape", C.c_int64 * MAX_DIMS), ("out_shape", C.c_int64 * MAX_DIMS),
                ("stride", C.c_int32 * MAX_DIMS), ("lo", C.c_int32 * MAX_DIMS)]


class KxoRegion(C.Structure):
    _fields_ = [("r0", C.c_int64), ("r1", C.c_int64), ("c0", C.c_int64), ("c1", C.c_int64),
                ("a", C.c_float * 3), ("bias", C.c_float)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        lib = C.CDLL(SO)
        lib.kxo_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads():
    return int(load().kxo_max_threads())


def _geom(in_shape, kernel_size, strides, border):
    g = KxoGeom()
    g.ndim = len(in_shape)
    for d, (n, k, s, (lo, hi)) in enumerate(zip(in_shape, kernel_size, strides, border)):
        g.in_shape[d], g.stride[d], g.lo[d] = n, s, lo
        g.out_shape[d] = max((n + lo + hi - k) // s + 1, 0)
    return g, tuple(g.out_shape[d] for d in range(g.ndim))


def affine_map(x, kernel_size, strides, border, taps, coefs, bias=0, nthreads=None, out=None):
    x = np.ascontiguousarray(x)
    g, out_shape = _geom(x.shape, kernel_size, strides, border)
    if out is None:
        out = np.empty(out_shape, dtype=x.dtype)
    tap = np.ascontiguousarray(np.asarray(taps, dtype=np.int32).reshape(len(taps), x.ndim))
    coef = np.ascontiguousarray(np.asarray(coefs, dtype=x.dtype))
    fn = {np.dtype(np.float32): "kxo_affine_f32", np.dtype(np.int32): "kxo_affine_i32"}[x.dtype]
    ct = C.c_float if x.dtype == np.float32 else C.c_int32
    getattr(load(), fn)(C.byref(g), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_int(len(taps)),
                        tap.ctypes.data_as(C.c_void_p), coef.ctypes.data_as(C.c_void_p), ct(bias),
                        C.c_int(nthreads or max_threads()))
    return out


def patch_map(x, kernel_size, strides, border, taps=None, nthreads=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    g, out_shape = _geom(x.shape, kernel_size, strides, border)
    if taps is None:
        taps = list(np.ndindex(*kernel_size))
    out = np.empty(out_shape + tuple(kernel_size), dtype=np.float32)
    tap = np.ascontiguousarray(np.asarray(taps, dtype=np.int32).reshape(len(taps), x.ndim))
    load().kxo_patch_f32(C.byref(g), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_int(len(taps)),
                         tap.ctypes.data_as(C.c_void_p), C.c_int(nthreads or max_threads()))
    return out


def rowmarch(nt, nx, regions, nthreads=None, out=None):
    """regions: list of (r0, r1, c0, c1, (a_left, a_mid, a_right), bias), later entries win."""
    arr = (KxoRegion * len(regions))()
    for i, (r0, r1, c0, c1, a, bias) in enumerate(regions):
        arr[i].r0, arr[i].r1, arr[i].c0, arr[i].c1, arr[i].bias = r0, r1, c0, c1, bias
        for j in range(3):
            arr[i].a[j] = a[j]
    if out is None:
        out = np.zeros((nt, nx), dtype=np.float32)
    load().kxo_rowmarch_f32(C.c_int64(nt), C.c_int64(nx), C.c_int(len(regions)), arr, out.ctypes.data_as(C.c_void_p),
                            C.c_int(nthreads or max_threads()))
    return out


def map_views(x, kernel_size, strides, border, taps, coefs, bias=0.0, relative=False, nthreads=None, out=None,
              scratch=None):
    """The reference's own kmap algorithm (pad at the end + materialised int32 views + per-window gather
    + roll + function), 2-D float32.  ``taps`` are PATCH coordinates after the roll (non-negative:
    relative index -1 is k-1) and list every term the user wrote.  ``scratch`` = (padded, views)
    arrays to reuse between calls."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2 and max(kernel_size) <= 5
    g, out_shape = _geom(x.shape, kernel_size, strides, border)
    if out is None:
        out = np.empty(out_shape, dtype=np.float32)
    nwin = int(np.prod(out_shape))
    pads = [max(lo, 0) + max(hi, 0) for lo, hi in border]
    if scratch is None:
        scratch = (np.empty((x.shape[0] + pads[0]) * (x.shape[1] + pads[1]), np.float32),
                   np.empty(nwin * (kernel_size[0] + kernel_size[1]), np.int32))
    padded, views = scratch
    ks = (C.c_int32 * 2)(*kernel_size)
    hi = (C.c_int32 * 2)(*[b[1] for b in border])
    tap = np.ascontiguousarray(np.asarray(taps, dtype=np.int32).reshape(len(taps), 2))
    coef = np.ascontiguousarray(np.asarray(coefs, dtype=np.float32))
    load().kxo_map_views_f32(C.byref(g), ks, hi, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                             C.c_int(len(taps)), tap.ctypes.data_as(C.c_void_p), coef.ctypes.data_as(C.c_void_p),
                             C.c_float(bias), C.c_int(int(bool(relative))), padded.ctypes.data_as(C.c_void_p),
                             views.ctypes.data_as(C.c_void_p), C.c_int(nthreads or max_threads()))
    return out
