"""ctypes binding of ``libkx_b200.so`` (the C ABI declared in ``include/kx_b200.h``).

There is deliberately no fallback: if the shared library is missing or cannot
be loaded every engine call raises :class:`EngineUnavailableError`.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

from .errors import EngineUnavailableError

KX_MAX_DIMS = 4
KX_MAX_TAPS = 1024
KX_MAX_FUNCS = 8
KX_MAX_KEYS = 32
KX_ABI_VERSION = 2

KX_F32, KX_I32 = 0, 1
KX_AFFINE, KX_REDUCE_MAX, KX_REDUCE_MIN, KX_GATHER = 0, 1, 2, 3
KX_KEY_ANY, KX_KEY_EQ, KX_KEY_RANGE, KX_KEY_RANGE_STEP = 0, 1, 2, 3

LIB_NAME = "libkx_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


class KxGeom(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32),
        ("in_dtype", C.c_int32),
        ("out_dtype", C.c_int32),
        ("reserved", C.c_int32),
        ("in_shape", C.c_int64 * KX_MAX_DIMS),
        ("out_shape", C.c_int64 * KX_MAX_DIMS),
        ("ksize", C.c_int32 * KX_MAX_DIMS),
        ("stride", C.c_int32 * KX_MAX_DIMS),
        ("lo", C.c_int32 * KX_MAX_DIMS),
        ("hi", C.c_int32 * KX_MAX_DIMS),
    ]


class KxKeyItem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("step", C.c_int32), ("a", C.c_int64), ("b", C.c_int64)]


class KxKey(C.Structure):
    _fields_ = [("n_items", C.c_int32), ("func", C.c_int32), ("item", KxKeyItem * KX_MAX_DIMS)]


class KxFunc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("tap_begin", C.c_int32),
        ("tap_end", C.c_int32),
        ("coef_is_int", C.c_int32),
        ("bias", C.c_double),
        ("post_div", C.c_double),
    ]


class KxProgram(C.Structure):
    _fields_ = [
        ("n_funcs", C.c_int32),
        ("n_keys", C.c_int32),
        ("n_taps", C.c_int32),
        ("reserved", C.c_int32),
        ("func", KxFunc * KX_MAX_FUNCS),
        ("key", KxKey * KX_MAX_KEYS),
        ("tap_off", (C.c_int16 * KX_MAX_DIMS) * KX_MAX_TAPS),
        ("coef", C.c_double * KX_MAX_TAPS),
    ]


KX_PEER_HANDLE_BYTES = 64
KX_PEER_FLAG_BYTES = 256


class KxHaloSide(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("src", C.c_void_p), ("bytes", C.c_int64), ("peer_ready", C.c_void_p),
                ("my_ready", C.c_void_p), ("peer_done", C.c_void_p)]


class KxHaloPull(C.Structure):
    _fields_ = [("n_sides", C.c_int32), ("accumulate", C.c_int32), ("epoch", C.c_uint64), ("counter", C.c_void_p),
                ("side", KxHaloSide * 2)]


_SIGNATURES = {
    "kx_abi_version": (C.c_int, []),
    "kx_last_error": (C.c_char_p, []),
    "kx_launch_count": (C.c_uint64, []),
    "kx_last_kernel": (C.c_char_p, []),
    "kx_map": (C.c_int, [C.POINTER(KxGeom), C.POINTER(KxProgram), C.c_void_p, C.c_void_p, C.c_int64,
                         C.c_void_p, C.c_int64, C.c_void_p]),
    "kx_cast_i32_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "kx_map_offset": (C.c_int, [C.POINTER(KxGeom), C.POINTER(KxProgram), C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    "kx_map_vjp_input": (C.c_int, [C.POINTER(KxGeom), C.POINTER(KxProgram), C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "kx_vjp_coef_scratch_bytes": (C.c_int64, [C.POINTER(KxGeom), C.POINTER(KxProgram)]),
    "kx_map_vjp_coef": (C.c_int, [C.POINTER(KxGeom), C.POINTER(KxProgram), C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "kx_scan_state_bytes": (C.c_int64, [C.POINTER(KxGeom)]),
    "kx_scan": (C.c_int, [C.POINTER(KxGeom), C.POINTER(KxProgram), C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_int32, C.c_void_p]),
    "kx_scan_offset": (C.c_int, [C.POINTER(KxGeom), C.POINTER(KxProgram), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                 C.c_void_p, C.c_void_p]),
    "kx_scan_rowmarch_shard": (C.c_int, [C.POINTER(KxGeom), C.POINTER(KxProgram), C.c_int64, C.c_int64,
                                         C.c_void_p, C.c_void_p]),
    "kx_scan_rowmarch_rows_per_launch": (C.c_int, []),
    "kx_scan_rowmarch_shard_rows": (C.c_int, [C.POINTER(KxGeom), C.POINTER(KxProgram), C.c_int64, C.c_int64,
                                              C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "kx_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "kx_peer_free": (C.c_int, [C.c_void_p]),
    "kx_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kx_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "kx_peer_close": (C.c_int, [C.c_void_p]),
    "kx_halo_pull": (C.c_int, [C.POINTER(KxHaloPull), C.c_void_p]),
    "kx_flag_wait": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load the shared library once; raise loudly when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise EngineUnavailableError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"(or `make -C kernex_b200/csrc`). kernex_b200 has no CPU / PyTorch fallback.")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:
            raise EngineUnavailableError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = ABI mismatch: fail loudly
            fn.restype, fn.argtypes = res, argtypes
        if lib.kx_abi_version() != KX_ABI_VERSION:
            raise EngineUnavailableError(
                f"{LIB_PATH} has ABI {lib.kx_abi_version()}, this package expects {KX_ABI_VERSION}")
        _lib = lib
    return _lib


def check(status: int):
    if status != 0:
        msg = load().kx_last_error()
        msg = msg.decode() if msg else "unknown error"
        if status == 2:
            from .errors import UnsupportedStencilError
            raise UnsupportedStencilError(msg)
        if status == 1:
            raise ValueError(msg)
        raise RuntimeError(f"kx_b200 error {status}: {msg}")


def launch_count() -> int:
    return int(load().kx_launch_count())


def last_kernel() -> str:
    s = load().kx_last_kernel()
    return s.decode() if s else ""
