"""kernex_b200 -- B200-native (sm_100a) engine behind the kernex kmap/kscan API.

Public surface = the reference's (``kernex/__init__.py:15-28``): ``kmap``, ``kscan``,
``smap``, ``sscan`` and the four engine factories ``kernel_map``,
``offset_kernel_map``, ``kernel_scan``, ``offset_kernel_scan``.
"""

from .engine import clear_plan_cache, kernel_map, kernel_scan, offset_kernel_map, offset_kernel_scan
from .errors import EngineUnavailableError, UnsupportedStencilError
from .interface import kmap, kscan, smap, sscan
from .batching import vmap

__all__ = (
    "kmap", "kscan", "smap", "sscan", "vmap",
    "kernel_map", "kernel_scan", "offset_kernel_map", "offset_kernel_scan",
    "UnsupportedStencilError", "EngineUnavailableError", "clear_plan_cache",
)

__version__ = "0.1.0"
