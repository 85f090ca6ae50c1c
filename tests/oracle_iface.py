"""Minimal mesh front-end (region assignment -> func_map) over the NumPy oracle."""
from __future__ import annotations

from oracle import kx_oracle as ko


class OracleIface:
    """Minimal mesh front-end over the oracle (region assignment -> func_map)."""

    def __init__(self, inplace, kernel_size, padding=0, relative=False, named_axis=None, strides=1):
        self.inplace, self.kernel_size, self.relative = inplace, kernel_size, relative
        self.padding, self.named_axis, self.strides = padding, named_axis, strides
        self.container = {}

    def __setitem__(self, index, func):
        self.container[func] = [*self.container.get(func, []), index]

    def __call__(self, x):
        from kernex_b200.named_axis import named_axis_wrapper
        border = ko.resolve_padding(self.padding, self.kernel_size)
        fmap = {}
        for fn, idx in self.container.items():
            keys = [ko.normalize_index(i, x.shape) for i in idx]
            if self.named_axis is not None:
                fn = named_axis_wrapper(self.kernel_size, self.named_axis)(fn)
            fmap[fn] = keys
        op = ko.kernel_scan if self.inplace else ko.kernel_map
        strides = (self.strides,) * x.ndim if isinstance(self.strides, int) else self.strides
        return op(fmap, x.shape, self.kernel_size, strides, border, self.relative)(x)
