"""Pipelined host -> device -> host execution of a kmap on a large pinned NumPy array.

When the caller hands the engine a HOST array (the reference's usual entry: a NumPy /
host-resident array goes in, an array comes out), the whole call is bound by the PCIe
link, not by HBM: 2 x 1.07 GB for BASELINE config 2 against a 0.35 ms kernel.  PCIe is
full duplex, so the array is cut into row slabs along axis 0 (windows are independent)
and three streams run concurrently:

    copy-in  stream : H2D of slab c+1 (its rows plus the k0-1 halo rows it shares)
    compute  stream : kx_map on slab c
    copy-out stream : D2H of the result rows of slab c-1 into one pinned output array

Steady state costs max(H2D, D2H) instead of their sum.  Only pinned inputs take this path
(pageable memory cannot be copied asynchronously); everything else uses the plain path
in ``engine.py``.
"""

from __future__ import annotations

import os

import numpy as np
import torch

from .engine import PreparedMap, calculate_output_shape

MIN_BYTES = 64 << 20        # below this the pipeline's fixed costs are not worth it
SLAB_BYTES = int(os.environ.get("KX_HOSTPIPE_SLAB_MB", "64")) << 20   # target size of one input slab (measured on B200:
                            # 4 MiB 28.9 ms, 16 MiB 23.8 ms, 32 MiB 23.6 ms, 64 MiB 23.3 ms per 16384^2 step; ramped sizes are slower)
DEPTH = 3                   # slabs in flight


def eligible(host: np.ndarray, t_host: torch.Tensor, strides, a, k, func_map=None) -> bool:
    # Region keys are GLOBAL window-centre coordinates while every slab launch sees slab-local ones, so only
    # key-less single-function maps are cut into slabs; meshes take the plain (whole-array) path.
    if func_map is not None and (len(func_map) != 1 or any(tuple(v) != () for v in func_map.values())):
        return False
    return (host.nbytes >= MIN_BYTES and host.ndim >= 1 and t_host.is_pinned() and not a and not k
            and host.flags.c_contiguous and host.dtype in (np.float32, np.int32))


def pipelined_map(func_map, shape, kernel_size, strides, border, relative, t_host: torch.Tensor):
    """Returns a pinned torch tensor (host) with the kmap result."""
    dev = torch.device("cuda", torch.cuda.current_device())
    n0, k0, s0 = shape[0], kernel_size[0], strides[0]
    lo0, hi0 = border[0]
    out_shape = tuple(max(o, 0) for o in calculate_output_shape(shape, kernel_size, strides, border))
    row_bytes = int(np.prod(shape[1:], dtype=np.int64)) * 4
    rows_per_slab = max(int(SLAB_BYTES // max(row_bytes, 1)), k0 * s0, 1)
    out_rows_per_slab = max(rows_per_slab // s0, 1)

    # slab c produces output rows [j0, j1); it needs input rows [j0*s0 - lo0, (j1-1)*s0 - lo0 + k0)
    slabs = []
    j0 = 0
    while j0 < out_shape[0]:
        j1 = min(j0 + out_rows_per_slab, out_shape[0])
        a0, a1 = j0 * s0 - lo0, (j1 - 1) * s0 - lo0 + k0
        pad_lo, pad_hi = max(0, -a0), max(0, a1 - n0)
        slabs.append((j0, j1, max(a0, 0), min(a1, n0), (pad_lo, pad_hi)))
        j0 = j1

    if any(a1 <= a0 for (_, _, a0, a1, _) in slabs):
        return None  # a slab made of padding only: let the plain path handle this corner
    plans: dict = {}
    probe = None
    for (j0, j1, a0, a1, b0) in slabs:
        key = (a1 - a0, b0)
        if key not in plans:
            plans[key] = PreparedMap(func_map, (a1 - a0,) + tuple(shape[1:]), kernel_size, strides,
                                     (b0,) + tuple(border[1:]), relative, dtype=t_host.dtype)
            probe = plans[key]
    if probe is None:  # no output rows at all
        return torch.empty(out_shape, dtype=t_host.dtype, pin_memory=True)
    full_out_shape = (out_shape[0],) + tuple(probe.out_shape[1:])
    out_host = torch.empty(full_out_shape, dtype=probe.out_dtype, pin_memory=True)

    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    max_in = max(a1 - a0 for (_, _, a0, a1, _) in slabs)
    max_out = max(j1 - j0 for (j0, j1, _, _, _) in slabs)
    in_bufs = [torch.empty((max_in,) + tuple(shape[1:]), dtype=t_host.dtype, device=dev) for _ in range(DEPTH)]
    out_bufs = [torch.empty((max_out,) + tuple(probe.out_shape[1:]), dtype=probe.out_dtype, device=dev)
                for _ in range(DEPTH)]
    in_free = [None] * DEPTH    # event: kernel that read in_bufs[i] finished
    out_free = [None] * DEPTH   # event: D2H that read out_bufs[i] finished
    s_in.wait_stream(main)
    s_out.wait_stream(main)
    for c, (j0, j1, a0, a1, b0) in enumerate(slabs):
        i = c % DEPTH
        xin = in_bufs[i][: a1 - a0]
        yout = out_bufs[i][: j1 - j0]
        with torch.cuda.stream(s_in):
            if in_free[i] is not None:
                s_in.wait_event(in_free[i])
            xin.copy_(t_host[a0:a1], non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(s_in)
        main.wait_event(ev_in)
        if out_free[i] is not None:
            main.wait_event(out_free[i])
        plans[(a1 - a0, b0)](xin, out=yout)
        ev_k = torch.cuda.Event()
        ev_k.record(main)
        in_free[i] = ev_k
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_k)
            out_host[j0:j1].copy_(yout, non_blocking=True)
            ev_o = torch.cuda.Event()
            ev_o.record(s_out)
        out_free[i] = ev_o
    main.wait_stream(s_out)
    main.wait_stream(s_in)
    main.synchronize()
    return out_host
