"""Large 1-D arrays as 2-D row problems (default for scalar window functions on device tensors of >= 2^22 elements;
``KX_NO_FOLD_1D=1`` disables it).  The fold is CPU-tested against the oracle (tests/test_fold1d_cpu.py) and bit-identical
to the one-row launch on the GPU (tests/test_gpu_fastpaths.py::test_folded_1d_map_equals_the_one_row_launch).

A 1-D array is a one-row problem for the row-marching kernel: every thread loads, computes four outputs, stores and
retires, so nothing overlaps the load latency (15-tap FIR: 0.51 of the HBM roofline, ``profiles/README.md``).  Folding
the array into ``H`` rows of ``W`` elements lets the same kernel march down rows with several row loads in flight.
Only the ``k-1`` windows that straddle the end of a row see the wrong neighbours (zero padding instead of the next
row); they are recomputed from a small gathered ``(H-1, 2(k-1))`` seam array and written over the main result:

    main : rows (H, W), window (1, k), borders ((0,0), (lo, k-1-lo))      -> (H, W), flat index i = window start + lo
    seam : rows x[(h+1)W-(k-1) : (h+1)W+(k-1)], window (1, k), 'valid'     -> (H-1, k-1) = flat outputs
           hW + (W-k+lo+1) ... (h+1)W + lo - 1

The zero padding of the true array ends (first row's left side, last row's right side) is the reference's padding
(``kernex/_src/utils.py:43-53``); the flat result is cut to ``N + lo + hi - k + 1`` windows (``utils.py:109-112``).
"""
from __future__ import annotations

_WIDTHS = (16384, 8192, 4096, 2048, 1024)


def plan_fold(n: int, k: int, lo: int, hi: int, stride: int = 1, min_elems: int = 1 << 22):
    """Row width ``W`` for folding an ``n``-element array, or ``None`` when the fold does not apply."""
    if stride != 1 or k < 1 or n < min_elems or lo < 0 or hi < 0 or lo > k - 1 or lo + hi > k - 1:
        return None
    for w in _WIDTHS:
        if n % w == 0 and n // w >= 64 and w >= 16 * k:
            return w
    return None


def folded_map_1d(x, k: int, lo: int, hi: int, w: int, run_main, run_seam):
    """``x``: contiguous 1-D torch tensor.  ``run_main((H, W) tensor)`` applies the window ``(1, k)`` with borders
    ``((0, 0), (lo, k-1-lo))`` and returns ``(H, W)``; ``run_seam((H-1, 2(k-1)) tensor)`` applies it 'valid' and
    returns ``(H-1, k-1)``."""
    n = x.numel()
    h = n // w
    assert h * w == n and x.dim() == 1 and x.is_contiguous()
    main = run_main(x.view(h, w))
    assert main.dim() == 2, "the 1-D fold needs a scalar window function"
    out = main.reshape(-1)
    if h > 1 and k > 1:
        seams = x.as_strided((h - 1, 2 * (k - 1)), (w, 1), x.storage_offset() + w - (k - 1)).contiguous()
        fixed = run_seam(seams)
        out.as_strided((h - 1, k - 1), (w, 1), out.storage_offset() + w - k + lo + 1).copy_(fixed)
    return out[: n + lo + hi - k + 1]
