"""The wide-grid ("fat", one 16K-column block per SM) variant of the row-marching kscan kernel:
oracle parity and bit-equality with the thin variant (reached through column shards)."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import kx_oracle_c as kc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def kex():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kernex_b200
    return kernex_b200


def _mesh(cls, nt, nx, kappa, odd):
    """Convection mesh whose region boundaries fall on columns that are NOT multiples of 4 when
    `odd` is set, so some 4-column groups mix functions (the exception-table path)."""
    q1, q2 = nx // 4 + (3 if odd else 0), nx // 2 + (1 if odd else 0)
    F = cls(kernel_size=(3, 3), padding=((1, 1), (1, 1)), named_axis={0: "n", 1: "i"}, relative=True)
    F[:, 0] = F[:, -1] = lambda u: 1
    F[:, :q1] = F[:, q2:] = lambda u: 1
    F[0:1, q1:q2] = lambda u: 2
    F[1:, 1:-1] = lambda u: u["i", "n-1"] - kappa * (u["i", "n-1"] - u["i-1", "n-1"])
    F[40:90, q2 + 5: q2 + 1234] = lambda u: 0.25 * u["i-1", "n-1"] + 0.5 * u["i", "n-1"] + 1.0
    regions = [(0, nt, 0, 1, (0, 0, 0), 1.0), (0, nt, nx - 1, nx, (0, 0, 0), 1.0),
               (0, nt, 0, q1, (0, 0, 0), 1.0), (0, nt, q2, nx, (0, 0, 0), 1.0),
               (0, 1, q1, q2, (0, 0, 0), 2.0), (1, nt, 1, nx - 1, (kappa, 1.0 - kappa, 0.0), 0.0),
               (40, min(90, nt), q2 + 5, q2 + 1234, (0.25, 0.5, 0.0), 1.0)]
    return F, regions


@pytest.mark.parametrize("odd", [False, True])
def test_fat_rowmarch_matches_oracle_and_thin_variant(kex, odd):
    from kernex_b200 import _lib
    from kernex_b200.dist import ColumnShardedScan
    nt, nx, kappa = 130, 1_700_000, 0.5          # >= 96 tiles of 16320 columns -> fat kernel
    import os
    F, regions = _mesh(kex.kscan, nt, nx, kappa, odd)
    os.environ["KX_RM_FAT"] = "1"        # opt-in: the fat variant is not the default
    try:
        whole = F(torch.ones((nt, nx), device="cuda"))
    finally:
        del os.environ["KX_RM_FAT"]
    assert _lib.last_kernel() == "scan_rowmarch_kernel"
    want = kc.rowmarch(nt, nx, regions)
    np.testing.assert_allclose(whole.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    # 6 column shards are ~283K columns wide -> thin kernel; stitched result must be bit-identical
    parts = []
    for rank in range(6):
        Fr, _ = _mesh(kex.kscan, nt, nx, kappa, odd)
        parts.append(ColumnShardedScan(Fr, nt, nx, rank=rank, world=6).run().clone())
    assert torch.equal(torch.cat(parts, dim=1), whole)
