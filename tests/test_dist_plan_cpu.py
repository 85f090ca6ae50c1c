"""CPU checks of the column-shard plan used by the multi-GPU kscan path."""
from __future__ import annotations

import pytest

from kernex_b200.dist import column_shard_plan


@pytest.mark.parametrize("nx,world,nt,rl,rr", [(1 << 21, 8, 4096, 1, 0), (6000, 3, 96, 1, 0), (1000, 4, 50, 2, 1),
                                                (17, 8, 5, 1, 1), (4096, 2, 4096, 1, 1)])
def test_column_shard_plan_covers_the_grid_with_full_cones(nx, world, nt, rl, rr):
    covered = 0
    for rank in range(world):
        base, width, lo, hi = column_shard_plan(nx, world, rank, nt, rl, rr)
        assert 0 <= base and base + width <= nx and base % 4 == 0
        c0, c1 = base + lo, base + hi
        assert c0 == covered            # shards tile the grid in order, no gaps
        covered = c1
        # the dependency cone of every kept column lies inside the held range or the global pad
        assert base <= max(0, c0 - nt * rl)
        assert base + width >= min(nx, c1 + nt * rr)
    assert covered == nx


@pytest.mark.parametrize("world,rows,k0,border0", [(2, 8, 3, (0, 0)), (3, 6, 3, (1, 1)), (4, 5, 2, (0, 1)), (3, 7, 4, (1, 2)),
                                                   (8, 9, 5, (2, 2)), (2, 4, 1, (0, 0))])
def test_halo_pull_plan_fills_the_halos_like_the_two_sided_exchange(world, rows, k0, border0):
    """The one-sided (peer-memory) plan must deliver exactly the rows the isend / irecv exchange delivers: after
    applying it, every rank's ext buffer equals the matching slice of the global array."""
    import numpy as np

    from kernex_b200.dist import SlabLayout, halo_pull_plan

    glob = np.arange(world * rows * 3, dtype=np.float32).reshape(world * rows, 3)
    layouts = [SlabLayout(r, world, rows, k0, border0) for r in range(world)]
    exts = []
    for r, L in enumerate(layouts):
        ext = np.full((L.ext_rows, 3), -1.0, dtype=np.float32)
        ext[L.own] = glob[r * rows:(r + 1) * rows]
        exts.append(ext)
    for r, L in enumerate(layouts):
        for mine, peer, theirs in halo_pull_plan(r, world, rows, k0, border0):
            assert abs(peer - r) == 1
            exts[r][mine] = exts[peer][theirs]
    for r, L in enumerate(layouts):
        g0 = r * rows - L.halo_top
        np.testing.assert_array_equal(exts[r], glob[g0:g0 + L.ext_rows])
