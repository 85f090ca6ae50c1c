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
