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


@pytest.mark.parametrize("world", [2, 3, 4])
def test_column_exchange_plan_rejects_shards_narrower_than_their_ghosts(world):
    """The exchange form refreshes a ghost zone from ONE neighbour: a ghost wider than the neighbouring shard, or a
    rank left without columns after rounding the shard width up to 4, must raise instead of hanging / returning
    garbage; and every accepted plan pairs each irecv with an isend of the same size."""
    from kernex_b200.dist import column_exchange_plan
    t_rows = 64
    for nx in (world * 64 - 3, world * 64 + 1, world * 129 + 2, 4099, 1000 + world):
        for rl, rr in ((1, 0), (1, 1), (2, 2), (0, 1)):
            try:
                plans = [column_exchange_plan(nx, world, r, t_rows, rl, rr) for r in range(world)]
            except ValueError:
                per = -(-nx // world)
                per += (-per) % 4
                need = max(-(-t_rows * rl // 4) * 4, -(-t_rows * rr // 4) * 4)
                assert need > per or nx - (world - 1) * per <= 0
                continue
            covered = 0
            for r, (base, width, k0, k1, gl, gr) in enumerate(plans):
                assert base >= 0 and base + width <= nx and k1 > k0      # no negative base, nobody is empty
                assert base + k0 == covered
                covered += k1 - k0
                if r > 0:                                                 # my left ghost = the left neighbour's last gl own columns
                    pb, pw, pk0, pk1, _, pgr = plans[r - 1]
                    assert gl <= pk1 - pk0 and base == pb + pk1 - gl
                    assert pgr <= k1 - k0                                 # its right ghost = my first pgr own columns
                    assert pb + pk1 + pgr == base + k0 + pgr
            assert covered == nx
    with pytest.raises(ValueError):
        column_exchange_plan(world * 8, world, 1, t_rows, 1, 0)           # 8-column shards, 64-column ghosts
    with pytest.raises(ValueError):
        column_exchange_plan(9, 4, 0, 1, 1, 0)                            # shards rounded up to 4: rank 3 is empty
