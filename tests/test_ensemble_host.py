"""Host side of the ensemble mosaic (no GPU): tiling member grids side by side must not connect them, and every
member must keep exactly the network a stand-alone run of the grid has (incl. set_dd's flat-index-0 quirk,
wflow/wflow_funcs.py:88)."""
import numpy as np
import pytest


@pytest.mark.parametrize("kind", ["forest", "tiles"])
def test_mosaic_keeps_every_members_network(kind):
    from wflow_b200 import core, ensemble, synth
    nrow, ncol, M = 40, 48, 3
    ldd = synth.ldd_random_forest(nrow, ncol, seed=11) if kind == "forest" else synth.ldd_basin_tiles(nrow, ncol, tile=20, seed=7)
    ldd[5, ncol - 1] = 6  # a cell draining out of the grid through the east edge: never part of the network
    ldd[9, 0] = 4         # ... and one through the west edge
    tile = ensemble._member_ldd(ldd)
    alone = set(core.Schedule(ldd).partition(1)[0].tolist())
    assert set(core.Schedule(tile).partition(1)[0].tolist()) == alone
    mosaic = core.Schedule(np.tile(tile, (1, M)))
    order = mosaic.partition(1)[0]
    assert order.size == M * len(alone)
    r, c = order // (M * ncol), order % (M * ncol)
    for m in range(M):
        sel = (c // ncol) == m
        assert set((r[sel] * ncol + c[sel] % ncol).tolist()) == alone


def test_shard_members_covers_all_members_once():
    from wflow_b200 import parallel
    for n, w in ((256, 8), (10, 4), (3, 8)):
        got = [m for r in range(w) for m in parallel.shard_members(n, r, w)]
        assert got == list(range(n))
