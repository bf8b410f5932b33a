"""Level-schedule builder of the product (host-only part of libwflowsbm) against the
reference's set_dd: golden vectors + the oracle, bit-exact network order and indices."""
import numpy as np

from oracle import oracle as O
from tests import common
from wflow_b200 import core, synth


def test_schedule_matches_reference_golden():
    g = np.load(common.GOLD + "/set_dd.npz")
    for n in sorted({k.split("/")[0] for k in g.files}):
        nodes, up, off = core.build_schedule(g[n + "/ldd"])
        assert np.array_equal(nodes, g[n + "/nodes"]), n
        assert np.array_equal(up, g[n + "/nodes_up"]), n
        assert np.array_equal(off, g[n + "/lev_off"]), n


def test_schedule_matches_oracle_on_larger_grids():
    for ldd in (synth.ldd_basin_tiles(256, 384, tile=128), synth.ldd_random_forest(60, 45, seed=4),
                synth.ldd_basin_tiles(200, 200, tile=64, mask_frac=0.3), synth.ldd_all_pits(7, 9)):
        net = O.Network(ldd)
        nodes, up, off = core.build_schedule(ldd)
        assert np.array_equal(nodes, net.flat_nodes)
        assert np.array_equal(up, net.flat_up)
        assert np.array_equal(off, net.lev_off)


def test_upstream_runs_are_contiguous():
    """The property the lateral kernel relies on: in network order the upstream neighbours of a
    cell are one contiguous run of the previous level, in the reference's neighbour order."""
    ldd = synth.ldd_basin_tiles(128, 128, tile=64)
    nodes, up, off = core.build_schedule(ldd)
    pos = {int(f): i for i, f in enumerate(nodes)}
    level_of = np.repeat(np.arange(off.size - 1), np.diff(off))
    for k in range(nodes.size):
        ups = [int(u) for u in up[k] if u >= 0]
        if not ups:
            continue
        p = [pos[u] for u in ups]
        assert p == list(range(p[0], p[0] + len(p)))
        assert all(level_of[q] == level_of[k] - 1 for q in p)


def test_empty_and_degenerate_inputs():
    nodes, up, off = core.build_schedule(np.zeros((5, 5), np.uint8))  # nothing valid
    assert nodes.size == 0 and list(off) == [0, 0]
    nodes, up, off = core.build_schedule(np.array([[5]], np.uint8))
    assert list(nodes) == [0] and list(off) == [0, 1]
    # a cycle is unreachable from any pit: left out, exactly as the reference does
    nodes, up, off = core.build_schedule(np.array([[6, 4, 5]], np.uint8))
    assert list(nodes) == [2]


def test_partition_keeps_basins_whole_and_upstream_runs_in_reference_order():
    """Storage layout (wf_schedule_partition): every group holds whole basins, level-major; the upstream
    neighbours of every cell stay one contiguous run in set_dd's neighbour order (reference goldens of
    nodes_up: wflow/wflow_funcs.py:43-95)."""
    from wflow_b200 import core, synth
    for ldd in (synth.ldd_basin_tiles(64, 96, tile=32, seed=3), synth.ldd_random_forest(40, 56, seed=11)):
        sched = core.Schedule(ldd)
        nodes, up, off = core.build_schedule(ldd)
        refup = {int(nodes[k]): [int(x) for x in up[k] if x >= 0] for k in range(len(nodes))}
        # basin (pit) of every cell, from the reference-form schedule
        pit_of = {}
        for k in range(len(nodes) - 1, -1, -1):
            c = int(nodes[k])
            pit_of.setdefault(c, c)
            for u in refup[c]:
                pit_of[u] = pit_of[c]
        for G in (1, 2, 3, 5, 1000):
            order, goff, ups, nup = sched.partition(G)
            assert sorted(order.tolist()) == sorted(nodes.tolist())
            assert goff[0] == 0 and goff[-1] == len(order) and np.all(np.diff(goff) >= 0)
            group_of = np.repeat(np.arange(len(goff) - 1), np.diff(goff))
            pits = {}
            for k, c in enumerate(order.tolist()):
                pits.setdefault(pit_of[c], set()).add(int(group_of[k]))
                assert [int(order[ups[k] + j]) for j in range(nup[k])] == refup[c]
                assert all(ups[k] + j < k for j in range(nup[k]))       # upstream cells come earlier (level-major)
            assert all(len(g) == 1 for g in pits.values())                # a basin never straddles groups
