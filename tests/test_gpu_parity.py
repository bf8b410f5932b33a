"""CUDA path vs CPU oracle on the same seeded inputs (-m gpu).

Tolerance (BASELINE.json north_star): rel 1e-5 on float32 fluxes and states.  Each quantity
gets an absolute floor of 1e-5 x its typical magnitude (a flux that is ~0 has no meaningful
relative error).
"""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu

RTOL = 1e-5
# absolute floors per quantity [unit of the field]
ATOL = {"SubsurfaceFlow": 1e3, "CellInFlow": 1e3, "default": 2e-5, "SurfaceWatbal": 5e-4, "InterceptionWatBal": 1e-5}

CASES = [
    dict(kind="forest", layer_thickness=(100, 300, 800)),
    dict(kind="tiles", layer_thickness=(100, 300, 800)),
    dict(kind="tiles", layer_thickness=(), tm=1),
    dict(kind="forest", layer_thickness=(50,), ust=1, sir=1),
    dict(kind="tiles", layer_thickness=(100, 300, 800), dt=3600, itL=3, itR=4),
    dict(kind="forest", layer_thickness=(100, 300), snow=0, itL=2, itR=2, dt=21600),
    dict(kind="pits", layer_thickness=(100, 300, 800)),
]


def run_case(nrow, ncol, nsteps, kw, check_every=1):
    from wflow_b200 import synth
    orc, mod, ldd, river = common.make_pair(nrow, ncol, **kw)
    mask = common.network_mask(orc)
    forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=kw.get("dt", 86400))
    names = common.state_names(orc.maxLayers, kw.get("snow", 1)) + common.DIAGS
    orc.want(*common.DIAGS)
    mod.request(*common.DIAGS)
    worst = {}
    for t in range(nsteps):
        P, PET, T = forc.step(t)
        for m in (orc, mod):
            m.set("P", P); m.set("PET", PET); m.set("TEMP", T)
        orc.step()
        mod.step()
        if (t + 1) % check_every and t != nsteps - 1:
            continue
        for k in names:
            a = mod.get(k)
            b = orc[k]
            e = common.compare(a, b, mask, RTOL, ATOL.get(k, ATOL["default"]))
            worst[k] = max(worst.get(k, 0.0), e)
    return worst


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
def test_states_and_fluxes_match_oracle(kw):
    worst = run_case(32, 32, 30, kw)
    bad = {k: v for k, v in worst.items() if v > 1.0}
    assert not bad, "fields outside tolerance (error / tolerance): %s" % bad


def test_network_matches_oracle_set_dd():
    from oracle import oracle as O
    from wflow_b200 import synth
    from wflow_b200.core import SbmModel
    for kind, n in (("forest", 40), ("tiles", 64)):
        ldd, river, _, _ = synth.make_case(n, n, kind=kind, tile=16, river_area=6)
        mod = SbmModel(ldd, river)
        for riv in (False, True):
            net = O.Network(np.where(river != 0, ldd, 0) if riv else ldd)
            nodes, up, off = mod.network(river=riv)
            assert np.array_equal(nodes, net.flat_nodes)
            assert np.array_equal(up, net.flat_up)
            assert np.array_equal(off, net.lev_off)
