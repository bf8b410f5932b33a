"""Kernel parity on many more cell-steps than the small cases see (~3 M per configuration), across the seasons of
the synthetic forcing and the whole range of the synthetic parameter classes (f*D from 0.1 to ~200): every
step starts from the oracle's float32 states, every state of every cell is compared at rel 1e-5.

The reference's algorithm is discontinuous at the float32-ulp level (exact-zero early-outs, its = ceil(..), strict
layer membership: tests/test_reference_sensitivity.py), so among millions of cell-steps a few sit on such an edge;
the bound on them is what these tests document (at most 1 in 100 000)."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu

CONFIGS = [
    dict(kind="tiles", tile=128, layer_thickness=(100, 300, 800), dt=86400, itR=2, river_area=60),
    dict(kind="tiles", tile=128, layer_thickness=(100, 300, 800), dt=21600, itL=2, itR=3, river_area=60),
    dict(kind="forest", layer_thickness=(100, 300), dt=86400, river_area=20, sir=1),
]


@pytest.mark.parametrize("kw", CONFIGS, ids=["daily", "6h", "forest"])
def test_many_cell_steps_from_identical_states(kw):
    from wflow_b200 import synth
    n, nsteps = 512, 12
    orc, mod, ldd, river = common.make_pair(n, n, init="random", block=32, **kw)
    mask = common.network_mask(orc)
    forc = synth.Forcing(n, n, seed=9, timestepsecs=kw["dt"], rain_mean=10.0)
    states = common.state_names(orc.maxLayers)
    per_year = int(round(365 * 86400 / kw["dt"]))
    outside = total = 0
    worst = {}
    for s in range(nsteps):
        t = (s * per_year) // nsteps + s  # walk through the year: frost, melt, dry spells
        P, PET, T = forc.step(t)
        for m in (orc, mod):
            m.set("P", P); m.set("PET", PET); m.set("TEMP", T)
        for k in states:
            mod.set(k, orc[k])
        orc.step()
        mod.step()
        bad = np.zeros(mask.shape, bool)
        for k in states:
            atol = 1e3 if k == "SubsurfaceFlow" else 2e-5
            a, b = mod.get(k).astype(np.float64), orc[k].astype(np.float64)
            e = mask & (np.abs(a - b) > atol + 1e-5 * np.abs(b))
            if e.any():
                worst[k] = worst.get(k, 0) + int(e.sum())
            bad |= e
        outside += int(bad.sum())
        total += int(mask.sum())
    print("%d of %d cell-steps outside rel 1e-5 %s" % (outside, total, worst))
    assert outside <= total * 1e-5, (outside, total, worst)
    mod.close()
