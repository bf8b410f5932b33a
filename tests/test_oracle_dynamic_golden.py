"""The CPU oracle against trajectories of the reference's whole, unmodified ``WflowModel.dynamic()`` --
PCRaster half included, run through the float32 shim of oracle/pcrshim.py (tests/golden/dyn_*.npz, generator
oracle/gen_golden_dynamic.py).  This is the pin of the map-algebra phases of the oracle: interception (Gash and
modified Rutter), HBV snow, glacier, LAI-driven canopy parameters, evaporation split, river inflow glue, the
Courant estimate of the sub-steps.  Bit-exact on every state and output the oracle produces."""
import numpy as np
import pytest

from tests import common

# cancellation residuals of ~1e3 mm terms: the reference forms them in another association than the oracle
RESIDUALS = {"SoilWatbal": 1e-9, "watbal": 2e-4, "SurfaceWatbal": 0.0, "InterceptionWatBal": 0.0}
# float32 sums over a whole catchment (pcr.catchmenttotal): the order in which the cells of a catchment are added is
# PCRaster's own and not documented; the shim and the oracle both use a topological (Kahn) order
SUM_ORDER = {"RunoffCoeff": 1e-6, "QCatchmentMM": 1e-6}


@pytest.mark.parametrize("name", common.golden_dyn_cases(objects=False))
def test_dynamic_trajectory_golden(name):
    g, kw, orc, _ = common.load_dyn_case(name)
    mask = common.network_mask(orc)
    # (Cmax / CanopyGapFraction / EoverR: inputs of the oracle; the reference overwrites them from LAI, wflow_sbm.py:2587-2603)
    outs = [f for f in common.dyn_outputs(g) if f in orc.names and f not in ("Cmax", "CanopyGapFraction", "EoverR")]
    missing = [f for f in common.dyn_outputs(g) if f not in orc.names]
    orc.want(*[f for f in outs if f not in orc.fields])
    assert len(outs) >= 100 and not missing, missing
    for t in range(kw["steps"]):
        for f in ("P", "PET", "TEMP", "Inflow"):
            key = "forcing/%d/%s" % (t, f)
            if key in g.files:
                orc.set(f, g[key])
        if kw.get("courant"):
            itl, itr = orc.estimate_iterations("land"), orc.estimate_iterations("river")
            assert (itl, itr) == tuple(int(v) for v in g["it/%d" % t]), (name, t)
            orc.set_substeps(itl, itr)
        orc.step()
        for f in outs:
            a, b = orc[f][mask], g["out/%d/%s" % (t, f)][mask]
            if RESIDUALS.get(f, 0.0) > 0.0:
                assert np.allclose(a, b, rtol=0, atol=RESIDUALS[f]), (name, t, f)
            elif f in SUM_ORDER:
                assert np.allclose(a, b, rtol=SUM_ORDER[f], atol=0), (name, t, f)
            else:
                assert np.array_equal(a, b), (name, t, f, int((a != b).sum()), float(np.abs(a - b).max()))


def test_dyn_goldens_exercise_the_map_algebra_branches():
    """Guard against vacuous goldens: both Gash storm sizes, the Cmax <= 0 / EoverR = 0 / TTI = 0 branches, Rutter
    drainage, snowfall, melt and refreezing, glacier melt on both sides of the 10 mm threshold, dry LAI cells."""
    g = np.load(common.GOLD + "/dyn_gash_snow_daily.npz")
    st = {k[7:]: g[k] for k in g.files if k.startswith("static/")}
    assert (st["Cmax"] <= 0).any() and (st["EoverR"] == 0).any() and (st["TTI"] == 0).any()
    seen = dict(snowfall=False, melt=False, intercept=False, overestimate=False)
    for t in range(6):
        seen["snowfall"] |= bool((g["out/%d/SnowFall" % t] > 0).any())
        seen["melt"] |= bool((g["out/%d/SnowMelt" % t] > 0).any())
        seen["intercept"] |= bool((g["out/%d/Interception" % t] > 0).any())
        seen["overestimate"] |= bool(np.isclose(g["out/%d/Interception" % t], g["out/%d/PotenEvap" % t]).any())
    assert all(seen.values()), seen
    g = np.load(common.GOLD + "/dyn_glacier_daily.npz")
    gm = np.concatenate([g["out/%d/GlacierMelt" % t].ravel() for t in range(6)])
    s2g = np.concatenate([g["out/%d/Snow2Glacier" % t].ravel() for t in range(6)])
    assert (gm > 0).any() and (s2g > 0).any()
    snow = np.concatenate([g["out/%d/Snow" % t].ravel() for t in range(6)])
    assert (snow > 10).any() and ((snow < 10) & (snow > 0)).any()
    g = np.load(common.GOLD + "/dyn_lai_daily.npz")
    assert any((g["out/%d/EoverR" % t] == 0).any() and (g["out/%d/EoverR" % t] > 0).any() for t in range(5))
    g = np.load(common.GOLD + "/dyn_rutter_hourly_substeps.npz")
    assert any((g["out/%d/CanopyStorage" % t] > 0).any() for t in range(6))
    g = np.load(common.GOLD + "/dyn_courant_hourly.npz")
    its = np.array([g["it/%d" % t] for t in range(5)])
    assert its[:, 1].max() > 4 and its[:, 0].max() > 1
