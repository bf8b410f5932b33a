"""The CUDA path against trajectories of the reference's whole, unmodified ``WflowModel.dynamic()`` (PCRaster half
included: tests/golden/dyn_*.npz, written by oracle/gen_golden_dynamic.py through the float32 shim).

Kernel parity per step: every step starts from the reference's own float32 states of the step before, and every
state and requested output of every cell must agree within rel 1e-5 (absolute floor 2e-5 of the field's unit,
1e3 mm3 for subsurface flow; cancellation residuals by their own bound)."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu

ATOL = {"SubsurfaceFlow": 1e3, "SoilWatbal": 1e-3, "watbal": 1e-3, "SurfaceWatbal": 1e-4, "InterceptionWatBal": 1e-4,
        "CellInFlow": 1e3, "CumCellInFlow": 1e3, "DeltaStorage": 2e-4, "MassBalKinWaveR": 1e-4, "MassBalKinWaveL": 1e-4,
        "KinWaveVolumeR": 1e-2, "OldKinWaveVolumeR": 1e-2, "KinWaveVolumeL": 1e-2, "OldKinWaveVolumeL": 1e-2,
        "ReservoirVolume": 1.0, "LakeVolume": 1.0}  # m3 of 1e7 .. 1e8
# pcr.catchmenttotal runs over the LDD map, the device over the network set_dd built: they differ downstream of a cell that
# set_dd's flat-index-0 quirk drops (wflow_funcs.py:88) -- such a cell is not on the device at all
CATCHMENT = ("QCatchmentMM", "RunoffCoeff")
OBJECT_FIELDS = ("IrriDemand", "IrriDemandm3", "ReservoirVolume", "OutflowSR", "ResPercFull", "ResPrecipSR", "ResEvapSR", "DemandRelease", "LakeWaterLevel",
                 "LakeOutflow", "LakeVolume", "LakePrecip", "LakeEvap")


def _downstream_of_dropped(ldd, mask):
    """cells of the network that have a valid-ldd cell OUTSIDE the network somewhere upstream"""
    nrow, ncol = ldd.shape
    dr = {7: (-1, -1), 8: (-1, 0), 9: (-1, 1), 4: (0, -1), 6: (0, 1), 1: (1, -1), 2: (1, 0), 3: (1, 1)}
    tainted = (ldd >= 1) & (ldd <= 9) & ~mask
    hit = np.zeros_like(mask)
    for r, c in zip(*np.nonzero(tainted)):
        while True:
            code = int(ldd[r, c])
            if code not in dr:
                break
            r, c = r + dr[code][0], c + dr[code][1]
            if not (0 <= r < nrow and 0 <= c < ncol) or not (1 <= ldd[r, c] <= 9) or hit[r, c]:
                break
            hit[r, c] = True
    return hit


@pytest.mark.parametrize("name", common.golden_dyn_cases())
def test_cuda_matches_reference_dynamic(name):
    g, kw, _, mod = common.load_dyn_case(name, oracle=False, gpu=True)
    try:
        nodes = mod.network()[0]
        mask = np.zeros(mod.shape, bool)
        mask.flat[nodes] = True
        from wflow_b200 import _lib
        known = set(_lib.field_names())
        skip = {"Cmax", "CanopyGapFraction", "EoverR",  # inputs here; the reference overwrites them from LAI
                "Inflow"}                                # likewise: the reference adds the objects' outflow to the map
        outs = [f for f in common.dyn_outputs(g) if f in known and f not in skip]
        states = [f for f in outs if "state0/" + f in g.files and f != "SatWaterDepth"] + ["SatWaterDepth"]  # the map itself, last
        irrigation = "irri/areas" in g.files or "paddy/areas" in g.files  # IRSupplymm is carried from step to step like a state (0 before the first)
        mod.request(*outs)
        cmask = mask & ~_downstream_of_dropped(np.asarray(g["ldd"]), mask)
        worst = (0.0, "")
        for t in range(kw["steps"]):
            for f in ("P", "PET", "TEMP", "Inflow", "CRPST"):
                key = "forcing/%d/%s" % (t, f)
                if key in g.files:
                    mod.set(f, g[key])
            for f in states:  # the reference's states of the step before, bit for bit
                mod.set(f, g["state0/" + f] if t == 0 else g["out/%d/%s" % (t - 1, f)])
            if irrigation and t > 0:
                mod.set("IRSupplymm", g["out/%d/IRSupplymm" % (t - 1)])
            if kw.get("courant"):
                itl, itr = mod.estimate_iterations("land"), mod.estimate_iterations("river")
                assert (itl, itr) == tuple(int(v) for v in g["it/%d" % t]), (name, t)
                mod.set_substeps(itl, itr)
            mod.step()
            if "nrit/%d" % t in g.files:  # iterations of the supply loop (self.nrit, wflow_sbm.py:3203-3300)
                assert mod.supply_iterations == int(g["nrit/%d" % t]), (name, t)
            for f in outs:
                ref = g["out/%d/%s" % (t, f)]
                msk = cmask if f in CATCHMENT else mask
                if f in OBJECT_FIELDS:  # defined at the objects' cells only (missing values elsewhere in the reference)
                    msk = msk & (ref > -999.0)
                e = common.compare(mod.get(f), ref, msk, 1e-5, ATOL.get(f, 2e-5))
                if e > worst[0]:
                    worst = (e, "%s step %d" % (f, t))
                assert e <= 1.0, (name, t, f, e)
        print("%s: %d outputs x %d steps, worst error/tolerance %.3g (%s)" % (name, len(outs), kw["steps"], worst[0], worst[1]))
    finally:
        mod.close()


END_OUTPUTS = ("MassBalKinWaveR", "MassBalKinWaveL", "KinWaveVolumeR", "KinWaveVolumeL", "RiverRunoffMM", "LandRunoffMM",
               "QCatchmentMM", "RunoffCoeff", "CellStorage", "DeltaStorage", "SatWaterFlux", "SnowCover", "RootStore", "vwcRoot",
               "CumPrec", "CumEvap", "CumLeakage", "CumSurfaceWater", "vwc_perc_0", "SurfaceWatbal", "watbal")


@pytest.mark.parametrize("field", END_OUTPUTS)
def test_an_output_requested_alone(field):
    """An ini usually lists ONE of these outputs: every one must pull in what it is formed from by itself (found by the
    reference's dynamic-vegetation case, whose [summary] asks for QCatchmentMM and nothing else of the end of dynamic())."""
    g, kw, _, mod = common.load_dyn_case("gash_snow_daily", oracle=False, gpu=True)
    try:
        nodes = mod.network()[0]
        mask = np.zeros(mod.shape, bool)
        mask.flat[nodes] = True
        if field in CATCHMENT:
            mask &= ~_downstream_of_dropped(np.asarray(g["ldd"]), mask)
        mod.request(field)
        for f in ("P", "PET", "TEMP"):
            mod.set(f, g["forcing/0/%s" % f])
        mod.step()
        e = common.compare(mod.get(field), g["out/0/%s" % field], mask, 1e-5, ATOL.get(field, 2e-5))
        assert e <= 1.0, (field, e)
    finally:
        mod.close()


def test_two_sweeps_around_the_paddy_demand_equal_one_sweep():
    """With paddy areas the step sweeps twice over the same schedule (subsurface task | the areas' demand | overland and river
    tasks).  Areas that neither pond nor ask (h_p = h_min = 0) must therefore leave every state bit-identical to the model
    without them, over several free-running hourly steps with river sub-steps -- the tasks read their own cell's hand-offs
    and the level above, whichever sweep wrote them."""
    from wflow_b200 import synth
    n, dt = 64, 3600
    kw = dict(kind="tiles", tile=32, layer_thickness=(100, 300, 800), dt=dt, itL=1, itR=4, river_area=20, oracle=False)
    _, a, ldd, river = common.make_pair(n, n, **kw)
    _, b, _, _ = common.make_pair(n, n, **kw)
    try:
        areas = np.zeros((n, n), np.int32)
        areas[4:20, 6:30] = 1
        areas[40:60, 34:50] = 2
        areas[ldd == 0] = 0
        intakes = np.zeros((n, n), np.int32)
        cand = np.argwhere((river != 0) & (ldd != 5))
        intakes[tuple(cand[len(cand) // 3])] = 1
        intakes[tuple(cand[2 * len(cand) // 3])] = 2
        zero = np.zeros((n, n), np.float32)
        for f in ("h_p", "h_max", "h_min"):
            b.set(f, zero)
        assert b.paddy_create(areas, intakes) == 2
        forc = synth.Forcing(n, n, seed=5, timestepsecs=dt, rain_mean=20.0)
        states = common.state_names(a.max_layers)
        for t in range(6):
            P, PET, T = forc.step(t)
            for m in (a, b):
                m.set("P", P); m.set("PET", PET); m.set("TEMP", T)
                m.step()
        for k in states:
            np.testing.assert_array_equal(b.get(k), a.get(k), err_msg=k)
        assert float(np.nanmax(b.get("PondingDepth"))) == 0.0 and float(np.nanmax(b.get("IRSupplymm"))) == 0.0
        assert float(a.get("RiverRunoff")[river != 0].max()) > 0.0
    finally:
        a.close()
        b.close()
