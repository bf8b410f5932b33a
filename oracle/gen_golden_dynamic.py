"""Write tests/golden/dyn_<case>.npz: trajectories of the REFERENCE's own, unmodified ``WflowModel.dynamic()``
(PCRaster half included, through oracle/pcrshim.py) -- run in the build container, where /root/reference
exists.  Usage:  python -m oracle.gen_golden_dynamic [case ...]

Every file holds the inputs (ldd, river, statics, initial states, forcing per step, options in ``meta``) and, per
step, the float32 maps the reference holds after the step: all states, the outputs of the map-algebra phases
(interception, snow, glacier, evaporation split, river inflow glue), the outputs of the numba half, and the
water-balance / kinematic-wave-volume / cumulative outputs of the end of ``dynamic()`` (wflow_sbm.py:3310-3526).
Where ``kinwaveIters = 1`` without fixed sub-steps, ``it/<t>`` holds what the reference's
``estimate_iterations_kin_wave`` returned for the step (land, river).

The cases put the edge branches of the map-algebra functions on the grid: Cmax <= 0, EoverR = 0, TTI = 0,
snow packs around the 10 mm glacier threshold, dry steps (Precipitation = 0 under LAI), water-covered cells.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_dynamic, refload  # noqa: E402
from wflow_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
F32 = np.float32

CASES = {
    # Gash interception (daily), snow, 4 layers; Cmax = 0 / EoverR = 0 / TTI = 0 cells
    "gash_snow_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=6, edge=1),
    # modified Rutter (hourly), fixed kinematic-wave sub-steps: it_kinL = 3, it_kinR = 4
    "rutter_hourly_substeps": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), dt=3600, kinwaveIters=1,
                                   landTstep=1200, riverTstep=900, steps=6, edge=1),
    # glacierHBV on a third of the cells, snow packs on both sides of the 10 mm melt threshold
    "glacier_daily": dict(n=20, kind="forest", layer_thickness=(100, 300), steps=6, glacier=1, edge=1),
    # LAI-driven canopy parameters (Cmax, CanopyGapFraction, EoverR from LAI), dry cells and bare cells
    "lai_daily": dict(n=20, kind="forest", layer_thickness=(100, 300, 800), steps=5, lai=1),
    # frozen-soil infiltration reduction + whole-UST transpiration, 2 layers
    "frozen_ust_l2": dict(n=20, kind="forest", layer_thickness=(50,), ust=1, sir=1, steps=5),
    # no snow, one layer, Topog transfer, 6-hourly (Rutter), sub-steps 2 / 2
    "nosnow_6h_topog": dict(n=24, kind="tiles", tile=12, layer_thickness=(), snow=0, tm=1, dt=21600, kinwaveIters=1,
                            landTstep=10800, riverTstep=10800, steps=5),
    # sub-steps from the Courant number, every step (estimate_iterations_kin_wave)
    "courant_hourly": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), dt=3600, kinwaveIters=1, steps=5,
                           courant=1, land_boost=400.0),
    # missing cells inside the basins
    "masked_tiles_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=4, mask_frac=0.1),
    # a positive lateral inflow at some river cells
    "inflow_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=4, inflow=1),
}

STATES = ["CanopyStorage", "Snow", "SnowWater", "TSoil", "zi", "SubsurfaceFlow", "LandRunoff", "LandRunoffsub", "WaterLevelL",
          "WaterLevelLsub", "RiverRunoff", "RiverRunoffsub", "WaterLevelR", "WaterLevelRsub", "GlacierStore"]
OUTPUTS = [
    # map-algebra phases
    "Precipitation", "PotenEvap", "Temperature", "ThroughFall", "StemFlow", "Interception", "PotTransSoil", "SnowMelt",
    "SnowFall", "RainFallPlusMelt", "AvailableForInfiltration", "RunoffRiverCells", "RunoffLandCells",
    "ActEvapOpenWaterRiver", "ActEvapOpenWaterLand", "RestEvap", "PotSoilEvap", "PotTrans", "InwaterMM", "Inwater",
    "Transpiration", "ActEvap", "Recharge", "InterceptionWatBal", "SurfaceWatbal", "SoilWatbal", "watbal", "Snow2Glacier",
    "GlacierMelt", "Cmax", "CanopyGapFraction", "EoverR",
    # numba half
    "SatWaterDepth", "UStoreDepth", "CapFlux", "Transfer", "ActEvapUStore", "ActEvapSat", "soilevap", "soilevapsat",
    "qo_toriver", "ssf_toriver", "ExfiltWater", "InfiltExcess", "ExcessWater", "ActInfilt", "ActLeakage", "SoilInf",
    "PathInf", "InfiltExcessSoil", "InfiltExcessPath", "ActInfiltSoil", "ActInfiltPath", "ExfiltFromUstore",
    "ExfiltFromSat", "InwaterL", "InflowKinWaveCellLand", "CellInFlow", "RootStore_unsat",
    # end of dynamic(): kinematic-wave volumes and mass balances, conversions, storage, cumulative sums
    "MassBalKinWaveR", "MassBalKinWaveL", "KinWaveVolumeR", "KinWaveVolumeL", "OldKinWaveVolumeR", "OldKinWaveVolumeL",
    "InflowKinWaveCell", "RiverRunoffMM", "LandRunoffMM", "QCatchmentMM", "RunoffCoeff", "CellStorage", "DeltaStorage",
    "SatWaterFlux", "sumUstore", "SnowCover", "RootStore", "RootStore_sat", "vwcRoot", "vwc_percRoot", "SurfaceWaterSupply",
    "ExfiltWaterCubic", "InfiltExcessCubic", "SumEvapSat", "SumEvapUstore", "CumOutFlow", "CumActInfilt", "CumCellInFlow",
    "CumPrec", "CumEvap", "CumPotenEvap", "CumInt", "CumLeakage", "CumExfiltWater", "CumSurfaceWater",
]


def build_inputs(n, kind, layer_thickness, dt=86400, seed=3, tile=12, mask_frac=0.0, glacier=0, lai=0, edge=0, **_):
    ldd, river, static, state = synth.make_case(n, n, kind=kind, tile=tile, seed=seed, timestepsecs=dt,
                                                layer_thickness=layer_thickness, river_area=8, init="random", block=4,
                                                mask_frac=mask_frac)
    static = {k: np.array(np.broadcast_to(v, (n, n)), F32) for k, v in static.items()}
    state = {k: np.array(np.broadcast_to(v, (n, n)), F32) for k, v in state.items()}
    rng = np.random.default_rng(seed + 1234)
    if edge:  # edge branches of the map-algebra functions, scattered over the grid
        u = rng.random((n, n))
        static["Cmax"][u < 0.08] = 0.0          # no canopy: Gash returns P as throughfall (wflow_funcs.py:356-360)
        static["EoverR"][(u > 0.08) & (u < 0.16)] = 0.0   # division by zero -> MV -> cover(0) (:336-342)
        static["TTI"][(u > 0.16) & (u < 0.3)] = 0.0       # threshold without interval (:516-520)
        static["CanopyGapFraction"][(u > 0.3) & (u < 0.34)] = 0.95  # 1 - gap - pt < 0
        static["WaterFrac"][(u > 0.34) & (u < 0.4)] = 0.3
    if glacier:
        gf = np.where(rng.random((n, n)) < 0.35, rng.uniform(0.05, 1.0, (n, n)), 0.0).astype(F32)
        static.update(GlacierFrac=gf, G_TT=np.full((n, n), 1.3, F32), G_Cfmax=np.full((n, n), 5.3 * dt / 86400.0, F32),
                      G_SIfrac=np.full((n, n), 0.002 * dt / 86400.0, F32))
        state["GlacierStore"] = np.where(gf > 0, rng.uniform(0.0, 5500.0, (n, n)), 0.0).astype(F32)
        state["Snow"] = np.where(rng.random((n, n)) < 0.6, rng.uniform(0, 25, (n, n)), state["Snow"]).astype(F32)
    if lai:
        blk = lambda lo, hi: np.kron(rng.uniform(lo, hi, ((n + 3) // 4, (n + 3) // 4)), np.ones((4, 4)))[:n, :n].astype(F32)  # noqa: E731
        static.update(Sl=blk(0.03, 0.13), Swood=blk(0.0, 0.5), Kext=blk(0.4, 0.8))
        static["LAI"] = np.where(rng.random((n, n)) < 0.1, 0.0, blk(0.1, 6.0)).astype(F32)
    if _.get("land_boost"):  # fast overland flow: the Courant estimate of the land wave exceeds 1 sub-step
        state["LandRunoffsub"] = (state["LandRunoffsub"] * F32(_["land_boost"])).astype(F32)
        state["LandRunoff"] = state["LandRunoffsub"].copy()
    return ldd, river, static, state


def gen_case(name, kw):
    funcs, _ = refload.load()
    n, dt = kw["n"], kw.get("dt", 86400)
    ldd, river, static, state = build_inputs(**kw)
    fields = dict(static)
    fields.update(state)
    lai = fields.pop("LAI", None)
    courant = kw.get("courant", 0)
    ref = ref_dynamic.RefDynamic(
        ldd, river, fields, timestepsecs=dt, layer_thickness=kw["layer_thickness"], modelSnow=kw.get("snow", 1),
        soilInfRedu=kw.get("sir", 0), transferMethod=kw.get("tm", 0), ust=kw.get("ust", 0),
        kinwaveIters=kw.get("kinwaveIters", 0), kinwaveLandTstep=kw.get("landTstep", 0),
        kinwaveRiverTstep=kw.get("riverTstep", 0))
    forc = synth.Forcing(n, n, seed=5, timestepsecs=dt, rain_mean=kw.get("rain_mean", 20.0))
    data = {"ldd": ldd, "river": river}
    for k, v in static.items():
        data["static/" + k] = v
    for k, v in state.items():
        data["state0/" + k] = v
    nl = ref.m.maxLayers
    outs = STATES + OUTPUTS + ["UStoreLayerDepth_%d" % k for k in range(nl)] + ["vwc_%d" % k for k in range(nl)] + \
        ["vwc_perc_%d" % k for k in range(nl)]
    rng = np.random.default_rng(99)
    inflow = None
    if kw.get("inflow"):
        inflow = np.where((river != 0) & (rng.random((n, n)) < 0.3), rng.uniform(0.05, 3.0, (n, n)), 0.0).astype(F32)
    for t in range(kw["steps"]):
        P, PET, T = forc.step(t)
        if kw.get("lai") and t % 2 == 1:
            P = np.where(rng.random((n, n)) < 0.5, F32(0), P).astype(F32)  # dry cells: the EoverR = 0 branch (:2595-2602)
        data["forcing/%d/P" % t], data["forcing/%d/PET" % t], data["forcing/%d/TEMP" % t] = P, PET, T
        if inflow is not None:
            data["forcing/%d/Inflow" % t] = inflow
        if courant:  # what the reference will derive inside dynamic() (:2911-2921, 3159-3169); the river estimate
            # uses the state BEFORE the step as well (RiverRunoffsub is not touched before kin_wave)
            m = ref.m
            itl = funcs.estimate_iterations_kin_wave(m.LandRunoffsub, m.Beta, m.AlphaL, m.timestepsecs, m.DL, m.mv)
            itr = funcs.estimate_iterations_kin_wave(m.RiverRunoffsub, m.Beta, m.AlphaR, m.timestepsecs, m.DCL, m.mv)
            data["it/%d" % t] = np.array([itl, itr], np.int64)
        ref.step(P, PET, T, Inflow=inflow, LAI=lai)
        for k in outs:
            if ref.has(k):
                data["out/%d/%s" % (t, k)] = ref.get(k)
    data["meta"] = np.array(repr(kw))
    np.savez_compressed(os.path.join(GOLD, "dyn_%s.npz" % name), **data)
    extra = ""
    if courant:
        extra = "  it = " + " ".join(str(tuple(data["it/%d" % t])) for t in range(kw["steps"]))
    print("wrote dyn_%s.npz (%d KB)%s" % (name, os.path.getsize(os.path.join(GOLD, "dyn_%s.npz" % name)) // 1024, extra))


if __name__ == "__main__":
    if not refload.available():
        raise SystemExit("reference checkout not found; goldens can only be generated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    want = sys.argv[1:] or list(CASES)
    for name in want:
        gen_case(name, CASES[name])
