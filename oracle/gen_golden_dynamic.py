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
    # simple reservoirs on the river (simplereservoir, wflow_funcs.py:864-949): LDD cut at the dams, P / PET filtered over
    # the reservoir areas, outflow into the cell below
    "reservoirs_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=6, reservoirs=1),
    "reservoirs_6h": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300), dt=21600, steps=5, reservoirs=1),
    # natural lakes (naturalLake, :669-861): modified Puls, rating curve b (H - H0)^e with S = A H, S-H and H-Q tables, a linked pair
    "lakes_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=6, lakes=1),
    # snow mass wasting (pcr.accucapacitystate of the dry pack, wflow_sbm.py:2700-2707): deep packs on steep cells
    "masswasting_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=5, masswasting=1),
    # abstractions: a negative Inflow at some river cells, met from the river through the supply iteration (:3131-3146, 3203-3308)
    "abstraction_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=5, inflow=-1),
    # irrigation areas fed from river intakes (wflow_sbm.py:927-945, 3014-3031, 3314-3328): the demand of an area (potential
    # minus actual transpiration) becomes a negative Inflow at its intakes, the supply iteration meets it from the river, and
    # what was taken is spread over the area as extra water of the NEXT step (IRSupplymm); an area with two intakes, one
    # without any and an intake without an area (idtoid's left-over values, wflow_lib.py:81-101)
    "irrigation_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=6, irrigation=1),
    # paddy areas (wflow_sbm.py:2812-2815 pond evaporation, :762-785 ponding of what would run off, :3033-3050 demand when the
    # pond falls below h_min, :3330-3346 supply of the intakes spread over the cells that asked); CRPST (the crop-stage
    # multiplier of the demand, a [modelparameters] map of the user) changes from step to step
    "irrigation_paddy_daily": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300, 800), steps=7, paddy=1),
    "abstraction_hourly_substeps": dict(n=24, kind="tiles", tile=12, layer_thickness=(100, 300), dt=3600, kinwaveIters=1,
                                        landTstep=3600, riverTstep=900, steps=5, inflow=-1),
}

STATES = ["CanopyStorage", "Snow", "SnowWater", "TSoil", "zi", "SubsurfaceFlow", "LandRunoff", "LandRunoffsub", "WaterLevelL",
          "WaterLevelLsub", "RiverRunoff", "RiverRunoffsub", "WaterLevelR", "WaterLevelRsub", "GlacierStore", "ReservoirVolume",
          "LakeWaterLevel"]
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
    "IRSupplymm", "IrriDemand", "IrriDemandm3", "PondingDepth", "ActEvapPond", "irr_depth",
    # reservoirs / lakes
    "OutflowSR", "ResPercFull", "ResPrecipSR", "ResEvapSR", "DemandRelease", "LakeOutflow", "LakeVolume", "LakePrecip",
    "LakeEvap", "Inflow",
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
    if _.get("masswasting"):
        state["Snow"] = np.where(rng.random((n, n)) < 0.7, rng.uniform(0, 6000, (n, n)), 0.0).astype(F32)
        static["landSlope"] = np.where(rng.random((n, n)) < 0.5, rng.uniform(0.5, 4.0, (n, n)), static["landSlope"]).astype(F32)
        # Flat index 0 is an active cell of these grids, but set_dd leaves it out of the numba network when it is the only
        # upstream neighbour of its receiver (np.any([0]) is False, wflow_funcs.py:88), while PCRaster's accucapacitystate
        # still routes its snow.  A real catchment never has its top-left corner inside the mask; here the cell is made flat,
        # so that no snow leaves it and the case does not hinge on that quirk.
        static["landSlope"][0, 0] = F32(0.0)
    if _.get("land_boost"):  # fast overland flow: the Courant estimate of the land wave exceeds 1 sub-step
        state["LandRunoffsub"] = (state["LandRunoffsub"] * F32(_["land_boost"])).astype(F32)
        state["LandRunoff"] = state["LandRunoffsub"].copy()
    return ldd, river, static, state


def _river_spots(ldd, river, count, seed, avoid=()):
    """`count` river cells that are not pits, spread over the basins, with the area (3x3 neighbourhood) around each."""
    rng = np.random.default_rng(seed)
    cand = np.argwhere((river != 0) & (ldd != 5) & (ldd > 0))
    rng.shuffle(cand)
    spots, taken = [], np.zeros(ldd.shape, bool)
    for r, c in avoid:
        taken[max(r - 2, 0):r + 3, max(c - 2, 0):c + 3] = True
    for r, c in cand:
        if len(spots) == count:
            break
        if taken[max(r - 2, 0):r + 3, max(c - 2, 0):c + 3].any():
            continue
        spots.append((int(r), int(c)))
        taken[max(r - 2, 0):r + 3, max(c - 2, 0):c + 3] = True
    return spots


def object_inputs(kw, ldd, river, static):
    """Reservoir / lake maps of a case: (cut ldd, reservoirs dict or None, lakes dict or None)."""
    n = ldd.shape[0]
    active = ldd > 0
    rng = np.random.default_rng(77)
    cut = ldd.copy()
    res = lak = None
    spots = []
    if kw.get("reservoirs"):
        spots = _river_spots(ldd, river, 4, 5)
        locs, areas = np.zeros((n, n), np.int32), np.zeros((n, n), np.int32)
        for k, (r, c) in enumerate(spots):
            locs[r, c] = k + 1
            areas[max(r - 1, 0):r + 2, max(c - 1, 0):c + 2] = k + 1
        areas[~active] = 0
        full = lambda lo, hi: rng.uniform(lo, hi, (n, n)).astype(F32)  # noqa: E731
        res = dict(locs=locs, areas=areas, ResSimpleArea=full(1e6, 9e6), ResMaxVolume=full(2e7, 8e7),
                   ResTargetFullFrac=full(0.5, 0.9), ResMaxRelease=full(2.0, 30.0), ResDemand=full(0.1, 2.0),
                   ResTargetMinFrac=full(0.1, 0.4))
        res["ReservoirVolume"] = (res["ResMaxVolume"] * rng.uniform(0.2, 1.05, (n, n))).astype(F32)
        cut[locs > 0] = 5
    if kw.get("lakes"):
        lspots = _river_spots(ldd, river, 5, 9, avoid=spots)
        locs, areas, linked = np.zeros((n, n), np.int32), np.zeros((n, n), np.int32), np.zeros((n, n), np.int32)
        for k, (r, c) in enumerate(lspots):
            locs[r, c] = k + 1
            areas[max(r - 1, 0):r + 2, max(c - 1, 0):c + 2] = k + 1
        areas[~active] = 0
        if len(lspots) >= 5:  # lake 4 exchanges water with lake 5
            linked[lspots[3]] = 5
        full = lambda lo, hi: rng.uniform(lo, hi, (n, n)).astype(F32)  # noqa: E731
        stor, outf = np.ones((n, n), F32), np.full((n, n), 3.0, F32)
        e = np.full((n, n), 2.0, F32)
        for k, (r, c) in enumerate(lspots):
            stor[r, c] = [1, 1, 2, 1, 1][k % 5]
            outf[r, c] = [3, 2, 1, 2, 2][k % 5]
            e[r, c] = [2.0, 1.5, 2.0, 2.0, 2.0][k % 5]
        lak = dict(locs=locs, areas=areas, linked=linked, LakeArea=full(2e7, 8e7), LakeThreshold=full(0.0, 0.5),
                   LakeStorFunc=stor, LakeOutflowFunc=outf, Lake_b=full(1.0, 6.0), Lake_e=e,
                   LakeWaterLevel=full(1.0, 3.0), sh={}, hq={})
        for k, (r, c) in enumerate(lspots):
            if stor[r, c] == 2:
                H = np.linspace(0.0, 8.0, 9)
                lak["sh"][k + 1] = np.stack([H, lak["LakeArea"][r, c] * H * (1.0 + 0.05 * H)], axis=1).astype(F32).astype(np.float64)
            if outf[r, c] == 1:
                H = np.linspace(0.0, 8.0, 9)
                q = np.stack([H] + [4.0 * np.maximum(H - 0.5, 0.0) ** 1.7 * (1.0 + 0.2 * np.sin(2 * np.pi * d / 366.0)) for d in range(366)], axis=1)
                lak["hq"][k + 1] = q.astype(F32).astype(np.float64)
        cut[locs > 0] = 5
    return cut, res, lak


def irrigation_inputs(kw, ldd, river):
    """(areas, intakes, DemandReturnFlowFraction) of an irrigation case: int32 id rasters (0 = none) and a float32 raster."""
    n = ldd.shape[0]
    active = ldd > 0
    areas, intakes = np.zeros((n, n), np.int32), np.zeros((n, n), np.int32)
    for k, (r0, c0) in enumerate([(1, 1), (9, 17), (14, 3), (16, 11)]):  # four 5 x 4 blocks
        areas[r0:r0 + 5, c0:c0 + 4] = k + 1
    areas[~active] = 0
    spots = _river_spots(ldd, river, 5, 21)
    for k, (r, c) in zip([1, 2, 2, 3, 7], spots):  # area 2: two intakes; area 4: none; intake 7: no area
        intakes[r, c] = k
    drff = np.full((n, n), 0.25, F32)
    return areas, intakes, drff


def gen_case(name, kw):
    funcs, _ = refload.load()
    n, dt = kw["n"], kw.get("dt", 86400)
    ldd, river, static, state = build_inputs(**kw)
    ldd_org = ldd
    ldd, res, lak = object_inputs(kw, ldd_org, river, static)
    fields = dict(static)
    fields.update(state)
    lai = fields.pop("LAI", None)
    courant = kw.get("courant", 0)
    ref = ref_dynamic.RefDynamic(
        ldd, river, fields, ldd_org=ldd_org, reservoirs=res, lakes=lak, timestepsecs=dt, layer_thickness=kw["layer_thickness"], modelSnow=kw.get("snow", 1),
        soilInfRedu=kw.get("sir", 0), transferMethod=kw.get("tm", 0), ust=kw.get("ust", 0),
        kinwaveIters=kw.get("kinwaveIters", 0), kinwaveLandTstep=kw.get("landTstep", 0),
        kinwaveRiverTstep=kw.get("riverTstep", 0), massWasting=kw.get("masswasting", 0))
    forc = synth.Forcing(n, n, seed=5, timestepsecs=dt, rain_mean=kw.get("rain_mean", 20.0))
    data = {"ldd": ldd, "river": river}
    if kw.get("irrigation"):
        from . import pcrshim as pcr
        areas, intakes, drff = irrigation_inputs(kw, ldd, river)
        m = ref.m
        idm = lambda a: pcr.numpy2pcr(pcr.Nominal, np.where(a > 0, a, -999).astype(np.float64), -999.0)  # noqa: E731
        m.IrrigationAreas, m.IrrigationSurfaceIntakes = idm(areas), idm(intakes)
        m.nrirri = int(areas.max())  # :1637-1638
        m.DemandReturnFlowFraction = m._map(drff)
        data["irri/areas"], data["irri/intakes"], data["static/DemandReturnFlowFraction"] = areas, intakes, drff
    if kw.get("paddy"):
        from . import pcrshim as pcr
        areas, intakes, _drff = irrigation_inputs(kw, ldd, river)
        m = ref.m
        idm = lambda a: pcr.numpy2pcr(pcr.Nominal, np.where(a > 0, a, -999).astype(np.float64), -999.0)  # noqa: E731
        m.IrrigationPaddyAreas, m.IrrigationSurfaceIntakes = idm(areas), idm(intakes)
        m.nrpaddyirri = int(areas.max())  # :1639-1641
        prng = np.random.default_rng(31)
        inside = areas > 0
        hp = np.where(inside, prng.uniform(40.0, 120.0, (n, n)), 0.0).astype(F32)
        hp[areas == 3] = 0.0  # an area whose bunds hold nothing: asks for water, ponds none
        hmax = np.where(inside, prng.uniform(30.0, 60.0, (n, n)), 0.0).astype(F32)
        hmin = np.where(inside, prng.uniform(5.0, 25.0, (n, n)), 0.0).astype(F32)
        pond0 = np.where(inside, prng.uniform(0.0, 45.0, (n, n)), 0.0).astype(F32)
        m.h_p, m.h_max, m.h_min, m.PondingDepth = m._map(hp), m._map(hmax), m._map(hmin), m._map(pond0)
        m.static["h_p"] = pcr.pcr2numpy(m.h_p, -999.0).ravel()  # :2284
        data["paddy/areas"], data["paddy/intakes"] = areas, intakes
        data["static/h_p"], data["static/h_max"], data["static/h_min"], data["state0/PondingDepth"] = hp, hmax, hmin, pond0
    if res or lak:
        data["ldd_org"] = ldd_org
        data["static/filter_P_PET"] = ref.get("filter_P_PET")
    for tag, d in (("res", res), ("lake", lak)):
        for k, v in (d or {}).items():
            if k in ("sh", "hq"):
                for i, t in v.items():
                    data["%s/%s/%d" % (tag, k, i)] = np.asarray(t, F32)
            elif k in ("ReservoirVolume", "LakeWaterLevel"):
                data["state0/" + k] = v
            elif k in ("locs", "areas", "linked"):
                data["%s/%s" % (tag, k)] = v
            else:
                data["static/" + k] = v
    for k, v in static.items():
        data["static/" + k] = v
    for k, v in state.items():
        data["state0/" + k] = v
    nl = ref.m.maxLayers
    outs = STATES + OUTPUTS + ["UStoreLayerDepth_%d" % k for k in range(nl)] + ["vwc_%d" % k for k in range(nl)] + \
        ["vwc_perc_%d" % k for k in range(nl)]
    rng = np.random.default_rng(99)
    inflow = None
    if kw.get("inflow", 0) > 0:
        inflow = np.where((river != 0) & (rng.random((n, n)) < 0.3), rng.uniform(0.05, 3.0, (n, n)), 0.0).astype(F32)
    elif kw.get("inflow", 0) < 0:  # demands: some small, some more than the river carries; a few sources beside them
        u = rng.random((n, n))
        inflow = np.where((river != 0) & (u < 0.25), -np.exp(rng.uniform(np.log(0.01), np.log(300.0), (n, n))), 0.0)
        inflow = np.where((river != 0) & (u > 0.9), rng.uniform(0.05, 3.0, (n, n)), inflow).astype(F32)
    for t in range(kw["steps"]):
        P, PET, T = forc.step(t)
        if kw.get("lai") and t % 2 == 1:
            P = np.where(rng.random((n, n)) < 0.5, F32(0), P).astype(F32)  # dry cells: the EoverR = 0 branch (:2595-2602)
        data["forcing/%d/P" % t], data["forcing/%d/PET" % t], data["forcing/%d/TEMP" % t] = P, PET, T
        if inflow is not None:
            data["forcing/%d/Inflow" % t] = inflow
        if kw.get("paddy"):  # the crop stage of the step (a timeseries parameter of the user's ini)
            crpst = np.full((n, n), [1.0, 0.6, 0.0, 1.0, 0.8, 1.0, 0.3][t % 7], F32)
            ref.m.CRPST = ref.m._map(crpst)
            data["forcing/%d/CRPST" % t] = crpst
        if courant:  # what the reference will derive inside dynamic() (:2911-2921, 3159-3169); the river estimate
            # uses the state BEFORE the step as well (RiverRunoffsub is not touched before kin_wave)
            m = ref.m
            itl = funcs.estimate_iterations_kin_wave(m.LandRunoffsub, m.Beta, m.AlphaL, m.timestepsecs, m.DL, m.mv)
            itr = funcs.estimate_iterations_kin_wave(m.RiverRunoffsub, m.Beta, m.AlphaR, m.timestepsecs, m.DCL, m.mv)
            data["it/%d" % t] = np.array([itl, itr], np.int64)
        ref.step(P, PET, T, Inflow=inflow, LAI=lai)
        if kw.get("inflow", 0) < 0 or kw.get("irrigation") or kw.get("paddy"):
            data["nrit/%d" % t] = np.array(ref.m.nrit, np.int64)
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
