"""Shared helpers of the parity tests: build the same synthetic case for the CPU oracle and
for the CUDA model, step both, compare."""
import numpy as np

from wflow_b200 import synth

DIAGS = ["SatWaterDepth", "UStoreDepth", "CapFlux", "Transfer", "ActEvapUStore", "ActEvapSat", "soilevap",
         "soilevapsat", "ActEvap", "Recharge", "InwaterMM", "qo_toriver", "ssf_toriver", "ExfiltWater",
         "InfiltExcess", "ExcessWater", "ActInfilt", "ActLeakage", "SoilInf", "PathInf", "InfiltExcessSoil",
         "InfiltExcessPath", "ActInfiltSoil", "ActInfiltPath", "ExfiltFromUstore", "ExfiltFromSat", "InwaterL",
         "InflowKinWaveCellLand", "CellInFlow", "RootStore_unsat", "Precipitation", "PotenEvap", "ThroughFall",
         "StemFlow", "Interception", "PotTransSoil", "RainFallPlusMelt", "AvailableForInfiltration",
         "RunoffRiverCells", "RunoffLandCells", "ActEvapOpenWaterRiver", "ActEvapOpenWaterLand", "RestEvap",
         "PotSoilEvap", "PotTrans", "Transpiration", "InterceptionWatBal", "SurfaceWatbal"]
RIVER_STATES = ["RiverRunoff", "RiverRunoffsub", "WaterLevelR", "WaterLevelRsub"]


def state_names(max_layers, snow=True):
    s = ["zi", "SubsurfaceFlow", "LandRunoff", "LandRunoffsub", "WaterLevelL", "WaterLevelLsub", "CanopyStorage"]
    s += ["UStoreLayerDepth_%d" % k for k in range(max_layers)]
    if snow:
        s += ["Snow", "SnowWater", "TSoil"]
    return s + RIVER_STATES


def make_pair(nrow, ncol, *, kind="forest", tile=16, layer_thickness=(100, 300, 800), dt=86400, snow=1, itL=1, itR=1,
              seed=3, river_area=8, tm=0, ust=0, sir=0, oracle=True, gpu=True, init="random", block=4, mask_frac=0.0,
              glacier=0, lai=0):
    """Returns (oracle_model or None, cuda_model or None, ldd, river)."""
    ldd, river, static, state = synth.make_case(nrow, ncol, kind=kind, tile=tile, seed=seed, timestepsecs=dt,
                                                layer_thickness=layer_thickness, river_area=river_area, init=init,
                                                block=block, mask_frac=mask_frac)
    if glacier:  # glacierHBV inputs (wflow/wflow_funcs.py:549-603; wflow_sbm.py:2715-2743): a third of the cells glaciated
        rng = np.random.default_rng(seed + 77)
        gf = np.where(rng.random((nrow, ncol)) < 0.35, rng.uniform(0.05, 1.0, (nrow, ncol)), 0.0).astype(np.float32)
        static = dict(static, GlacierFrac=gf, G_TT=np.float32(1.3), G_Cfmax=np.float32(5.3 * dt / 86400.0),
                      G_SIfrac=np.float32(0.002 * dt / 86400.0))
        state = dict(state, GlacierStore=np.where(gf > 0, rng.uniform(0.0, 5500.0, (nrow, ncol)), 0.0).astype(np.float32))
    if lai:  # LAI-driven canopy parameters (wflow_sbm.py:2587-2603): land-cover style blocks, some cells bare (LAI = 0)
        rng = np.random.default_rng(seed + 99)
        blk = lambda lo, hi: np.kron(rng.uniform(lo, hi, ((nrow + 7) // 8, (ncol + 7) // 8)), np.ones((8, 8)))[:nrow, :ncol].astype(np.float32)
        static = dict(static, LAI=np.where(rng.random((nrow, ncol)) < 0.1, 0.0, blk(0.1, 6.0)).astype(np.float32),
                      Sl=blk(0.03, 0.13), Swood=blk(0.0, 0.5), Kext=blk(0.4, 0.8))
    orc = mod = None
    if oracle:
        from oracle import oracle as O
        orc = O.Oracle(ldd, river, timestepsecs=dt, layer_thickness=layer_thickness, modelSnow=snow, it_kinL=itL,
                       it_kinR=itR, transferMethod=tm, ust=ust, soilInfRedu=sir, glacier=glacier)
        for k, v in static.items():
            orc.set(k, v)
        for k, v in state.items():
            orc.set(k, v)
    if gpu:
        from wflow_b200.core import SbmModel
        mod = SbmModel(ldd, river, timestepsecs=dt, layer_thickness=layer_thickness, model_snow=snow, it_kin_l=itL,
                       it_kin_r=itR, transfer_method=tm, ust=ust, soil_inf_redu=sir, glacier=glacier)
        skip = {"River", "Beta", "SatWaterDepth"}
        for k, v in static.items():
            if k not in skip:
                mod.set(k, v)
        for k, v in state.items():
            if k not in skip:
                mod.set(k, v)
    return orc, mod, ldd, river


def network_mask(orc):
    m = np.zeros(orc.shape, bool)
    m.flat[orc.net.flat_nodes] = True
    return m


def compare(a, b, mask, rtol, atol):
    """max over masked cells of |a-b| / (atol + rtol*|b|); <= 1 passes."""
    a = a[mask].astype(np.float64)
    b = b[mask].astype(np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / (atol + rtol * np.abs(b))))


# ---- golden fixtures (tests/golden, written by oracle/gen_golden.py from the reference) ----
import ast
import glob
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_step_cases():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLD, "step_*.npz")))


def load_golden_case(name, oracle=True, gpu=False):
    """Rebuild the oracle / CUDA model of a golden trajectory from its stored inputs."""
    g = np.load(os.path.join(GOLD, "step_%s.npz" % name))
    kw = ast.literal_eval(str(g["meta"]))
    lt = kw["layer_thickness"]
    opt = dict(dt=kw.get("dt", 86400), snow=kw.get("snow", 1), itL=kw.get("itL", 1), itR=kw.get("itR", 1),
               tm=kw.get("tm", 0), ust=kw.get("ust", 0), sir=kw.get("sir", 0))
    static = {k[7:]: g[k] for k in g.files if k.startswith("static/")}
    state = {k[7:]: g[k] for k in g.files if k.startswith("state0/")}
    orc = mod = None
    if oracle:
        from oracle import oracle as O
        orc = O.Oracle(g["ldd"], g["river"], timestepsecs=opt["dt"], layer_thickness=lt, modelSnow=opt["snow"],
                       it_kinL=opt["itL"], it_kinR=opt["itR"], transferMethod=opt["tm"], ust=opt["ust"],
                       soilInfRedu=opt["sir"])
        for k, v in {**static, **state}.items():
            orc.set(k, v)
    if gpu:
        from wflow_b200.core import SbmModel
        mod = SbmModel(g["ldd"], g["river"], timestepsecs=opt["dt"], layer_thickness=lt, model_snow=opt["snow"],
                       it_kin_l=opt["itL"], it_kin_r=opt["itR"], transfer_method=opt["tm"], ust=opt["ust"],
                       soil_inf_redu=opt["sir"])
        for k, v in {**static, **state}.items():
            if k not in ("River", "Beta", "SatWaterDepth"):
                mod.set(k, v)
    return g, kw, orc, mod


# ---- goldens of the reference's whole dynamic() (tests/golden/dyn_*.npz, oracle/gen_golden_dynamic.py) ----
def golden_dyn_cases(objects=True):
    """objects=False: without the reservoir / lake / mass-wasting / abstraction / irrigation cases (the C oracle does not restate those
    optional modules; the CUDA path is checked against the reference's own goldens for them)."""
    names = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLD, "dyn_*.npz")))
    return [n for n in names if objects or not n.startswith(("reservoirs", "lakes", "masswasting", "abstraction", "irrigation"))]


def dyn_substeps(kw):
    """(it_kinL, it_kinR) of a dyn golden with fixed sub-steps (wflow_sbm.py:2911-2923, 3159-3171)."""
    dt = kw.get("dt", 86400)
    if not kw.get("kinwaveIters", 0):
        return 1, 1
    itl = int(np.ceil(dt / kw["landTstep"])) if kw.get("landTstep", 0) else 1
    itr = int(np.ceil(dt / kw["riverTstep"])) if kw.get("riverTstep", 0) else 1
    return itl, itr


def load_dyn_case(name, oracle=True, gpu=False):
    """Rebuild the oracle / CUDA model of a dyn golden from its stored inputs.  Returns (g, kw, orc, mod)."""
    g = np.load(os.path.join(GOLD, "dyn_%s.npz" % name))
    kw = ast.literal_eval(str(g["meta"]))
    lt = kw["layer_thickness"]
    dt = kw.get("dt", 86400)
    itl, itr = dyn_substeps(kw)
    static = {k[7:]: g[k] for k in g.files if k.startswith("static/")}
    state = {k[7:]: g[k] for k in g.files if k.startswith("state0/")}
    glacier = int(kw.get("glacier", 0))
    orc = mod = None
    if oracle:
        from oracle import oracle as O
        orc = O.Oracle(g["ldd"], g["river"], timestepsecs=dt, layer_thickness=lt, modelSnow=kw.get("snow", 1), it_kinL=itl,
                       it_kinR=itr, transferMethod=kw.get("tm", 0), ust=kw.get("ust", 0), soilInfRedu=kw.get("sir", 0),
                       glacier=glacier)
        for k, v in {**static, **state}.items():
            if k in orc.names:
                orc.set(k, v)
    if gpu:
        from wflow_b200.core import SbmModel
        mod = SbmModel(g["ldd"], g["river"], timestepsecs=dt, layer_thickness=lt, model_snow=kw.get("snow", 1),
                       it_kin_l=itl, it_kin_r=itr, transfer_method=kw.get("tm", 0), ust=kw.get("ust", 0),
                       soil_inf_redu=kw.get("sir", 0), glacier=glacier)
        for k, v in {**static, **state}.items():
            if k not in ("River", "Beta", "SatWaterDepth", "ReservoirVolume", "LakeWaterLevel", "PondingDepth"):
                mod.set(k, v)
        if kw.get("masswasting"):
            mod.set_option("MassWasting", 1)
        if kw.get("inflow", 0) < 0:
            mod.set_option("Abstractions", 1)
        if "irri/areas" in g.files:  # irrigation areas and intakes (:1096-1115, 3014-3031, 3314-3328)
            assert mod.irrigation_create(g["irri/areas"], g["irri/intakes"]) == len(np.unique(g["irri/areas"][g["irri/areas"] > 0]))
        if "paddy/areas" in g.files:  # paddy areas (:762-785, 2811-2815, 3033-3050, 3330-3346)
            assert mod.paddy_create(g["paddy/areas"], g["paddy/intakes"]) == len(np.unique(g["paddy/areas"][g["paddy/areas"] > 0]))
            mod.set("PondingDepth", state["PondingDepth"])
        if "res/locs" in g.files:  # simple reservoirs (wflow_sbm.py:1804-1826, 2822-2852)
            assert mod.reservoirs_create(g["res/locs"], g["res/areas"], g["ldd_org"]) == int((g["res/locs"] > 0).sum())
            mod.set("ReservoirVolume", state["ReservoirVolume"])
        if "lake/locs" in g.files:  # natural lakes (:1827-1893, 2853-2883)
            sh = {int(k.split("/")[2]): g[k] for k in g.files if k.startswith("lake/sh/")}
            hq = {int(k.split("/")[2]): g[k] for k in g.files if k.startswith("lake/hq/")}
            assert mod.lakes_create(g["lake/locs"], g["lake/linked"], g["lake/areas"], g["ldd_org"], sh, hq) == int((g["lake/locs"] > 0).sum())
            mod.set("LakeWaterLevel", state["LakeWaterLevel"])
    return g, kw, orc, mod


def dyn_outputs(g, t=0):
    return sorted(k.split("/")[2] for k in g.files if k.startswith("out/%d/" % t))
