"""Golden trajectories of the REFERENCE's wflow_hbv ``dynamic()`` (TEST INFRASTRUCTURE ONLY; build container only).

    python -m oracle.gen_golden_hbv [case ...]      ->  tests/golden/hbv_<case>.npz

Every case: synthetic maps (wflow_b200/synth.hbv_case), forcing (synth.Forcing), then the reference's unmodified
``wflow_hbv.WflowModel.dynamic()`` (wflow/wflow_hbv.py:1222-1767) stepped through oracle/ref_hbv.py on the float32-numpy
PCRaster stand-in; states and outputs of every timestep are stored.  The CPU restatement (oracle/hbv_oracle.py) must
reproduce them bit for bit, the CUDA path within the tolerance written in tests/test_gpu_hbv.py.
"""
import os
import sys

import numpy as np

from wflow_b200 import synth

from . import ref_hbv, refload

F32 = np.float32
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # daily, basins of 12 x 12, edge branches (TTI = 0, ICF = 0, Cflux = 0), one kinematic-wave pass
    "tiles_daily": dict(n=24, kind="tiles", tile=12, steps=8, edge=1),
    # hourly with fixed sub-steps of 900 s (kinwaveIters = 1, kinwaveTstep = 900: four passes), missing cells inside the basins
    "tiles_hourly_sub": dict(n=24, kind="tiles", tile=12, steps=6, dt=3600, kinwaveIters=1, kinwaveTstep=900, mask_frac=0.1),
    # a random forest of small trees with pits everywhere; K0 / SUZ quick flow (SetKquickFlow = 1); inflow at some cells
    "forest_kquick": dict(n=20, kind="forest", steps=6, setKquickFlow=1, inflow=1),
    # seepage from an external model instead of the lower zone (ExternalQbase = 1), glacier module
    "tiles_qbase_glacier": dict(n=24, kind="tiles", tile=12, steps=6, externalQbase=1, glacier=1, cold=1),
    # snow mass wasting (MassWasting = 1: pcr.accucapacitystate of the dry pack and its liquid water): deep packs on steep cells
    "tiles_masswasting": dict(n=24, kind="tiles", tile=12, steps=5, massWasting=1, cold=1),
    # the Courant estimate of the sub-steps (kinwaveIters = 1, kinwaveTstep = 0)
    "tiles_courant": dict(n=24, kind="tiles", tile=12, steps=5, kinwaveIters=1, kinwaveTstep=0, boost=3.0),
}

STATES = ["FreeWater", "SoilMoisture", "UpperZoneStorage", "LowerZoneStorage", "InterceptionStorage", "SurfaceRunoff",
          "WaterLevel", "DrySnow", "SurfaceRunoffsub", "WaterLevelsub", "GlacierStore"]
OUTPUTS = ["Precipitation", "PotEvaporation", "Temperature", "IntEvap", "SnowMelt", "SnowCover", "SoilEvap", "ActEvap",
           "HBVSeepage", "DirectRunoff", "InUpperZone", "Percolation", "CapFlux", "QuickFlow", "RealQuickFlow", "BaseFlow",
           "InSoil", "RainAndSnowmelt", "NetInSoil", "InwaterMM", "Inwater", "QuickFlowCubic", "BaseFlowCubic",
           "SurfaceWaterSupply", "SurfaceRunoffMM", "SurfaceRunoffsubMM", "KinWaveVolume", "OldKinWaveVolume",
           "MassBalKinWave", "QCatchmentMM", "storage", "watbal", "sumprecip", "sumevap", "sumrunoff", "sumlevel",
           "sumpotevap", "sumsoilevap", "sumtemp", "suminflow", "Snow2Glacier", "GlacierMelt"]


def gen_case(name, kw):
    funcs, _ = refload.load()
    n, dt = kw["n"], kw.get("dt", 86400)
    ldd, static, state = synth.hbv_case(n, n, kind=kw["kind"], tile=kw.get("tile", 12), seed=3, timestepsecs=dt,
                                        mask_frac=kw.get("mask_frac", 0.0), glacier=kw.get("glacier", 0), edge=kw.get("edge", 0))
    if kw.get("boost"):
        state["SurfaceRunoff"] = (state["SurfaceRunoff"] * F32(kw["boost"])).astype(F32)
    # Flat index 0 is an active cell of these grids, but set_dd leaves it out of the numba network when it is the only upstream
    # neighbour of its receiver (np.any([0]) is False, wflow_funcs.py:88): kin_wave never routes it, while the map algebra
    # (Courant estimate, upstream()) still sees its initial discharge.  A real catchment never has its top-left corner inside
    # the mask; here the cell starts without discharge so that no case hinges on that quirk.
    state["SurfaceRunoff"][0, 0] = F32(0.0)
    if kw.get("massWasting"):
        mrng = np.random.default_rng(41)
        state["DrySnow"] = np.where(mrng.random((n, n)) < 0.7, mrng.uniform(0, 6000, (n, n)), 0.0).astype(F32)
        state["FreeWater"] = (state["DrySnow"] * mrng.uniform(0, 0.1, (n, n))).astype(F32)
        static["Slope"] = np.where(mrng.random((n, n)) < 0.5, mrng.uniform(0.5, 4.0, (n, n)), static["Slope"]).astype(F32)
        static["Slope"][0, 0] = F32(0.0)  # (the cell set_dd may leave out of the network keeps its snow: see above)
    fields = dict(static)
    fields.update(state)
    ref = ref_hbv.RefHbv(ldd, fields, timestepsecs=dt, kinwaveIters=kw.get("kinwaveIters", 0), kinwaveTstep=kw.get("kinwaveTstep", 0),
                         setKquickFlow=kw.get("setKquickFlow", 0), externalQbase=kw.get("externalQbase", 0),
                         massWasting=kw.get("massWasting", 0))
    forc = synth.Forcing(n, n, seed=5, timestepsecs=dt, rain_mean=kw.get("rain_mean", 14.0), t_mean=-2.0 if kw.get("cold") else 4.0)
    data = {"ldd": ldd}
    for k, v in static.items():
        data["static/" + k] = v
    for k, v in state.items():
        data["state0/" + k] = v
    rng = np.random.default_rng(99)
    inflow = None
    if kw.get("inflow"):  # sources and demands (Inflow < 0: SurfaceWaterSupply, :1599-1606)
        u = rng.random((n, n))
        inflow = np.where(u < 0.15, rng.uniform(0.05, 3.0, (n, n)), 0.0)
        inflow = np.where(u > 0.85, -np.exp(rng.uniform(np.log(0.01), np.log(30.0), (n, n))), inflow).astype(F32)
    courant = kw.get("kinwaveIters", 0) == 1 and kw.get("kinwaveTstep", 0) == 0
    for t in range(kw["steps"]):
        P, PET, T = forc.step(t)
        data["forcing/%d/P" % t], data["forcing/%d/PET" % t], data["forcing/%d/TEMP" % t] = P, PET, T
        seep = None
        if kw.get("externalQbase"):
            seep = rng.uniform(0.0, 2.0, (n, n)).astype(F32)
            data["forcing/%d/Seepage" % t] = seep
        if inflow is not None:
            data["forcing/%d/Inflow" % t] = inflow
        if courant:  # what the reference derives inside dynamic() (:1617-1626) from the state before the step
            m = ref.m
            data["it/%d" % t] = np.array(funcs.estimate_iterations_kin_wave(m.SurfaceRunoffsub, m.Beta, m.Alpha, m.timestepsecs, m.DCL, m.mv), np.int64)
        ref.step(P, PET, T, Inflow=inflow, Seepage=seep)
        for k in STATES + OUTPUTS:
            if ref.has(k):
                data["out/%d/%s" % (t, k)] = ref.get(k)
    data["meta"] = np.array(repr(kw))
    path = os.path.join(GOLD, "hbv_%s.npz" % name)
    np.savez_compressed(path, **data)
    extra = ("  it = " + " ".join(str(int(data["it/%d" % t])) for t in range(kw["steps"]))) if courant else ""
    print("wrote %s (%d KB)%s" % (os.path.basename(path), os.path.getsize(path) // 1024, extra))


if __name__ == "__main__":
    if not refload.available():
        raise SystemExit("reference checkout not found; goldens can only be generated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    for name in (sys.argv[1:] or list(CASES)):
        gen_case(name, CASES[name])
