"""Write tests/golden/*.npz from the REFERENCE's own numba kernels (run in the build
container, where /root/reference exists).  Usage:  python -m oracle.gen_golden

Golden sets
  set_dd.npz          nodes / nodes_up / level offsets of reference set_dd on hand-made and
                      random LDDs, including the flat-index-0 quirk (wflow_funcs.py:88)
  kinwave.npz         reference kinematic_wave on log-uniform sweeps + zero / NaN branches
  kinwave_ssf.npz     reference kinematic_wave_ssf on log-uniform sweeps + early-out branch
  step_<case>.npz     multi-step trajectories: inputs (ldd, river, statics, initial states,
                      forcing per step) and, per step, the float32 outputs of the reference's
                      sbm_cell + kin_wave driven through the reference's own pcr2numpy /
                      numpy2pcr glue (oracle/ref_driver.py).  The PCRaster phases feeding
                      them come from the C oracle (PCRaster is not installable; parity of
                      those phases is unpinned, see oracle/wflow_sbm_oracle.c header).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import ref_driver, refload  # noqa: E402
from wflow_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

STEP_CASES = {
    "forest_l4_daily": dict(n=20, kind="forest", layer_thickness=(100, 300, 800), steps=12),
    "tiles_l4_daily": dict(n=32, kind="tiles", layer_thickness=(100, 300, 800), steps=12),
    "tiles_l1_topog": dict(n=32, kind="tiles", layer_thickness=(), tm=1, steps=10),
    "forest_l2_ust_frozen": dict(n=20, kind="forest", layer_thickness=(50,), ust=1, sir=1, steps=10),
    "tiles_l4_hourly_substeps": dict(n=32, kind="tiles", layer_thickness=(100, 300, 800), dt=3600, itL=3, itR=4, steps=10),
    "forest_l3_nosnow_6h": dict(n=20, kind="forest", layer_thickness=(100, 300), snow=0, itL=2, itR=2, dt=21600, steps=10),
    "pits_l4_daily": dict(n=16, kind="pits", layer_thickness=(100, 300, 800), steps=10),
    "tiles_l4_coldstart": dict(n=32, kind="tiles", layer_thickness=(100, 300, 800), steps=8, init="cold"),
}

OUT_FIELDS = ["SubsurfaceFlow", "LandRunoffsub", "WaterLevelLsub", "LandRunoff", "WaterLevelL", "zi", "SatWaterDepth",
              "UStoreDepth", "CapFlux", "Transfer", "ActEvapUStore", "ActEvapSat", "soilevap", "soilevapsat",
              "qo_toriver", "ssf_toriver", "ExfiltWater", "InfiltExcess", "ExcessWater", "ActInfilt", "ActLeakage",
              "InwaterL", "CellInFlow", "SoilWatbal", "Inwater", "RiverRunoff", "WaterLevelR", "RiverRunoffsub",
              "WaterLevelRsub"]
PHASE_DIAGS = ["RunoffLandCells", "ActEvapOpenWaterLand", "ActEvapOpenWaterRiver", "AvailableForInfiltration",
               "PotSoilEvap", "PotTrans", "InwaterMM"]


def build_case(n, kind, layer_thickness, dt=86400, snow=1, itL=1, itR=1, tm=0, ust=0, sir=0, seed=3, init="random", **_):
    ldd, river, static, state = synth.make_case(n, n, kind=kind, tile=16, seed=seed, timestepsecs=dt,
                                                layer_thickness=layer_thickness, river_area=8, init=init, block=4)
    orc = O.Oracle(ldd, river, timestepsecs=dt, layer_thickness=layer_thickness, modelSnow=snow, it_kinL=itL,
                   it_kinR=itR, transferMethod=tm, ust=ust, soilInfRedu=sir)
    for k, v in static.items():
        orc.set(k, v)
    for k, v in state.items():
        orc.set(k, v)
    return orc, ldd, river, static, state


def gen_steps():
    for name, kw in STEP_CASES.items():
        orc, ldd, river, static, state = build_case(**kw)
        ref = ref_driver.ReferenceModel(orc)
        forc = synth.Forcing(kw["n"], kw["n"], seed=5, timestepsecs=kw.get("dt", 86400), rain_mean=kw.get("rain_mean", 20.0))
        out_fields = OUT_FIELDS + ["UStoreLayerDepth_%d" % k for k in range(orc.maxLayers)]
        orc.want(*PHASE_DIAGS)
        orc.want(*[f for f in out_fields if f not in orc.fields])
        data = {"ldd": ldd, "river": river}
        for k, v in static.items():
            data["static/" + k] = v
        for k, v in state.items():
            data["state0/" + k] = v
        pre_names = ["zi", "SubsurfaceFlow", "WaterLevelL", "LandRunoffsub", "RiverRunoffsub"] + \
                    ["UStoreLayerDepth_%d" % k for k in range(orc.maxLayers)]
        for t in range(kw["steps"]):
            P, PET, T = forc.step(t)
            data["forcing/%d/P" % t], data["forcing/%d/PET" % t], data["forcing/%d/TEMP" % t] = P, PET, T
            orc.set("P", P); orc.set("PET", PET); orc.set("TEMP", T)
            pre = {k: orc[k].copy() for k in pre_names}
            orc.step()
            out = ref.step(pre, {k: orc[k] for k in PHASE_DIAGS})
            for k in out_fields:
                # the trajectory continues from the ORACLE's states; they must equal the reference's
                # on the network cells or the golden would drift: assert that here
                mask = np.zeros(orc.shape, bool)
                mask.flat[orc.net.flat_nodes] = True
                if k == "SoilWatbal":  # cancellation residual (~1e-6 mm of ~1e3 mm terms): compare absolutely
                    ok = np.allclose(orc[k][mask], out[k][mask], rtol=0, atol=1e-9)
                else:
                    ok = np.array_equal(orc[k][mask], out[k][mask])
                if not ok:
                    raise SystemExit("oracle != reference for %s at step %d of %s" % (k, t, name))
                data["out/%d/%s" % (t, k)] = out[k]
        data["meta"] = np.array(repr(kw))
        np.savez_compressed(os.path.join(GOLD, "step_%s.npz" % name), **data)
        print("wrote step_%s.npz" % name)


def gen_set_dd():
    funcs, _ = refload.load()
    cases = {
        "quirk2x2": np.array([[6, 5], [5, 5]], np.uint8),
        "quirk_chain": np.array([[6, 6, 5], [0, 0, 0]], np.uint8),
        "quirk_sibling": np.array([[3, 0, 0], [6, 5, 0], [0, 0, 0]], np.uint8),
        "hand3x3": np.array([[5, 4, 4], [8, 7, 4], [9, 8, 5]], np.uint8),
        "nopit": np.array([[6, 4], [0, 0]], np.uint8),
        "forest": synth.ldd_random_forest(23, 31, seed=1),
        "forest2": synth.ldd_random_forest(40, 17, seed=2, missing_frac=0.2),
        "tiles": synth.ldd_basin_tiles(48, 64, tile=16),
    }
    data = {}
    for name, ldd in cases.items():
        l = ldd.astype(np.float64)
        l[l == 0] = -999.0
        nodes, up = funcs.set_dd(l)
        data[name + "/ldd"] = ldd
        data[name + "/nodes"] = np.concatenate([np.asarray(a, np.int64) for a in nodes]) if len(nodes) else np.zeros(0, np.int64)
        data[name + "/nodes_up"] = np.concatenate([np.asarray(a, np.int64).reshape(-1, 8) for a in up])
        data[name + "/lev_off"] = np.concatenate([[0], np.cumsum([len(a) for a in nodes])]).astype(np.int64)
    np.savez_compressed(os.path.join(GOLD, "set_dd.npz"), **data)
    print("wrote set_dd.npz")


def gen_kinwave():
    funcs, _ = refload.load()
    rng = np.random.default_rng(11)
    n = 4000
    # river regime and overland regime, log-uniform; plus degenerate rows
    Qin = np.exp(rng.uniform(np.log(1e-6), np.log(3e3), n)); Qold = np.exp(rng.uniform(np.log(1e-6), np.log(3e3), n))
    q = np.exp(rng.uniform(np.log(1e-9), np.log(1e-2), n)); alpha = np.exp(rng.uniform(np.log(0.1), np.log(60), n))
    beta = np.full(n, float(np.float32(0.6))); dT = rng.choice([900.0, 3600.0, 86400.0], n)
    dX = np.exp(rng.uniform(np.log(100), np.log(5000), n))
    z = rng.random(n)
    Qin[z < 0.1] = 0.0
    Qold[(z > 0.05) & (z < 0.2)] = 0.0
    q[(z > 0.15) & (z < 0.3)] = 0.0
    q[(z > 0.3) & (z < 0.33)] *= -1.0
    allzero = z > 0.97
    Qin[allzero] = 0.0; Qold[allzero] = 0.0; q[allzero] = 0.0
    out = np.array([funcs.kinematic_wave(*a) for a in zip(Qin, Qold, q, alpha, beta, dT, dX)])
    np.savez_compressed(os.path.join(GOLD, "kinwave.npz"), Qin=Qin, Qold=Qold, q=q, alpha=alpha, beta=beta, dT=dT, dX=dX, out=out)
    n = 3000
    ssf_in = np.exp(rng.uniform(np.log(1e3), np.log(1e13), n)); ssf_old = np.exp(rng.uniform(np.log(1e3), np.log(1e13), n))
    D = rng.uniform(400, 7000, n); zi_old = rng.uniform(0, 1, n) * D
    r = rng.normal(0, 5, n) * 1e6; Kh = np.exp(rng.uniform(0, np.log(100), n)); Ks = np.exp(rng.uniform(np.log(5), np.log(8000), n))
    slope = np.exp(rng.uniform(np.log(1e-3), np.log(0.3), n)); neff = rng.uniform(0.3, 0.55, n)
    f = neff / np.exp(rng.uniform(np.log(17), np.log(12000), n)); dx = rng.choice([1e6, 1.4142e6], n); w = 1e12 / dx
    z = rng.random(n)
    ssf_in[z < 0.15] = 0.0
    ssf_old[(z > 0.1) & (z < 0.25)] = 0.0
    with np.errstate(all="ignore"):
        ssfmax = (Kh * Ks * slope) / f * (1.0 - np.exp(-f * D))
    args = [ssf_in, ssf_old, zi_old, r, Kh, Ks, slope, neff, f, D, np.ones(n), dx, w, ssfmax]
    out = np.array([funcs.kinematic_wave_ssf(*a) for a in zip(*args)])
    np.savez_compressed(os.path.join(GOLD, "kinwave_ssf.npz"), args=np.array(args), out=out)
    print("wrote kinwave.npz, kinwave_ssf.npz")


def gen_kinwave_ssf_dry():
    """kinwave_ssf_dry.npz: reference kinematic_wave_ssf on dry and almost dry columns (water table at or near
    the soil bottom, no or vanishing subsurface flow, vanishing recharge) up to f*D ~ 200 -- the regime where the
    reference returns from its first-guess evaluation before the Newton loop starts (wflow_funcs.py:224-253) or
    sits at its 1e-30 floor for all 3000 iterations.  Found by the full-size config 2 parity test."""
    funcs, _ = refload.load()
    rng = np.random.default_rng(20261020)
    n = 3000
    D = rng.uniform(400, 7000, n).astype(np.float32).astype(np.float64)
    z = rng.random(n)
    zi_old = D.copy()
    sel = (z > 0.70) & (z <= 0.85)
    zi_old[sel] = (D[sel] * (1.0 - np.exp(rng.uniform(np.log(1e-7), np.log(1e-3), sel.sum())))).astype(np.float32)
    sel = z > 0.85
    zi_old[sel] = (D[sel] * rng.uniform(0.8, 1.0, sel.sum())).astype(np.float32)
    z = rng.random(n)
    ssf_old = np.zeros(n)
    sel = (z > 0.40) & (z <= 0.70)
    ssf_old[sel] = float(np.float32(1e-30))
    sel = z > 0.70
    ssf_old[sel] = np.exp(rng.uniform(np.log(1e-40), np.log(1e3), sel.sum())).astype(np.float32)
    z = rng.random(n)
    ssf_in = np.zeros(n)
    sel = z > 0.70
    ssf_in[sel] = np.exp(rng.uniform(np.log(1e-30), np.log(1e3), sel.sum()))
    z = rng.random(n)
    r = np.exp(rng.uniform(np.log(1e-14), np.log(1e4), n))
    r[(z > 0.70) & (z <= 0.85)] *= -1e-3
    r[z > 0.85] = 0.0
    Kh = np.exp(rng.uniform(0, np.log(100), n)); Ks = np.exp(rng.uniform(np.log(5), np.log(8000), n))
    slope = np.exp(rng.uniform(np.log(1e-3), np.log(0.3), n)); neff = rng.uniform(0.3, 0.55, n)
    f = neff / np.exp(rng.uniform(np.log(17), np.log(12000), n)); dx = rng.choice([1e6, 1.4142e6], n); w = 1e12 / dx
    with np.errstate(all="ignore"):
        ssfmax = (Kh * Ks * slope) / f * (1.0 - np.exp(-f * D))
    args = [ssf_in, ssf_old, zi_old, r, Kh, Ks, slope, neff, f, D, np.ones(n), dx, w, ssfmax]
    out = np.array([funcs.kinematic_wave_ssf(*a) for a in zip(*args)])
    np.savez_compressed(os.path.join(GOLD, "kinwave_ssf_dry.npz"), args=np.array(args), out=out)
    print("wrote kinwave_ssf_dry.npz: f*D up to %.0f, %d early-outs, %d at the floor" %
          ((f * D).max(), int((out[:, 0] == 0).sum()), int((out[:, 0] == 1e-30).sum())))


if __name__ == "__main__":
    if not refload.available():
        raise SystemExit("reference checkout not found; goldens can only be generated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "dry":  # add the dry-column set without touching the others
        gen_kinwave_ssf_dry()
        raise SystemExit(0)
    gen_set_dd()
    gen_kinwave()
    gen_kinwave_ssf_dry()
    gen_steps()
