import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common
from wflow_b200 import core
name = sys.argv[1] if len(sys.argv) > 1 else "tiles_l1_topog"
g, kw, orc, mod = common.load_golden_case(name, oracle=True, gpu=True)
nodes, _, _ = core.build_schedule(g["ldd"])
mask = np.zeros(g["ldd"].shape, bool); mask.flat[nodes] = True
outs = sorted({k.split("/")[2] for k in g.files if k.startswith("out/0/")})
outs = [f for f in outs if f != "SoilWatbal"]
carried = [f for f in outs if f in common.state_names(8) and f != "SatWaterDepth"]
mod.request(*outs); orc.want(*[f for f in outs if f not in orc.fields])
for t in range(kw["steps"]):
    for f in ("P", "PET", "TEMP"):
        mod.set(f, g["forcing/%d/%s" % (t, f)]); orc.set(f, g["forcing/%d/%s" % (t, f)])
    if t > 0:
        for f in carried:
            mod.set(f, g["out/%d/%s" % (t - 1, f)])
    mod.step(); orc.step()
    for f in outs:
        a = mod.get(f); b = g["out/%d/%s" % (t, f)]
        atol = 1e3 if f in ("SubsurfaceFlow", "CellInFlow") else 2e-5
        err = np.abs(a.astype(np.float64) - b) / (atol + 1e-5 * np.abs(b)) * mask
        if err.max() > 1:
            c = np.unravel_index(np.argmax(err), err.shape)
            print("step", t, f, "err", err.max(), "cell", c, "gpu %.9g gold %.9g orc %.9g" % (a[c], b[c], orc[f][c]))
            for k in ["zi", "SubsurfaceFlow", "UStoreLayerDepth_0", "ExfiltWater", "Transfer", "CellInFlow", "CapFlux", "ActEvapSat", "ActLeakage"]:
                if "out/%d/%s" % (t, k) in g.files:
                    prev = g["out/%d/%s" % (t - 1, k)][c] if t > 0 else (g["state0/" + k][c] if "state0/" + k in g.files else np.nan)
                    print("   %-20s gpu %.9g gold %.9g prev %.9g" % (k, mod.get(k)[c], g["out/%d/%s" % (t, k)][c], prev))
            for k in ["SoilThickness", "thetaS", "thetaR", "KsatVer", "f", "KsatHorFrac", "landSlope", "DW", "DL"]:
                print("   static %-14s %.9g" % (k, g["static/" + k][c]))
