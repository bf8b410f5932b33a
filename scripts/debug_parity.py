import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common
from wflow_b200 import synth
np.set_printoptions(precision=9, linewidth=200)
kw = dict(kind="tiles", tile=32, layer_thickness=(100, 300, 800), dt=3600, itL=1, itR=4, river_area=20)
N = 64
orc, mod, ldd, river = common.make_pair(N, N, **kw)
mask = common.network_mask(orc)
forc = synth.Forcing(N, N, seed=5, timestepsecs=3600, rain_mean=20.0)
names = common.state_names(orc.maxLayers, 1) + common.DIAGS
orc.want(*common.DIAGS); mod.request(*common.DIAGS)
for t in range(4):
    P, PET, T = forc.step(t)
    for m in (orc, mod):
        m.set("P", P); m.set("PET", PET); m.set("TEMP", T)
    pre = {k: orc[k].copy() for k in common.state_names(orc.maxLayers, 1)}
    for k in pre: mod.set(k, orc[k])
    orc.step(); mod.step()
    res = []
    for k in names:
        a = mod.get(k); b = orc[k]
        atol = 1e3 if k in ("SubsurfaceFlow", "CellInFlow") else 2e-5
        err = np.abs(a.astype(np.float64) - b) / (atol + 1e-5 * np.abs(b)) * mask
        res.append((float(err.max()), k, np.unravel_index(np.argmax(err), err.shape), int((err > 1).sum())))
    res.sort(reverse=True)
    print("step", t, [(round(e, 2), k, tuple(int(x) for x in c), n) for e, k, c, n in res[:6]])
    if res[0][0] > 1:
        # earliest-phase failing field: print cell details
        bad_cells = {}
        for e, k, c, n in res:
            if e > 1:
                bad_cells.setdefault(c, []).append(k)
        c = res[0][2]
        print("cell", c, "ldd", ldd[c], "river", river[c])
        for k in names:
            print("  %-28s gpu %-16.9g orc %-16.9g pre %s" % (k, mod.get(k)[c], orc[k][c], pre.get(k, [[None]])[c] if k in pre else ""))
        for k in ["SoilThickness", "thetaS", "thetaR", "KsatVer", "f", "RootingDepth", "PathFrac", "InfiltCapSoil", "InfiltCapPath", "c_0", "c_1", "KsatVerFrac_0", "CapScale", "MaxLeakage", "Cmax", "CanopyGapFraction", "EoverR", "WaterFrac", "RiverFrac", "DW", "DL", "SW", "landSlope"]:
            print("  static %-20s %.9g" % (k, orc[k][c]))
        break
