"""Quick device timing of the two kernels on synthetic cases (development aid)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wflow_b200 import synth
from wflow_b200.core import SbmModel


def run(n, kind, lt=(100, 300, 800), dt=86400, itL=1, itR=1, steps=20, tile=1024):
    t0 = time.time()
    ldd, river, static, state = synth.make_case(n, n, kind=kind, tile=tile, timestepsecs=dt, layer_thickness=lt)
    t1 = time.time()
    mod = SbmModel(ldd, river, timestepsecs=dt, layer_thickness=lt, it_kin_l=itL, it_kin_r=itR)
    t2 = time.time()
    for k, v in {**static, **state}.items():
        if k not in ("River", "Beta", "SatWaterDepth"):
            mod.set(k, v)
    forc = synth.Forcing(n, n, timestepsecs=dt)
    P, PET, T = forc.step(0)
    mod.set("P", P); mod.set("PET", PET); mod.set("TEMP", T)
    t3 = time.time()
    for w in range(3):
        mod.step()
    mod.synchronize()
    vs, ls = [], []
    tt = time.time()
    for s in range(steps):
        if s % 5 == 0:
            P, PET, T = forc.step(s // 5 + 1)
            mod.set("P", P)
        mod.step()
        v, l, tot = mod.last_step_ms()
        vs.append(v); ls.append(l)
    mod.synchronize()
    nc = mod.num_cells
    v, l = np.median(vs), np.median(ls)
    print("%s n=%d cells=%d levels=%d river=%d | gen %.1fs build %.1fs upload %.1fs | vertical %.3f ms (%.2f Gcell/s, %.0f GB/s @236B) lateral %.3f ms (%.2f us/level) | step %.3f ms = %.2f Gcell-steps/s" % (
        kind, n, nc, mod.num_levels, mod.num_river_cells, t1 - t0, t2 - t1, t3 - t2, v, nc / v / 1e6, nc * 236 / v / 1e6, l,
        l * 1e3 / mod.num_levels, v + l, nc / (v + l) / 1e6), flush=True)
    mod.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run(1024, "pits"); run(1024, "tiles", tile=256)
    elif which == "mid":
        run(4096, "pits"); run(2048, "tiles"); run(4096, "tiles", dt=3600, itR=4)
