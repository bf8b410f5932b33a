"""Development aid: run the bench workload a few steps and print state statistics."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from wflow_b200 import core, synth

class A: pass
a = A(); a.tile = 1024; a.timestepsecs = 3600; a.basins = 2
nrow, ncol = bench.shard_shape(a.basins, a.tile)
ldd, river, sched, static, state, lt = bench.make_model_inputs(a, 0, nrow, ncol, core.Schedule)
mod = core.SbmModel(ldd, river, schedule=sched, timestepsecs=3600, layer_thickness=lt, model_snow=1, it_kin_l=1, it_kin_r=4)
for k, v in {**static, **state}.items():
    if k not in ("River", "Beta", "SatWaterDepth"):
        mod.set(k, v)
forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=3600, rain_mean=20.0)
mask = ldd != 0
for s in range(8):
    P, PET, T = forc.step(s % 4)
    mod.set("P", P); mod.set("PET", PET); mod.set("TEMP", T)
    mod.step()
    out = []
    for k in range(4):
        u = mod.get("UStoreLayerDepth_%d" % k)[mask]
        out.append("U%d min %.3g neg %d nan %d" % (k, np.nanmin(u), (u < 0).sum(), np.isnan(u).sum()))
    zi = mod.get("zi")[mask]
    print(s, " | ".join(out), "| zi nan", np.isnan(zi).sum(), "P mean %.3f" % P.mean(), flush=True)
