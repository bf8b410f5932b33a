"""Development aid: time the lateral sweep on the bench shard (32 basins of 1024^2) under run-time switches of the
library (environment variables read at every launch), one model build for all of them.
usage: python scripts/lat_bench.py [basins] VAR=a,b [VAR2=c,d ...]   e.g.  WF_LAT_INTERLEAVE=0,1"""
import itertools, os, sys, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from wflow_b200 import core, synth

argv = sys.argv[1:]
basins = int(argv.pop(0)) if argv and argv[0].isdigit() else 32
switches = [(a.split("=")[0], a.split("=")[1].split(",")) for a in argv]
args = types.SimpleNamespace(tile=1024, timestepsecs=3600)
nrow, ncol = bench.shard_shape(basins, args.tile)
t0 = time.time()
ldd, river, sched, static, state, lt = bench.make_model_inputs(args, 0, nrow, ncol, core.Schedule)
it_l, it_r = bench.it_substeps(args)
mod = core.SbmModel(ldd, river, schedule=sched, timestepsecs=args.timestepsecs, layer_thickness=lt, model_snow=1,
                    it_kin_l=it_l, it_kin_r=it_r)
for k, v in {**static, **state}.items():
    if k not in ("River", "Beta", "SatWaterDepth"):
        mod.set(k, v)
forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=args.timestepsecs, rain_mean=20.0)
P, PET, T = forc.step(0)
mod.set("P", P); mod.set("PET", PET); mod.set("TEMP", T)
print("built in %.0f s: %d cells, plan %s" % (time.time() - t0, mod.num_cells, mod.lateral_plan), flush=True)
for _ in range(5):
    mod.step()
for rep in range(2):
    for combo in itertools.product(*[vals for _, vals in switches]):
        for (name, _), v in zip(switches, combo):
            os.environ[name] = v
        vs, ls = [], []
        for s in range(8):
            mod.step()
            v_, l_, _t = mod.last_step_ms()
            vs.append(v_); ls.append(l_)
        print(" ".join("%s=%s" % (n, v) for (n, _), v in zip(switches, combo)),
              "| vertical %.3f ms lateral %.3f ms (min %.3f) = %.2f us/level" % (np.median(vs), np.median(ls), min(ls), np.median(ls) * 1e3 / mod.num_levels), flush=True)
chk = float(np.nansum(np.where(mod.get("zi") > -999, mod.get("zi"), 0).astype(np.float64)))
print("checksum zi %.6f" % chk)
