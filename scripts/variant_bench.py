"""Development aid: time both hot kernels on the bench shard for several builds of the library (make dev TAG=x DEVFLAGS=...).
usage: python scripts/variant_bench.py compare [basins] TAG [TAG ...] [-- ENV=VAL ...]   ('-' = the in-tree library)
The shard's inputs are generated once and cached under /tmp for the other builds of the same call."""
import os, subprocess, sys, time, types, pickle
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "compare":
    rest = sys.argv[2:]
    basins = rest.pop(0) if rest and rest[0].isdigit() else "32"
    envs = []
    if "--" in rest:
        k = rest.index("--")
        rest, envs = rest[:k], rest[k + 1:]
    for tag in rest:
        lib = os.path.join(ROOT, "wflow_b200", "libwflowsbm%s.so" % ("" if tag == "-" else "_" + tag))
        env = dict(os.environ, WFLOW_B200_LIB=lib)
        for e in envs:
            env[e.split("=")[0]] = e.split("=")[1]
        out = subprocess.run([sys.executable, os.path.abspath(__file__), basins], env=env, capture_output=True, text=True)
        print("%-8s %s" % (tag, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-600:]), flush=True)
    sys.exit(0)

import bench
from wflow_b200 import core, synth
basins = int(sys.argv[1]) if len(sys.argv) > 1 else 32
args = types.SimpleNamespace(tile=1024, timestepsecs=3600)
nrow, ncol = bench.shard_shape(basins, args.tile)
cache = "/tmp/variant_case_%d.pkl" % basins
if os.path.isfile(cache):
    ldd, river, static, state, lt = pickle.load(open(cache, "rb"))
    sched = None
else:
    ldd, river, sched, static, state, lt = bench.make_model_inputs(args, 0, nrow, ncol, core.Schedule)
    pickle.dump((ldd, river, static, state, lt), open(cache, "wb"), protocol=4)
it_l, it_r = bench.it_substeps(args)
mod = core.SbmModel(ldd, river, schedule=sched, timestepsecs=args.timestepsecs, layer_thickness=lt, model_snow=1,
                    it_kin_l=it_l, it_kin_r=it_r)
for k, v in {**static, **state}.items():
    if k not in ("River", "Beta", "SatWaterDepth"):
        mod.set(k, v)
forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=args.timestepsecs, rain_mean=20.0)
trips = [forc.step(f) for f in range(4)]
vs, ls = [], []
for s in range(5 + 16):
    for name, a in zip(("P", "PET", "TEMP"), trips[s % 4]):
        mod.set(name, a)
    mod.step()
    v_, l_, _t = mod.last_step_ms()
    if s >= 5:
        vs.append(v_); ls.append(l_)
zi = mod.get("zi"); q = mod.get("RiverRunoff")
chk = float(np.nansum(np.where(zi > -999, zi, 0).astype(np.float64))), float(np.nansum(np.where(q > -999, q, 0).astype(np.float64)))
print("plan %s | vertical %.3f ms lateral %.3f ms (min %.3f) sum %.3f | checksum zi %.6f Q %.6f"
      % (mod.lateral_plan.get("mode"), np.median(vs), np.median(ls), min(ls), np.median(vs) + np.median(ls), chk[0], chk[1]))
mod.close()
