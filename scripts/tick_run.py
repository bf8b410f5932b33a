"""Development aid: run a few steps of the bench shard with the profiling build and write the per-tick csv.
usage: WFLOW_B200_LIB=wflow_b200/libwflowsbm_prof.so WF_TICK_PROFILE=gpurun_out/ticks.csv python scripts/tick_run.py [basins]"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from wflow_b200 import core, synth
basins = int(sys.argv[1]) if len(sys.argv) > 1 else 32
args = types.SimpleNamespace(tile=1024, timestepsecs=3600)
nrow, ncol = bench.shard_shape(basins, args.tile)
ldd, river, sched, static, state, lt = bench.make_model_inputs(args, 0, nrow, ncol, core.Schedule)
it_l, it_r = bench.it_substeps(args)
mod = core.SbmModel(ldd, river, schedule=sched, timestepsecs=args.timestepsecs, layer_thickness=lt, model_snow=1, it_kin_l=it_l, it_kin_r=it_r)
for k, v in {**static, **state}.items():
    if k not in ("River", "Beta", "SatWaterDepth"):
        mod.set(k, v)
forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=args.timestepsecs, rain_mean=20.0)
P, PET, T = forc.step(0)
mod.set("P", P); mod.set("PET", PET); mod.set("TEMP", T)
for _ in range(6):
    mod.step()
print(mod.lateral_plan, mod.last_step_ms())
mod.close()
