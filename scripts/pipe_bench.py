"""Development aid: per-step stepping vs the pipelined multi-step sweep on the bench workload, per sweep mode.
usage: python scripts/pipe_bench.py [--basins 32] [--modes default,grid] [--ks 10,20]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from wflow_b200 import core, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--basins", type=int, default=32)
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--timestepsecs", type=int, default=3600)
    ap.add_argument("--modes", default="default,grid")
    ap.add_argument("--ks", default="10,20")
    args = ap.parse_args()
    nrow, ncol = bench.shard_shape(args.basins, args.tile)
    it_l, it_r = bench.it_substeps(args)
    t0 = time.time()
    ldd, river, sched0, static, state, lt = bench.make_model_inputs(args, 0, nrow, ncol, core.Schedule)
    print("gen %.1fs" % (time.time() - t0), flush=True)
    forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=args.timestepsecs, rain_mean=20.0)
    trips = [forc.step(f) for f in range(4)]
    ks = [int(k) for k in args.ks.split(",")]
    for mode in args.modes.split(","):
        for k in ("WF_LATERAL_MODE", "WF_CLUSTER_SIZE", "WF_CLUSTER_WIDE", "WF_GROUPS", "WF_PIPE_OVERLAP", "WF_AHEAD_CLUSTERS"):
            os.environ.pop(k, None)
        for kv in mode.split("+"):
            if kv == "default":
                continue
            if kv.startswith("lib:"):  # another build of the library (development variants)
                from wflow_b200 import _lib
                _lib.LIB_PATH = os.path.join(ROOT, "wflow_b200", kv[4:])
                _lib._lib = None
                continue
            if "=" in kv:
                a, b = kv.split("=")
                os.environ[a] = b
            else:
                os.environ["WF_LATERAL_MODE"] = kv
        t0 = time.time()
        mod = core.SbmModel(ldd, river, timestepsecs=args.timestepsecs, layer_thickness=lt, model_snow=1,
                            it_kin_l=it_l, it_kin_r=it_r, device=0)
        for k, v in {**static, **state}.items():
            if k not in ("River", "Beta", "SatWaterDepth"):
                mod.set(k, v)
        print("mode %s: plan %s, init %.1fs" % (mode, mod.lateral_plan, time.time() - t0), flush=True)
        kmax = max(ks)
        mod.pipeline_reserve(kmax)
        for s in range(kmax):
            for name, a in zip(("P", "PET", "TEMP"), trips[s % 4]):
                mod.pipeline_set_forcing(name, s, a)
        stream = torch.cuda.ExternalStream(mod.stream)
        for s in range(3):
            for name, a in zip(("P", "PET", "TEMP"), trips[s % 4]):
                mod.set(name, a)
            mod.step()
        mod.synchronize()
        ncell = mod.num_cells
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mod.timing_reset()
        e0.record(stream)
        mod.step(10)
        e1.record(stream)
        mod.synchronize()
        ms = e0.elapsed_time(e1) / 10
        (ms_v, ms_l, _), nst = mod.timing_get()
        print("  per-step: %.3f ms/step  %.3f G cell-steps/s  (column %.3f ms, sweep %.3f ms)"
              % (ms, ncell / ms / 1e6, ms_v / max(nst, 1), ms_l / max(nst, 1)), flush=True)
        for K in ks:
            mod.step_pipelined(K)  # warm
            mod.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            mod.step_pipelined(K)
            e1.record(stream)
            mod.synchronize()
            ms = e0.elapsed_time(e1) / K
            print("  multi-step launch (%s) K=%d: %.3f ms/step  %.3f G cell-steps/s" % (mod.pipeline_mode, K, ms, ncell / ms / 1e6), flush=True)
        mod.close()


if __name__ == "__main__":
    main()
