"""wflow_hbv measured like the headline: one JSON line for the HBV step (column kernel + channel sweep) on a shard of
independent basins, device-resident forcing (`value`) and host rasters through HbvModel.dynamic() (`e2e`), with the CPU
baselines beside it (the float32-numpy restatement, and the reference's own dynamic() where baseline/_ref holds it).

usage: python scripts/hbv_bench.py [--basins 32] [--tile 1024] [--timestepsecs 3600] [--steps 20] [--warmup 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

# Algorithmic bytes per cell and timestep (float32, every array touched once), wflow_hbv.py:1222-1767:
# column: forcing 3 (P PET TEMP) + parameters 23 (Pcorr TempCor TTI TT SFCF RFCF ICF EPF ECORR CEVPF Cfmax CFR WHC FieldCapacity
# Treshold BetaSeepage PERC Cflux KQuickFlow AlphaNL K4 xl yl) + states in 6 + out 6 + SurfaceRunoff 1 + lateral inflow out 1 = 40 words
COLUMN_BYTES = 4 * 40
# channel, per sub-step: Qold, q, Alpha, DCL, Bw in (5), one gathered upstream outflow (1), Qnew out (1) = 7 words + 5 B topology;
# the step's means (SurfaceRunoff, WaterLevel, WaterLevelsub) out once = 3 words
def channel_bytes(it):
    return it * (4 * 7 + 5) + 4 * 3


def main():
    import torch
    from wflow_b200 import hbv, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--basins", type=int, default=32)
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--timestepsecs", type=int, default=3600)
    ap.add_argument("--river-tstep", type=int, default=900)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cpu-tile", type=int, default=256)
    ap.add_argument("--cpu-steps", type=int, default=5)
    args = ap.parse_args()
    nrow, ncol = bench.shard_shape(args.basins, args.tile)
    dt = args.timestepsecs
    t0 = time.time()
    ldd, static, state = synth.hbv_case(nrow, ncol, kind="tiles", tile=args.tile, seed=11, timestepsecs=dt, block=64)
    gen_s = time.time() - t0
    mod = hbv.HbvModel(ldd, timestepsecs=dt, kinwaveIters=1, kinwaveTstep=args.river_tstep)
    mod.set_many({k: v for k, v in static.items() if k in hbv.PARAMETERS})
    mod.set_many(state)
    ncell = mod.num_cells
    forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=dt, rain_mean=14.0, t_mean=3.0)
    trips = [forc.step(f) for f in range(4)]
    pinned = [[torch.from_numpy(a).pin_memory() for a in tr] for tr in trips]
    stream = torch.cuda.ExternalStream(mod.stream)
    # ---- device-resident forcing: the forcing of step 0 stays bound, K timed steps ----
    mod.dynamic(*trips[0])
    for _ in range(args.warmup):
        mod.step()
    mod.synchronize()
    mod.timing_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with bench.ClockSampler(0) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            mod.step()
        e1.record(stream)
        mod.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    (col_ms, sweep_ms, _), nst = mod.timing_get()
    col_ms, sweep_ms = col_ms / max(nst, 1), sweep_ms / max(nst, 1)
    # ---- end to end: pinned host rasters -> device -> step -> outlet discharge back, every step ----
    pits = np.flatnonzero(ldd.ravel() == 5)
    g = mod.gauges_create(pits)
    out = torch.zeros(len(pits), dtype=torch.float32).pin_memory()
    for w in range(3):
        for name, t in zip(("P", "PET", "TEMP"), pinned[w % 4]):
            mod.set_pinned(name, t.data_ptr())
        mod.step()
    mod.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for s in range(args.steps):
        for name, t in zip(("P", "PET", "TEMP"), pinned[s % 4]):
            mod.set_pinned(name, t.data_ptr())
        mod.step()
        mod.gauges_sample(g, "SurfaceRunoff", out.data_ptr(), wait=False)
    e1.record(stream)
    mod.synchronize()
    e2e_ms = e0.elapsed_time(e1) / args.steps
    peak, peak_src = bench.peaks()
    it = mod.it_kin
    line = {
        "metric": "cell_timesteps_per_second", "value": ncell / ms * 1e3, "unit": "cell-steps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 column, f64 channel solve", "data": "synthetic",
        "config": {"workload": "wflow_hbv, %d basins of %d^2 (%d cells), timestep %d s, %d channel sub-steps; inputs %.1f x L2"
                               % (args.basins, args.tile, ncell, dt, it, ncell * COLUMN_BYTES / 126e6),
                   "lateral_plan": mod.lateral_plan},
        "e2e": {"value": ncell / e2e_ms * 1e3, "unit": "cell-steps/s", "h2d_bytes_per_step": int(3 * 4 * nrow * ncol),
                "d2h_bytes_per_step": int(4 * len(pits))},
        "gpu_launches": 2 * args.steps,
        "kernels_ms_per_step": {"k_hbv_column": col_ms, "k_channel_wave": sweep_ms},
        "roofline": {"bound": "hbm", "kernel": "k_channel_wave (kin_wave on the whole network)" if sweep_ms > col_ms else "k_hbv_column",
                     "achieved": (channel_bytes(it) if sweep_ms > col_ms else COLUMN_BYTES) * ncell / (max(sweep_ms, col_ms) * 1e-3) / 1e9,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": None,
                     "kernels": [
                         {"kernel": "k_hbv_column", "ms": col_ms, "algorithmic_bytes_per_cell": COLUMN_BYTES,
                          "achieved": COLUMN_BYTES * ncell / (col_ms * 1e-3) / 1e9, "frac": COLUMN_BYTES * ncell / (col_ms * 1e-3) / 1e9 / peak},
                         {"kernel": "k_channel_wave (kin_wave on the whole network)", "ms": sweep_ms, "algorithmic_bytes_per_cell": channel_bytes(it),
                          "achieved": channel_bytes(it) * ncell / (sweep_ms * 1e-3) / 1e9,
                          "frac": channel_bytes(it) * ncell / (sweep_ms * 1e-3) / 1e9 / peak}]},
        "clocks": clk.summary(), "gen_s": gen_s,
    }
    line["roofline"]["frac"] = line["roofline"]["achieved"] / peak
    mod.close()
    if args.cpu_steps <= 0:  # (profiling runs: the device part only)
        print(json.dumps(line))
        return
    # ---- CPU beside it: one basin tile, one core ----
    n = args.cpu_tile
    ldd_c, st_c, state_c = synth.hbv_case(n, n, kind="tiles", tile=n, seed=11, timestepsecs=dt, block=64)
    fields = dict(st_c)
    fields.update(state_c)
    forc_c = synth.Forcing(n, n, seed=5, timestepsecs=dt, rain_mean=14.0, t_mean=3.0)
    tr = [forc_c.step(f) for f in range(2)]
    from oracle import hbv_oracle
    from oracle import oracle as orc_mod
    orc_mod.set_threads(1)
    o = hbv_oracle.HbvOracle(ldd_c, fields, timestepsecs=dt, kinwaveIters=1, kinwaveTstep=args.river_tstep)
    o.step(*tr[0])
    t0 = time.perf_counter()
    for s in range(args.cpu_steps):
        o.step(*tr[s % 2])
    port_s = (time.perf_counter() - t0) / args.cpu_steps
    line["cpu_baseline"] = {"value": int(o.active.sum()) / port_s, "unit": "cell-steps/s", "cores": 1, "kind": "port",
                            "sample": "%d steps of one %d^2 basin, float32-numpy column + C kin_wave (oracle/hbv_oracle.py)" % (args.cpu_steps, n)}
    try:
        from oracle import ref_hbv, refload
        if refload.available() and os.path.isfile(os.path.join(refload.REFERENCE_ROOT, "wflow", "wflow_hbv.py")):
            r = ref_hbv.RefHbv(ldd_c, fields, timestepsecs=dt, kinwaveIters=1, kinwaveTstep=args.river_tstep)
            r.step(*tr[0])
            t0 = time.perf_counter()
            for s in range(args.cpu_steps):
                r.step(*tr[s % 2])
            ref_s = (time.perf_counter() - t0) / args.cpu_steps
            line["cpu_baseline_reference"] = {"value": int(o.active.sum()) / ref_s, "unit": "cell-steps/s", "cores": 1, "kind": "reference",
                                              "sample": "%d steps of one %d^2 basin, the reference's unmodified wflow_hbv dynamic() "
                                                        "(numba kin_wave; map algebra on the numpy stand-in for PCRaster)" % (args.cpu_steps, n)}
    except Exception as e:  # the reference is optional here
        line["cpu_baseline_reference"] = {"unavailable": str(e)[:200]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
