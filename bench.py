#!/usr/bin/env python
"""bench.py -- cell-timesteps/s of the full wflow_sbm step (SBM column + subsurface / overland /
river kinematic wave) on synthetic independent D8 basins (BASELINE.json config 3).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the
                                                             # reference, all host threads

Workload.  BASELINE config 3 is a 16384x16384 grid of 256 independent 1024x1024 basins, hourly
steps (modified-Rutter interception, river sub-step 900 s => it_kinR = 4), sharded by basin over
8 GPUs.  Each rank holds `--basins` (default 32 = the config's per-GPU share; 8 ranks = the
full 16k x 16k grid) whole basins: weak scaling, no data-path collective; the only collective
is the gather of the basin-outlet discharges (NCCL all_gather) once per timed region.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_VERTICAL = 236  # algorithmic bytes per cell-step of the vertical kernel, 4 layers, snow on
BYTES_FULL = 333      # full step incl. lateral statics, routing states and topology
# DRAM traffic of one k_vertical launch per cell, from the round's `ncu --set full` capture of the default
# workload (profiles/r1_vertical_ncu_details.txt: dram__bytes_read.sum + dram__bytes_write.sum = 11.75 GB for
# 33 554 431 cells).  Above the algorithmic 236 B: float64 hand-offs, derived statics, register spills.
TRAFFIC_VERTICAL_PER_CELL = 350.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--basins", type=int, default=32, help="1024^2 basins per GPU (32 = config 3's per-GPU share)")
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--timestepsecs", type=int, default=3600)
    ap.add_argument("--cpu-sample-steps", type=int, default=40, help="timesteps of one basin tile timed for cpu_baseline (~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-basin", action="store_true", help="skip the auxiliary single-basin per-step / pipelined measurement")
    return ap.parse_args()


def shard_shape(basins, tile):
    """Arrange `basins` tiles in a near-square block of whole tiles."""
    r = int(np.floor(np.sqrt(basins)))
    while basins % r:
        r -= 1
    return r * tile, (basins // r) * tile


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_model_inputs(args, rank, nrow, ncol, schedule_factory):
    from wflow_b200 import params, synth
    lt = (100, 300, 800)
    ldd = synth.ldd_basin_tiles(nrow, ncol, tile=args.tile, seed=7 + 1000 * rank)
    sched = schedule_factory(ldd)
    up = sched.catchment_total(None)
    river = (up >= 500).astype(np.uint8)  # river = upstream area >= 500 cells
    base = synth.base_parameters(nrow, ncol, seed=42 + rank, n_layers=4)
    upf = up.astype(np.float32)
    base["RiverWidth"] = np.where(river != 0, np.float32(2.0) * np.power(np.maximum(upf, 1), np.float32(0.35)), 0).astype(np.float32)
    base["BankfullDepth"] = np.where(river != 0, np.float32(1.0), np.float32(0.0)).astype(np.float32)
    static = params.derive_static(base, ldd, river, args.timestepsecs, cellsize=1000.0, n_layers=4)
    state = synth.random_state(static, 4, lt, seed=99 + rank, smooth=48)
    return ldd, river, sched, static, state, lt


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def it_substeps(args):
    # kinwaveIters=1, kinwaveRiverTstep=900, kinwaveLandTstep=3600
    it_r = int(np.ceil(args.timestepsecs / 900.0))
    it_l = int(np.ceil(args.timestepsecs / 3600.0))
    return max(it_l, 1), max(it_r, 1)


def cpu_baseline(args, steps, threads, warmup=1):
    """The oracle port of the reference on ONE basin tile of the same workload (bounded sample)."""
    from oracle import oracle as O
    from wflow_b200 import core, synth
    n = args.tile
    ldd, river, sched, static, state, lt = make_model_inputs(args, 0, n, n, core.Schedule)
    it_l, it_r = it_substeps(args)
    orc = O.Oracle(ldd, river, timestepsecs=args.timestepsecs, layer_thickness=lt, modelSnow=1, it_kinL=it_l, it_kinR=it_r)
    for k, v in {**static, **state}.items():
        orc.set(k, v)
    used = O.set_threads(threads)
    forc = synth.Forcing(n, n, seed=5, timestepsecs=args.timestepsecs, rain_mean=20.0)
    P, PET, T = forc.step(0)
    orc.set("P", P); orc.set("PET", PET); orc.set("TEMP", T)
    for _ in range(max(1, warmup)):
        orc.step()  # warm-up (page faults, thread pool)
    t0 = time.time()
    for s in range(steps):
        orc.step()
    dt = time.time() - t0
    cells = int(orc.net.flat_nodes.size)
    return {"value": cells * steps / dt, "unit": "cell-timesteps/s", "cores": used, "kind": "port",
            "sample": "%d steps of one %dx%d basin tile (%d cells) of the same workload, C oracle port of the "
                      "reference (oracle/wflow_sbm_oracle.c), OpenMP over the cells of a level" % (steps, n, n, cells),
            "ms_per_step": dt / steps * 1e3, "cells": cells}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # every step of this arm is one timestep of ONE basin tile (1/32 of the GPU arm's per-GPU shard): K = 20 takes ~5 s
    steps = max(1, min(args.steps, 200))
    warm = max(1, min(args.warmup, 5))
    cb = cpu_baseline(args, steps, threads, warm)
    it_l, it_r = it_substeps(args)
    line = {
        "impl": "reference", "metric": "cell-timesteps/sec (SBM+kinwave)", "value": cb["value"],
        "unit": "cell-timesteps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (column, kinematic waves) + f32 (map-algebra phases)", "data": "synthetic",
        "config": workload_config(args, it_l, it_r, sample=cb["sample"]),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, it_l, it_r, **extra):
    nrow, ncol = shard_shape(args.basins, args.tile)
    cfg = {"workload": "BASELINE config 3: synthetic D8 grid of independent %dx%d basins, full SBM + subsurface/overland/"
                       "river kinematic wave, %d s steps, 4 soil layers, snow on; %d basins (%dx%d cells) per GPU "
                       "(8 GPUs = the 16384x16384 grid)" % (args.tile, args.tile, args.timestepsecs, args.basins, nrow, ncol),
           "basins_per_gpu": args.basins, "tile": args.tile, "timestepsecs": args.timestepsecs,
           "it_kinL": it_l, "it_kinR": it_r, "layers": 4, "snow": 1,
           "state": "spatially correlated random initial state (synth.random_state, 48-cell correlation length: water "
                    "tables from surface to bottom, partly filled layers, snow, flowing rivers) + warm-up steps",
           "l2": "inputs larger than L2 (per-step working set >> 126 MB), no flush needed"}
    cfg.update(extra)
    return cfg


def single_basin_pipelined(args, local, steps_per_launch=50):
    """Auxiliary measurement (not the headline): ONE basin tile of the workload -- the regime of a typical
    single-catchment model, where a per-step sweep leaves most of the GPU idle -- stepped one timestep per
    launch and `steps_per_launch` timesteps per launch (wf_step_pipelined: step s+1 trails step s down the
    network inside one kernel; bit-identical results, tests/test_gpu_pipeline.py)."""
    import torch
    from wflow_b200 import core, synth
    n = args.tile
    ldd, river, sched, static, state, lt = make_model_inputs(args, 0, n, n, core.Schedule)
    it_l, it_r = it_substeps(args)
    mod = core.SbmModel(ldd, river, schedule=sched, timestepsecs=args.timestepsecs, layer_thickness=lt, model_snow=1,
                        it_kin_l=it_l, it_kin_r=it_r, device=local)
    for k, v in {**static, **state}.items():
        if k not in ("River", "Beta", "SatWaterDepth"):
            mod.set(k, v)
    forc = synth.Forcing(n, n, seed=5, timestepsecs=args.timestepsecs, rain_mean=20.0)
    trips = [forc.step(f) for f in range(4)]
    K = steps_per_launch
    mod.pipeline_reserve(K)
    for s in range(K):
        for name, a in zip(("P", "PET", "TEMP"), trips[s % 4]):
            mod.pipeline_set_forcing(name, s, a)
    for name, a in zip(("P", "PET", "TEMP"), trips[0]):
        mod.set(name, a)
    stream = torch.cuda.ExternalStream(mod.stream)
    mod.step(3)
    mod.step_pipelined(K)
    mod.synchronize()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        mod.synchronize()
        return e0.elapsed_time(e1)
    ms_step = timed(lambda: mod.step(10)) / 10
    ms_pipe = timed(lambda: mod.step_pipelined(K)) / K
    cells = mod.num_cells
    out = {"workload": "one %dx%d basin of the same workload (%d cells, %d levels)" % (n, n, cells, mod.num_levels),
           "per_step": {"ms_per_step": ms_step, "value": cells / ms_step * 1e3},
           "pipelined": {"steps_per_launch": K, "ms_per_step": ms_pipe, "value": cells / ms_pipe * 1e3},
           "unit": "cell-timesteps/s"}
    mod.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from wflow_b200 import core, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nrow, ncol = shard_shape(args.basins, args.tile)
    it_l, it_r = it_substeps(args)

    t0 = time.time()
    ldd, river, sched, static, state, lt = make_model_inputs(args, rank, nrow, ncol, core.Schedule)
    t_gen = time.time() - t0
    mod = core.SbmModel(ldd, river, schedule=sched, timestepsecs=args.timestepsecs, layer_thickness=lt, model_snow=1,
                        it_kin_l=it_l, it_kin_r=it_r, device=local)
    for k, v in {**static, **state}.items():
        if k not in ("River", "Beta", "SatWaterDepth"):
            mod.set(k, v)
    del static, state
    ncell = mod.num_cells
    # gauges: the basin outlets (pits)
    gauges = np.flatnonzero(ldd.ravel() == 5).astype(np.int64)
    t_init = time.time() - t0

    # forcing: a ring of F distinct steps; host copies pinned (e2e), device copies in network order (value)
    F = 4
    forc = synth.Forcing(nrow, ncol, seed=5 + rank, timestepsecs=args.timestepsecs, rain_mean=20.0)
    host_ring, dev_ring = [], []
    for f in range(F):
        trip = forc.step(f)
        pinned = [torch.from_numpy(a).pin_memory() for a in trip]
        host_ring.append(pinned)
        devs = []
        for name, a in zip(("P", "PET", "TEMP"), trip):
            mod.set(name, a)
            ptr, n = mod.device_ptr(name)
            mod.synchronize()
            # clone the model's network-order array into our ring slot (plumbing only)
            t = torch.as_tensor(_DevArray(ptr, n), device="cuda").clone()
            torch.cuda.synchronize()
            devs.append(t)
        dev_ring.append(devs)
    mod.synchronize()
    stream = torch.cuda.ExternalStream(mod.stream)

    def bind(f):
        for name, t in zip(("P", "PET", "TEMP"), dev_ring[f % F]):
            mod.bind(name, t.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also relaxes the random initial state) ----
    for s in range(max(args.warmup, 3)):
        bind(s)
        mod.step()
    mod.synchronize()

    # ---- value: device-resident forcing ----
    barrier()
    mod.timing_reset()
    l0 = mod.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        for s in range(args.steps):
            bind(s)
            mod.step()
        e1.record(stream)
        mod.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = mod.launch_count - l0
    (ms_v, ms_l, ms_tot), nst = mod.timing_get()
    for name in ("P", "PET", "TEMP"):
        mod.bind(name, None)

    # ---- e2e: host forcing in pinned memory -> set_field -> step -> gauge read-back, every step ----
    # Public API only.  Nothing blocks the host inside the loop: the uploads run on the library's copy
    # stream through its staging ring (so step s+1's forcing crosses PCIe while step s computes) and the
    # gauge values of every step land in pinned host memory; one synchronize closes the timed region.
    gh = mod.gauges_create(gauges)
    gq_all = torch.zeros((args.steps, max(int(gauges.size), 1)), dtype=torch.float32).pin_memory()
    mod.gauges_sample(gh, "RiverRunoff", gq_all[0].data_ptr(), wait=True)  # warm the path
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for s in range(args.steps):
        for name, t in zip(("P", "PET", "TEMP"), host_ring[s % F]):
            mod.set_pinned(name, t.data_ptr())
        mod.step()
        mod.gauges_sample(gh, "RiverRunoff", gq_all[s].data_ptr(), wait=False)
    e3.record(stream)
    mod.synchronize()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    gq = gq_all[args.steps - 1].numpy()[: int(gauges.size)].copy()
    if not np.all(np.isfinite(gq_all.numpy())):
        raise SystemExit("bench.py: non-finite outlet discharge in the e2e run")
    h2d = 3 * nrow * ncol * 4
    d2h = int(gauges.size) * 4

    # max over ranks; gather gauge series (the only collective of the path)
    tms = torch.tensor([ms, ms_e2e, ms_v, ms_l], device="cuda", dtype=torch.float64)
    cells = torch.tensor([float(ncell)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
        from wflow_b200 import parallel
        gathered = parallel.gather_series(gq)  # the path's only collective: outlet discharges of all ranks
        assert len(gathered) == world
    ms, ms_e2e, ms_v, ms_l = [float(x) for x in tms.tolist()]
    total_cells = float(cells.item())

    if rank == 0:
        peak, peak_src = peaks()
        value = total_cells * args.steps / (ms / 1e3)
        e2e_v = total_cells * args.steps / (ms_e2e / 1e3)
        v_launch_ms = ms_v / max(nst, 1)
        achieved = ncell * BYTES_VERTICAL / (v_launch_ms / 1e3) / 1e9
        line = {
            "metric": "cell-timesteps/sec (SBM+kinwave)", "value": value, "unit": "cell-timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (column, kinematic waves) + f32 (map-algebra phases, storage)", "data": "synthetic",
            "config": workload_config(args, it_l, it_r, cells_per_gpu=ncell, levels=mod.num_levels,
                                      river_cells=mod.num_river_cells, init_s=round(t_init, 1), gen_s=round(t_gen, 1)),
            "e2e": {"value": e2e_v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "path": "SbmModel.set_pinned(P,PET,TEMP) from pinned host rasters (staging ring, copy stream) -> step() -> gauges_sample(RiverRunoff at the basin outlets) into pinned host memory, every step; one synchronize at the end of the timed region"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_vertical<4> (SBM vertical column)", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "bytes_per_cell": BYTES_VERTICAL, "cells_per_launch": ncell, "ms_per_launch": v_launch_ms,
                         "traffic": TRAFFIC_VERTICAL_PER_CELL * ncell,
                         "traffic_source": "ncu --set full, profiles/r1_vertical_ncu_details.txt (per cell x cells of this launch)",
                         "bound_note": "the kernel is instruction-issue bound (ncu: issue slots busy 66 %, DRAM 31 %), "
                                       "not HBM-bound; the HBM fraction is reported as the contract asks",
                         "full_step_gbs_at_333B": ncell * BYTES_FULL / ((ms_v + ms_l) / max(nst, 1) / 1e3) / 1e9},
            "kernels_ms_per_step": {"vertical": ms_v / max(nst, 1), "lateral": ms_l / max(nst, 1),
                                    "lateral_us_per_level": ms_l / max(nst, 1) * 1e3 / max(mod.num_levels, 1)},
            "lateral_plan": mod.lateral_plan,
            "clocks": clk.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_baseline(args, args.cpu_sample_steps, os.cpu_count() or 1)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:  # the checker must never take the product number down with it
                line["cpu_baseline"] = {"error": repr(ex)}
        if world == 1 and not args.no_single_basin:
            try:
                line["single_basin"] = single_basin_pipelined(args, local)
            except Exception as ex:  # auxiliary: never takes the headline down
                line["single_basin"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
    mod.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
