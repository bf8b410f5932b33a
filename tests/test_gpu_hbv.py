"""wflow_hbv on the device (wf_hbv_create: the HBV column of wflow_b200/csrc/hbv.cuh + the river task of the level sweep on
the whole network) against trajectories of the reference's own, unmodified ``wflow_hbv.WflowModel.dynamic()``
(tests/golden/hbv_*.npz, written by oracle/gen_golden_hbv.py through the float32 PCRaster stand-in).

Kernel parity per step: every step starts from the reference's float32 states of the step before; every state and every
requested output of every cell must agree within rel 1e-5 (absolute floor 2e-5 of the field's unit).  The free run starts
once and is compared after the last step, with the CPU restatement (oracle/hbv_oracle.py, bit-identical to the goldens)
beside it on a larger grid."""
import ast
import glob
import os

import numpy as np
import pytest

from tests import common
from wflow_b200 import hbv, synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLD, "hbv_*.npz")))
ATOL = {"SurfaceRunoff": 2e-5, "SurfaceRunoffsub": 2e-5}


def _load(name):
    g = np.load(os.path.join(GOLD, "hbv_%s.npz" % name))
    kw = ast.literal_eval(str(g["meta"]))
    mod = hbv.HbvModel(g["ldd"], timestepsecs=kw.get("dt", 86400), kinwaveIters=kw.get("kinwaveIters", 0),
                       kinwaveTstep=kw.get("kinwaveTstep", 0), set_kquick_flow=kw.get("setKquickFlow", 0),
                       external_qbase=kw.get("externalQbase", 0), glacier=kw.get("glacier", 0), mass_wasting=kw.get("massWasting", 0))
    known = set(hbv.PARAMETERS)
    mod.set_many({k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("static/") and k.split("/", 1)[1] in known})
    mod.set_many({k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("state0/")})
    return g, kw, mod


def _forcing(g, t):
    fk = lambda k: g["forcing/%d/%s" % (t, k)] if "forcing/%d/%s" % (t, k) in g.files else None  # noqa: E731
    return fk("P"), fk("PET"), fk("TEMP"), fk("Inflow"), fk("Seepage")


def test_goldens_present():
    assert len(CASES) >= 6


@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_hbv_dynamic(name):
    g, kw, mod = _load(name)
    try:
        mask = np.zeros(mod.shape, bool)
        mask.flat[mod.network()[0]] = True
        assert mod.num_river_cells == mod.num_cells  # every cell carries a channel
        states = [s for s in hbv.STATES + ["GlacierStore"] if "out/0/" + s in g.files]
        outs = [o for o in hbv.OUTPUTS if "out/0/" + o in g.files]
        mod.request(*outs)
        worst = (0.0, "")
        for t in range(kw["steps"]):
            if t > 0:  # the reference's states of the step before, bit for bit
                for s in states:
                    mod.set(s, g["out/%d/%s" % (t - 1, s)])
            P, PET, T, Inflow, Seep = _forcing(g, t)
            mod.dynamic(P, PET, T, Inflow=Inflow, Seepage=Seep)
            if "it/%d" % t in g.files:
                assert mod.it_kin == int(g["it/%d" % t]), (name, t)
            for f in states + outs:
                e = common.compare(mod.get(f), g["out/%d/%s" % (t, f)], mask, 1e-5, ATOL.get(f, 2e-5))
                if e > worst[0]:
                    worst = (e, "%s step %d" % (f, t))
                assert e <= 1.0, (name, t, f, e)
        print("%s: %d fields x %d steps, worst error/tolerance %.3g (%s)" % (name, len(states + outs), kw["steps"], worst[0], worst[1]))
    finally:
        mod.close()


@pytest.mark.parametrize("name", ["tiles_daily", "tiles_hourly_sub"])
def test_free_run_ends_where_the_reference_does(name):
    """No re-seeding between steps: the device model runs the whole trajectory by itself."""
    g, kw, mod = _load(name)
    try:
        mask = np.zeros(mod.shape, bool)
        mask.flat[mod.network()[0]] = True
        # the water balance of the run (wflow_hbv.py:1208-1215, 1731-1766): sums since the start, storage, watbal
        s0 = {k: g["state0/" + k] for k in ("FreeWater", "DrySnow", "SoilMoisture", "UpperZoneStorage", "LowerZoneStorage", "InterceptionStorage")}
        mod.set("initstorage", ((((s0["FreeWater"] + s0["DrySnow"]) + s0["SoilMoisture"]) + s0["UpperZoneStorage"]) + s0["LowerZoneStorage"]) + s0["InterceptionStorage"])
        sums = ["sumprecip", "sumevap", "sumsoilevap", "sumpotevap", "sumtemp", "sumrunoff", "sumlevel", "suminflow", "storage", "watbal", "SurfaceRunoffMM"]
        mod.request(*sums)
        for t in range(kw["steps"]):
            P, PET, T, Inflow, Seep = _forcing(g, t)
            mod.dynamic(P, PET, T, Inflow=Inflow, Seepage=Seep)
        last = kw["steps"] - 1
        for f in hbv.STATES + sums:
            # (watbal is a cancellation residual of sums of a few hundred mm: compared by its own bound)
            e = common.compare(mod.get(f), g["out/%d/%s" % (last, f)], mask, 1e-4, 2e-3 if f == "watbal" else 1e-4)
            assert e <= 1.0, (name, f, e)
    finally:
        mod.close()


def test_larger_grid_against_the_cpu_restatement():
    """256 x 256 cells in basins of 64 x 64 (cluster / block sweeps over several groups), 5 hourly steps with 4 sub-steps."""
    from oracle import hbv_oracle
    n, dt = 256, 3600
    ldd, static, state = synth.hbv_case(n, n, kind="tiles", tile=64, seed=11, timestepsecs=dt, block=16)
    state["SurfaceRunoff"][0, 0] = 0.0
    fields = dict(static)
    fields.update(state)
    orc = hbv_oracle.HbvOracle(ldd, fields, timestepsecs=dt, kinwaveIters=1, kinwaveTstep=900)
    mod = hbv.HbvModel(ldd, timestepsecs=dt, kinwaveIters=1, kinwaveTstep=900)
    try:
        mod.set_many({k: v for k, v in static.items() if k in hbv.PARAMETERS})
        mod.set_many(state)
        mask = orc.active
        forc = synth.Forcing(n, n, seed=5, timestepsecs=dt, rain_mean=14.0, t_mean=3.0)
        for t in range(5):
            P, PET, T = forc.step(t)
            orc.step(P, PET, T)
            mod.dynamic(P, PET, T)
            for f in hbv.STATES:
                e = common.compare(mod.get(f), orc.get(f), mask, 1e-5, 2e-5)
                assert e <= 1.0, (t, f, e)
        assert float(mod.get("SurfaceRunoff")[mask].max()) > 1.0
    finally:
        mod.close()


def test_refused_options_fail_loudly():
    ldd, static, state = synth.hbv_case(24, 24, kind="tiles", tile=12)
    mod = hbv.HbvModel(ldd, glacier=1)
    try:
        with pytest.raises(Exception):
            mod.set_option("MassWasting", 1)  # (not together with the glacier module)
        with pytest.raises(Exception):
            mod.set_option("Abstractions", 1)
        with pytest.raises(Exception):
            mod.step()  # parameters not set
    finally:
        mod.close()


@pytest.mark.parametrize("cluster_size,case", [("2", "tiles_hourly_sub"), ("4", "tiles_hourly_sub"), ("2", "tiles_courant")])
def test_channel_kernel_is_bit_identical_to_the_shared_sweep(monkeypatch, cluster_size, case):
    """Under a cluster plan the channel wave runs as k_channel_wave (1024 threads per block); the river task of the shared
    sweep (WF_HBV_SHARED_SWEEP=1) must give the same bits: same device function, same schedule of dependencies.  Both
    against the reference's golden as well.  Hourly steps with four sub-steps, missing cells inside the basins."""
    g = np.load(os.path.join(GOLD, "hbv_%s.npz" % case))  # (the Courant case: sub-steps re-sized every step, 80 ... 304 of them)
    kw = ast.literal_eval(str(g["meta"]))
    monkeypatch.setenv("WF_LATERAL_MODE", "cluster")
    monkeypatch.setenv("WF_CLUSTER_SIZE", cluster_size)
    got = {}
    for shared in ("0", "1"):
        monkeypatch.setenv("WF_HBV_SHARED_SWEEP", shared)
        _, _, mod = _load(case)
        try:
            assert mod.lateral_plan["mode"] in ("cluster", "cluster-wide"), mod.lateral_plan
            for t in range(kw["steps"]):
                P, PET, T, Inflow, Seep = _forcing(g, t)
                mod.dynamic(P, PET, T, Inflow=Inflow, Seepage=Seep)
            got[shared] = {f: mod.get(f) for f in hbv.STATES}
            mask = np.zeros(mod.shape, bool)
            mask.flat[mod.network()[0]] = True
        finally:
            mod.close()
    last = kw["steps"] - 1
    for f in hbv.STATES:
        np.testing.assert_array_equal(got["0"][f], got["1"][f], err_msg=f)
        assert common.compare(got["0"][f], g["out/%d/%s" % (last, f)], mask, 1e-4, 1e-4) <= 1.0, f


def test_ensemble_members_are_bit_identical_to_stand_alone_models():
    """Calibration trials of a wflow_hbv case as ONE device model (ensemble.HbvEnsemble: members side by side, what they share
    uploaded once): every member must equal the stand-alone model with its parameters, bit for bit."""
    from wflow_b200.ensemble import HbvEnsemble
    n, dt, M = 36, 3600, 4
    ldd, static, state = synth.hbv_case(n, n, kind="tiles", tile=12, seed=9, timestepsecs=dt)
    factors = {"FieldCapacity": [1.0, 0.7, 1.3, 1.1], "KQuickFlow": [1.0, 1.5, 0.6, 1.0], "Alpha": [1.0, 1.0, 1.2, 0.8]}
    ens = HbvEnsemble(ldd, M, timestepsecs=dt, kinwaveIters=1, kinwaveTstep=900)
    singles = []
    try:
        for k, v in static.items():
            if k in hbv.PARAMETERS and k not in factors:
                ens.set(k, v)
        for k, f in factors.items():
            ens.scale(k, static[k], f)
        ens.set_states(state)
        for m in range(M):
            mod = hbv.HbvModel(ens.ldd, timestepsecs=dt, kinwaveIters=1, kinwaveTstep=900)
            par = {k: v for k, v in static.items() if k in hbv.PARAMETERS}
            for k, f in factors.items():
                par[k] = (static[k] * np.float32(f[m])).astype(np.float32)
            mod.set_many(par)
            mod.set_many(state)
            singles.append(mod)
        forc = synth.Forcing(n, n, seed=5, timestepsecs=dt, rain_mean=14.0, t_mean=3.0)
        for t in range(4):
            P, PET, T = forc.step(t)
            ens.dynamic(P, PET, T)
            for mod in singles:
                mod.dynamic(P, PET, T)
        for f in hbv.STATES:
            got = ens.get(f)
            for m in range(M):
                np.testing.assert_array_equal(got[m], singles[m].get(f), err_msg="%s, member %d" % (f, m))
        q = ens.get("SurfaceRunoff")
        assert not np.array_equal(q[0], q[1]) and float(q[q > -999].max()) > 0.0
    finally:
        ens.close()
        for mod in singles:
            mod.close()
