"""BASELINE config 4's shape: a parameter ensemble (wflow_fit trials: wflow/wflow_fit.py:153-158, 235-257) as one
device model (wflow_b200.ensemble).  Every member must come out BIT-IDENTICAL to a stand-alone run of the model
with that member's parameters -- members are independent full runs in the reference."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu

LT = (100, 300, 800)


def _member_static(base, ldd, river, dt, factors):
    """Derived statics of one trial: the table parameters a calibration multiplies, then initial()'s derivations."""
    from wflow_b200 import params
    b = dict(base)
    for name, f in factors.items():
        b[name] = (np.asarray(b[name], dtype=np.float32) * np.float32(f)).astype(np.float32)
    return params.derive_static(b, ldd, river, dt, cellsize=1000.0, n_layers=len(LT) + 1)


@pytest.mark.parametrize("kind", ["forest", "tiles"])
def test_members_match_stand_alone_runs(kind):
    from wflow_b200 import core, ensemble, params, synth
    nrow, ncol, M, dt, nsteps = 40, 48, 6, 86400, 5
    ldd = synth.ldd_random_forest(nrow, ncol, seed=11) if kind == "forest" else synth.ldd_basin_tiles(nrow, ncol, tile=20, seed=7)
    river = (synth.upstream_count(ldd) >= 8).astype(np.uint8)
    base = synth.base_parameters(nrow, ncol, seed=5, n_layers=len(LT) + 1, block=8)
    up = synth.upstream_count(ldd).astype(np.float32)
    base["RiverWidth"] = np.where(river != 0, np.float32(2.0) * np.power(np.maximum(up, 1), np.float32(0.35)), 0).astype(np.float32)
    base["BankfullDepth"] = np.where(river != 0, np.float32(1.0), np.float32(0.0)).astype(np.float32)
    rng = np.random.default_rng(11)
    trials = [dict(M=f[0], KsatVer=f[1], N=f[2], N_River=f[3], SoilThickness=f[4]) for f in rng.uniform(0.5, 2.0, (M, 5))]
    trials[0] = dict(M=1.0, KsatVer=1.0, N=1.0, N_River=1.0, SoilThickness=1.0)
    statics = [_member_static(base, ldd, river, dt, t) for t in trials]
    states = [synth.random_state(s, len(LT) + 1, LT, seed=21) for s in statics]
    kw = dict(timestepsecs=dt, layer_thickness=LT, model_snow=1, it_kin_l=1, it_kin_r=2)
    ens = ensemble.SbmEnsemble(ldd, river, M, **kw)
    skip = {"River", "Beta", "SatWaterDepth"}
    for name in statics[0]:
        if name not in skip:
            ens.set_members(name, np.stack([np.broadcast_to(np.asarray(s[name], np.float32), (nrow, ncol)) for s in statics]))
    for name in states[0]:
        if name not in skip:
            ens.set_members(name, np.stack([np.broadcast_to(np.asarray(s[name], np.float32), (nrow, ncol)) for s in states]))
    forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=dt, rain_mean=12.0)
    trips = [forc.step(t) for t in range(nsteps)]
    gauges = np.flatnonzero(ldd.ravel() == 5)[:4]
    gh = ens.gauges_create(gauges)
    series = []
    for t in range(nsteps):
        for name, a in zip(("P", "PET", "TEMP"), trips[t]):
            ens.set(name, a)
        ens.step()
        series.append(ens.gauges_sample(gh, "RiverRunoff"))
    names = common.state_names(len(LT) + 1)
    got = {k: ens.get(k) for k in names}
    ens.close()
    differ = 0
    for m in range(M):
        one = core.SbmModel(ldd, river, **kw)
        for src in (statics[m], states[m]):
            for name, v in src.items():
                if name not in skip:
                    one.set(name, v)
        for t in range(nsteps):
            for name, a in zip(("P", "PET", "TEMP"), trips[t]):
                one.set(name, a)
            one.step()
            np.testing.assert_array_equal(series[t][m], one.sample("RiverRunoff", gauges))
        for k in names:
            np.testing.assert_array_equal(got[k][m], one.get(k), err_msg="member %d: %s" % (m, k))
        if m > 0:
            differ += int(not np.array_equal(got["zi"][m], got["zi"][0]))
        one.close()
    assert differ == M - 1  # the trials really are different runs


RHINE = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))),
                                   "baseline", "_ref", "wflow_rhine_sbm")


@pytest.mark.skipif(not __import__("os").path.isdir(RHINE), reason="Rhine example data not present (tests/golden/make_ref_cases.py)")
def test_rhine_ensemble_of_256_trials_matches_stand_alone_runs(tmp_path):
    """BASELINE config 4: the Rhine example x 256 parameter trials (wflow_fit shapes: factors on M, KsatVer, N, N_River,
    SoilThickness) as ONE device model on the real Rhine maps; trials picked from the ensemble are bit-identical to
    stand-alone runs of the model with their parameters."""
    import shutil
    from wflow_b200 import core, ensemble
    M, nsteps = 256, 5
    rng = np.random.default_rng(11)
    f = rng.uniform(0.5, 2.0, (M, 5)).astype(np.float32)
    f[0] = 1.0
    factors = dict(M=f[:, 0], KsatVer=f[:, 1], N=f[:, 2], N_River=f[:, 3], SoilThickness=f[:, 4])
    run = "ens_%s" % tmp_path.name
    ce = ensemble.CaseEnsemble(RHINE, factors, run_dir=run)
    try:
        fw = ce.fw
        assert ce.members == M and ce.model.num_cells == M * fw._mod.num_cells
        assert set(ce.varied) >= {"KsatVer", "f", "AlphaL", "SoilThickness"} and "thetaS" not in ce.varied
        gauges = np.flatnonzero(np.nan_to_num(fw._staticmap("staticmaps/wflow_gauges.map"), nan=0.0).ravel() > 0)
        ce.gauges(gauges)
        names = common.state_names(fw.maxLayers, snow=bool(fw.modelSnow))
        start = {k: ce.ens.get(k) for k in names}
        forc = [ce.forcing(t + 1) for t in range(nsteps)]
        series = ce.run(nsteps)
        assert series.shape == (nsteps, M, gauges.size) and np.isfinite(series).all() and series.max() > 0
        got = {k: ce.ens.get(k) for k in names}
        it_l, it_r = fw._it_fixed
        for m in (0, 100, 255):
            one = core.SbmModel(fw.ldd, fw.River.astype(np.uint8), timestepsecs=fw.timestepsecs, layer_thickness=fw.layer_thickness,
                                model_snow=fw.modelSnow, soil_inf_redu=fw.soilInfReduction, transfer_method=fw.TransferMethod,
                                ust=fw.UST, it_kin_l=it_l, it_kin_r=it_r)
            for k, v in ce.statics[m].items():
                if k not in ("River", "Beta"):
                    one.set(k, np.nan_to_num(np.broadcast_to(np.asarray(v, np.float32), fw.shape), nan=0.0))
            for k in names:
                one.set(k, np.where(start[k][m] == np.float32(-999.0), np.float32(0.0), start[k][m]))
            for t in range(nsteps):
                for name, a in forc[t]:
                    one.set(name, a)
                one.step()
                np.testing.assert_array_equal(series[t][m], one.sample("RiverRunoff", gauges))
            for k in names:
                np.testing.assert_array_equal(got[k][m], one.get(k), err_msg="trial %d: %s" % (m, k))
            one.close()
        assert not np.array_equal(series[:, 100], series[:, 255])
    finally:
        ce.close()
        shutil.rmtree(__import__("os").path.join(RHINE, run), ignore_errors=True)
