"""Several timesteps in one launch (wf_step_pipelined, wflow_b200/csrc/pipeline.cuh) against the
per-step path: the reference's time loop (wf_DynamicFramework._runDynamic,
wflow/wf_DynamicFramework.py:2968-3042) run as a wavefront skewed in time must leave every state
BIT-IDENTICAL to calling wf_step once per timestep, and hand out every step's discharge at the gauges."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu

ENV_KEYS = ("WF_LATERAL_MODE", "WF_CLUSTER_SIZE", "WF_CLUSTER_WIDE", "WF_GROUPS", "WF_PIPE_OVERLAP")
MODES = {
    "grid": {"WF_LATERAL_MODE": "grid"},
    "cta": {"WF_LATERAL_MODE": "cta"},
    "cluster": {"WF_LATERAL_MODE": "cluster", "WF_CLUSTER_SIZE": "2", "WF_CLUSTER_WIDE": "0"},
    "cluster-wide": {"WF_LATERAL_MODE": "cluster", "WF_CLUSTER_SIZE": "4", "WF_CLUSTER_WIDE": "1"},
    # cluster sweeps of several groups default to the per-step kernels with the next column on the idle SMs
    # (lateral.cuh column_ahead); "-skewed": the time-skewed kernel (k_pipeline) under the same plan
    "cluster-skewed": {"WF_LATERAL_MODE": "cluster", "WF_CLUSTER_SIZE": "2", "WF_CLUSTER_WIDE": "0", "WF_PIPE_OVERLAP": "0"},
    "cluster-wide-skewed": {"WF_LATERAL_MODE": "cluster", "WF_CLUSTER_SIZE": "4", "WF_CLUSTER_WIDE": "1", "WF_PIPE_OVERLAP": "0"},
}
CASES = [
    # (mode, grid, steps, make_pair keywords)
    ("grid", (48, 64), 7, dict(kind="tiles", tile=16, layer_thickness=(100, 300, 800), dt=3600, itL=2, itR=3)),
    ("cta", (48, 64), 7, dict(kind="tiles", tile=16, layer_thickness=(100, 300, 800), dt=3600, itL=2, itR=3)),
    ("cluster", (48, 64), 7, dict(kind="tiles", tile=16, layer_thickness=(100, 300, 800), dt=3600, itL=2, itR=3)),
    ("cluster-wide", (48, 64), 7, dict(kind="tiles", tile=16, layer_thickness=(100, 300, 800), dt=3600, itL=1, itR=4)),
    ("cluster-skewed", (48, 64), 7, dict(kind="tiles", tile=16, layer_thickness=(100, 300, 800), dt=3600, itL=2, itR=3)),
    ("cluster-wide-skewed", (48, 64), 7, dict(kind="tiles", tile=16, layer_thickness=(100, 300, 800), dt=3600, itL=1, itR=4)),
    ("cluster-wide", (96, 128), 6, dict(kind="tiles", tile=32, layer_thickness=(100, 300, 800), dt=3600, itL=1, itR=4, mask_frac=0.1)),
    ("grid", (40, 40), 12, dict(kind="forest", layer_thickness=(), tm=1)),                    # one layer, daily (Gash), Topog transfer
    ("cta", (40, 40), 12, dict(kind="forest", layer_thickness=(100, 300), snow=0, itL=2, itR=2, dt=21600)),
    ("cluster", (64, 64), 5, dict(kind="tiles", tile=32, layer_thickness=(100, 300, 800), glacier=1, mask_frac=0.1)),
    ("grid", (32, 32), 40, dict(kind="tiles", tile=32, layer_thickness=(100, 300, 800), dt=3600, itR=4)),  # more steps than levels
    # one basin under a single-block / cluster plan: the pipelined launch sweeps it with the whole grid
    ("cta", (32, 32), 9, dict(kind="tiles", tile=32, layer_thickness=(100, 300, 800), dt=3600, itR=4)),
    ("cluster", (32, 32), 9, dict(kind="tiles", tile=32, layer_thickness=(100, 300), dt=3600, itL=2, itR=2)),
]


def _env(monkeypatch, mode):
    for k in ENV_KEYS:
        monkeypatch.delenv(k, raising=False)
    for k, v in MODES[mode].items():
        monkeypatch.setenv(k, v)


def _gauge_cells(mod, river, shape):
    """A handful of river cells, one cell of the network outside the river, and (if any) one outside the network."""
    rng = np.random.default_rng(1)
    riv = np.flatnonzero(np.asarray(river).ravel() != 0)
    cells = list(rng.choice(riv, size=min(6, riv.size), replace=False)) if riv.size else []
    order = mod.cell_order()
    nonriv = np.setdiff1d(order, riv)
    if nonriv.size:
        cells.append(int(nonriv[0]))
    outside = np.setdiff1d(np.arange(shape[0] * shape[1]), order)
    if outside.size:
        cells.append(int(outside[0]))
    if cells:
        cells.append(cells[0])  # the same cell twice
    return np.asarray(cells, np.int64)


@pytest.mark.parametrize("mode,shape,nsteps,kw", CASES, ids=["%s-%d" % (c[0], i) for i, c in enumerate(CASES)])
def test_pipelined_steps_are_bit_identical_to_single_steps(monkeypatch, mode, shape, nsteps, kw):
    from wflow_b200 import synth
    _env(monkeypatch, mode)
    nrow, ncol = shape
    _, a, ldd, river = common.make_pair(nrow, ncol, oracle=False, **kw)
    _, b, _, _ = common.make_pair(nrow, ncol, oracle=False, **kw)
    assert a.lateral_plan["mode"] == mode.replace("-skewed", "")
    forc = synth.Forcing(nrow, ncol, seed=5, timestepsecs=kw.get("dt", 86400), rain_mean=12.0)
    states = common.state_names(a.max_layers, kw.get("snow", 1)) + (["GlacierStore"] if kw.get("glacier") else [])
    cells = _gauge_cells(a, river, shape)
    ga, gb = a.gauges_create(cells), b.gauges_create(cells)
    # per-step path
    want = np.zeros((nsteps, cells.size), np.float32)
    for t in range(nsteps):
        P, PET, T = forc.step(t)
        a.set("P", P); a.set("PET", PET); a.set("TEMP", T)
        a.step()
        a.gauges_sample(ga, "RiverRunoff", want[t].ctypes.data, wait=True)
    # the same steps in two launches (the second continues where the first stopped)
    b.pipeline_reserve(nsteps)
    b.pipeline_capture(gb)
    got = np.full((nsteps, cells.size), 7.0, np.float32)
    first = max(1, nsteps // 3)
    for lo, hi in ((0, first), (first, nsteps)):
        if hi <= lo:
            continue
        for t in range(lo, hi):
            P, PET, T = forc.step(t)
            b.pipeline_set_forcing("P", t - lo, P)
            b.pipeline_set_forcing("PET", t - lo, PET)
            b.pipeline_set_forcing("TEMP", t - lo, T)
        b.step_pipelined(hi - lo, got[lo:hi])
    for k in states:
        np.testing.assert_array_equal(b.get(k), a.get(k), err_msg="%s differs between pipelined and per-step runs" % k)
    np.testing.assert_array_equal(got, want)
    if mode in ("cluster", "cluster-wide") and a.lateral_plan["groups"] > 1:
        assert b.pipeline_mode == "ahead"   # several groups under a cluster sweep: per-step kernels + the next column on idle SMs
    else:
        assert b.pipeline_mode == "skewed"
    assert np.isfinite(a.get("RiverRunoff")[np.asarray(river) != 0]).all()
    assert float(np.abs(want).max()) > 0.0  # the gauges saw water
    a.close()
    b.close()


def test_pipelined_run_matches_the_oracle(monkeypatch):
    """The pipelined path against the CPU oracle directly (free run, a few steps: rel 1e-4 like the
    sweep-mode test, the per-step resynchronised parity being the per-step path's own test)."""
    from wflow_b200 import synth
    for k in ENV_KEYS:
        monkeypatch.delenv(k, raising=False)
    kw = dict(kind="tiles", tile=16, layer_thickness=(100, 300, 800), dt=3600, itL=1, itR=4)
    orc, mod, ldd, river = common.make_pair(48, 64, **kw)
    mask = common.network_mask(orc)
    forc = synth.Forcing(48, 64, seed=5, timestepsecs=3600, rain_mean=8.0)
    nsteps = 4
    mod.pipeline_reserve(nsteps)
    for t in range(nsteps):
        P, PET, T = forc.step(t)
        orc.set("P", P); orc.set("PET", PET); orc.set("TEMP", T)
        orc.step()
        for name, v in (("P", P), ("PET", PET), ("TEMP", T)):
            mod.pipeline_set_forcing(name, t, v)
    mod.step_pipelined(nsteps)
    for k in common.state_names(4):
        atol = 1e4 if k == "SubsurfaceFlow" else 2e-4
        assert common.compare(mod.get(k), orc[k], mask, 1e-4, atol) <= 1.0, k
    mod.close()


def test_pipelined_argument_errors():
    from wflow_b200._lib import WflowLibError
    _, mod, ldd, river = common.make_pair(32, 32, oracle=False, kind="tiles", tile=16)
    with pytest.raises(WflowLibError):
        mod.step_pipelined(2)                      # nothing reserved
    mod.pipeline_reserve(2)
    with pytest.raises(WflowLibError):
        mod.step_pipelined(2)                      # forcing missing
    with pytest.raises(WflowLibError):
        mod.pipeline_set_forcing("KsatVer", 0, np.zeros((32, 32), np.float32))  # not a forcing
    with pytest.raises(WflowLibError):
        mod.pipeline_set_forcing("P", 2, np.zeros((32, 32), np.float32))        # outside the window
    mod.close()
