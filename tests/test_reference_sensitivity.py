"""Characterisation of the reference algorithm (CPU only): it is DISCONTINUOUS at the float32
ulp level.  `unsatzone_flow_layer` picks its sub-step count with a ceil() of a flux ratio
(wflow/wflow_sbm.py:264), layer membership is a strict `zi > top` (:381) and the subsurface
solver has an exact-zero early-out (wflow/wflow_funcs.py:218).  Nudging the float32 states of the
ORACLE by one ulp therefore moves some cells by far more than rel 1e-5 within a few steps.

This is why the GPU parity tests compare single steps from identical states cell by cell
(tests/test_gpu_parity.py::test_single_step_parity_from_identical_states) and judge free-running
trajectories statistically."""
import numpy as np

from tests import common
from wflow_b200 import synth


def _run(perturb_seed, nsteps=12):
    orc, _, _, _ = common.make_pair(24, 24, gpu=False, kind="forest", layer_thickness=(100, 300, 800))
    states = common.state_names(orc.maxLayers, 1)
    if perturb_seed is not None:
        rng = np.random.default_rng(perturb_seed)
        for k in states:
            a = orc.fields[k]
            d = rng.integers(-1, 2, a.shape)
            up, dn = np.nextafter(a, np.float32(np.inf)), np.nextafter(a, np.float32(-np.inf))
            a[...] = np.where(a > 0, np.where(d > 0, up, np.where(d < 0, dn, a)), a)
    forc = synth.Forcing(24, 24, seed=5, rain_mean=20.0)
    out = []
    for t in range(nsteps):
        P, PET, T = forc.step(t)
        orc.set("P", P); orc.set("PET", PET); orc.set("TEMP", T)
        orc.step()
        out.append({k: orc[k].copy() for k in ("zi", "UStoreLayerDepth_0", "SubsurfaceFlow")})
    return orc, out


def test_reference_algorithm_amplifies_one_ulp_perturbations():
    orc, base = _run(None)
    mask = common.network_mask(orc)
    _, pert = _run(1)
    worst, outside, total = 0.0, 0, 0
    for t in range(len(base)):
        for k in ("zi", "UStoreLayerDepth_0"):
            a, b = pert[t][k][mask].astype(np.float64), base[t][k][mask].astype(np.float64)
            err = np.abs(a - b) / (2e-5 + 1e-5 * np.abs(b))
            worst = max(worst, float(err.max()))
            outside += int((err > 1).sum())
            total += err.size
    # the perturbation is 6e-8 relative; the response exceeds 1e-5 relative in some cells ...
    assert worst > 1.0
    # ... but the bulk of the cells stays put
    assert outside / total < 0.2
