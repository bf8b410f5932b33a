"""The calibration driver (wflow_b200.fit, the reference's wflow/wflow_fit.py:99-347) on the Rhine example: a synthetic twin --
the "observations" are the model's own discharge for known multipliers -- must be recovered by the least-squares fit whose
Jacobian columns run as members of one device model."""
import os
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RHINE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "wflow_rhine_sbm")


def _case_with_fit_section(tmp_path, extra):
    """The Rhine case through symbolic links, with an ini that carries a [fit] section."""
    case = tmp_path / "case"
    case.mkdir()
    for sub in ("staticmaps", "intbl", "inmaps", "instate"):
        os.symlink(os.path.join(RHINE, sub), case / sub)
    ini = open(os.path.join(RHINE, "wflow_sbm.ini")).read()
    (case / "wflow_sbm.ini").write_text(ini + "\n[fit]\n" + extra)
    return str(case)


@pytest.mark.skipif(not os.path.isdir(RHINE), reason="Rhine case not present (tests/golden/make_ref_cases.py)")
def test_fit_recovers_the_multipliers_of_a_synthetic_twin(tmp_path):
    from wflow_b200 import fit
    case = _case_with_fit_section(tmp_path, "parameter_0 = N_River\nparameter_1 = N\nQ = calib.tss\nColSim = [3]\nColMeas = [1]\n"
                                            "areacode = [1]\nareamap = staticmaps/none.map\nWarmUpSteps = 2\nepsfcn = 0.0001\n"
                                            "ftol = 1e-10\nxtol = 1e-10\ngtol = 1e-10\n")
    truth = np.array([1.30, 0.70])  # river and land roughness: what a month of discharge at a gauge can tell apart
    try:
        wf = fit.WflowFit(case, nsteps=29)
        assert wf.settings.parameters == ["N_River", "N"] and wf.settings.col_sim == [3] and wf.settings.warmup == 2
        wf.settings.warmup = 0
        q_true = wf.simulate([truth, [1.0, 1.0]], 1, 3)          # the twin and the starting point in one run
        wf.settings.warmup = 2
        assert q_true.shape == (29, 2) and np.isfinite(q_true).all() and q_true[:, 0].max() > 0
        obs = q_true[:, :1].copy()
        obs[10, 0] = np.nan                                       # a missing observation is left out of the residual
        fit.write_tss(os.path.join(case, "calib.tss"), np.where(np.isnan(obs), 1e31, obs))
        start_cost = float(np.nansum((q_true[2:, 1] - obs[2:, 0]) ** 2))
        res = wf.fit()[0]
        print("fit: multipliers %s (truth %s), cost %.3g -> %.3g, NS %.6f, %d ensemble runs of %d trials"
              % (np.round(res["multipliers"], 4), truth, start_cost, res["cost"], res["nash_sutcliffe"], res["ensemble_runs"], 3))
        assert start_cost > 0 and res["cost"] < 1e-4 * start_cost
        assert res["nash_sutcliffe"] > 0.9999
        np.testing.assert_allclose(res["multipliers"][0], truth[0], rtol=0.03)
        # the land roughness only where the twin's discharge depends on it at all (otherwise any value fits)
        q_n = wf.simulate([[truth[0], 1.0]], 1, 3)[:, 0]
        if float(np.nansum((q_n - obs[2:, 0]) ** 2)) > 1e-3 * start_cost:
            np.testing.assert_allclose(res["multipliers"][1], truth[1], rtol=0.05)
    finally:
        shutil.rmtree(os.path.join(case, "fit"), ignore_errors=True)
