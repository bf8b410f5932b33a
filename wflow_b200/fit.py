"""Calibration driver: wflow_fit's least-squares fit on top of the parameter ensemble.

The reference (wflow/wflow_fit.py) calibrates a model by ``scipy.optimize.leastsq`` over multipliers of parameter
maps: every evaluation of the residual -- and MINPACK needs one per parameter for each forward-difference Jacobian --
is a complete serial run of the model (:235-257 ``run``, :315-347 ``errfuncFIT``).  Here the trials of one Jacobian are
the members of ONE device model (ensemble.CaseEnsemble): the base point and its P perturbed neighbours advance
together, a Levenberg-Marquardt iteration costs one ensemble run instead of P + 1 model runs.

The ``[fit]`` section of the model's ini is the reference's (:153-184):
    parameter_0 = M, parameter_1 = KsatVer, ...   names of the maps to scale (start value 1.0 each)
    Q = calib.tss          observed discharge (PCRaster timeseries) under the case
    ColSim = [1]  ColMeas = [1]   1-based columns of the simulated gauges / the observations, fitted one after the other
    areamap = staticmaps/wflow_catchment.map   areacode = [1]   where the multipliers apply
    WarmUpSteps = 1        rows left out of the residual
    epsfcn, ftol, xtol, gtol, factor           MINPACK's controls
Residual: simulated minus observed RiverRunoff at the gauge, rows with a missing observation dropped (:316-319).
"""
import configparser
import os

import numpy as np

from .ensemble import CaseEnsemble


def read_tss(path):
    """PCRaster timeseries file -> array [rows, 1 + columns] (first column: the timestep), wflow/pcrut.py readtss."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    ncol = int(lines[1].split()[0])
    rows = [ln.split() for ln in lines[2 + ncol:] if ln.strip()]
    a = np.array(rows, dtype=np.float64)
    a[a == 1e31] = np.nan
    return a


def write_tss(path, values, title="timeseries scalar"):
    """values [rows, columns] -> PCRaster timeseries file with the timestep as first column."""
    values = np.atleast_2d(np.asarray(values, np.float64))
    with open(path, "w") as fh:
        fh.write("%s\n%d\ntimestep\n" % (title, values.shape[1] + 1))
        for k in range(values.shape[1]):
            fh.write("%d\n" % (k + 1))
        for r, row in enumerate(values):
            fh.write(" ".join(["%d" % (r + 1)] + ["%.9g" % v for v in row]) + "\n")


class FitSettings:
    def __init__(self, case_dir, inifile="wflow_sbm.ini"):
        cp = configparser.ConfigParser()
        cp.optionxform = str
        cp.read(os.path.join(case_dir, inifile))
        get = lambda k, d: cp.get("fit", k) if cp.has_section("fit") and cp.has_option("fit", k) else d  # noqa: E731
        self.parameters = []
        while get("parameter_%d" % len(self.parameters), "__") != "__":
            self.parameters.append(get("parameter_%d" % len(self.parameters), "M").strip())
        self.qmeas = get("Q", "calib.tss").strip()
        aslist = lambda s: [int(v) for v in s.strip().strip("[]").split(",") if v.strip()]  # noqa: E731
        self.col_sim, self.col_meas = aslist(get("ColSim", "[1]")), aslist(get("ColMeas", "[1]"))
        self.area_codes = aslist(get("areacode", "[1]"))
        self.area_map = get("areamap", "staticmaps/wflow_catchment.map").strip()
        self.warmup = int(get("WarmUpSteps", "1"))
        self.epsfcn = float(get("epsfcn", "0.00001"))
        self.ftol, self.xtol, self.gtol = (float(get(k, "0.0001")) for k in ("ftol", "xtol", "gtol"))
        self.factor = float(get("factor", "100.0"))


class WflowFit:
    """Fit the multipliers of ``[fit] parameter_i`` to the observed discharge of a case; one ensemble run per Jacobian."""

    def __init__(self, case_dir, inifile="wflow_sbm.ini", run_dir="fit", nsteps=None, device=0, gauges=None):
        self.case_dir, self.inifile, self.run_dir, self.device = case_dir, inifile, run_dir, device
        self.settings = FitSettings(case_dir, inifile)
        if not self.settings.parameters:
            raise ValueError("[fit] holds no parameter_0 ... entries")
        self.nsteps = nsteps
        self._gauges = gauges
        self.runs = 0          # ensemble runs made
        self.trials = 0        # model trials inside them

    def simulate(self, pars_list, area_code, col_sim):
        """Discharge series [rows after the warm-up, trials] at simulated column `col_sim` (1-based) for every parameter
        vector of `pars_list`, all trials in one device model."""
        st = self.settings
        pars = np.atleast_2d(np.asarray(pars_list, np.float64))
        factors = {name: pars[:, k].astype(np.float32) for k, name in enumerate(st.parameters)}
        area = (st.area_map, area_code) if os.path.isfile(os.path.join(self.case_dir, st.area_map)) else None
        ce = CaseEnsemble(self.case_dir, factors, inifile=self.inifile, run_dir=self.run_dir, device=self.device, area=area)
        try:
            fw = ce.fw
            g = self._gauges
            if g is None:  # the gauges map, ordered by id like the columns of the reference's run.tss
                loc = np.asarray(fw.OutputLoc).ravel()
                cells = np.flatnonzero(loc > 0)
                g = cells[np.argsort(loc[cells], kind="stable")]
            ce.gauges(np.asarray(g, np.int64))
            n = self.nsteps or fw.nrTimeSteps
            q = ce.run(n)                                    # [steps, trials, gauges]
        finally:
            ce.close()
        self.runs += 1
        self.trials += pars.shape[0]
        return q[st.warmup:, :, col_sim - 1].astype(np.float64)

    def fit(self, x0=None):
        """Levenberg-Marquardt per (ColSim, ColMeas, areacode) triplet like the reference's main loop (:440-470).
        Returns a list of dicts (parameters, multipliers, residual norm, Nash-Sutcliffe, ensemble runs used)."""
        import scipy.optimize
        st = self.settings
        obs = read_tss(os.path.join(self.case_dir, st.qmeas))
        results = []
        for col_sim, col_meas, code in zip(st.col_sim, st.col_meas, st.area_codes):
            P = len(st.parameters)
            cache = {}

            def residuals_of(q, qm):
                res = q - qm
                return res, ~np.isnan(res)

            def base_and_neighbours(x):
                key = tuple(np.round(x, 14))
                if key not in cache:
                    h = np.sqrt(st.epsfcn) * np.where(x != 0.0, np.abs(x), 1.0)   # MINPACK's forward-difference step
                    trials = [x] + [x + h[k] * np.eye(P)[k] for k in range(P)]
                    q = self.simulate(trials, code, col_sim)
                    cache.clear()
                    cache[key] = (q, h)
                return cache[key]

            def fun(x):
                q, _ = base_and_neighbours(x)
                n = q.shape[0]
                res, ok = residuals_of(q[:, 0], obs[st.warmup:st.warmup + n, col_meas])
                return res[ok]

            def jac(x):
                q, h = base_and_neighbours(x)
                n = q.shape[0]
                _, ok = residuals_of(q[:, 0], obs[st.warmup:st.warmup + n, col_meas])
                return ((q[:, 1:] - q[:, :1]) / h[None, :])[ok]

            start = np.ones(P) if x0 is None else np.asarray(x0, np.float64)
            runs0 = self.runs
            sol = scipy.optimize.least_squares(fun, start, jac=jac, method="lm", ftol=st.ftol, xtol=st.xtol, gtol=st.gtol,
                                               x_scale=1.0)
            r = sol.fun
            q, _ = base_and_neighbours(sol.x)
            qm = obs[st.warmup:st.warmup + q.shape[0], col_meas]
            ok = ~np.isnan(q[:, 0] - qm)
            ns = 1.0 - np.sum((q[ok, 0] - qm[ok]) ** 2) / max(np.sum((qm[ok] - qm[ok].mean()) ** 2), 1e-300)
            results.append(dict(parameters=list(st.parameters), multipliers=sol.x.tolist(), cost=float(np.sum(r ** 2)),
                                nash_sutcliffe=float(ns), areacode=code, col_sim=col_sim, ensemble_runs=self.runs - runs0,
                                status=int(sol.status)))
        return results


def main(argv=None):
    import argparse
    import json
    ap = argparse.ArgumentParser(description="least-squares fit of a wflow_sbm case (wflow_fit), trials as one device model")
    ap.add_argument("-C", dest="case", required=True)
    ap.add_argument("-c", dest="ini", default="wflow_sbm.ini")
    ap.add_argument("-R", dest="run", default="fit")
    ap.add_argument("-T", dest="last", type=int, default=None)
    a = ap.parse_args(argv)
    for r in WflowFit(a.case, a.ini, a.run, nsteps=a.last).fit():
        print(json.dumps(r))


if __name__ == "__main__":
    main()
