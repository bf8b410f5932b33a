"""Host-side mirror of wflow_hbv's model class and of the framework services around it (wflow/wflow_hbv.py,
wflow/wf_DynamicFramework.py), on the CUDA library: `python -m wflow_b200.framework_hbv -C case -R run` runs a wflow_hbv
case directory as the reference's `python -m wflow.wflow_hbv` does.

Everything the framework does around a model -- ini handling, time loop, map stacks, [modelparameters], [outputcsv_N] /
[outputtss_N] / [outputmaps] / [summary*], the API calls -- is `framework.WflowSbm`'s; this class replaces what is the
model's own: ``initial()`` (wflow_hbv.py:324-1140: tables with HBV's defaults, rates scaled by timestepsecs / basetimestep,
river width and length, the channel's Alpha), ``resume()`` (:1146-1220: cold state, initstorage, KQuickFlow from KHQ / HQ /
AlphaNL), ``stateVariables()`` and the per-timestep call (``dynamic()``, :1222-1767 = wf_hbv_create's step).
Not covered: reservoirs / lakes, updating from observed discharge, scalar (tss) input, SubCatchFlowOnly
-- refused loudly.  The slope and window-average restatements are framework.py's (not pinned against
PCRaster; supply Slope.map / wflow_riverwidth.map to bypass them).
"""
import os

import numpy as np

from . import hbv, params, tbl
from .core import Schedule
from .framework import F32, MV, WflowSbm, _repair_ldd, configget, lattometres, slope_from_dem, window_average

STATE_FILES = ["FreeWater", "SoilMoisture", "UpperZoneStorage", "LowerZoneStorage", "InterceptionStorage", "SurfaceRunoff",
               "WaterLevel", "DrySnow"]
# readtblDefault defaults of initial() (wflow_hbv.py:560-900); True: multiplied by timestepsecs / basetimestep
TABLES = {"FC": (260.0, False), "BetaSeepage": (1.8, False), "LP": (0.53, False), "K4": (0.02307, True),
          "KQuickFlow": (0.09880, True), "SUZ": (100.0, False), "K0": (0.3, True), "KHQ": (0.09880, True), "HQ": (3.27, True),
          "AlphaNL": (1.1, False), "PERC": (0.4, True), "CFR": (0.05, False), "Pcorr": (1.0, False), "RFCF": (1.0, False),
          "SFCF": (1.0, False), "Cflux": (2.0, True), "ICF": (2.0, False), "CEVPF": (1.0, False), "EPF": (0.0, False),
          "ECORR": (1.0, False), "TTI": (1.0, False), "TT": (-1.41934, False), "Cfmax": (3.75653, True), "WHC": (0.1, False),
          "N": (0.072, False), "N_River": (0.036, False)}


def _pow(x, y):
    """``**`` of REAL4 maps: evaluated in float64, rounded once (as the goldens' PCRaster stand-in does)."""
    with np.errstate(all="ignore"):
        return np.power(np.asarray(x, np.float64), np.asarray(y, np.float64)).astype(F32)


class WflowHbv(WflowSbm):
    _ALIAS = {"Temperature": "Temperature"}

    def __init__(self, cloneMap="wflow_subcatch.map", Dir=".", RunDir="run_default", configfile="wflow_hbv.ini", **kw):
        super().__init__(cloneMap, Dir, RunDir, configfile, **kw)

    # ---- initial(), wflow_hbv.py:324-1140 ---------------------------------------------------------------
    def _runInitial(self):
        cfg = lambda k, d: configget(self.config, "model", k, d)  # noqa: E731
        self._check_netcdf_options()
        for opt, what in (("ScalarInput", "scalar (tss) input"), ("updating", "updating from observed discharge"),
                          ("SubCatchFlowOnly", "a subcatchment-only drainage network")):
            if int(cfg(opt, "0")):
                raise ValueError("[model] %s: %s is not part of this implementation of wflow_hbv" % (opt, what))
        nrivermethod = int(cfg("nrivermethod", "1"))
        if nrivermethod not in (1, 2):
            raise ValueError("[model] nrivermethod = %d (1: N_River by land use / subcatchment / soil, 2: by stream order)" % nrivermethod)
        mapname = lambda k: cfg(k, "staticmaps/%s.map" % k)  # noqa: E731
        self.modelSnow = 1  # (the temperature stack is always read, wflow_hbv.py:1275)
        self.kinwaveIters = int(cfg("kinwaveIters", "0"))
        self.kinwaveTstep = int(cfg("kinwaveTstep", "0"))
        self.SetKquickFlow = int(cfg("SetKquickFlow", "0"))
        self.ExternalQbase = int(cfg("ExternalQbase", "0"))
        self.MassWasting = int(cfg("MassWasting", "0"))
        self.intbl = cfg("intbl", "intbl")
        sizeinmetres = int(configget(self.config, "layout", "sizeinmetres", "0"))
        self.maxLayers = 1
        self.layer_thickness = ()
        dt, base_dt = F32(self.timestepsecs), F32(86400)

        sub = self._staticmap(mapname("wflow_subcatch"), fail=True)
        self.shape = sub.shape
        info = self._info
        active = sub > 0
        self.TopoId = np.where(active, sub, 0).astype(np.int32)
        ldd = self._staticmap(mapname("wflow_ldd"), fail=True)
        dem = self._staticmap(mapname("wflow_dem"), fail=True).astype(F32)
        river = self._staticmap(mapname("wflow_river"), fail=True)
        landuse = self._staticmap(mapname("wflow_landuse"), fail=True)
        soil = self._staticmap(mapname("wflow_soil"), fail=True)
        gauges = self._staticmap(mapname("wflow_gauges"), fail=True)
        for name in ("staticmaps/wflow_reservoirlocs.map", "staticmaps/wflow_lakelocs.map"):
            v = self._staticmap(name)
            if v is not None and (np.nan_to_num(v, nan=0.0) > 0).any():
                raise ValueError("%s: reservoirs / lakes are not part of this implementation of wflow_hbv" % name)
        landuse = np.where(landuse <= -999, active.astype(np.int32), landuse)
        soil = np.where(soil <= -999, active.astype(np.int32), soil)
        self.OutputLoc = np.where(gauges > 0, gauges, 0).astype(np.int32)
        self.LandUse, self.Soil = landuse, soil
        ldd_map = _repair_ldd(np.where((ldd >= 1) & (ldd <= 9), ldd, 0).astype(np.uint8))  # TopoLdd as read: the whole map
        ldd = np.where(active & (ldd >= 1) & (ldd <= 9), ldd, 0).astype(np.uint8)  # lddmask(TopoLdd, TopoId), :995
        ldd = _repair_ldd(ldd)
        self.ldd = ldd
        river_b = (np.nan_to_num(river.astype(np.float64), nan=0.0) > 0) & active

        p = {}
        for name, (default, scaled) in TABLES.items():
            if name == "N_River" and nrivermethod == 2:
                continue  # (looked up by stream order below: its table has one key column)
            v = tbl.read_tbl_default(self.Dir, self.intbl, name, landuse, sub, soil, default)
            p[name] = ((v * dt) / base_dt).astype(F32) if scaled else v.astype(F32)
        if nrivermethod == 2:  # readtblFlexDefault(N_River.tbl, 0.036, wflow_streamorder), :572-577
            so = self._staticmap(mapname("wflow_streamorder"))
            path = os.path.join(self.Dir, self.intbl, "N_River.tbl")
            if so is None or not os.path.isfile(path):
                p["N_River"] = np.full(self.shape, 0.036, F32)
            else:
                p["N_River"] = np.nan_to_num(tbl.lookupscalar(path, so), nan=0.036).astype(F32)
        tc = self._staticmap(cfg("TemperatureCorrectionMap", "staticmap/swflow_tempcor.map"))  # (the reference's default path)
        p["TempCor"] = np.zeros(self.shape, F32) if tc is None else np.nan_to_num(tc, nan=0.0).astype(F32)
        # cell sizes, wflow/pcrut.py:53-73
        if sizeinmetres:
            xl = yl = np.full(self.shape, info.cellsize, F32)
        else:
            latlen, longlen = lattometres(np.repeat(info.ycoordinate()[:, None], self.shape[1], 1).astype(F32))
            xl, yl = (longlen * F32(info.cellsize)).astype(F32), (latlen * F32(info.cellsize)).astype(F32)
        reallength = ((xl + yl) * F32(0.5)).astype(F32)
        slope_map = self._staticmap("staticmaps/Slope.map")
        slope = slope_map if slope_map is not None else slope_from_dem(np.where(active, dem, np.nan), info.cellsize)
        slope = np.maximum(F32(0.001), np.nan_to_num(slope, nan=0.0) * F32(info.cellsize) / reallength).astype(F32)  # :911-915
        # [variable_change_once] (wf_multparameters in initial(), :927)
        self._read_variable_changes()
        for var, fac in self._changes_once:
            if var in p:
                p[var] = (p[var] * F32(fac)).astype(F32)
                self.logger.warning("Variable change (apply_once) applied to %s with factor %g" % (var, fac))
            else:
                self.logger.error("Variable change (apply_once) could not be applied to " + var)
        N = np.where(river_b, p["N_River"], p["N"]).astype(F32)  # :928
        # river width from the upstream area and the yearly discharge, :936-949 (window average restated, unpinned)
        # (upstr is taken from TopoLdd BEFORE initial() masks it with the catchment, :936-941 against :995: cells that
        # receive water from outside the mask count that area, and mapmaximum runs over the whole map)
        upcells = Schedule(ldd_map).catchment_total(None).astype(F32)
        rw = self._staticmap(mapname("wflow_riverwidth"))
        rw = np.zeros(self.shape, F32) if rw is None else np.nan_to_num(rw, nan=0.0).astype(F32)
        alf, Qmax = float(cfg("Alpha", "60")), float(cfg("AnnualDischarge", "300"))
        Qscale = (upcells / F32(max(float(upcells.max()), 1.0)) * F32(Qmax)).astype(F32)
        wavg = np.nan_to_num(window_average(np.where(active, slope, np.nan), 4.0), nan=0.0001).astype(F32)
        c0 = F32((alf * (alf + 2.0) ** 0.6666666667) ** 0.375)  # (Python floats in the reference, REAL4 where it meets the map)
        W = (c0 * _pow(Qscale, F32(0.375)) * _pow(np.maximum(F32(0.0001), wavg), F32(-0.1875))
             * _pow(N, F32(0.375))).astype(F32)
        rw = np.where(rw <= 0.0, W, rw).astype(F32)
        bf = self._staticmap(mapname("wflow_bankfulldepth"))
        bf = np.ones(self.shape, F32) if bf is None else np.nan_to_num(bf, nan=0.0).astype(F32)
        bankfull = np.where(river_b, bf, F32(0.0)).astype(F32)
        # channel length and width of every cell, :1049-1062
        rl = self._staticmap(mapname("wflow_riverlength"))
        rl = np.zeros(self.shape, F32) if rl is None else np.nan_to_num(rl, nan=0.0).astype(F32)
        rlf = self._staticmap(mapname("wflow_riverlength_fact"))
        rlf = np.ones(self.shape, F32) if rlf is None else np.nan_to_num(rlf, nan=1.0).astype(F32)
        dcl = (np.maximum(params.drain_length(ldd, xl, yl), rl) * np.maximum(F32(1.0), rlf)).astype(F32)
        bw = np.where(river_b, rw, params.drain_width(ldd, xl, yl)).astype(F32)
        beta = F32(0.6)
        with np.errstate(all="ignore"):
            alp_term = _pow(N / np.sqrt(slope.astype(np.float64)).astype(F32), beta)  # :1064
            alp_pow = F32(F32(2.0 / 3.0) * beta)
            alpha = (alp_term * _pow(bw + bankfull, alp_pow)).astype(F32)  # :1068
        st = {k: p[k] for k in ("Pcorr", "TempCor", "TTI", "TT", "SFCF", "RFCF", "ICF", "EPF", "ECORR", "CEVPF", "Cfmax", "CFR", "WHC",
                                "BetaSeepage", "PERC", "Cflux", "AlphaNL", "K4", "K0", "SUZ")}
        st["FieldCapacity"] = p["FC"]
        st["Treshold"] = (p["LP"] * p["FC"]).astype(F32)  # :1033-1036
        if self.SetKquickFlow:
            st["KQuickFlow"] = p["KQuickFlow"]
        else:  # resume(), :1217-1220
            st["KQuickFlow"] = (_pow(p["KHQ"], F32(1.0) + p["AlphaNL"]) * _pow(p["HQ"], -p["AlphaNL"])).astype(F32)
        st.update(xl=xl, yl=yl, DCL=dcl, Bw=bw, Alpha=alpha, Slope=slope)
        self.static = st
        # maps of initial() that outputs, summaries and the API may name and that are not rasters of the library
        # (wflow_hbv.py: self.Altitude, self.River, self.Slope, self.N, self.RiverWidth, self.reallength; DLC is the [API]
        # section's name of DCL in the shipped examples)
        inside = active & np.isin(ldd, (1, 2, 3, 4, 5, 6, 7, 8, 9))
        self._host_maps = {k: np.where(inside, np.nan_to_num(v, nan=0.0).astype(F32), F32(MV)).astype(F32) for k, v in
                           dict(Altitude=dem, River=river_b.astype(F32), Slope=slope, N=N, RiverWidth=rw, reallength=reallength,
                                DLC=dcl).items()}
        self.River = river_b
        self._read_modelparameters()
        it = 1
        if self.kinwaveIters == 1 and self.kinwaveTstep > 0:
            it = int(np.ceil(self.timestepsecs / self.kinwaveTstep))
        self._estimate_l, self._estimate_r = False, self.kinwaveIters == 1 and self.kinwaveTstep == 0
        self._it_fixed = (1, it)
        self._requested = set()
        self._mod = hbv.HbvModel(ldd, timestepsecs=self.timestepsecs, kinwaveIters=self.kinwaveIters, kinwaveTstep=self.kinwaveTstep,
                                 set_kquick_flow=self.SetKquickFlow, external_qbase=self.ExternalQbase, mass_wasting=self.MassWasting,
                                 device=self.device)
        for k, v in st.items():
            self._mod.set(k, np.nan_to_num(v, nan=0.0).astype(F32))
        self._update_parameters(0, initial=True)
        saved, self._api_values = self._api_values, {}
        self._mod.set("P", self._forcing("P", "Precipitation", "/inmaps/P", 1, 0.0))
        self._mod.set("PET", self._forcing("PET", "EvapoTranspiration", "/inmaps/PET", 1, 0.0))
        self._mod.set("TEMP", self._forcing("TEMP", "Temperature", "/inmaps/TEMP", 1, 10.0))
        self._api_values = saved
        self._setup_outputs()

    _DIRECT = ("Inflow", "Seepage", "TempCor", "Pcorr", "ICF", "CEVPF", "EPF", "ECORR")

    # ---- resume(), wflow_hbv.py:1146-1220 ------------------------------------------------------------------
    def _state_names(self):
        return list(STATE_FILES)

    def _runResume(self):
        if self.reinit == 1:
            zero = np.zeros(self.shape, F32)
            fc = self.static["FieldCapacity"]
            with np.errstate(all="ignore"):
                lz = (F32(1.0) / (F32(3.0) * self.static["K4"])).astype(F32)
            self._load_state({"FreeWater": zero, "SoilMoisture": fc, "UpperZoneStorage": (F32(0.2) * fc).astype(F32),
                              "LowerZoneStorage": lz, "InterceptionStorage": zero, "SurfaceRunoff": zero, "WaterLevel": zero,
                              "DrySnow": zero})
        else:
            self.wf_resume(os.path.join(self.Dir, "instate"))

    def _load_state(self, state):
        s = {k: np.nan_to_num(np.where(v == F32(MV), np.nan, v), nan=0.0).astype(F32) for k, v in state.items()}
        self._mod.set_many(s)  # (SurfaceRunoffsub / WaterLevelsub start from their step means, :1202)
        init = (((((s["FreeWater"] + s["DrySnow"]) + s["SoilMoisture"]) + s["UpperZoneStorage"]) + s["LowerZoneStorage"])
                + s["InterceptionStorage"]).astype(F32)  # :1208-1215
        self._mod.set("initstorage", init)

    def _get(self, name):
        if name in self._host_maps:
            return self._host_maps[name]
        return self._mod.get(name, MV)

    def _request(self, name):
        if name not in self._host_maps:
            super()._request(name)

    # ---- dynamic(), wflow_hbv.py:1222-1767 -----------------------------------------------------------------
    def _runDynamic(self, firststep, laststep):
        first = max(int(firststep), 1) if firststep else self._currentTimeStep + 1
        last = int(laststep) if laststep else self.nrTimeSteps
        for step in range(first, last + 1):
            self._currentTimeStep = step
            self._update_parameters(step)
            inflow = self._forcing("Inflow", "Inflow", "/inmaps/IF", step, 0.0) if self._has_stack("Inflow", "/inmaps/IF", step) else None
            seep = self._forcing("Seepage", "Seepage", "/inmaps/SE", step, 0.0) if self.ExternalQbase else None
            self.wf_QuickSuspend()
            self._mod.dynamic(self._forcing("P", "Precipitation", "/inmaps/P", step, 0.0),
                              self._forcing("PET", "EvapoTranspiration", "/inmaps/PET", step, 0.0),
                              self._forcing("TEMP", "Temperature", "/inmaps/TEMP", step, 10.0), Inflow=inflow, Seepage=seep)
            self._write_outputs(step)

    def _has_stack(self, key, default_stack, step):
        from .framework import generate_name_t
        if key in self._api_values:
            return True
        prefix = configget(self.config, "inputmapstacks", key, default_stack).lstrip("/")
        if getattr(self, "_ncin", None) is not None:
            return os.path.basename(prefix) in self._ncin.alldat
        return os.path.isfile(generate_name_t(os.path.join(self.Dir, prefix), step))

    def run_pipelined(self, *a, **kw):
        raise NotImplementedError("wflow_hbv models run the per-step path")


def main(argv=None):
    from . import framework
    return framework.main(argv, model_class=WflowHbv, configfile="wflow_hbv.ini")


if __name__ == "__main__":
    main()
