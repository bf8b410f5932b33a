"""Host-side mirror of the reference's framework surface for the wflow_sbm hot path.

`WflowSbm` plays the roles of the reference's `WflowModel` (``wflow/wflow_sbm.py:903-3528``) and of
`wf_DynamicFramework` (``wflow/wf_DynamicFramework.py:502-3548``) for this one model: ini-driven
set-up, static maps and lookup tables, state files, the time loop and the per-step outputs -- same
method names, argument meaning and error behaviour -- while every per-timestep computation runs in
libwflowsbm on the GPU (`core.SbmModel`).  Nothing here is executed per cell per step on the host.

Covered: case directory layout (staticmaps/, intbl/, inmaps/ map stacks, instate/, <runId>/outstate,
outmaps, outsum), [run] [model] [layout] [inputmapstacks] [modelparameters] [variable_change_once]
[variable_change_timestep] [outputmaps] [outputcsv_N] [outputtss_N] [summary] [summary_sum|avg|min|max] [API],
cold start (reinit = 1) and warm start, wf_suspend / wf_resume, wf_setValues*, wf_supplyMapAsNumpy,
wf_supplyVariableNamesAndRoles, wf_QuickSuspend/Resume, LAI-driven canopy parameters, the command line.
netCDF: forcing input, [outputmaps] output and states in / out on the classic netCDF-3 formats (ncio.py); netCDF-4 files
and static maps through netCDF are refused loudly (the netCDF4 / HDF5 libraries are not in the image).
Not covered (out of scope): state updating, [rollingmean].

The derivations of `initial()` that need PCRaster neighbourhood operators (slope from the DEM, window
average for the river width) are restated in numpy and are NOT pinned against PCRaster (it cannot be
installed in the build environment); supply `Slope.map` / `RiverWidth.map` to bypass them.
"""
import configparser
import datetime as _dt
import logging
import os

import numpy as np

from . import csf, params, tbl
from .core import Schedule, SbmModel

F32 = np.float32
MV = -999.0

# reference defaults of the parameter tables (wflow_sbm.py:1458-1790)
TABLE_DEFAULTS = {
    "MaxCanopyStorage": 1.0, "CanopyGapFraction": 0.1, "EoverR": 0.1, "RootingDepth": 750.0, "AirEntryPressure": 10.0,
    "rootdistpar": -8000.0, "InfiltCapSoil": 100.0, "CapScale": 100.0, "InfiltCapPath": 10.0, "MaxLeakage": 0.0,
    "PathFrac": 0.01, "SoilThickness": 2000.0, "thetaR": 0.01, "thetaS": 0.6, "SoilMinThickness": 500.0,
    "KsatVer": 3000.0, "KsatHorFrac": 1.0, "M": 300.0, "N": 0.072, "N_River": 0.036, "WaterFrac": 0.0,
    "et_reftopot": 1.0, "TTI": 1.0, "TT": -1.41934, "TTM": -1.41934, "Cfmax": 3.75653, "WHC": 0.1, "w_soil": 0.1125,
    "cf_soil": 0.038,
}

STATE_FILES = ["RiverRunoffsub", "RiverRunoff", "WaterLevelR", "WaterLevelRsub", "LandRunoff", "LandRunoffsub",
               "WaterLevelL", "WaterLevelLsub", "SatWaterDepth", "Snow", "TSoil", "UStoreLayerDepth", "SnowWater",
               "CanopyStorage", "SubsurfaceFlow"]  # stateVariables(), wflow_sbm.py:955-1010


def configget(config, section, var, default):
    """wflow_lib.configget (wflow/wflow_lib.py:141-168): missing options take and record the default."""
    try:
        return config.get(section, var)
    except (configparser.NoSectionError, configparser.NoOptionError):
        if not config.has_section(section):
            config.add_section(section)
        config.set(section, var, str(default))
        return str(default)


def generate_name_t(prefix, step):
    """PCRaster map-stack file name (8.3), e.g. ('inmaps/P', 1) -> 'inmaps/P0000000.001'."""
    d, stem = os.path.split(prefix)
    digits = "%0*d" % (11 - len(stem), step)
    full = stem + digits
    return os.path.join(d, full[:8] + "." + full[8:])


def lattometres(lat):
    """Length of a degree of latitude and longitude in metres (wflow/pcrut.py:17-50), REAL4."""
    r = np.deg2rad(lat.astype(np.float64))
    latlen = 111132.92 + (-559.82 * np.cos(2.0 * r)) + (1.175 * np.cos(4.0 * r)) + (-0.0023 * np.cos(6.0 * r))
    longlen = (111412.84 * np.cos(r)) + (-93.5 * np.cos(3.0 * r)) + (0.118 * np.cos(5.0 * r))
    return latlen.astype(F32), longlen.astype(F32)


def slope_from_dem(dem, cellsize):
    """pcr.slope restated (third-order finite difference / Horn on the 3x3 window, edges and missing
    neighbours replaced by the centre value).  NOT pinned against PCRaster."""
    z = dem.astype(np.float64)
    zc = np.where(np.isnan(z), 0.0, z)
    valid = ~np.isnan(z)
    zp = np.pad(zc, 1, mode="edge")
    vp = np.pad(valid, 1, mode="constant", constant_values=False)

    def nb(dr, dc):
        zz = zp[1 + dr: 1 + dr + z.shape[0], 1 + dc: 1 + dc + z.shape[1]]
        vv = vp[1 + dr: 1 + dr + z.shape[0], 1 + dc: 1 + dc + z.shape[1]]
        return np.where(vv, zz, zc)

    dzdx = ((nb(-1, 1) + 2 * nb(0, 1) + nb(1, 1)) - (nb(-1, -1) + 2 * nb(0, -1) + nb(1, -1))) / (8.0 * cellsize)
    dzdy = ((nb(1, -1) + 2 * nb(1, 0) + nb(1, 1)) - (nb(-1, -1) + 2 * nb(-1, 0) + nb(-1, 1))) / (8.0 * cellsize)
    s = np.sqrt(dzdx * dzdx + dzdy * dzdy)
    return np.where(valid, s, np.nan).astype(F32)


def window_average(a, ncells):
    """pcr.windowaverage(a, ncells * celllength) restated: square window centred on the cell, cells cut
    by the window edge weighted by their covered fraction, missing values skipped.  NOT pinned."""
    half = ncells / 2.0
    k = int(np.ceil(half - 0.5))
    w1 = np.array([min(1.0, half - (abs(d) - 0.5)) for d in range(-k, k + 1)])
    w1 = np.clip(w1, 0.0, 1.0)
    valid = ~np.isnan(a)
    v = np.where(valid, a, 0.0).astype(np.float64)
    vp, mp = np.pad(v, k), np.pad(valid.astype(np.float64), k)
    num = np.zeros(a.shape)
    den = np.zeros(a.shape)
    for i, wi in enumerate(w1):
        for j, wj in enumerate(w1):
            w = wi * wj
            if w > 0:
                num += w * vp[i:i + a.shape[0], j:j + a.shape[1]]
                den += w * mp[i:i + a.shape[0], j:j + a.shape[1]]
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(den > 0, num / den, np.nan).astype(F32)


def area_majority(a, idmap, ids):
    """pcr.areamajority: the value occurring most often within each area; of several such values the highest."""
    out = np.zeros(len(ids), F32)
    for k, i in enumerate(ids):
        u, c = np.unique(a[idmap == i], return_counts=True)
        out[k] = u[len(c) - 1 - int(np.argmax(c[::-1]))]
    return out


class WflowSbm:
    """One wflow_sbm run on one GPU, driven like the reference's WflowModel + wf_DynamicFramework."""

    def __init__(self, cloneMap="wflow_subcatch.map", Dir=".", RunDir="run_default", configfile="wflow_sbm.ini",
                 device=0):
        self.Dir = Dir
        self.caseName = Dir
        self.runId = RunDir
        self.configfile = configfile
        self.cloneMap = cloneMap
        self.device = device
        self.SaveDir = os.path.join(Dir, RunDir)
        self.logger = logging.getLogger("wflow_b200")
        self.config = configparser.ConfigParser()
        self.config.optionxform = str
        ini = os.path.join(Dir, configfile)
        if not os.path.isfile(ini):
            raise FileNotFoundError("Cannot open config file: " + ini)  # reference: logger.error + sys.exit(1)
        self.config.read(ini)
        self._api_values = {}      # forcing set through the API overrides the map stack for one step
        self._currentTimeStep = 0
        self._mod = None
        self._outputs = []
        self._files = {}

    # ---- configuration -------------------------------------------------------------------------
    def _read_time(self):
        run = lambda k, d: configget(self.config, "run", k, d)
        self.timestepsecs = int(run("timestepsecs", configget(self.config, "model", "timestepsecs", "86400")))
        self.reinit = int(run("reinit", configget(self.config, "model", "reinit", "0")))
        fmt = "%Y-%m-%d %H:%M:%S"
        st = run("starttime", "None").replace(" GMT", "").strip()
        et = run("endtime", "None").replace(" GMT", "").strip()
        if st != "None" and et != "None":
            self.datetime_start = _dt.datetime.strptime(st, fmt)
            self.datetime_end = _dt.datetime.strptime(et, fmt)
            # [run] runlengthdetermination (wf_DynamicFramework.py:74-96, 1257): 'steps' (default) -- the state time is one
            # step before starttime; 'intervals' -- the state time IS starttime, every timestep is the interval that ends
            # at its forcing, so the run has one step fewer.  datetime_start is the time of the first forcing either way.
            mode = run("runlengthdetermination", "steps").strip()
            if mode not in ("steps", "intervals"):
                raise ValueError("[run] runlengthdetermination = %s (steps or intervals)" % mode)
            self.runlengthdetermination = mode
            if mode == "intervals":
                self.datetime_start += _dt.timedelta(seconds=self.timestepsecs)
            self.nrTimeSteps = int((self.datetime_end - self.datetime_start).total_seconds() // self.timestepsecs) + 1
        else:
            self.datetime_start = _dt.datetime(1990, 1, 1)
            self.nrTimeSteps = int(run("laststep", "1"))
            self.datetime_end = self.datetime_start + _dt.timedelta(seconds=self.timestepsecs * (self.nrTimeSteps - 1))

    def createRunId(self, **kw):
        """Run directory layout (wf_DynamicFramework.py:1175-1280)."""
        for sub in ("outmaps", "outstate", "outsum", "intbl", "runinfo"):
            os.makedirs(os.path.join(self.SaveDir, sub), exist_ok=True)
        self._read_time()

    # ---- initial() -------------------------------------------------------------------------------
    def _staticmap(self, name, default=None, fail=False):
        path = os.path.join(self.Dir, name)
        if os.path.isfile(path):
            a, info = csf.read_map(path)
            if not hasattr(self, "_info"):
                self._info = info
            elif not info.same_grid(self._info):
                raise ValueError("map %s does not match the clone grid" % path)
            return a
        if fail:
            raise ValueError("Could not find required map: " + path)  # wf_DynamicFramework.py:3372
        return None if default is None else np.full(self.shape, default, F32)

    # ---- netCDF (wf_DynamicFramework.py:1405-1580; wflow_b200/ncio.py) ---------------------------------------
    def _check_netcdf_options(self):
        """[framework] netcdfinput / netcdfoutput / netcdfstatesinput / netcdfstatesoutput are served on the classic netCDF-3
        formats (ncio.py); static maps through netCDF are not part of this implementation."""
        for opt in ("netcdfstaticinput",):
            if configget(self.config, "framework", opt, "None") != "None":
                # (the reference's own reader of static maps cannot serve a map: netcdfinputstatic.gettimestep reads
                #  self.alldat, which that class never sets, wf_netcdfio.py:960-1013)
                raise ValueError("[framework] %s: static maps read from netCDF are not part of this implementation "
                                 "(forcing input, [outputmaps] / summary output and states are: netcdfinput, netcdfoutput, "
                                 "netcdfstaticoutput, netcdfstatesinput, netcdfstatesoutput)" % opt)
        self._ncin = self._ncout = None

    def _grid_xy(self):
        info = self._info
        return info.x_ul + (np.arange(info.ncol) + 0.5) * info.cellsize, info.ycoordinate()

    def _setup_netcdf(self):
        """Called at the end of initial(), when the grid is known."""
        from . import ncio
        x, y = self._grid_xy()
        ncfile = configget(self.config, "framework", "netcdfinput", "None")
        if ncfile != "None":
            names = [os.path.basename(v) for _, v in self.config.items("inputmapstacks")] if self.config.has_section("inputmapstacks") else []
            names += [os.path.basename(p["stack"]) for p in getattr(self, "_modelparameters", []) if p["type"] == "timeseries"]
            for dflt in ("P", "PET", "TEMP", "IF", "SE"):
                if dflt not in names:
                    names.append(dflt)
            self._ncin = ncio.NetcdfInput(os.path.join(self.Dir, ncfile), self.logger, names, x, y)
        ncout = configget(self.config, "framework", "netcdfoutput", "None")
        if ncout != "None":
            self._ncout = ncio.NetcdfOutput(
                os.path.join(self.SaveDir, ncout), self.logger, self.datetime_start, self.nrTimeSteps, self.timestepsecs, x, y,
                maxbuf=int(configget(self.config, "framework", "netcdfwritebuffer", "50")),
                metadata={"caseName": self.caseName, "runId": self.runId, "wflow_ini": self.configfile},
                fmt=configget(self.config, "framework", "netcdf_format", "NETCDF4"))

    def _runInitial(self):
        cfg = lambda k, d: configget(self.config, "model", k, d)
        self._check_netcdf_options()
        mapname = lambda k: cfg(k, "staticmaps/%s.map" % k)
        self.modelSnow = int(cfg("ModelSnow", "1"))
        self.soilInfReduction = int(cfg("soilInfRedu", "0"))
        self.TransferMethod = int(cfg("transfermethod", "0"))
        self.UST = int(cfg("Whole_UST_Avail", "0"))
        self.kinwaveIters = int(cfg("kinwaveIters", "0"))
        self.kinwaveRiverTstep = int(cfg("kinwaveRiverTstep", "0"))
        self.kinwaveLandTstep = int(cfg("kinwaveLandTstep", "0"))
        self.intbl = cfg("intbl", "intbl")
        sizeinmetres = int(configget(self.config, "layout", "sizeinmetres", "0"))
        WIMaxScale = float(cfg("WIMaxScale", "0.8"))
        thick = cfg("UStoreLayerThickness", "0")
        self.layer_thickness = tuple(float(t) for t in thick.split(",")) if thick != "0" else ()
        self.maxLayers = len(self.layer_thickness) + 1 if self.layer_thickness else 1

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
        landuse = np.where(landuse <= -999, active.astype(np.int32), landuse)
        soil = np.where(soil <= -999, active.astype(np.int32), soil)
        self.OutputLoc = np.where(gauges > 0, gauges, 0).astype(np.int32)
        self.LandUse, self.Soil = landuse, soil
        # lddmask(TopoLdd, TopoId): the network lives on the subcatchment cells; cells that drain out of
        # the mask become pits (wflow_sbm.py:1969)
        ldd = np.where(active & (ldd >= 1) & (ldd <= 9), ldd, 0).astype(np.uint8)
        ldd = _repair_ldd(ldd)
        self.ldd = ldd
        river_b = (np.nan_to_num(river.astype(np.float64), nan=0.0) > 0) & active

        base = {}
        T = lambda name, key=None: tbl.read_tbl_default(self.Dir, self.intbl, name, landuse, sub, soil,
                                                        TABLE_DEFAULTS[key or name])
        base["Cmax"] = T("MaxCanopyStorage")
        for n in ("CanopyGapFraction", "EoverR", "RootingDepth", "AirEntryPressure", "rootdistpar", "InfiltCapSoil",
                  "CapScale", "InfiltCapPath", "MaxLeakage", "PathFrac", "SoilThickness", "thetaR", "thetaS",
                  "KsatVer", "KsatHorFrac", "M", "N", "N_River", "WaterFrac"):
            base[n] = T(n)
        base["et_RefToPot"] = T("et_reftopot")
        soil_min = T("SoilMinThickness")
        if self.modelSnow:
            for n in ("TTI", "TT", "TTM", "Cfmax", "WHC", "w_soil"):
                base[n] = T(n)
            if self.soilInfReduction:
                base["cf_soil"] = T("cf_soil")
        for k in range(self.maxLayers):
            base["c_%d" % k] = tbl.read_tbl_default(self.Dir, self.intbl, "c", landuse, sub, soil, 10.0, layer=k)
            base["KsatVerFrac_%d" % k] = tbl.read_tbl_default(self.Dir, self.intbl, "KsatVerFrac", landuse, sub, soil, 1.0, layer=k)
        tc = self._staticmap(cfg("TemperatureCorrectionMap", "staticmaps/wflow_tempcor.map"))
        base["TempCor"] = np.zeros(self.shape, F32) if tc is None else np.nan_to_num(tc, nan=0.0).astype(F32)

        # cell sizes, wflow/pcrut.py:53-73
        if sizeinmetres:
            xl = yl = np.full(self.shape, info.cellsize, F32)
        else:
            latlen, longlen = lattometres(np.repeat(info.ycoordinate()[:, None], self.shape[1], 1).astype(F32))
            xl, yl = (longlen * F32(info.cellsize)).astype(F32), (latlen * F32(info.cellsize)).astype(F32)
        reallength = ((xl + yl) * F32(0.5)).astype(F32)
        base["xl"], base["yl"] = xl, yl
        slope_map = self._staticmap("staticmaps/Slope.map")
        slope = slope_map if slope_map is not None else slope_from_dem(np.where(active, dem, np.nan), info.cellsize)
        slope = np.maximum(F32(0.00001), np.nan_to_num(slope, nan=0.0) * F32(info.cellsize) / reallength).astype(F32)  # :1800
        base["landSlope"] = slope
        # soil thickness scaled by the topographic wetness index, wflow_sbm.py:1926-1933
        sched = Schedule(ldd)
        upcells = sched.catchment_total(None).astype(F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            WI = np.log(np.where(active, upcells / np.maximum(slope, F32(0.00001)), np.nan)).astype(F32)
            WIMax = np.zeros(self.shape, F32)
            for i in np.unique(self.TopoId[active]):
                m = self.TopoId == i
                WIMax[m] = np.nanmax(WI[m]) * F32(WIMaxScale)
            st = np.maximum(np.minimum(base["SoilThickness"], (WI / WIMax) * base["SoilThickness"]), soil_min)
        base["SoilThickness"] = np.where(active, st, base["SoilThickness"]).astype(F32)
        # river geometry, wflow_sbm.py:2044-2105
        rlf = self._staticmap(mapname("wflow_riverlength_fact"))
        base["RiverLengthFac"] = np.ones(self.shape, F32) if rlf is None else np.nan_to_num(rlf, nan=1.0).astype(F32)
        rl = self._staticmap(mapname("wflow_riverlength"))
        if rl is not None:
            base["DCL"] = np.nan_to_num(rl, nan=0.0).astype(F32)
        dl = params.drain_length(ldd, xl, yl)
        dcl = base.get("DCL", (dl * np.maximum(F32(1.0), base["RiverLengthFac"])).astype(F32))
        rslope = self._staticmap("staticmaps/RiverSlope.map")
        base["riverSlope"] = (slope * (dl / np.maximum(dcl, F32(1e-6)))).astype(F32) if rslope is None else rslope
        rw = self._staticmap(mapname("wflow_riverwidth"))
        rw = np.zeros(self.shape, F32) if rw is None else np.nan_to_num(rw, nan=0.0).astype(F32)
        if (rw[river_b] <= 0).any():  # Q-scaled width, wflow_sbm.py:2073-2086 (window average restated, unpinned)
            alf = float(cfg("Alpha", "60"))
            Qmax = float(cfg("AnnualDischarge", "300"))
            Qscale = upcells / max(float(upcells.max()), 1.0) * Qmax
            W = ((alf * (alf + 2.0) ** 0.6666666667) ** 0.375 * Qscale ** 0.375
                 * np.maximum(0.0001, np.nan_to_num(window_average(np.where(active, base["riverSlope"], np.nan), 4.0), nan=0.0001)) ** (-0.1875)
                 * base["N_River"] ** 0.375).astype(F32)
            rw = np.where(rw <= 0.0, W, rw).astype(F32)
        base["RiverWidth"] = rw
        bf = self._staticmap(mapname("wflow_bankfulldepth"))
        bf = np.ones(self.shape, F32) if bf is None else np.nan_to_num(bf, nan=0.0).astype(F32)
        base["BankfullDepth"] = np.where(river_b, bf, F32(0.0)).astype(F32)

        # [variable_change_once]: self.X = self.X * c on the parameters read so far (wf_multparameters in initial(),
        # wflow_sbm.py:1792; wf_DynamicFramework.py:651-690)
        self._read_variable_changes()
        for var, fac in self._changes_once:
            key = {"MaxCanopyStorage": "Cmax", "et_reftopot": "et_RefToPot"}.get(var, var)
            if key in base:
                base[key] = (base[key] * F32(fac)).astype(F32)
                self.logger.warning("Variable change (apply_once) applied to %s with factor %g" % (var, fac))
            else:
                self.logger.error("Variable change (apply_once) could not be applied to " + var)
        # [modelparameters] of static type replace the table-derived rasters before the derivations (:719-764)
        self._read_modelparameters()
        for par in self._modelparameters:
            if par["type"] in self._STATIC_TYPES and par["name"] in base:
                base[par["name"]] = self._parameter_raster(par, 0)
        self.static = params.derive_static(base, ldd, river_b, self.timestepsecs, cellsize=info.cellsize,
                                           n_layers=self.maxLayers)
        self.River = river_b
        self._base, self._cellsize = base, info.cellsize  # kept for parameter ensembles (ensemble.CaseEnsemble)
        it_l = it_r = 1
        if self.kinwaveIters == 1:  # wflow_sbm.py:2911-2923, 3159-3171 (fixed sub-steps only)
            if self.kinwaveLandTstep > 0:
                it_l = int(np.ceil(self.timestepsecs / self.kinwaveLandTstep))
            if self.kinwaveRiverTstep > 0:
                it_r = int(np.ceil(self.timestepsecs / self.kinwaveRiverTstep))
        # kinwaveIters = 1 without a fixed sub-step: estimated from the Courant number before every step
        self._estimate_l = self.kinwaveIters == 1 and self.kinwaveLandTstep == 0
        self._estimate_r = self.kinwaveIters == 1 and self.kinwaveRiverTstep == 0
        self._it_fixed = (it_l, it_r)
        self._requested = set()  # diagnostic outputs the library materialises for this run (outputmaps, csv / tss, summary ...)
        self._mod = SbmModel(ldd, river_b, schedule=sched, timestepsecs=self.timestepsecs,
                             layer_thickness=self.layer_thickness, model_snow=self.modelSnow,
                             soil_inf_redu=self.soilInfReduction, transfer_method=self.TransferMethod, ust=self.UST,
                             it_kin_l=it_l, it_kin_r=it_r, device=self.device)
        active_cells = np.asarray(active, bool) & np.isin(ldd, (1, 2, 3, 4, 5, 6, 7, 8, 9))
        for k, v in self.static.items():
            if k not in ("River", "Beta"):
                if np.ndim(v) == 2 and np.isnan(v[active_cells]).any():
                    # the reference's lookups log "Not all catchment cells have a value" (wflow_lib / pcrut readtbl checks)
                    self.logger.error("Not all catchment cells have a value for %s: %d undefined cells are set to 0"
                                      % (k, int(np.isnan(v[active_cells]).sum())))
                self._mod.set(k, np.where(np.isnan(v), F32(0.0), v))
        self._setup_optional_modules(cfg)
        self._update_parameters(0, initial=True)
        # initial() ends with wf_updateparameters(): the forcing of the first timestep is in place (and can be read
        # back through the API) before the first step
        saved, self._api_values = self._api_values, {}
        self._mod.set("P", self._forcing("P", "Precipitation", "/inmaps/P", 1, 0.0))
        self._mod.set("PET", self._forcing("PET", "EvapoTranspiration", "/inmaps/PET", 1, 0.0))
        if self.modelSnow:
            self._mod.set("TEMP", self._forcing("TEMP", "Temperature", "/inmaps/TEMP", 1, 10.0))
        self._api_values = saved
        self._setup_outputs()

    def _setup_optional_modules(self, cfg):
        """Modules of dynamic() that the ini or the case's maps switch on (wflow_sbm.py:1241, 1290, 1096-1135, 1636-1641)."""
        self._mod.set_option("maxitsupply", int(cfg("maxitsupply", "5")))
        if int(cfg("MassWasting", "0")):
            self._mod.set_option("MassWasting", 1)
        ids = lambda name: None if self._staticmap(name) is None else np.nan_to_num(self._staticmap(name), nan=0.0).astype(np.int32)  # noqa: E731
        paddy = ids("staticmaps/wflow_irrigationpaddyareas.map")
        areas, intakes = ids("staticmaps/wflow_irrigationareas.map"), ids("staticmaps/wflow_irrisurfaceintake.map")
        self._paddy = False
        if paddy is not None and (paddy > 0).any():  # :1116-1165, 1639-1641 (nrpaddyirri > 0)
            if areas is not None and (areas > 0).any():
                raise ValueError("paddy areas together with irrigation areas are not part of this implementation "
                                 "(the reference lets the paddy supply overwrite the other, wflow_sbm.py:3330-3346)")
            for fld, fname in (("h_max", "wflow_hmax.map"), ("h_min", "wflow_hmin.map"), ("h_p", "wflow_hp.map")):
                v = self._staticmap("staticmaps/" + fname)
                self._mod.set(fld, np.zeros(self.shape, F32) if v is None else np.nan_to_num(v, nan=0.0).astype(F32))
            n = self._mod.paddy_create(paddy, np.zeros(self.shape, np.int32) if intakes is None else intakes)
            self._abstractions = True
            self._paddy = True
            self.logger.info("%d paddy areas (CRPST: the crop-stage parameter of [modelparameters], 1 when absent)" % n)
            return
        if areas is not None and (areas > 0).any():
            returns = ids("staticmaps/wflow_irrisurfacereturns.map")
            if returns is not None and (returns > 0).any():
                raise ValueError("staticmaps/wflow_irrisurfacereturns.map: irrigation return points are not part of this implementation")
            if intakes is None:
                intakes = np.zeros(self.shape, np.int32)
            n = self._mod.irrigation_create(areas, intakes)
            self._abstractions = True
            self.logger.info("%d irrigation areas" % n)

    # ---- resume / suspend ------------------------------------------------------------------------
    def _state_names(self):
        out = []
        for s in STATE_FILES:
            if s == "UStoreLayerDepth":
                out += ["UStoreLayerDepth_%d" % k for k in range(self.maxLayers)]
            elif s in ("Snow", "SnowWater", "TSoil") and not self.modelSnow and False:
                continue
            else:
                out.append(s)
        if getattr(self, "_paddy", False):  # stateVariables(), wflow_sbm.py:1007-1009
            out.append("PondingDepth")
        return out

    def _runResume(self):
        if self.reinit == 1:
            state = params.cold_state(self.static, self.maxLayers)
            self._load_state(state)
        else:
            self.wf_resume(os.path.join(self.Dir, "instate"))

    def _load_state(self, state):
        st, neff = self.static["SoilThickness"], (self.static["thetaS"] - self.static["thetaR"]).astype(F32)
        if "SatWaterDepth" in state:  # zi is the live saturated-zone state, wflow_sbm.py:2512
            state = dict(state)
            state["zi"] = np.maximum(F32(0.0), st - state["SatWaterDepth"] / neff).astype(F32)
        for k, v in state.items():
            if k == "SatWaterDepth":
                continue
            self._mod.set(k, np.nan_to_num(np.where(v == F32(MV), np.nan, v), nan=0.0).astype(F32))

    def _get(self, name):
        """A model map by (library) field name.  SatWaterDepth is not a stored state here -- zi is the live saturated-zone
        state -- so unless an output asked the library to materialise it, it is what the reference forms from zi at the
        start of every step: (thetaS - thetaR) * (SoilThickness - zi), REAL4 (wflow_sbm.py:2751)."""
        if name == "SatWaterDepth" and "SatWaterDepth" not in self._requested:
            zi = self._mod.get("zi", MV)
            ok = zi != F32(MV)
            neff = (self.static["thetaS"] - self.static["thetaR"]).astype(F32)
            sat = (neff * (self.static["SoilThickness"] - zi)).astype(F32)
            return np.where(ok, sat, F32(MV)).astype(F32)
        return self._mod.get(name, MV)

    def _request(self, name):
        self._requested.add(name)
        self._mod.request(name)

    def wf_resume(self, directory):
        """Read the state maps (wf_DynamicFramework.py:2065-2129): list-valued states as Name_<k>.map.  With [framework]
        netcdfstatesinput the first time slice of the variables of that file, the map files for what it does not hold
        (wf_readmap in resume, :3419-3436; netcdfinputstates, wf_netcdfio.py:793-957)."""
        state = {}
        ncst = None
        ncfile = configget(self.config, "framework", "netcdfstatesinput", "None")
        if ncfile != "None":
            from . import ncio
            x, y = self._grid_xy()
            ncst = ncio.NetcdfInput(os.path.join(self.Dir, ncfile), self.logger, self._state_names(), x, y)
        for name in self._state_names():
            if ncst is not None:
                a, ok = ncst.gettimestep(1, var=name)
                if ok:
                    state[name] = np.nan_to_num(a, nan=0.0).astype(F32)
                    continue
            path = os.path.join(directory, name + ".map")
            if not os.path.isfile(path):
                # wf_readmap(path, 0.0) hands back the default map for a missing file (:2109)
                self.logger.warning("state file %s not found, using 0.0" % path)
                state[name] = np.zeros(self.shape, F32)
                continue
            a, _ = csf.read_map(path)
            state[name] = np.nan_to_num(a.astype(F32), nan=0.0)
        self._load_state(state)

    def wf_suspend(self, directory):
        """Write one REAL4 map per state (wf_DynamicFramework.py:1780-1827)."""
        os.makedirs(directory, exist_ok=True)
        for name in self._state_names():
            csf.write_map(os.path.join(directory, name + ".map"), self._get(name), self._info, mv=MV)

    def wf_savesummarymaps(self):
        """[summary]: end values; [summary_*]: the statistics gathered over the run (wf_DynamicFramework.py:1886-1951)."""
        os.makedirs(os.path.join(self.SaveDir, "outsum"), exist_ok=True)
        ncfile = configget(self.config, "framework", "netcdfstaticoutput", "None")
        nc = None
        if ncfile != "None":  # _reportStatic with a netCDF file (wf_DynamicFramework.py:3069-3090): one time slice, the run's end
            from . import ncio
            x, y = self._grid_xy()
            nc = ncio.NetcdfOutput(os.path.join(self.SaveDir, ncfile), self.logger, self.datetime_end, 1, self.timestepsecs, x, y,
                                   maxbuf=1, metadata={"caseName": self.caseName, "runId": self.runId},
                                   fmt=configget(self.config, "framework", "netcdf_format", "NETCDF4"))

        def save(path, a, longname):
            if nc is not None:
                nc.savetimestep(1, np.where(a == F32(MV), F32(np.nan), a), var=os.path.splitext(os.path.basename(path))[0], name=longname)
            else:
                csf.write_map(path, a, self._info, mv=MV)

        if self.config.has_section("summary"):
            for var, fname in self.config.items("summary"):
                name = var.replace("self.", "")
                try:
                    a = self._get(self._ALIAS.get(name, name))
                except Exception:
                    self.logger.warning("Could not find or save the configured summary map:" + var)
                    continue
                save(os.path.join(self.SaveDir, "outsum", fname), a, var)
        for st in getattr(self, "_stats", []):
            if st["count"] == 0:
                continue
            d = st["data"] / F32(st["count"]) if st["mode"] == "avg" else st["data"]
            save(st["path"], np.where(np.isnan(d), F32(MV), d).astype(F32), "self." + st["name"])
        if nc is not None:
            nc.finish()

    def _runSuspend(self):
        ncfile = configget(self.config, "framework", "netcdfstatesoutput", "None")
        if ncfile != "None":  # reportState with a netCDF state file (wf_DynamicFramework.py:3099-3119): one time slice, the run's end
            from . import ncio
            x, y = self._grid_xy()
            out = ncio.NetcdfOutput(os.path.join(self.SaveDir, ncfile), self.logger, self.datetime_end, 1, self.timestepsecs, x, y,
                                    maxbuf=1, metadata={"caseName": self.caseName, "runId": self.runId},
                                    fmt=configget(self.config, "framework", "netcdf_format", "NETCDF4"))
            for name in self._state_names():
                a = self._get(name)
                out.savetimestep(1, np.where(a == F32(MV), F32(np.nan), a), var=name, name=name)
            out.finish()
        else:
            self.wf_suspend(os.path.join(self.SaveDir, "outstate"))
        self.wf_savesummarymaps()
        if int(configget(self.config, "model", "OverWriteInit", "0")):
            self.wf_suspend(os.path.join(self.Dir, "instate"))

    def wf_QuickSuspend(self):
        self._mod.quick_suspend()

    def wf_QuickResume(self):
        self._mod.quick_resume()

    # ---- dynamic ----------------------------------------------------------------------------------
    # ---- [modelparameters] (wf_updateparameters, wf_DynamicFramework.py:708-874, 1624-1690) ---------------
    _STATIC_TYPES = ("staticmap", "statictbl", "tbl", "tblsparse")
    _DYNAMIC_TYPES = ("timeseries", "monthlyclim", "dailyclim", "tblmonthlyclim")
    # parameters that are library fields as they stand (no derived static depends on them)
    _DIRECT = ("LAI", "Sl", "Swood", "Kext", "RootingDepth", "Cmax", "CanopyGapFraction", "EoverR", "et_RefToPot",
               "PathFrac", "rootdistpar", "CapScale", "AirEntryPressure", "MaxLeakage", "WaterFrac", "TempCor",
               "Inflow", "DemandReturnFlowFraction", "CRPST", "h_p", "h_max", "h_min")

    def _read_modelparameters(self):
        """name = stack,type,default,verbose[,lookupmap_1,...]"""
        self._modelparameters = []
        self._par_loaded = {}
        if not self.config.has_section("modelparameters"):
            return
        for name, line in self.config.items("modelparameters"):
            vals = [v.strip() for v in line.split(",")]
            if len(vals) < 4:
                self.logger.error("Parameter line in ini not valid: " + line)  # :1688
                continue
            self._modelparameters.append(dict(name=name, stack=vals[0], type=vals[1], default=float(vals[2]),
                                              verbose=vals[3], lookupmaps=vals[4:]))

    def _climate_datetime(self, step):
        """DT.currentDateTime inside dynamic(): in 'steps' mode the clock starts at the state time, one step before
        starttime, and is advanced AFTER each step (wf_DynamicFramework.py:136-171, 2991-3021)."""
        return self.datetime_start + _dt.timedelta(seconds=self.timestepsecs * (max(step, 1) - 2))

    def _parameter_source(self, par, step):
        """(kind, path) the parameter is read from at this step."""
        t, stack = par["type"], par["stack"].lstrip("/")
        if t in ("tbl", "tblsparse", "statictbl"):
            return "tbl", os.path.join(self.Dir, stack)
        if t == "staticmap":
            return "map", os.path.join(self.Dir, stack)
        if t == "timeseries":
            return "map", generate_name_t(os.path.join(self.Dir, stack), max(step, 1))
        if t == "monthlyclim":
            return "map", generate_name_t(os.path.join(self.Dir, stack), self._climate_datetime(step).month)
        if t == "dailyclim":
            return "map", generate_name_t(os.path.join(self.Dir, stack), self._climate_datetime(step).timetuple().tm_yday)
        if t == "tblmonthlyclim":
            root, ext = os.path.splitext(stack)
            return "tbl", os.path.join(self.Dir, "%s_%d%s" % (root, self._climate_datetime(step).month, ext))
        return None, None

    def _parameter_raster(self, par, step):
        kind, path = self._parameter_source(par, step)
        default = np.float32(par["default"])
        if par["type"] == "timeseries" and getattr(self, "_ncin", None) is not None:  # wf_readmap: from the netCDF input file
            when = self.datetime_start + _dt.timedelta(seconds=self.timestepsecs * (max(step, 1) - 1))
            a, ok = self._ncin.gettimestep(max(step, 1), tsdatetime=when, var=os.path.basename(par["stack"]))
            return np.nan_to_num(a, nan=float(default)).astype(F32) if ok else np.full(self.shape, default, F32)
        if kind == "map":
            if os.path.isfile(path):
                a, _ = csf.read_map(path)
                return np.nan_to_num(a.astype(F32), nan=float(default)).astype(F32)  # pcr.cover(map, default)
            if par["verbose"] not in ("0", ""):
                self.logger.warning("map %s not found, using default %g" % (path, default))
            return np.full(self.shape, default, F32)
        if kind == "tbl":
            if not os.path.isfile(path):  # readtblFlexDefault: the default
                return np.full(self.shape, default, F32)
            if par["type"] == "statictbl":  # readtblDefault: landuse, subcatchment, soil
                keys = [self.LandUse, self.TopoId, self.Soil]
            else:                           # readtblFlexDefault: the listed lookup maps
                keys = []
                for lm in par["lookupmaps"]:
                    lp = os.path.join(self.Dir, lm)
                    if not os.path.isfile(lp):
                        return np.full(self.shape, default, F32)
                    keys.append(csf.read_map(lp)[0])
            a = tbl.lookupscalar(path, *keys)
            return np.nan_to_num(a, nan=float(default)).astype(F32)
        return None

    def _update_parameters(self, step, initial=False):
        """Static types once (initial), time-dependent types whenever their source file changes."""
        for par in self._modelparameters:
            name, t = par["name"], par["type"]
            if t in self._STATIC_TYPES:
                if not initial or name not in self._DIRECT:
                    continue
            elif t in self._DYNAMIC_TYPES:
                if name not in self._DIRECT:
                    if initial:
                        self.logger.warning("[modelparameters] %s: time-dependent values of this parameter are not supported" % name)
                    continue
            else:
                if initial:
                    self.logger.warning("[modelparameters] %s: type %s is not supported" % (name, t))
                continue
            src = self._parameter_source(par, step)
            if self._par_loaded.get(name) == src:
                continue
            self._par_loaded[name] = src
            a = self._parameter_raster(par, step)
            if name == "Inflow" and not getattr(self, "_abstractions", False) and float(np.nanmin(a)) < 0.0:
                # a negative Inflow is an abstraction: met from the river through the supply iteration (wflow_sbm.py:3203-3308)
                self._mod.set_option("Abstractions", 1)
                self._abstractions = True
            self._mod.set(name, a)

    def _read_variable_changes(self):
        """[variable_change_once] / [variable_change_timestep]: 'self.X = self.X * c' (wf_DynamicFramework.py:1691-1707)."""
        def section(name):
            out = []
            if self.config.has_section(name):
                for key, val in self.config.items(name):
                    try:
                        out.append((key.replace("self.", "").strip(), float(val.split("*")[1])))
                    except (IndexError, ValueError):
                        self.logger.error("Variable change not understood: %s = %s" % (key, val))
            return out
        self._changes_once = section("variable_change_once")
        # per timestep the reference multiplies the attribute right after the forcing corrections (wflow_sbm.py:2619);
        # for the forcing that is a factor on the raster handed to the library (the corrections are products themselves)
        self._changes_step = {}
        for var, fac in section("variable_change_timestep"):
            field = {"Precipitation": "P", "PotenEvap": "PET"}.get(var)
            if field is None or fac < 0:
                self.logger.error("Variable change (apply_timestep) could not be applied to " + var)
            else:
                self._changes_step[field] = self._changes_step.get(field, 1.0) * fac

    def _forcing(self, var, stack_key, default_stack, step, default):
        a = self._forcing_raw(var, stack_key, default_stack, step, default)
        fac = getattr(self, "_changes_step", {}).get(var)
        return a if fac is None else (a * F32(fac)).astype(F32)

    def _forcing_raw(self, var, stack_key, default_stack, step, default):
        if var in self._api_values:  # set through the API: skip the disk for this step (:3300-3306)
            return self._api_values.pop(var)
        prefix = configget(self.config, "inputmapstacks", stack_key, default_stack).lstrip("/")
        if getattr(self, "_ncin", None) is not None:  # wf_readmap with a netCDF input file (wf_DynamicFramework.py:3340-3353)
            when = self.datetime_start + _dt.timedelta(seconds=self.timestepsecs * (max(step, 1) - 1))  # DT.nextDateTime
            a, ok = self._ncin.gettimestep(max(step, 1), tsdatetime=when, var=os.path.basename(prefix))
            return np.nan_to_num(a, nan=default).astype(F32) if ok else np.full(self.shape, default, F32)
        path = generate_name_t(os.path.join(self.Dir, prefix), step)
        if os.path.isfile(path):
            a, _ = csf.read_map(path)
            return np.nan_to_num(a.astype(F32), nan=default)
        return np.full(self.shape, default, F32)

    def _runDynamic(self, firststep, laststep):
        """The time loop (wf_DynamicFramework.py:2968-3042); steps are 1-based like the map stacks."""
        first = max(int(firststep), 1) if firststep else self._currentTimeStep + 1
        last = int(laststep) if laststep else self.nrTimeSteps
        for step in range(first, last + 1):
            self._currentTimeStep = step
            self._update_parameters(step)
            self._mod.set("P", self._forcing("P", "Precipitation", "/inmaps/P", step, 0.0))
            self._mod.set("PET", self._forcing("PET", "EvapoTranspiration", "/inmaps/PET", step, 0.0))
            if self.modelSnow:
                self._mod.set("TEMP", self._forcing("TEMP", "Temperature", "/inmaps/TEMP", step, 10.0))
            self.wf_QuickSuspend()
            if self._estimate_l or self._estimate_r:  # wflow_sbm.py:2911-2923, 3159-3171
                it_l = self._mod.estimate_iterations("land") if self._estimate_l else self._it_fixed[0]
                it_r = self._mod.estimate_iterations("river") if self._estimate_r else self._it_fixed[1]
                # the library holds sub-step buffers for at most 1024 sub-steps: a Courant estimate beyond that (a fine grid
                # under a long timestep) is clamped with a warning where the reference would run ever more sub-steps
                if max(it_l, it_r) > 1024:
                    self.logger.warning("kinematic-wave sub-steps clamped to 1024 (Courant estimate: land %d, river %d)" % (it_l, it_r))
                    it_l, it_r = min(it_l, 1024), min(it_r, 1024)
                if (it_l, it_r) != getattr(self, "_it_now", None):  # (a change re-sizes the sweep: only when it is one)
                    self._mod.set_substeps(it_l, it_r)
                    self._it_now = (it_l, it_r)
            self._mod.step()
            self._write_outputs(step)

    def run_pipelined(self, firststep=None, laststep=None, gauges=None, chunk=64):
        """Extension without a counterpart in the reference: the same timesteps as _runDynamic, but `chunk`
        of them per kernel launch, timestep s+1 following s down the drainage network (SbmModel.step_pipelined;
        bit-identical states).  The forcing of a chunk is read ahead from the map stacks.  Per-step outputs
        other than the river discharge cannot leave a pipelined launch, so [outputcsv] / [outputtss] /
        [outputmaps] are NOT written; returns RiverRunoff [steps, gauges] at `gauges` (flat cell indices;
        default: the cells of the gauges map), each row what _runDynamic would have sampled after that step.
        Needs fixed kinematic-wave sub-steps (kinwaveIters = 0, or kinwaveRiverTstep / kinwaveLandTstep set)."""
        if self._estimate_l or self._estimate_r:
            raise ValueError("run_pipelined needs fixed kinematic-wave sub-steps (they are re-estimated per step otherwise)")
        first = max(int(firststep), 1) if firststep else self._currentTimeStep + 1
        last = int(laststep) if laststep else self.nrTimeSteps
        if gauges is None:
            g = getattr(self, "OutputLoc", None)
            gauges = np.flatnonzero(np.asarray(g).ravel() > 0) if g is not None else np.zeros(0, np.int64)
        gauges = np.ascontiguousarray(gauges, dtype=np.int64)
        handle = self._mod.gauges_create(gauges)
        chunk = max(1, min(int(chunk), 64, last - first + 1))  # a launch holds at most 64 steps
        self._mod.pipeline_reserve(chunk)
        self._mod.pipeline_capture(handle)
        rows = []
        for lo in range(first, last + 1, chunk):
            hi = min(lo + chunk - 1, last)
            for step in range(lo, hi + 1):
                self._mod.pipeline_set_forcing("P", step - lo, self._forcing("P", "Precipitation", "/inmaps/P", step, 0.0))
                self._mod.pipeline_set_forcing("PET", step - lo, self._forcing("PET", "EvapoTranspiration", "/inmaps/PET", step, 0.0))
                if self.modelSnow:
                    self._mod.pipeline_set_forcing("TEMP", step - lo, self._forcing("TEMP", "Temperature", "/inmaps/TEMP", step, 10.0))
            out = np.zeros((hi - lo + 1, max(gauges.size, 1)), F32)
            self._mod.step_pipelined(hi - lo + 1, out, mv=MV)
            rows.append(out[:, : gauges.size])
            self._currentTimeStep = hi
        return np.concatenate(rows, axis=0) if rows else np.zeros((0, gauges.size), F32)

    def currentTimeStep(self):
        return self._currentTimeStep

    # ---- API (wf_DynamicFramework.py:2203-2327, 2582-2760) -------------------------------------------------
    _ALIAS = {"Precipitation": "P", "PotenEvap": "PET", "Temperature": "TEMP"}

    def wf_supplyVariableNamesAndRoles(self):
        """[API] section: name = role,unit (0 in, 1 out, 2 state, 3 parameter)."""
        res = []
        if self.config.has_section("API"):
            for k, v in self.config.items("API"):
                role, unit = (v.split(",") + ["-"])[:2]
                res.append([k, int(role), unit])
        return res

    def wf_supplyMapAsNumpy(self, name):
        """float32 (nrow, ncol), flipped south-up, -999 for missing (wf_DynamicFramework.py:2582-2613);
        Name_k addresses element k of a list-valued variable."""
        if name in ("timestepsecs",):
            return np.array(getattr(self, name))
        field = self._ALIAS.get(name, name)
        try:
            a = self._get(field)
        except Exception:
            self.logger.error("%s is not defined in the model" % name)
            return None
        return np.flipud(a).copy()

    def wf_setValuesAsNumpy(self, name, values):
        field = self._ALIAS.get(name, name)
        a = np.flipud(np.asarray(values, dtype=F32)).copy()
        a = np.where(a == F32(MV), F32(0.0), a)
        if field in ("P", "PET", "TEMP"):
            self._api_values[field] = a
        else:
            self._mod.set(field, a)

    def wf_setValues(self, name, values):
        """Scalar -> uniform map; list/array -> map WITHOUT the south-up flip (:2284-2327)."""
        field = self._ALIAS.get(name, name)
        v = np.asarray(values, dtype=F32)
        a = np.full(self.shape, float(v), F32) if v.size == 1 else v.reshape(self.shape)
        if field in ("P", "PET", "TEMP"):
            self._api_values[field] = a
        else:
            self._mod.set(field, a)

    # ---- outputs ----------------------------------------------------------------------------------
    def _setup_outputs(self):
        """[outputcsv_N] / [outputtss_N]: samplemap, function, self.Var = file (wf_DynamicFramework.py:1709-1778)."""
        self._setup_netcdf()
        for kind in ("csv", "tss"):
            n = 0
            while self.config.has_section("output%s_%d" % (kind, n)):
                sec = "output%s_%d" % (kind, n)
                n += 1
                samplemap = self.config.get(sec, "samplemap")
                func = configget(self.config, sec, "function", "average")
                tformat = configget(self.config, sec, "timeformat", "steps")
                ids, _ = csf.read_map(os.path.join(self.Dir, samplemap))
                ids = np.where(ids > 0, ids, 0).astype(np.int32)
                # the reference samples np.unique(pcr2numpy(area, 0)): missing cells show up as a first column "0"
                zero_col = bool((ids == 0).any())
                area, uniq = self._mod.area_create(ids)
                for var, fname in self.config.items(sec):
                    if var in ("samplemap", "function", "timeformat"):
                        continue
                    terms = [t.strip().replace("self.", "") for t in var.split("+")]
                    try:
                        for t in terms:
                            self._request(self._ALIAS.get(t, t))
                    except Exception as ex:
                        self.logger.warning("output %s skipped: %s" % (var, ex))
                        continue
                    self._outputs.append(dict(kind=kind, area=area, ids=uniq, func=func, terms=terms, zero_col=zero_col,
                                              idmap=ids, tformat=tformat, path=os.path.join(self.SaveDir, fname)))
        # [summary_sum] / [summary_avg] / [summary_min] / [summary_max]: wf_sumavg, wf_DynamicFramework.py:362-401, 1596-1617
        self._stats = []
        for mode in ("sum", "avg", "min", "max"):
            sec = "summary_" + mode
            if not self.config.has_section(sec):
                continue
            for var, fname in self.config.items(sec):
                if "self." not in var:
                    raise ValueError("Entry in ini invalid: " + var)
                name = var.split("self.")[1]
                try:
                    self._request(self._ALIAS.get(name, name))
                    self._mod.device_ptr(self._ALIAS.get(name, name))
                except Exception as ex:
                    self.logger.warning("summary map %s skipped: %s" % (var, ex))
                    continue
                self._stats.append(dict(name=name, mode=mode, data=None, count=0,
                                        path=os.path.join(self.SaveDir, "outsum", fname)))
        self._outmaps = []
        if self.config.has_section("outputmaps"):
            for var, prefix in self.config.items("outputmaps"):
                name = var.replace("self.", "")
                try:
                    self._request(self._ALIAS.get(name, name))
                    self._mod.device_ptr(self._ALIAS.get(name, name))
                except Exception as ex:
                    self.logger.warning("output map %s skipped: %s" % (var, ex))
                    continue
                self._outmaps.append((name, prefix))

    def _write_outputs(self, step):
        for o in self._outputs:
            vals = None
            if o["func"] == "majority":  # pcr.areamajority of the (summed) raster: host side, rarely used
                a = sum(self._get(self._ALIAS.get(t, t)).astype(F32) for t in o["terms"])
                vals = area_majority(a, o["idmap"], o["ids"])
            else:
                for t in o["terms"]:
                    v = self._mod.area_reduce(o["area"], self._ALIAS.get(t, t), o["func"])
                    vals = v if vals is None else vals + v
            ids = ([0] if o["zero_col"] else []) + [int(i) for i in o["ids"]]
            vals = ([0.0] if o["zero_col"] else []) + [float(x) for x in vals]
            fh = self._files.get(o["path"])
            if fh is None:
                fh = self._files[o["path"]] = open(o["path"], "w")
                if o["kind"] == "csv":
                    fh.write("# Timestep," + ",".join(str(i) for i in ids) + "\n")
                else:  # PCRaster timeseries header
                    fh.write("timeseries scalar\n%d\ntimestep\n" % (len(ids) + 1))
                    fh.write("".join("%d\n" % i for i in ids))
            sep = "," if o["kind"] == "csv" else " "
            # row label: currentTimeStep - 1, or with timeformat = datetime the clock of dynamic() (:493-496, 1879-1884)
            label = str(self._climate_datetime(step)) if o["tformat"] == "datetime" else str(step - 1)
            fh.write(label + sep + sep.join(repr(x) for x in vals) + "\n")
        for st in self._stats:  # add_one, after every timestep (:3010-3012)
            a = self._get(self._ALIAS.get(st["name"], st["name"]))
            a = np.where(a == F32(MV), F32(np.nan), a).astype(F32)
            if st["count"] == 0:
                st["data"] = a
            elif st["mode"] in ("sum", "avg"):
                st["data"] = (st["data"] + a).astype(F32)
            elif st["mode"] == "max":
                st["data"] = np.where(np.isnan(st["data"]) | np.isnan(a), F32(np.nan), np.maximum(st["data"], a))
            else:
                st["data"] = np.where(np.isnan(st["data"]) | np.isnan(a), F32(np.nan), np.minimum(st["data"], a))
            st["count"] += 1
        for name, prefix in self._outmaps:
            a = self._get(self._ALIAS.get(name, name))
            if getattr(self, "_ncout", None) is not None:  # _reportNew with a netCDF output file (:3179-3189)
                self._ncout.savetimestep(step, np.where(a == F32(MV), F32(np.nan), a), var=os.path.basename(prefix), name="self." + name)
                continue
            path = generate_name_t(os.path.join(self.SaveDir, "outmaps", prefix), step)
            csf.write_map(path, a, self._info, mv=MV)

    def _wf_shutdown(self):
        for fh in self._files.values():
            fh.close()
        self._files = {}
        if getattr(self, "_ncout", None) is not None:
            self._ncout.finish()
        if getattr(self, "_ncin", None) is not None:
            self._ncin.close()
            self._ncin = None
        with open(os.path.join(self.SaveDir, "runinfo", "configofrun.ini"), "w") as fh:
            self.config.write(fh)
        if self._mod is not None:
            self._mod.close()
            self._mod = None


def _repair_ldd(ldd):
    """Cells whose receiver is missing or outside the map become pits (pcr.lddrepair / lddmask)."""
    from .synth import LDD_DC, LDD_DR
    nrow, ncol = ldd.shape
    out = ldd.copy()
    rr, cc = np.nonzero((ldd >= 1) & (ldd <= 9) & (ldd != 5))
    for r, c in zip(rr, cc):
        k = int(ldd[r, c])
        r2, c2 = r + LDD_DR[k], c + LDD_DC[k]
        if not (0 <= r2 < nrow and 0 <= c2 < ncol) or ldd[r2, c2] == 0:
            out[r, c] = 5
    return out


def main(argv=None, model_class=None, configfile="wflow_sbm.ini"):
    """python -m wflow_b200.framework -C case [-R runId] [-c ini] ... -- the reference's command line for the path
    (wflow/wflow_sbm.py:3531-3645): -C case, -R run id, -c ini file, -S / -T start / end time ("YYYY-MM-DD HH:MM:SS"),
    -s timestepsecs, -I cold start (reinit), -i intbl directory, -X overwrite instate with the end state,
    -P / -p 'self.X=self.X * c' as [variable_change_once] / [variable_change_timestep], -l log level, -f overwrite the run.
    Options that address subsystems out of scope here (-x, -M, -u, -U, -E, -W, -O, -v, -L) are accepted and ignored."""
    import getopt
    import sys
    args = sys.argv[1:] if argv is None else argv
    if "-h" in args:
        print(main.__doc__)
        return 0
    opts, _ = getopt.getopt(args, "XL:hC:Ii:v:S:T:WR:u:s:EP:p:x:U:fOc:l:M")
    o = dict(opts)
    m = (model_class or WflowSbm)(Dir=o.get("-C", "."), RunDir=o.get("-R", "run_default"), configfile=o.get("-c", configfile))

    def configset(section, var, value):
        if not m.config.has_section(section):
            m.config.add_section(section)
        m.config.set(section, var, value)

    for k, a in opts:
        if k in ("-P", "-p"):
            left, right = a.split("=", 1)
            configset("variable_change_once" if k == "-P" else "variable_change_timestep", left.strip(), right.strip())
        elif k == "-X":
            configset("model", "OverWriteInit", "1")
        elif k == "-I":
            configset("run", "reinit", "1")
        elif k == "-i":
            configset("model", "intbl", a)
        elif k == "-s":
            configset("run", "timestepsecs", a)
        elif k == "-T":
            configset("run", "endtime", a)
        elif k == "-S":
            configset("run", "starttime", a)
        elif k == "-l":
            m.logger.setLevel(getattr(logging, a, logging.INFO))
        elif k in ("-x", "-M", "-u", "-U", "-E", "-W", "-O", "-v", "-L"):
            m.logger.warning("option %s addresses a part of the reference that is out of scope here: ignored" % k)
    m.createRunId()
    m._runInitial()
    m._runResume()
    m._runDynamic(1, m.nrTimeSteps)
    m._runSuspend()
    m._wf_shutdown()
    return 0


if __name__ == "__main__":
    main()
