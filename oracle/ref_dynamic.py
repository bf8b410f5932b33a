"""Run the REFERENCE's own ``WflowModel.dynamic()`` -- unmodified, PCRaster half included -- on the inputs
of an oracle-shaped case (TEST INFRASTRUCTURE ONLY; needs /root/reference, i.e. the build container).

``wflow/wflow_sbm.py`` is loaded where it lies by oracle/refload.py with ``pcraster`` = the float32-numpy
shim of oracle/pcrshim.py.  This module only does what ``initial()`` / ``resume()`` and the framework do
before the first timestep: it hands the model its maps (wflow_sbm.py:1168-2394 builds them from files and
tables -- the framework mirror's job, not the hot path's) and the numpy mirrors ``static`` / ``layer`` / ``dyn``
exactly as :2176-2380 lay them out, then calls the reference's ``dynamic`` once per timestep.  Everything a
timestep computes -- interception, snow, glacier, evaporation split, sbm_cell, river inflow glue, kin_wave,
abstraction iteration, reservoirs, lakes, the water-balance and cumulative outputs -- is reference code.
"""
import numpy as np

from . import pcrshim as pcr
from . import refload

F32 = np.float32
MV = -999.0

STATE_MAPS = ["CanopyStorage", "Snow", "SnowWater", "TSoil", "zi", "SubsurfaceFlow", "LandRunoff", "LandRunoffsub",
              "WaterLevelL", "WaterLevelLsub", "RiverRunoff", "RiverRunoffsub", "WaterLevelR", "WaterLevelRsub"]
CUMULATIVE = ["CumOutFlow", "CumActInfilt", "CumCellInFlow", "CumPrec", "CumEvap", "CumPotenEvap", "CumInt",
              "CumLeakage", "CumExfiltWater", "CumSurfaceWater"]


def _model_class():
    _, sbm = refload.load()

    class Model(sbm.WflowModel):
        """The reference model with the framework's per-timestep services replaced by in-memory ones."""

        def __init__(self):  # the reference's __init__ wants a case directory and a clone map file
            self.UStoreLayerDepth = []

        def wf_updateparameters(self):  # wf_DynamicFramework.py:708-874: forcing of the timestep
            f = self._forcing
            self.Precipitation = self._map(f["P"])
            self.PotenEvap = self._map(f["PET"])
            self.Temperature = self._map(f["TEMP"]) if f.get("TEMP") is not None else self.ZeroMap
            self.Inflow = self._map(f["Inflow"]) if f.get("Inflow") is not None else self.ZeroMap
            if f.get("LAI") is not None:
                self.LAI = self._map(f["LAI"])

        def wf_multparameters(self):
            return None

        def wf_supplyJulianDOY(self):
            return int(self._forcing.get("JDOY", 1))

        def _map(self, a, vs=pcr.Scalar):
            a = np.array(np.broadcast_to(np.asarray(a), self.shape), dtype=np.float64)
            a[~self._active] = MV
            return pcr.numpy2pcr(vs, a, MV)

    return Model, sbm


class RefDynamic:
    """fields: dict of float32 rasters by reference attribute name (statics and states, the names of
    wflow_b200/csrc/fields.h; ``c_k`` / ``KsatVerFrac_k`` / ``UStoreLayerDepth_k`` per layer)."""

    def __init__(self, ldd, river, fields, *, timestepsecs=86400, layer_thickness=(), modelSnow=1, soilInfRedu=0,
                 transferMethod=0, ust=0, kinwaveIters=0, kinwaveLandTstep=0, kinwaveRiverTstep=0, massWasting=0,
                 topo_id=None, extra=None, ldd_org=None, reservoirs=None, lakes=None):
        Model, self.sbm = _model_class()
        funcs, _ = refload.load()
        ldd = np.asarray(ldd)
        self.shape = ldd.shape
        pcr.set_shape(self.shape)
        active = (ldd >= 1) & (ldd <= 9)
        m = self.m = Model()
        m._active = active
        m.shape = self.shape
        m.mv = MV
        m._forcing = {}
        mp = m._map
        fl = {k: np.asarray(v, F32) for k, v in fields.items()}
        # ---- options (initial(), wflow_sbm.py:1196-1290) ----
        m.timestepsecs = int(timestepsecs) if float(timestepsecs).is_integer() else float(timestepsecs)
        m.basetimestep = 86400
        m.modelSnow = int(modelSnow)
        m.soilInfReduction = int(soilInfRedu)
        m.MassWasting = int(massWasting)
        m.TransferMethod = int(transferMethod)
        m.UST = int(ust)
        m.kinwaveIters = int(kinwaveIters)
        m.kinwaveLandTstep = kinwaveLandTstep
        m.kinwaveRiverTstep = kinwaveRiverTstep
        m.nrresSimple = m.nrlake = m.nrpaddyirri = m.nrirri = 0
        m.filterResArea = 0
        m.updating = 0
        m.maxitsupply = 5
        m.DemandReturnFlowFraction = pcr.scalar(0.0)  # defaults of the irrigation maps: none (:1096-1135, 1481-1482)
        nlay = len(layer_thickness)
        m.maxLayers = nlay + 1 if nlay > 0 else 1  # :1757-1766
        nrLayers = nlay if nlay > 0 else 1
        # ---- maps ----
        zero = np.zeros(self.shape, F32)
        m.ZeroMap = mp(zero)
        lddf = np.where(active, ldd, 0).astype(np.float64)
        lddf[~active] = MV
        m.TopoLdd = pcr.numpy2pcr(pcr.Ldd, lddf, MV)
        m.TopoLddOrg = m.TopoLdd
        m.TopoId = mp(np.ones(self.shape) if topo_id is None else topo_id, pcr.Ordinal)
        for k in ["Cmax", "EoverR", "CanopyGapFraction", "et_RefToPot", "TempCor", "TTI", "TT", "TTM", "Cfmax", "WHC",
                  "w_soil", "cf_soil", "WaterFrac", "SoilThickness", "thetaS", "thetaR", "KsatVer", "f", "InfiltCapSoil",
                  "InfiltCapPath", "PathFrac", "CapScale", "rootdistpar", "RootingDepth", "AirEntryPressure", "MaxLeakage",
                  "xl", "yl", "KsatHorFrac", "landSlope", "DL", "DW", "SW", "AlphaL", "GlacierFrac", "G_TT", "G_Cfmax",
                  "G_SIfrac", "Sl", "Swood", "Kext"]:
            if k in fl:
                setattr(m, k, mp(fl[k]))
        riv = (np.asarray(river) != 0) & active
        m.River = mp(riv.astype(F32))
        # river-only parameters: the reference holds them on every cell (maps); outside the river any value does
        for k, dflt in [("RiverFrac", 0.0), ("AlphaR", 1.0), ("DCL", 1.0), ("Bw", 1.0)]:
            a = np.array(np.broadcast_to(fl.get(k, F32(dflt)), self.shape), F32)
            if k == "RiverFrac":
                a = np.where(riv, a, F32(0))
            setattr(m, k, mp(a))
        m.Beta = pcr.scalar(0.6)  # :1643
        m.neff = m.thetaS - m.thetaR  # :2158
        m.SoilWaterCapacity = m.SoilThickness * (m.thetaS - m.thetaR)  # :1935
        m.IRSupplymm = m.ZeroMap
        m.c = [mp(fl["c_%d" % k]) for k in range(m.maxLayers)]
        m.KsatVerFrac = [mp(fl["KsatVerFrac_%d" % k]) for k in range(m.maxLayers)]
        # :1985-1998
        m.QMMConv = m.timestepsecs / (m.xl * m.yl * 0.001)
        temp = pcr.catchmenttotal(pcr.cover(1.0), m.TopoLdd) * m.xl * 0.001 * 0.001 * m.yl
        m.QMMConvUp = pcr.cover(m.timestepsecs * 0.001) / temp
        m.ToCubic = (m.xl * m.yl * 0.001) / m.timestepsecs
        # ---- states (resume(), :2443-2516) ----
        for k in STATE_MAPS:
            setattr(m, k, mp(fl.get(k, zero)))
        if "GlacierStore" in fl:
            m.GlacierStore = mp(fl["GlacierStore"])
        m.UStoreLayerDepth = [mp(fl.get("UStoreLayerDepth_%d" % k, zero)) for k in range(m.maxLayers)]
        m.SatWaterDepth = (m.thetaS - m.thetaR) * (m.SoilThickness - m.zi)
        m.KinWaveVolumeL = m.WaterLevelLsub * m.SW * m.DL
        m.OldKinWaveVolumeL = m.KinWaveVolumeL
        m.KinWaveVolumeR = m.WaterLevelRsub * m.Bw * m.DCL
        m.OldKinWaveVolumeR = m.KinWaveVolumeR
        m.IrrigationSurfaceIntakes = m.IrrigationSurfaceReturn = m.ZeroMap
        m.DemandReturnFlowFraction = m.ZeroMap
        for k in CUMULATIVE:
            setattr(m, k, m.ZeroMap)
        m.count = 0
        # ---- reservoirs / lakes (initial(), :1804-1907): id maps are missing outside the objects / their areas ----
        def idmap(a):
            a = np.asarray(a, np.float64).copy()
            a[(a <= 0) | ~active] = MV
            return pcr.numpy2pcr(pcr.Nominal, a, MV)
        if reservoirs or lakes:
            lo = np.where(active, np.asarray(ldd if ldd_org is None else ldd_org), 0).astype(np.float64)
            lo[lo == 0] = MV
            m.TopoLddOrg = pcr.numpy2pcr(pcr.Ldd, lo, MV)
            m.filter_P_PET = m.ZeroMap + 1.0
        if reservoirs:
            m.ReserVoirSimpleLocs, m.ReservoirSimpleAreas = idmap(reservoirs["locs"]), idmap(reservoirs["areas"])
            m.nrresSimple = int((np.asarray(reservoirs["locs"]) > 0).sum())
            for k in ("ResSimpleArea", "ResMaxVolume", "ResTargetFullFrac", "ResMaxRelease", "ResDemand", "ResTargetMinFrac"):
                setattr(m, k, mp(reservoirs[k]))
            m.ReservoirVolume = mp(reservoirs["ReservoirVolume"]) if "ReservoirVolume" in reservoirs else m.ResMaxVolume * m.ResTargetFullFrac
            res_area = pcr.cover(pcr.scalar(m.ReservoirSimpleAreas), 0.0)
            m.filter_P_PET = pcr.ifthenelse(pcr.boolean(pcr.cover(res_area, pcr.scalar(0.0))), res_area * 0.0, m.filter_P_PET)
        if lakes:
            m.LakeLocs, m.LakeAreasMap = idmap(lakes["locs"]), idmap(lakes["areas"])
            m.LinkedLakeLocs = idmap(lakes.get("linked", np.zeros(self.shape)))
            m.nrlake = int((np.asarray(lakes["locs"]) > 0).sum())
            for k in ("LakeArea", "LakeThreshold", "LakeStorFunc", "LakeOutflowFunc", "Lake_b", "Lake_e", "LakeWaterLevel"):
                setattr(m, k, mp(lakes[k]))
            m.sh, m.hq = dict(lakes.get("sh", {})), dict(lakes.get("hq", {}))
            lake_area = pcr.cover(pcr.scalar(m.LakeAreasMap), 0.0)
            m.filter_P_PET = pcr.ifthenelse(lake_area > 0, lake_area * 0.0, m.filter_P_PET)
        if reservoirs or lakes:
            m.filterResArea = pcr.pcr2numpy(m.filter_P_PET, 1.0).min()
        for k, v in (extra or {}).items():
            setattr(m, k, v)
        # ---- numpy mirrors, :2176-2380 ----
        n = self.shape[0] * self.shape[1]
        g = lambda x: pcr.pcr2numpy(x, MV).ravel()  # noqa: E731
        static_names = ["KsatVer", "KsatHorFrac", "f", "thetaS", "thetaR", "SoilThickness", "InfiltCapSoil", "PathFrac",
                        "InfiltCapPath", "cf_soil", "neff", "landSlope", "riverSlope", "SoilWaterCapacity", "MaxLeakage",
                        "CanopyGapFraction", "CapScale", "rootdistpar", "River", "DL", "DW", "SW", "Beta", "DCL",
                        "reallength", "RootingDepth", "AirEntryPressure", "ActRootingDepth", "ssfmax", "xl", "yl", "h_p",
                        "Bw", "AlpPow", "AlpTermR", "AlpTermL", "AlphaL", "AlphaR"]
        st = np.zeros(n, dtype=np.dtype([(k, np.float64) for k in static_names]))
        for k in ["KsatVer", "KsatHorFrac", "f", "thetaS", "thetaR", "SoilThickness", "InfiltCapSoil", "InfiltCapPath",
                  "PathFrac", "neff", "landSlope", "MaxLeakage", "SoilWaterCapacity", "CanopyGapFraction", "CapScale",
                  "rootdistpar", "River", "DL", "DW", "SW", "Beta", "DCL", "xl", "yl", "RootingDepth",
                  "AirEntryPressure", "Bw", "AlphaR", "AlphaL"]:
            st[k] = g(getattr(m, k))
        if m.modelSnow & m.soilInfReduction:
            st["cf_soil"] = g(m.cf_soil)
        with np.errstate(all="ignore"):
            st["ssfmax"] = ((st["KsatHorFrac"] * st["KsatVer"] * st["landSlope"]) / st["f"]
                            * (np.exp(0) - np.exp(-st["f"] * st["SoilThickness"])))
        m.static = st
        layer_names = ["c", "UStoreLayerDepth", "st", "KsatVerFrac", "ActEvapUstore", "vwc", "vwc_perc",
                       "UStoreLayerThickness", "UStest"]
        layer = np.zeros((m.maxLayers, n), dtype=np.dtype([(k, np.float64) for k in layer_names]))
        for k in range(m.maxLayers):
            layer["c"][k] = g(m.c[k])
            layer["KsatVerFrac"][k] = g(m.KsatVerFrac[k])
        np_zeros = np.zeros(n)  # layer thickness, :2301-2326
        sumT = np_zeros
        for k in range(m.maxLayers):
            sumL = sumT
            if nrLayers > 1 and k < nrLayers:
                tmp = float(layer_thickness[k]) + np_zeros
                th = np.minimum(tmp, np.maximum(st["SoilThickness"] - sumL, 0.0))
            else:
                tmp = np.maximum(st["SoilThickness"] - sumL, 0.0)
                th = np.maximum(st["SoilThickness"] - sumL, 0.0)
            sumT = tmp + sumT
            layer["UStoreLayerThickness"][k] = th
        m.layer = layer
        from .ref_driver import DYN_FIELDS
        m.dyn = np.zeros(n, dtype=np.dtype([(k, np.float64) for k in DYN_FIELDS]))
        m.np_ldd = lddf
        m.nodes, m.nodes_up = funcs.set_dd(lddf)
        lr = lddf.copy()
        lr.flat[st["River"] == 0] = MV  # :2387-2388
        m.rnodes, m.rnodes_up = funcs.set_dd(lr)
        self.active = active
        self.it_kinL = self.it_kinR = None

    # ------------------------------------------------------------------------------------------------
    def set_state(self, name, a):
        m = self.m
        if name.startswith("UStoreLayerDepth_"):
            m.UStoreLayerDepth[int(name.rsplit("_", 1)[1])] = m._map(a)
        else:
            setattr(m, name, m._map(a))
        if name in ("WaterLevelLsub", "WaterLevelRsub"):  # resume(): the kinematic-wave volumes follow the levels
            m.KinWaveVolumeL = m.WaterLevelLsub * m.SW * m.DL
            m.KinWaveVolumeR = m.WaterLevelRsub * m.Bw * m.DCL

    def step(self, P, PET, TEMP=None, Inflow=None, LAI=None, JDOY=1):
        self.m._forcing = dict(P=P, PET=PET, TEMP=TEMP, Inflow=Inflow, LAI=LAI, JDOY=JDOY)
        self.m.dynamic()

    def get(self, name):
        """float32 raster of a model attribute (MV = -999), list attributes as ``name_k``."""
        m = self.m
        base, _, idx = name.rpartition("_")
        if base in ("UStoreLayerDepth", "vwc", "vwc_perc") and idx.isdigit():
            v = getattr(m, base)[int(idx)]
        else:
            v = getattr(m, name)
        if isinstance(v, pcr.Field):
            return pcr.pcr2numpy(pcr.scalar(v), MV).astype(F32)
        return np.asarray(v)

    def has(self, name):
        base, _, idx = name.rpartition("_")
        if base in ("UStoreLayerDepth", "vwc", "vwc_perc") and idx.isdigit():
            return hasattr(self.m, base) and int(idx) < len(getattr(self.m, base))
        return hasattr(self.m, name)
