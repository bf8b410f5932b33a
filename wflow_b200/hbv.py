"""wflow_hbv on the device: `HbvModel`, the raster-level owner of a ``wf_hbv_create`` handle.

The reference's wflow_hbv (wflow/wflow_hbv.py) shares everything around its column with wflow_sbm: the LDD network of
``set_dd``, ``kin_wave`` for the channel of every cell, gauges and area statistics.  So this class is `SbmModel` with another
constructor and the timestep of ``WflowModel.dynamic()`` (wflow_hbv.py:1222-1767): fields are set and read by the reference's
attribute names (``SurfaceRunoff``, ``WaterLevel``, ``Alpha`` ... map to the channel rasters of the shared sweep), ``dynamic``
uploads the forcing of a timestep, derives the sub-steps of the kinematic wave the way the reference does
([model] kinwaveIters / kinwaveTstep, :1614-1626) and steps.  No CPU fallback: the constructor fails without the CUDA library.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from .core import SbmModel

#: stateVariables() of the reference (wflow_hbv.py:131-163) + the sub-step discharge the kinematic wave carries
STATES = ["FreeWater", "SoilMoisture", "UpperZoneStorage", "LowerZoneStorage", "InterceptionStorage", "SurfaceRunoff",
          "WaterLevel", "DrySnow", "SurfaceRunoffsub", "WaterLevelsub"]
#: parameter maps of the timestep (readtblDefault / wf_readmap in initial(), wflow_hbv.py:560-900)
PARAMETERS = ["Pcorr", "TempCor", "TTI", "TT", "SFCF", "RFCF", "ICF", "EPF", "ECORR", "CEVPF", "Cfmax", "CFR", "WHC",
              "FieldCapacity", "Treshold", "BetaSeepage", "PERC", "Cflux", "KQuickFlow", "AlphaNL", "K4", "K0", "SUZ",
              "xl", "yl", "DCL", "Bw", "Alpha", "Slope", "GlacierFrac", "G_TT", "G_Cfmax", "G_SIfrac"]
#: outputs of the column that wf_request_output materialises
OUTPUTS = ["Precipitation", "PotEvaporation", "Temperature", "IntEvap", "SnowMelt", "SnowCover", "SoilEvap", "ActEvap",
           "HBVSeepage", "DirectRunoff", "InUpperZone", "Percolation", "CapFlux", "QuickFlow", "RealQuickFlow", "BaseFlow",
           "InSoil", "RainAndSnowmelt", "NetInSoil", "InwaterMM", "Inwater", "QuickFlowCubic", "BaseFlowCubic",
           "SurfaceWaterSupply", "Snow2Glacier", "GlacierMelt"]


class HbvModel(SbmModel):
    def __init__(self, ldd, *, timestepsecs=86400, kinwaveIters=0, kinwaveTstep=0, set_kquick_flow=0, external_qbase=0,
                 mass_wasting=0, glacier=0, device=0, beta=0.0):
        self._l = _lib.lib()
        ldd = np.ascontiguousarray(ldd, dtype=np.uint8)
        if ldd.ndim != 2:
            raise ValueError("ldd must be a 2-D raster")
        self.shape = ldd.shape
        self._ldd = ldd
        self.timestepsecs = timestepsecs
        self.kinwaveIters, self.kinwaveTstep = int(kinwaveIters), kinwaveTstep
        it = 1  # :1614-1626: one pass unless kinwaveIters = 1
        if self.kinwaveIters == 1 and kinwaveTstep:
            it = int(math.ceil(timestepsecs / kinwaveTstep))
        cfg = _lib.WfConfig()
        cfg.nrow, cfg.ncol = self.shape
        cfg.n_layer_thickness = 0
        cfg.model_snow, cfg.glacier, cfg.it_kin_l, cfg.it_kin_r = 1, int(glacier), 1, it
        cfg.device, cfg.timestepsecs, cfg.beta = int(device), float(timestepsecs), float(beta)
        self.cfg = cfg
        self.max_layers = 1
        h = C.c_void_p()
        _lib.check(self._l.wf_hbv_create(C.byref(cfg), ldd.ctypes.data_as(C.c_void_p), C.byref(h)))
        self._h = h
        if set_kquick_flow:
            self.set_option("SetKquickFlow", 1)
        if external_qbase:
            self.set_option("ExternalQbase", 1)
        if mass_wasting:  # (needs the Slope raster)
            self.set_option("MassWasting", 1)
        self.it_kin = it

    def set_many(self, fields):
        """Rasters by name; a case that gives SurfaceRunoff / WaterLevel without their sub-step twins starts those from
        them, as resume() does (wflow_hbv.py:1202)."""
        super().set_many(fields)
        if "SurfaceRunoff" in fields and "SurfaceRunoffsub" not in fields:
            self.set("SurfaceRunoffsub", fields["SurfaceRunoff"])
        if "WaterLevel" in fields and "WaterLevelsub" not in fields:
            self.set("WaterLevelsub", fields["WaterLevel"])

    def dynamic(self, P, PET, TEMP=None, Inflow=None, Seepage=None):
        """One timestep (WflowModel.dynamic, wflow_hbv.py:1222-1767) with the forcing rasters of the step."""
        self.set("P", P)
        self.set("PET", PET)
        if TEMP is not None:
            self.set("TEMP", TEMP)
        if Inflow is not None:
            self.set("Inflow", Inflow)
        if Seepage is not None:
            self.set("Seepage", Seepage)
        self.advance()

    def advance(self):
        """The timestep itself, forcing already in place: sub-steps of the kinematic wave the way the reference derives them,
        then the device step."""
        if self.kinwaveIters == 1 and not self.kinwaveTstep:  # Courant estimate from the state before the step (:1617-1624)
            it = min(self.estimate_iterations("river"), 1024)
            if it != self.it_kin:
                self.set_substeps(1, it)
                self.it_kin = it
        self.step()
