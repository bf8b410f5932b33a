"""Run the REFERENCE's own ``wflow_hbv.WflowModel.dynamic()`` -- unmodified -- on in-memory inputs
(TEST INFRASTRUCTURE ONLY; needs /root/reference, i.e. the build container).

``wflow/wflow_hbv.py`` is loaded where it lies (oracle/refload.load_hbv) with ``pcraster`` = the float32-numpy shim
of oracle/pcrshim.py.  This module does what ``initial()`` / ``resume()`` (wflow_hbv.py:324-1220) and the framework do
before the first timestep -- hand the model its parameter maps, states, the ``static`` numpy mirror (:1100-1123) and the
``set_dd`` network (:1125-1132) -- and then calls the reference's ``dynamic`` (:1222-1767) once per timestep: snow routine,
interception, soil moisture, upper / lower zone, the lateral inflow glue and ``kin_wave`` are all reference code.
"""
import numpy as np

from . import pcrshim as pcr
from . import refload

F32 = np.float32
MV = -999.0

# parameter maps of the timestep (readtblDefault / wf_readmap in initial(), :560-900)
PARAMS = ["Pcorr", "TempCor", "TTI", "TT", "SFCF", "RFCF", "ICF", "EPF", "ECORR", "CEVPF", "Cfmax", "CFR", "WHC",
          "FieldCapacity", "Treshold", "BetaSeepage", "PERC", "Cflux", "KQuickFlow", "AlphaNL", "K4", "K0", "SUZ",
          "xl", "yl", "DCL", "Bw", "Alpha", "Slope", "GlacierFrac", "G_TT", "G_Cfmax", "G_SIfrac"]
# stateVariables(), :131-163
STATES = ["FreeWater", "SoilMoisture", "UpperZoneStorage", "LowerZoneStorage", "InterceptionStorage", "SurfaceRunoff",
          "WaterLevel", "DrySnow"]
SUMS = ["sumprecip", "sumevap", "sumrunoff", "sumlevel", "sumpotevap", "sumsoilevap", "sumtemp", "suminflow"]


def _model_class():
    hbv = refload.load_hbv()

    class Model(hbv.WflowModel):
        """The reference model with the framework's per-timestep services replaced by in-memory ones."""

        def __init__(self):  # the reference's __init__ wants a case directory and a clone map file
            pass

        def wf_updateparameters(self):  # wf_DynamicFramework.py:708-874: forcing of the timestep
            f = self._forcing
            self.Precipitation = self._map(f["P"])
            self.PotEvaporation = self._map(f["PET"])
            self.Inflow = self._map(f["Inflow"]) if f.get("Inflow") is not None else self.ZeroMap

        def wf_readmap(self, name, default, **_):  # the temperature and seepage stacks are read directly (:1275-1278)
            f = self._forcing
            if name == "TEMP":
                return self._map(f["TEMP"]) if f.get("TEMP") is not None else pcr.scalar(default)
            if name == "SEEPAGE":
                return self._map(f["Seepage"]) if f.get("Seepage") is not None else pcr.scalar(default)
            raise KeyError(name)

        def wf_multparameters(self):
            return None

        def wf_supplyJulianDOY(self):
            return int(self._forcing.get("JDOY", 1))

        def _map(self, a, vs=pcr.Scalar):
            a = np.array(np.broadcast_to(np.asarray(a), self.shape), dtype=np.float64)
            a[~self._active] = MV
            return pcr.numpy2pcr(vs, a, MV)

    return Model, hbv


class RefHbv:
    """fields: dict of float32 rasters by reference attribute name (PARAMS and STATES above)."""

    def __init__(self, ldd, fields, *, timestepsecs=86400, kinwaveIters=0, kinwaveTstep=0, setKquickFlow=0,
                 externalQbase=0, massWasting=0, topo_id=None):
        Model, self.hbv = _model_class()
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
        # ---- options (initial(), :383-470) ----
        m.timestepsecs = int(timestepsecs) if float(timestepsecs).is_integer() else float(timestepsecs)
        m.basetimestep = 86400
        m.kinwaveIters = int(kinwaveIters)
        m.kinwaveTstep = kinwaveTstep
        m.SetKquickFlow = int(setKquickFlow)
        m.ExternalQbase = int(externalQbase)
        m.MassWasting = int(massWasting)
        m.updating = 0
        m.nrresSimple = m.nrlake = 0
        m.filterResArea = 0
        m.TEMP_mapstack, m.Seepage_mapstack = "TEMP", "SEEPAGE"
        # ---- maps ----
        zero = np.zeros(self.shape, F32)
        m.ZeroMap = mp(zero)
        lddf = np.where(active, ldd, 0).astype(np.float64)
        lddf[~active] = MV
        m.TopoLdd = pcr.numpy2pcr(pcr.Ldd, lddf, MV)
        m.TopoLddOrg = m.TopoLdd
        m.TopoId = mp(np.ones(self.shape) if topo_id is None else topo_id, pcr.Ordinal)
        for k in PARAMS:
            if k in fl:
                setattr(m, k, mp(fl[k]))
        m.Beta = pcr.scalar(0.6)  # :561
        m.River = mp(np.ones(self.shape, F32))  # only kin_wave's signature takes it
        # :1013-1031
        m.QMMConv = m.timestepsecs / (m.xl * m.yl * 0.001)
        temp = pcr.catchmenttotal(pcr.cover(1.0), m.TopoLdd) * m.xl * 0.001 * 0.001 * m.yl
        m.QMMConvUp = pcr.cover(m.timestepsecs * 0.001) / temp
        m.ToCubic = (m.xl * m.yl * 0.001) / m.timestepsecs
        m.ForecQ_qmec = m.ZeroMap
        for k in SUMS:
            setattr(m, k, m.ZeroMap)
        # ---- states (resume(), :1146-1200) ----
        for k in STATES:
            setattr(m, k, mp(fl.get(k, zero)))
        if "GlacierStore" in fl:
            m.GlacierStore = mp(fl["GlacierStore"])
        m.SurfaceRunoffsub = mp(fl["SurfaceRunoffsub"]) if "SurfaceRunoffsub" in fl else m.SurfaceRunoff  # :1202
        m.WaterLevelsub = m.WaterLevel
        m.KinWaveVolume = m.WaterLevel * m.Bw * m.DCL
        m.OldKinWaveVolume = m.KinWaveVolume
        m.initstorage = (m.FreeWater + m.DrySnow + m.SoilMoisture + m.UpperZoneStorage + m.LowerZoneStorage
                         + m.InterceptionStorage)  # :1208-1215
        # ---- numpy mirror and network, :1100-1132 ----
        n = self.shape[0] * self.shape[1]
        g = lambda x: pcr.pcr2numpy(x, MV).ravel()  # noqa: E731
        names = ["Alpha", "Beta", "DCL", "River", "Bw", "AlpTerm", "AlpPow"]
        st = np.zeros(n, dtype=np.dtype([(k, np.float64) for k in names]))
        for k in ["Alpha", "Beta", "DCL", "River", "Bw"]:
            st[k] = g(getattr(m, k))
        m.static = st
        m.np_ldd = lddf
        m.nodes, m.nodes_up = funcs.set_dd(lddf)
        self.active = active

    def set_state(self, name, a):
        setattr(self.m, name, self.m._map(a))

    def step(self, P, PET, TEMP=None, Inflow=None, Seepage=None, JDOY=1):
        self.m._forcing = dict(P=P, PET=PET, TEMP=TEMP, Inflow=Inflow, Seepage=Seepage, JDOY=JDOY)
        self.m.dynamic()

    def get(self, name):
        """float32 raster of a model attribute (MV = -999)."""
        v = getattr(self.m, name)
        if isinstance(v, pcr.Field):
            return pcr.pcr2numpy(pcr.scalar(v), MV).astype(F32)
        return np.asarray(v)

    def has(self, name):
        return hasattr(self.m, name)
