"""ctypes wrapper around the C oracle (TEST INFRASTRUCTURE ONLY -- see
wflow_sbm_oracle.c).  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from wflow_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwflow_sbm_oracle.so")
MAXL = 8


def build(force=False):
    src = os.path.join(_HERE, "wflow_sbm_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


class _Params(C.Structure):
    _fields_ = [
        ("nrow", C.c_int32), ("ncol", C.c_int32), ("maxLayers", C.c_int32), ("nrLayers", C.c_int32),
        ("modelSnow", C.c_int32), ("soilInfRedu", C.c_int32), ("transferMethod", C.c_int32),
        ("ust", C.c_int32), ("it_kinL", C.c_int32), ("it_kinR", C.c_int32), ("glacier", C.c_int32),
        ("reserved", C.c_int32), ("timestepsecs", C.c_double), ("basetimestep", C.c_double),
        ("layerThickness", C.c_double * MAXL),
    ]


class _Network(C.Structure):
    _fields_ = [
        ("nlev", C.c_int64), ("ncells", C.c_int64), ("lev_off", C.POINTER(C.c_int64)),
        ("nodes", C.POINTER(C.c_int64)), ("nodes_up", C.POINTER(C.c_int64)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.wfo_field_name.restype = C.c_char_p
        _lib.wfo_kinematic_wave.restype = C.c_double
        _lib.wfo_kinematic_wave.argtypes = [C.c_double] * 7
        _lib.wfo_kinematic_wave_ssf.argtypes = [C.c_double] * 14 + [C.POINTER(C.c_double)] * 3
        _lib.wfo_kinematic_wave_ssf.restype = None
        _lib.wfo_set_threads.argtypes = [C.c_int]
    return _lib


def set_threads(n):
    """Threads used by Oracle.step (OpenMP over the cells of a level); returns the count in use."""
    # cap by the CPUs this process may run on, not by OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1);
    # the OpenMP loops carry an explicit num_threads() clause
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    n = max(1, min(int(n), avail)) if lib().wfo_has_openmp() else 1
    lib().wfo_set_threads(n)
    return n


def field_names():
    l = lib()
    return [l.wfo_field_name(i).decode() for i in range(l.wfo_nfields())]


class Network:
    """Result of set_dd in the reference's form: nodes[level] / nodes_up[level]."""

    def __init__(self, ldd2d):
        ldd2d = np.ascontiguousarray(ldd2d)
        l8 = np.where((ldd2d >= 1) & (ldd2d <= 9), ldd2d, 0).astype(np.uint8)
        self.ldd = np.ascontiguousarray(l8)
        self.shape = ldd2d.shape
        self._net = _Network()
        rc = lib().wfo_set_dd(self.ldd.ctypes.data_as(C.c_void_p), C.c_int(self.shape[0]),
                              C.c_int(self.shape[1]), C.byref(self._net))
        if rc != 0:
            raise ValueError("wfo_set_dd failed (%d): ldd has a cycle" % rc)
        n, nl = self._net.ncells, self._net.nlev
        self.lev_off = np.ctypeslib.as_array(self._net.lev_off, shape=(nl + 1,)).copy()
        self.flat_nodes = np.ctypeslib.as_array(self._net.nodes, shape=(max(n, 1),))[:n].copy()
        self.flat_up = np.ctypeslib.as_array(self._net.nodes_up, shape=(max(n, 1) * 8,))[: n * 8].reshape(n, 8).copy()

    @property
    def nodes(self):
        return [self.flat_nodes[self.lev_off[i]: self.lev_off[i + 1]] for i in range(len(self.lev_off) - 1)]

    @property
    def nodes_up(self):
        return [self.flat_up[self.lev_off[i]: self.lev_off[i + 1]] for i in range(len(self.lev_off) - 1)]

    def __del__(self):
        try:
            lib().wfo_free_network(C.byref(self._net))
        except Exception:
            pass


def kinematic_wave(Qin, Qold, q, alpha, beta, dT, dX):
    return lib().wfo_kinematic_wave(Qin, Qold, q, alpha, beta, dT, dX)


def kin_wave(net, Qold, q, Alpha, Beta, DCL, Bw, deltaT, it):
    """kin_wave (wflow_funcs.py:155-207) over a Network; Qold / q float32 flat arrays (Qold is updated in place like the
    reference's), Alpha / Beta / DCL / Bw float64.  Returns (acc_flow, Qnew, waterlevel_av, waterlevel), float64 flat."""
    n = Qold.size
    assert Qold.dtype == np.float32 and Qold.flags.c_contiguous
    q = np.ascontiguousarray(q, np.float32)
    st = [np.ascontiguousarray(a, np.float64) for a in (Alpha, Beta, DCL, Bw)]
    out = [np.zeros(n, np.float64) for _ in range(4)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib().wfo_kin_wave(C.byref(net._net), C.c_int64(n), p(Qold), p(q), *[p(a) for a in st], C.c_double(deltaT), C.c_int(int(it)),
                            *[p(a) for a in out])
    if rc != 0:
        raise MemoryError("wfo_kin_wave failed (%d)" % rc)
    return tuple(out)


def kinematic_wave_ssf(*a):
    o = [C.c_double() for _ in range(3)]
    lib().wfo_kinematic_wave_ssf(*[C.c_double(float(x)) for x in a], *[C.byref(x) for x in o])
    return tuple(x.value for x in o)


class Oracle:
    """One model instance: full-grid float32 fields by reference attribute name."""

    def __init__(self, ldd2d, river2d=None, active2d=None, *, timestepsecs=86400, layer_thickness=(),
                 modelSnow=1, soilInfRedu=0, transferMethod=0, ust=0, it_kinL=1, it_kinR=1, glacier=0):
        self.net = Network(ldd2d)
        self.shape = self.net.shape
        ldd = self.net.ldd
        if river2d is None:
            river2d = np.zeros(self.shape, np.uint8)
        lr = np.where(np.asarray(river2d) != 0, ldd, 0)  # wflow_sbm.py:2387-2389
        self.rnet = Network(lr)
        self.active = np.ascontiguousarray((ldd > 0) if active2d is None else np.asarray(active2d) != 0).astype(np.uint8)
        nlay = len(layer_thickness)
        p = _Params()
        p.nrow, p.ncol = self.shape
        if nlay > 0:
            p.nrLayers, p.maxLayers = nlay, nlay + 1  # wflow_sbm.py:1757-1766
        else:
            p.nrLayers, p.maxLayers = 1, 1
        p.modelSnow, p.soilInfRedu, p.transferMethod, p.ust = modelSnow, soilInfRedu, transferMethod, ust
        p.it_kinL, p.it_kinR, p.glacier = it_kinL, it_kinR, glacier
        p.timestepsecs, p.basetimestep = float(timestepsecs), 86400.0
        for i, t in enumerate(layer_thickness):
            p.layerThickness[i] = float(t)
        self.p = p
        self.maxLayers = p.maxLayers
        self.names = field_names()
        self.fields = {}

    def set(self, name, arr):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(arr, dtype=np.float32), self.shape)).copy()
        self.fields[name] = a
        return a

    # maps that the end-of-dynamic() outputs are formed from (wflow_sbm.py:3310-3526): requested along with them
    END_DEPS = {
        "MassBalKinWaveR": ["Inwater"], "MassBalKinWaveL": ["InwaterL", "InflowKinWaveCellLand"],
        "RunoffCoeff": ["RainFallPlusMelt"], "CellStorage": ["SatWaterDepth"], "DeltaStorage": ["SatWaterDepth"],
        "RootStore": ["RootStore_unsat"], "vwcRoot": ["RootStore_unsat"], "vwc_percRoot": ["RootStore_unsat"],
        "ExfiltWaterCubic": ["ExfiltWater"], "InfiltExcessCubic": ["InfiltExcess"], "CumActInfilt": ["ActInfilt"],
        "CumCellInFlow": ["CellInFlow"], "CumPrec": ["Precipitation"], "CumEvap": ["ActEvap"], "CumPotenEvap": ["PotenEvap"],
        "CumInt": ["Interception"], "CumLeakage": ["ActLeakage"], "CumExfiltWater": ["ExfiltWater"],
    }

    def want(self, *names):
        """Request diagnostic outputs."""
        for n in names:
            deps = list(self.END_DEPS.get(n, []))
            if n.startswith("vwc_perc_"):
                deps.append("vwc_" + n.rsplit("_", 1)[1])
            for d in deps + [n]:
                if d not in self.fields:
                    self.fields[d] = np.zeros(self.shape, np.float32)

    def __getitem__(self, name):
        return self.fields[name]

    def step(self):
        tab = (C.c_void_p * len(self.names))()
        for i, n in enumerate(self.names):
            a = self.fields.get(n)
            tab[i] = a.ctypes.data if a is not None else None
        rc = lib().wfo_step(C.byref(self.p), C.byref(self.net._net), C.byref(self.rnet._net),
                            self.net.ldd.ctypes.data_as(C.c_void_p), self.active.ctypes.data_as(C.c_void_p), tab)
        if rc != 0:
            raise RuntimeError("wfo_step failed with code %d" % rc)

    # -- estimate_iterations_kin_wave, wflow/wflow_funcs.py:103-116 --------------------------------
    # PCRaster half of the reference (REAL4 map algebra + np.nanpercentile on the float32 raster):
    # restated with float32 numpy; pinned against the reference's own function run through
    # oracle/pcrshim.py (tests/golden/dyn_courant_hourly.npz, tests/test_oracle_dynamic_golden.py).
    def estimate_iterations(self, which):
        """which = "land": LandRunoffsub / AlphaL / DL (wflow_sbm.py:2911-2921);
        "river": RiverRunoffsub / AlphaR / DCL (:3159-3169)."""
        f32 = np.float32
        qn, an, dn = {"land": ("LandRunoffsub", "AlphaL", "DL"), "river": ("RiverRunoffsub", "AlphaR", "DCL")}[which]
        valid = np.zeros(self.shape, bool)
        valid.flat[self.net.flat_nodes] = True                # cells of the model domain; MV elsewhere
        Q = np.where(valid, self.fields[qn], f32(0)).astype(f32)
        if not (Q.max() > 0):
            return 1                                          # all-NaN percentile raises: it_kin = 1 (:113-114)
        beta = f32(0.6)                                        # self.Beta, REAL4 (wflow_sbm.py:1643)
        flow = Q > 0
        with np.errstate(all="ignore"):
            cel = f32(1.0) / ((self.fields[an].astype(f32) * beta) * np.power(Q, beta - f32(1.0), dtype=f32))
            courant = (f32(self.p.timestepsecs) / self.fields[dn].astype(f32)) * cel
        c = np.where(flow, courant, f32(np.nan)).astype(f32)
        try:
            return int(np.ceil(f32(1.25) * np.nanpercentile(c, 95)))
        except Exception:
            return 1

    def set_substeps(self, it_kinL, it_kinR):
        self.p.it_kinL, self.p.it_kinR = int(it_kinL), int(it_kinR)
