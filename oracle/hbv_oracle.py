"""CPU restatement of one timestep of wflow_hbv (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Follows ``WflowModel.dynamic()`` of wflow/wflow_hbv.py:1222-1767 line by line: REAL4 map algebra as float32 numpy, one
operation at a time in the reference's order (snow routine :1288-1376, glacier :1378-1404 = glacierHBV
wflow/wflow_funcs.py:549-603, soil moisture :1412-1442, upper / lower zone :1444-1520, lateral inflow :1522-1612),
then ``kin_wave`` (wflow/wflow_funcs.py:155-207) through the C oracle (oracle/wflow_sbm_oracle.c, wfo_kin_wave) on the
full network, every cell a channel.  Transcendentals (``pcr.exp``, ``**``) are evaluated in float64 and rounded to REAL4
once, like oracle/pcrshim.py.

PINNED: tests/test_hbv_oracle.py compares every state and output of every timestep with the goldens of the reference's
own, unmodified ``dynamic()`` (tests/golden/hbv_*.npz, written by oracle/gen_golden_hbv.py) bit for bit.

Not restated (the reference's optional parts that the CUDA path refuses as well): reservoirs / lakes (:1533-1594),
state updating from observed discharge (:1684-1726).
"""
import numpy as np

from . import oracle as _orc

F32 = np.float32


def _pow(x, y):
    with np.errstate(all="ignore"):
        return np.power(np.asarray(x, np.float64), np.asarray(y, np.float64)).astype(F32)


def _exp(x):
    with np.errstate(all="ignore"):
        return np.exp(np.asarray(x, np.float64)).astype(F32)


class HbvOracle:
    STATES = ["FreeWater", "SoilMoisture", "UpperZoneStorage", "LowerZoneStorage", "InterceptionStorage", "SurfaceRunoff",
              "WaterLevel", "DrySnow", "SurfaceRunoffsub", "WaterLevelsub"]

    def __init__(self, ldd, fields, *, timestepsecs=86400, kinwaveIters=0, kinwaveTstep=0, setKquickFlow=0, externalQbase=0,
                 massWasting=0):
        ldd = np.asarray(ldd)
        self.shape = ldd.shape
        self.net = _orc.Network(ldd)
        self.active = np.zeros(self.shape, bool)
        self.active.flat[self.net.flat_nodes] = True
        self.valid = (ldd >= 1) & (ldd <= 9)
        self.dt = int(timestepsecs) if float(timestepsecs).is_integer() else float(timestepsecs)
        self.kinwaveIters, self.kinwaveTstep = int(kinwaveIters), kinwaveTstep
        self.setKquickFlow, self.externalQbase = int(setKquickFlow), int(externalQbase)
        self.massWasting = int(massWasting)
        self.f = {k: np.array(np.broadcast_to(np.asarray(v, F32), self.shape)) for k, v in fields.items()}
        f = self.f
        zero = np.zeros(self.shape, F32)
        for k in self.STATES[:8]:
            f.setdefault(k, zero.copy())
        f.setdefault("SurfaceRunoffsub", f["SurfaceRunoff"].copy())  # :1202
        f.setdefault("WaterLevelsub", f["WaterLevel"].copy())
        self.glacier = "GlacierFrac" in f
        dt = F32(self.dt)
        f["QMMConv"] = dt / ((f["xl"] * f["yl"]) * F32(0.001))      # :1013
        f["ToCubic"] = ((f["xl"] * f["yl"]) * F32(0.001)) / dt      # :1014
        f["KinWaveVolume"] = (f["WaterLevel"] * f["Bw"]) * f["DCL"]
        # resume(), :1208-1215, and the sums of the water balance (:1015-1031)
        f.setdefault("initstorage", (((((f["FreeWater"] + f["DrySnow"]) + f["SoilMoisture"]) + f["UpperZoneStorage"])
                                      + f["LowerZoneStorage"]) + f["InterceptionStorage"]).astype(F32))
        for k in ("sumprecip", "sumevap", "sumrunoff", "sumlevel", "sumpotevap", "sumsoilevap", "sumtemp", "suminflow"):
            f[k] = zero.copy()
        self.it_kin = None
        self.out = {}

    def set(self, name, a):
        self.f[name] = np.array(np.broadcast_to(np.asarray(a, F32), self.shape))

    def get(self, name):
        return self.out[name] if name in self.out else self.f[name]

    def _accucapacitystate(self, material, capacity):
        """pcr.accucapacitystate(ldd, material, capacity): what a cell holds plus what arrives from upstream leaves it up to
        `capacity`; the state is what stays (REAL4, neighbours summed in the network's order)."""
        flux = np.zeros(material.size, F32)
        state = material.ravel().copy()
        cap = capacity.ravel()
        for kk, idx in enumerate(self.net.flat_nodes):  # level-major: upstream cells come first
            tot = state[idx]
            for nb in self.net.flat_up[kk]:
                if nb >= 0:
                    tot = F32(tot + flux[nb])
            out = min(tot, cap[idx])
            flux[idx] = out
            state[idx] = F32(tot - out)
        return state.reshape(self.shape)

    def estimate_iterations(self):
        """estimate_iterations_kin_wave(SurfaceRunoffsub, Beta, Alpha, timestepsecs, DCL), wflow_funcs.py:103-116."""
        f = self.f
        valid = self.valid  # the MAP's domain (every cell with an ldd code), not set_dd's network
        Q = np.where(valid, f["SurfaceRunoffsub"], F32(0)).astype(F32)
        if not (Q.max() > 0):
            return 1
        beta = F32(0.6)
        with np.errstate(all="ignore"):
            cel = F32(1.0) / ((f["Alpha"] * beta) * _pow(Q, beta - F32(1.0)))
            courant = (F32(self.dt) / f["DCL"]) * cel
        c = np.where((Q > 0) & valid, courant, F32(np.nan)).astype(F32)
        try:
            return int(np.ceil(F32(1.25) * np.nanpercentile(c, 95)))
        except Exception:
            return 1

    def step(self, P, PET, TEMP=None, Inflow=None, Seepage=None):
        f, o = self.f, self.out
        mn, mx, wh = np.minimum, np.maximum, np.where
        zero = np.zeros(self.shape, F32)
        one = F32(1.0)
        with np.errstate(all="ignore"):
            Precipitation = mx(F32(0.0), np.asarray(P, F32)) * f["Pcorr"]                                    # :1264
            PotEvaporation = np.asarray(PET, F32)
            Seep = np.asarray(Seepage, F32) if (self.externalQbase and Seepage is not None) else zero       # :1271-1274
            Temperature = (np.asarray(TEMP, F32) if TEMP is not None else zero + F32(10.0)) + f["TempCor"]  # :1275-1276
            TT, TTI = f["TT"], f["TTI"]
            RainFrac = wh(one * TTI == F32(0.0), wh(Temperature <= TT, F32(0.0), F32(1.0)),
                          mn((Temperature - (TT - TTI / F32(2.0))) / TTI, one))                              # :1282-1291
            RainFrac = mx(RainFrac, F32(0.0)).astype(F32)
            SnowFrac = one - RainFrac
            Precipitation = f["SFCF"] * SnowFrac * Precipitation + f["RFCF"] * RainFrac * Precipitation     # :1297-1300
            Interception = mn(Precipitation, f["ICF"] - f["InterceptionStorage"])                           # :1303-1305
            IS = f["InterceptionStorage"] + Interception
            Precipitation = Precipitation - Interception
            PotEvaporation = _exp(-f["EPF"] * Precipitation) * f["ECORR"] * PotEvaporation                  # :1311-1313
            PotEvaporation = f["CEVPF"] * PotEvaporation
            IntEvap = mn(IS, PotEvaporation)
            IS = IS - IntEvap
            RestEvap = mx(F32(0.0), PotEvaporation - IntEvap)                                               # :1322
            SnowFall = SnowFrac * Precipitation
            RainFall = RainFrac * Precipitation
            PotSnowMelt = wh(Temperature > TT, f["Cfmax"] * (Temperature - TT), F32(0.0)).astype(F32)       # :1334-1338
            PotRefreezing = wh(Temperature < TT, f["Cfmax"] * f["CFR"] * (TT - Temperature), F32(0.0)).astype(F32)
            Refreezing = wh(Temperature < TT, mn(PotRefreezing, f["FreeWater"]), F32(0.0)).astype(F32)
            SnowMelt = mn(PotSnowMelt, f["DrySnow"])
            DrySnow = f["DrySnow"] + SnowFall + Refreezing - SnowMelt
            FreeWater = f["FreeWater"] - Refreezing
            MaxFreeWater = DrySnow * f["WHC"]
            FreeWater = FreeWater + SnowMelt + RainFall
            InSoil = mx(FreeWater - MaxFreeWater, F32(0.0))
            FreeWater = FreeWater - InSoil
            RainAndSnowmelt = RainFall + SnowMelt
            SnowCover = wh(DrySnow > 0, F32(1), F32(0)).astype(F32)
            NetInSoil = InSoil
            SM = f["SoilMoisture"] + NetInSoil                                                               # :1372
            FC = f["FieldCapacity"]
            DirectRunoff = mx(SM - FC, F32(0.0))
            SM = SM - DirectRunoff
            NetInSoil = NetInSoil - DirectRunoff
            if self.massWasting:  # :1351-1364, pcr.accucapacitystate along the network (level by level, upstream first)
                frac = (mn(F32(0.5), f["Slope"] / F32(5.67)) * mn(F32(1.0), DrySnow / F32(10000.0))).astype(F32)
                DrySnow = self._accucapacitystate(DrySnow.astype(F32), (frac * DrySnow).astype(F32))
                FreeWater = self._accucapacitystate(FreeWater.astype(F32), (frac * FreeWater).astype(F32))
            if self.glacier:  # glacierHBV, wflow_funcs.py:549-603 (timestepsecs / basetimestep is a Python float)
                ts = F32(self.dt / 86400)
                gf = f["GlacierFrac"]
                S2G = f["G_SIfrac"] * ts * DrySnow
                S2G = wh(gf > F32(0.0), S2G, F32(0.0)).astype(F32)
                S2G = mn(S2G, F32(8.0 * (self.dt / 86400)))
                DrySnow = DrySnow - (S2G * gf)
                GS = f["GlacierStore"] + S2G
                PotMelt = wh(Temperature > f["G_TT"], f["G_Cfmax"] * ts * (Temperature - f["G_TT"]), F32(0.0)).astype(F32)
                GM = wh(DrySnow < F32(10.0), mn(PotMelt, GS), F32(0.0)).astype(F32)
                GS = GS - GM
                GM = GM * gf                                                                                 # :1402
                FreeWater = FreeWater + GM
                f["GlacierStore"] = GS.astype(F32)
                o["Snow2Glacier"], o["GlacierMelt"] = S2G, GM
            Tr = f["Treshold"]
            SoilEvap = wh(SM > Tr, mn(SM, RestEvap), mn(SM, mn(RestEvap, PotEvaporation * (SM / Tr)))).astype(F32)  # :1412-1422
            SM = SM - SoilEvap
            ActEvap = IntEvap + SoilEvap
            HBVSeepage = _pow(mn(SM / FC, F32(1)), f["BetaSeepage"]) * NetInSoil                            # :1431-1433
            SM = SM - HBVSeepage
            Backtosoil = mn(FC - SM, DirectRunoff)
            DirectRunoff = DirectRunoff - Backtosoil
            SM = SM + Backtosoil
            InUpperZone = DirectRunoff + HBVSeepage
            UZ = f["UpperZoneStorage"] + InUpperZone                                                         # :1448-1450
            Percolation = mn(f["PERC"], UZ - InUpperZone / F32(2))
            UZ = UZ - Percolation
            CapFlux = f["Cflux"] * ((FC - SM) / FC)
            CapFlux = mn(UZ, CapFlux)
            CapFlux = mn(FC - SM, CapFlux)
            UZ = UZ - CapFlux
            SM = SM + CapFlux
            if not self.setKquickFlow:                                                                       # :1466-1500
                QuickFlow = mn(wh(Percolation < f["PERC"], F32(0),
                                  f["KQuickFlow"] * _pow(UZ - mn(InUpperZone / F32(2), UZ), one + f["AlphaNL"])).astype(F32), UZ)
                UZ = mx(wh(Percolation < f["PERC"], UZ, UZ - QuickFlow).astype(F32), F32(0))
                RealQuickFlow = zero
            else:
                QuickFlow = f["KQuickFlow"] * UZ
                RealQuickFlow = mx(F32(0), f["K0"] * (UZ - f["SUZ"]))
                UZ = UZ - QuickFlow - RealQuickFlow
            LZ = f["LowerZoneStorage"] + Percolation
            BaseFlow = mn(LZ, f["K4"] * LZ)
            LZ = LZ - BaseFlow
            if self.externalQbase:
                DirectRunoffStorage = QuickFlow + Seep + RealQuickFlow
            else:
                DirectRunoffStorage = QuickFlow + BaseFlow + RealQuickFlow
            InwaterMM = mx(F32(0.0), DirectRunoffStorage)
            Inwater = InwaterMM * f["ToCubic"]
            Infl = np.asarray(Inflow, F32) if Inflow is not None else zero
            SWS = wh(Infl < F32(0.0), mx(F32(-1.0) * Inwater, f["SurfaceRunoff"]), F32(0)).astype(F32)      # :1599-1603
            Inwater = Inwater + wh(SWS > 0, F32(-1.0) * SWS, Infl).astype(F32)
            q = Inwater / f["DCL"] + zero / f["DCL"]                                                         # :1612
        f["InterceptionStorage"], f["DrySnow"], f["FreeWater"] = IS.astype(F32), DrySnow.astype(F32), FreeWater.astype(F32)
        f["SoilMoisture"], f["UpperZoneStorage"], f["LowerZoneStorage"] = SM.astype(F32), UZ.astype(F32), LZ.astype(F32)
        # ---- kin_wave (:1614-1660) ----
        it = 1
        if self.kinwaveIters == 1:
            it = self.estimate_iterations() if self.kinwaveTstep == 0 else int(np.ceil(self.dt / self.kinwaveTstep))
        if self.it_kin is not None:
            it = int(self.it_kin)
        self.last_it = it
        qold = np.ascontiguousarray(f["SurfaceRunoffsub"].ravel().copy())
        g64 = lambda k: f[k].ravel().astype(np.float64)  # noqa: E731
        acc, qnew, wl_av, wl = _orc.kin_wave(self.net, qold, q.ravel(), g64("Alpha"), np.full(qold.size, np.float64(F32(0.6))),
                                             g64("DCL"), g64("Bw"), self.dt, it)
        r = lambda a: a.reshape(self.shape).astype(F32)  # noqa: E731
        f["SurfaceRunoff"] = r(acc / self.dt)
        f["SurfaceRunoffsub"] = r(qnew)
        f["WaterLevel"] = r(wl_av)
        f["WaterLevelsub"] = r(wl)
        old_kw = f["KinWaveVolume"]
        f["KinWaveVolume"] = (f["WaterLevel"] * f["Bw"]) * f["DCL"]                                          # :124-129
        up = np.zeros(self.shape, F32)  # pcr.upstream(TopoLdd, SurfaceRunoff): neighbour order of the network
        sr = f["SurfaceRunoff"].ravel()
        for kk, idx in enumerate(self.net.flat_nodes):
            s = F32(0)
            for nb in self.net.flat_up[kk]:
                if nb >= 0:
                    s = F32(s + sr[nb])
            up.flat[idx] = s
        with np.errstate(all="ignore"):
            massbal = ((f["KinWaveVolume"] - old_kw) / F32(self.dt) + up + Inwater) - f["SurfaceRunoff"]
        o.update(Precipitation=Precipitation, PotEvaporation=PotEvaporation, Temperature=Temperature, IntEvap=IntEvap,
                 SnowMelt=SnowMelt, SnowCover=SnowCover, SoilEvap=SoilEvap, ActEvap=ActEvap, HBVSeepage=HBVSeepage,
                 DirectRunoff=DirectRunoff, InUpperZone=InUpperZone, Percolation=Percolation, CapFlux=CapFlux,
                 QuickFlow=QuickFlow, RealQuickFlow=RealQuickFlow, BaseFlow=BaseFlow, InSoil=InSoil,
                 RainAndSnowmelt=RainAndSnowmelt, NetInSoil=NetInSoil, InwaterMM=InwaterMM, Inwater=Inwater,
                 QuickFlowCubic=(QuickFlow + RealQuickFlow) * f["ToCubic"], BaseFlowCubic=BaseFlow * f["ToCubic"],
                 SurfaceWaterSupply=SWS, SurfaceRunoffMM=f["SurfaceRunoff"] * f["QMMConv"],
                 SurfaceRunoffsubMM=f["SurfaceRunoffsub"] * f["QMMConv"], OldKinWaveVolume=old_kw, MassBalKinWave=massbal)
        # ---- water balance, :1731-1766 ----
        f["sumprecip"] = f["sumprecip"] + Precipitation
        f["sumevap"] = f["sumevap"] + ActEvap
        f["sumsoilevap"] = f["sumsoilevap"] + SoilEvap
        f["sumpotevap"] = f["sumpotevap"] + PotEvaporation
        f["sumtemp"] = f["sumtemp"] + Temperature
        f["sumrunoff"] = f["sumrunoff"] + InwaterMM
        f["sumlevel"] = f["sumlevel"] + f["WaterLevel"]
        f["suminflow"] = f["suminflow"] + Infl
        o["storage"] = (((f["FreeWater"] + f["DrySnow"]) + f["SoilMoisture"]) + f["UpperZoneStorage"]) + f["LowerZoneStorage"]
        o["watbal"] = (((f["initstorage"] - o["storage"]) + f["sumprecip"]) - f["sumsoilevap"]) - f["sumrunoff"]
        for k, v in list(o.items()):
            o[k] = np.asarray(v, F32)
