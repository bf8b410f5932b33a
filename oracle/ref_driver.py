"""Drive the REFERENCE's own numba kernels (loaded by oracle/refload.py) on the inputs of
an oracle model, to validate the C restatement and to write golden vectors.

TEST INFRASTRUCTURE ONLY; needs /root/reference (this container).

The reference's ``dynamic()`` glue between the PCRaster maps and the numba kernels
(wflow/wflow_sbm.py:2885-2909 in, :2944-3129 and :3189-3201 out) is reproduced here with
numpy float32 <-> float64 casts; the PCRaster phases themselves (interception, snow,
evap split) come from the oracle's diagnostics, since PCRaster cannot be installed.
"""
import numpy as np

from . import refload

LAYER_FIELDS = ["c", "UStoreLayerDepth", "st", "KsatVerFrac", "ActEvapUstore", "vwc", "vwc_perc",
                "UStoreLayerThickness", "UStest"]
STATIC_FIELDS = ["KsatVer", "KsatHorFrac", "f", "thetaS", "thetaR", "SoilThickness", "InfiltCapSoil", "PathFrac",
                 "InfiltCapPath", "cf_soil", "neff", "landSlope", "riverSlope", "SoilWaterCapacity", "MaxLeakage",
                 "CanopyGapFraction", "CapScale", "rootdistpar", "River", "DL", "DW", "SW", "Beta", "DCL",
                 "reallength", "RootingDepth", "AirEntryPressure", "ActRootingDepth", "ssfmax", "xl", "yl", "h_p",
                 "Bw", "AlpPow", "AlpTermR", "AlpTermL", "AlphaL", "AlphaR"]
DYN_FIELDS = ["AvailableForInfiltration", "SoilInf", "PathInf", "TSoil", "zi", "SatWaterDepth", "ssf",
              "LandRunoffsub", "WaterLevelLsub", "PotTrans", "PotSoilEvap", "ActEvapOpenWaterLand",
              "ActEvapOpenWaterRiver", "RiverRunoff", "InfiltExcess", "ExcessWater", "ExcessWaterSoil",
              "ExcessWaterPath", "ExfiltWater", "ExfiltSatWater", "ExfiltFromUstore", "soilevap", "soilevapunsat",
              "soilevapsat", "Transfer", "CapFlux", "qo_toriver", "ssf_toriver", "ActEvapUStore", "ActEvapSat",
              "ActInfilt", "ActInfiltSoil", "ActInfiltPath", "RunoffLandCells", "ActLeakage", "PondingDepth",
              "RootStore_unsat", "CellInFlow", "InwaterO", "SoilWatbal", "Qo_in", "WaterLevelL", "InfiltSoilPath",
              "sumUStoreLayerDepth", "SatWaterDepthOld", "sumUStoreLayerDepthOld", "InfiltWater"]

F32 = np.float32


class ReferenceModel:
    """Holds the reference's structured arrays for one oracle-shaped case."""

    def __init__(self, orc):
        self.funcs, self.sbm = refload.load()
        self.orc = orc
        p = orc.p
        self.shape = orc.shape
        n = self.shape[0] * self.shape[1]
        self.maxLayers = p.maxLayers
        ldd = orc.net.ldd.astype(np.float64)
        ldd[ldd == 0] = -999.0
        self.ldd2d = ldd
        self.nodes, self.nodes_up = self.funcs.set_dd(ldd)
        fl = orc.fields
        river = fl["River"].ravel()
        lr = ldd.copy()
        lr.flat[river == 0] = -999.0  # wflow_sbm.py:2387-2388
        self.rnodes, self.rnodes_up = self.funcs.set_dd(lr)

        st = np.zeros(n, dtype=np.dtype([(k, np.float64) for k in STATIC_FIELDS]))
        for k in ["KsatVer", "KsatHorFrac", "f", "thetaS", "thetaR", "SoilThickness", "InfiltCapSoil", "PathFrac",
                  "InfiltCapPath", "landSlope", "MaxLeakage", "CanopyGapFraction", "CapScale", "rootdistpar", "River",
                  "DL", "DW", "SW", "Beta", "DCL", "RootingDepth", "AirEntryPressure", "xl", "yl", "Bw", "AlphaL",
                  "AlphaR"]:
            st[k] = fl[k].ravel()
        if "cf_soil" in fl:
            st["cf_soil"] = fl["cf_soil"].ravel()
        st["neff"] = (fl["thetaS"] - fl["thetaR"]).astype(F32).ravel()  # :2158
        st["SoilWaterCapacity"] = (fl["SoilThickness"] * (fl["thetaS"] - fl["thetaR"])).astype(F32).ravel()  # :1935
        with np.errstate(all="ignore"):
            st["ssfmax"] = ((st["KsatHorFrac"] * st["KsatVer"] * st["landSlope"]) / st["f"]
                            * (np.exp(0) - np.exp(-st["f"] * st["SoilThickness"])))  # :2291
        self.static = st
        layer = np.zeros((self.maxLayers, n), dtype=np.dtype([(k, np.float64) for k in LAYER_FIELDS]))
        for k in range(self.maxLayers):
            layer["c"][k] = fl["c_%d" % k].ravel()
            layer["KsatVerFrac"][k] = fl["KsatVerFrac_%d" % k].ravel()
        # layer thickness, :2301-2326
        np_zeros = np.zeros(n)
        sumT = np_zeros
        for k in range(self.maxLayers):
            sumL = sumT
            if p.nrLayers > 1 and k < p.nrLayers:
                tmp = float(p.layerThickness[k]) + np_zeros
                th = np.minimum(tmp, np.maximum(st["SoilThickness"] - sumL, 0.0))
            else:
                tmp = np.maximum(st["SoilThickness"] - sumL, 0.0)
                th = np.maximum(st["SoilThickness"] - sumL, 0.0)
            sumT = tmp + sumT
            layer["UStoreLayerThickness"][k] = th
        self.layer = layer
        self.dyn = np.zeros(n, dtype=np.dtype([(k, np.float64) for k in DYN_FIELDS]))

    def step(self, pre, diag):
        """pre: dict of float32 state maps BEFORE the step; diag: the oracle's PCRaster-phase
        diagnostics of this step.  Returns dict of float32 outputs of the numba half."""
        p = self.orc.p
        fl = self.orc.fields
        dyn, layer, st = self.dyn, self.layer, self.static
        shape = self.shape
        dt = float(p.timestepsecs)
        sat = ((fl["thetaS"] - fl["thetaR"]) * (fl["SoilThickness"] - pre["zi"])).astype(F32)  # :2751
        st["ActRootingDepth"] = np.minimum(fl["SoilThickness"] * F32(0.99), fl["RootingDepth"]).astype(F32).ravel()
        dyn["RunoffLandCells"] = diag["RunoffLandCells"].ravel()
        dyn["ActEvapOpenWaterLand"] = diag["ActEvapOpenWaterLand"].ravel()
        dyn["ActEvapOpenWaterRiver"] = diag["ActEvapOpenWaterRiver"].ravel()
        dyn["AvailableForInfiltration"] = diag["AvailableForInfiltration"].ravel()
        dyn["zi"] = pre["zi"].ravel()
        dyn["SatWaterDepth"] = sat.ravel()
        dyn["PotSoilEvap"] = diag["PotSoilEvap"].ravel()
        dyn["PotTrans"] = diag["PotTrans"].ravel()
        dyn["TSoil"] = fl["TSoil"].ravel() if "TSoil" in fl else 0.0  # after the snow phase
        dyn["ssf"] = pre["SubsurfaceFlow"].ravel()
        dyn["WaterLevelL"] = pre["WaterLevelL"].ravel()
        dyn["Qo_in"] = dyn["Qo_in"] * 0.0
        usum = np.zeros(shape, F32)
        for k in range(self.maxLayers):
            layer["UStoreLayerDepth"][k] = pre["UStoreLayerDepth_%d" % k].ravel()
            usum = usum + pre["UStoreLayerDepth_%d" % k]
        dyn["LandRunoffsub"] = pre["LandRunoffsub"].ravel()
        dyn["sumUStoreLayerDepth"] = usum.ravel()
        ssf, qav, wl_av, dyn, layer = self.sbm.sbm_cell(
            self.nodes, self.nodes_up, self.ldd2d.ravel(), layer, st, dyn, int(p.modelSnow), int(p.soilInfRedu),
            dt, 86400.0, 1.0, 0, shape, int(p.transferMethod), int(p.it_kinL), int(p.ust))
        self.dyn, self.layer = dyn, layer
        out = {}
        out["SubsurfaceFlow"] = ssf.astype(F32).reshape(shape)
        out["LandRunoffsub"] = dyn["LandRunoffsub"].astype(F32).reshape(shape)
        out["WaterLevelLsub"] = dyn["WaterLevelLsub"].astype(F32).reshape(shape)
        out["LandRunoff"] = qav.astype(F32).reshape(shape)
        out["WaterLevelL"] = wl_av.astype(F32).reshape(shape)
        out["zi"] = dyn["zi"].astype(F32).reshape(shape)
        out["SatWaterDepth"] = dyn["SatWaterDepth"].astype(F32).reshape(shape)
        for k in range(self.maxLayers):
            out["UStoreLayerDepth_%d" % k] = layer["UStoreLayerDepth"][k].astype(F32).reshape(shape)
            out["vwc_%d" % k] = layer["vwc"][k].astype(F32).reshape(shape)
        for k_out, k_dyn in [("UStoreDepth", "sumUStoreLayerDepth"), ("CapFlux", "CapFlux"), ("Transfer", "Transfer"),
                             ("ActEvapUStore", "ActEvapUStore"), ("ActEvapSat", "ActEvapSat"), ("soilevap", "soilevap"),
                             ("soilevapsat", "soilevapsat"), ("qo_toriver", "qo_toriver"), ("ssf_toriver", "ssf_toriver"),
                             ("ExfiltWater", "ExfiltWater"), ("InfiltExcess", "InfiltExcess"),
                             ("ExcessWater", "ExcessWater"), ("ActInfilt", "ActInfilt"), ("ActLeakage", "ActLeakage"),
                             ("SoilInf", "SoilInf"), ("PathInf", "PathInf"), ("InfiltExcessSoil", "ExcessWaterSoil"),
                             ("InfiltExcessPath", "ExcessWaterPath"), ("ActInfiltSoil", "ActInfiltSoil"),
                             ("ActInfiltPath", "ActInfiltPath"), ("ExfiltFromUstore", "ExfiltFromUstore"),
                             ("ExfiltFromSat", "ExfiltSatWater"), ("InwaterL", "InwaterO"),
                             ("InflowKinWaveCellLand", "Qo_in"), ("CellInFlow", "CellInFlow"),
                             ("SoilWatbal", "SoilWatbal"), ("RootStore_unsat", "RootStore_unsat")]:
            out[k_out] = dyn[k_dyn].astype(F32).reshape(shape)
        # river, :3052-3065, 3144-3201
        tocubic = ((fl["xl"] * fl["yl"] * F32(0.001)) / F32(dt)).astype(F32)
        inwater = (diag["InwaterMM"] * tocubic + out["qo_toriver"] + out["ssf_toriver"]).astype(F32)
        if "Inflow" in fl:
            inwater = (inwater + fl["Inflow"]).astype(F32)
        out["Inwater"] = inwater
        qr = (inwater / fl["DCL"]).astype(F32)
        rr_sub = pre["RiverRunoffsub"].astype(F32).ravel().copy()  # pcr2numpy of a REAL4 map: float32 (:3157)
        acc, qnew, wl_r_av, wl_r = self.funcs.kin_wave(
            self.rnodes, self.rnodes_up, rr_sub, qr.astype(F32).ravel(), st["AlphaR"], st["Beta"], st["DCL"],
            st["River"], st["Bw"], st["AlphaR"], st["Beta"], dt, int(p.it_kinR))
        out["RiverRunoff"] = (acc / dt).astype(F32).reshape(shape)
        out["WaterLevelR"] = wl_r_av.astype(F32).reshape(shape)
        out["RiverRunoffsub"] = qnew.astype(F32).reshape(shape)
        out["WaterLevelRsub"] = wl_r.astype(F32).reshape(shape)
        return out
