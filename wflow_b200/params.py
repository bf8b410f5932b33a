"""Host-side derivation of the kernel-level parameter maps and cold-start states.

Mirrors, in float32 numpy (PCRaster Scalar = REAL4), the part of
``WflowModel.initial()`` / ``resume()`` that turns base parameter maps into the maps the
per-timestep path consumes (reference ``wflow/wflow_sbm.py:1985-2174`` and
``:2443-2516``).  Everything here runs once per model; the per-timestep work is the
CUDA path.

All functions take and return ``dict[str, np.ndarray(float32, shape=(nrow, ncol))]``.
"""
import numpy as np

F32 = np.float32

# defaults of wflow_sbm's parameter tables, SURVEY.md Appendix B (wflow_sbm.py:1458-1790)
DEFAULTS = {
    "Cmax": 1.0, "CanopyGapFraction": 0.1, "EoverR": 0.1, "RootingDepth": 750.0,
    "AirEntryPressure": 10.0, "rootdistpar": -8000.0, "InfiltCapSoil": 100.0, "CapScale": 100.0,
    "InfiltCapPath": 10.0, "MaxLeakage": 0.0, "PathFrac": 0.01, "SoilThickness": 2000.0,
    "thetaR": 0.01, "thetaS": 0.6, "KsatVer": 3000.0, "KsatHorFrac": 1.0, "M": 300.0, "N": 0.072,
    "N_River": 0.036, "WaterFrac": 0.0, "et_RefToPot": 1.0, "TTI": 1.0, "TT": -1.41934,
    "TTM": -1.41934, "Cfmax": 3.75653, "WHC": 0.1, "w_soil": 0.1125, "cf_soil": 0.038,
    "TempCor": 0.0, "KsatVerFrac": 1.0, "c": 10.0, "RiverWidth": 0.0, "BankfullDepth": 0.0,
    "RiverLengthFac": 1.0,
}

# parameters that are rates per day and are rescaled to the model timestep
# (wflow_sbm.py:1519-1563, 1617-1619, 1721-1743)
RATE_PARAMS = ("InfiltCapSoil", "InfiltCapPath", "MaxLeakage", "KsatVer", "Cfmax", "w_soil")

BETA = F32(0.6)  # wflow_sbm.py:1643


def drain_length(ldd, xl, yl):
    """detdrainlength, wflow/wflow_lib.py:832-861."""
    slant = np.sqrt(xl * xl + yl * yl).astype(F32)
    d = ldd
    return np.where((d == 2) | (d == 8), yl, np.where((d == 4) | (d == 6), xl, slant)).astype(F32)


def drain_width(ldd, xl, yl):
    """detdrainwidth, wflow/wflow_lib.py:864-893."""
    slant = ((xl + yl) * F32(0.5)).astype(F32)
    d = ldd
    return np.where((d == 2) | (d == 8), xl, np.where((d == 4) | (d == 6), yl, slant)).astype(F32)


def derive_static(base, ldd, river, timestepsecs, *, cellsize=1000.0, n_layers=1):
    """Kernel-level static maps from base parameter maps.

    base: dict of base maps / scalars (names as in the reference's intbl tables; missing
    ones take DEFAULTS).  Needs at least nothing: every parameter has a default.
    Optional spatial inputs: ``xl``, ``yl`` (cell size in m, else ``cellsize``),
    ``landSlope``, ``riverSlope``, ``RiverWidth``, ``BankfullDepth``, ``DCL``.
    Per-layer tables: ``c_k`` / ``KsatVerFrac_k`` (fall back to ``c`` / ``KsatVerFrac``).
    """
    shape = ldd.shape
    dt_ratio = F32(timestepsecs) / F32(86400.0)

    def get(name, default=None):
        v = base.get(name, DEFAULTS.get(name, default))
        return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=F32), shape))

    out = {}
    for name in ("Cmax", "EoverR", "CanopyGapFraction", "et_RefToPot", "TempCor", "TTI", "TT", "TTM", "WHC",
                 "thetaS", "thetaR", "PathFrac", "CapScale", "rootdistpar", "AirEntryPressure", "KsatHorFrac"):
        out[name] = get(name).copy()
    for name in RATE_PARAMS:
        out[name] = (get(name) * dt_ratio).astype(F32)  # x * timestepsecs / basetimestep
    out["cf_soil"] = np.minimum(F32(0.99), get("cf_soil")).astype(F32)  # wflow_sbm.py:1745
    out["SoilThickness"] = get("SoilThickness").copy()
    # limit roots to top 99% of the soil, wflow_sbm.py:1938
    out["RootingDepth"] = np.minimum(out["SoilThickness"] * F32(0.99), get("RootingDepth")).astype(F32)
    # f = |thetaS - thetaR| / M, wflow_sbm.py:2159
    out["f"] = np.abs((out["thetaS"] - out["thetaR"]) / get("M")).astype(F32)

    xl = get("xl", cellsize).copy()
    yl = get("yl", cellsize).copy()
    out["xl"], out["yl"] = xl, yl
    river_b = np.asarray(river) != 0
    out["River"] = river_b.astype(F32)
    land_slope = np.maximum(F32(0.00001), get("landSlope", 0.01)).astype(F32)  # :1918
    river_slope = np.maximum(F32(0.00001), get("riverSlope", base.get("landSlope", 0.01))).astype(F32)  # :2055
    out["landSlope"] = land_slope

    dl = drain_length(ldd, xl, yl)  # :2161
    out["DL"] = dl
    out["DW"] = ((xl * yl) / dl).astype(F32)  # :2162
    dcl = base.get("DCL")
    if dcl is None:
        dcl = (dl * np.maximum(F32(1.0), get("RiverLengthFac"))).astype(F32)  # :2047-2049
    out["DCL"] = np.ascontiguousarray(np.broadcast_to(np.asarray(dcl, F32), shape)).copy()
    river_width = get("RiverWidth")
    bw = np.where(river_b, river_width, drain_width(ldd, xl, yl)).astype(F32)  # :2101-2105
    out["Bw"] = bw
    river_frac = np.minimum(F32(1.0), np.where(river_b, (river_width * out["DCL"]) / (xl * yl), F32(0.0))).astype(F32)
    out["RiverFrac"] = river_frac  # :2108
    out["WaterFrac"] = np.maximum(get("WaterFrac") - river_frac, F32(0.0)).astype(F32)  # :2115
    out["SW"] = np.where(river_b, np.maximum(out["DW"] - river_width, F32(0.0)), out["DW"]).astype(F32)  # :2165

    beta = BETA
    alp_pow = F32(F32(2.0) / F32(3.0)) * beta  # :2121
    with np.errstate(divide="ignore", invalid="ignore"):
        alp_term_r = np.power(get("N_River") / np.sqrt(river_slope), beta).astype(F32)  # :2118
        alp_term_l = np.power(get("N") / np.sqrt(land_slope), beta).astype(F32)  # :2170
        out["AlphaL"] = (alp_term_l * np.power(out["SW"], alp_pow)).astype(F32)  # :2171
        out["AlphaR"] = (alp_term_r * np.power(bw + get("BankfullDepth"), alp_pow)).astype(F32)  # :2174
    out["Beta"] = np.full(shape, beta, F32)

    for k in range(n_layers):
        out["c_%d" % k] = get("c_%d" % k, None) if ("c_%d" % k) in base else get("c")
        out["KsatVerFrac_%d" % k] = get("KsatVerFrac_%d" % k, None) if ("KsatVerFrac_%d" % k) in base else get("KsatVerFrac")
        out["c_%d" % k] = out["c_%d" % k].copy()
        out["KsatVerFrac_%d" % k] = out["KsatVerFrac_%d" % k].copy()
    return out


def cold_state(static, n_layers, ldd=None):
    """Default initial conditions (``reinit = 1``), reference wflow_sbm.py:2443-2516."""
    shape = static["SoilThickness"].shape
    zero = np.zeros(shape, F32)
    st = static["SoilThickness"]
    neff = (static["thetaS"] - static["thetaR"]).astype(F32)
    swc = (st * neff).astype(F32)  # :1935
    sat = (swc * F32(0.85)).astype(F32)  # :2445
    zi = np.maximum(F32(0.0), st - sat / neff).astype(F32)  # :2447
    fabs = np.abs(static["f"])
    ssf = (((static["KsatHorFrac"] * static["KsatVer"] * static["landSlope"]) / fabs)
           * (np.exp(-fabs * zi) - np.exp(-fabs * st)) * static["DW"] * F32(1000.0)).astype(F32)  # :2452
    state = {
        "SatWaterDepth": sat, "zi": zi, "SubsurfaceFlow": ssf,
        "WaterLevelR": zero.copy(), "WaterLevelRsub": zero.copy(), "WaterLevelL": zero.copy(),
        "WaterLevelLsub": zero.copy(), "RiverRunoff": zero.copy(), "RiverRunoffsub": zero.copy(),
        "LandRunoff": zero.copy(), "LandRunoffsub": zero.copy(), "Snow": zero.copy(),
        "SnowWater": zero.copy(), "TSoil": (zero + F32(10.0)).astype(F32), "CanopyStorage": zero.copy(),
    }
    for k in range(n_layers):
        state["UStoreLayerDepth_%d" % k] = zero.copy()
    # zi recomputed from SatWaterDepth at the end of resume(), :2512
    state["zi"] = np.maximum(F32(0.0), st - state["SatWaterDepth"] / neff).astype(F32)
    return state


STATE_NAMES = (
    "CanopyStorage", "Snow", "SnowWater", "TSoil", "zi", "SubsurfaceFlow", "LandRunoff", "LandRunoffsub",
    "WaterLevelL", "WaterLevelLsub", "RiverRunoff", "RiverRunoffsub", "WaterLevelR", "WaterLevelRsub",
)
