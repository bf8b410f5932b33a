"""Synthetic wflow_sbm cases: drainage networks,
parameter maps in the value ranges of ``examples/wflow_rhine_sbm/intbl``, forcing.

Host-side, numpy.  Deterministic for a given seed.
"""
import numpy as np

from . import params

F32 = np.float32

# PCRaster LDD codes: 7 8 9 / 4 5 6 / 1 2 3, pit = 5  (wflow/wflow_funcs.py:43)
LDD_DR = {7: -1, 8: -1, 9: -1, 4: 0, 5: 0, 6: 0, 1: 1, 2: 1, 3: 1}
LDD_DC = {7: -1, 8: 0, 9: 1, 4: -1, 5: 0, 6: 1, 1: -1, 2: 0, 3: 1}
_CODE = np.array([[7, 8, 9], [4, 5, 6], [1, 2, 3]], dtype=np.uint8)


def _hash32(seed, r, c):
    """Cheap integer hash (uint32) of (seed, row, col), vectorised."""
    h = (r.astype(np.uint64) * np.uint64(0x9E3779B1) + c.astype(np.uint64) * np.uint64(0x85EBCA77)
         + np.uint64(seed) * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    return h


def ldd_all_pits(nrow, ncol):
    """Config 2: every cell is a pit (one level, no lateral connection)."""
    return np.full((nrow, ncol), 5, np.uint8)


def ldd_basin_tiles(nrow, ncol, tile=1024, seed=7, mask_frac=0.0):
    """Config 3/5: independent tile-sized basins, each a random D8 tree.

    Outlet (pit) at the bottom-centre of each tile; every other cell drains to one of
    its in-tile neighbours that is strictly closer (Chebyshev) to the outlet, chosen by
    hash(seed,row,col).  Acyclic, <= tile-1 levels, hop distance = Chebyshev distance.
    ``mask_frac`` > 0 removes (sets to 0 = missing) a hashed fraction of *leaf-side*
    cells: a masked cell takes its whole upstream subtree with it is avoided by only
    masking cells on the tile's outer ring.
    """
    rr, cc = np.meshgrid(np.arange(nrow, dtype=np.int64), np.arange(ncol, dtype=np.int64), indexing="ij")
    tr, tc = rr % tile, cc % tile
    th = np.minimum(tile, nrow - (rr - tr))  # actual tile height/width at grid edges
    tw = np.minimum(tile, ncol - (cc - tc))
    ro, co = th - 1, tw // 2
    d = np.maximum(np.abs(tr - ro), np.abs(tc - co))
    h = _hash32(seed, rr, cc)
    ncand = np.zeros((nrow, ncol), np.int64)
    cand_ok = []
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr == 0 and dc == 0:
                continue
            r2, c2 = tr + dr, tc + dc
            ok = (r2 >= 0) & (r2 < th) & (c2 >= 0) & (c2 < tw)
            d2 = np.maximum(np.abs(r2 - ro), np.abs(c2 - co))
            ok &= d2 == d - 1
            cand_ok.append((dr, dc, ok))
            ncand += ok
    pick = (h % np.maximum(ncand, 1).astype(np.uint64)).astype(np.int64)
    ldd = np.full((nrow, ncol), 5, np.uint8)
    seen = np.zeros((nrow, ncol), np.int64)
    for dr, dc, ok in cand_ok:
        sel = ok & (seen == pick) & (d > 0)
        ldd[sel] = _CODE[dr + 1, dc + 1]
        seen += ok
    if mask_frac > 0:
        ring = (tr == 0) | (tc == 0) | (tc == tw - 1)
        ring &= d > 0
        hm = _hash32(seed + 101, rr, cc)
        ldd[ring & ((hm % np.uint64(10000)) < np.uint64(int(mask_frac * 10000)))] = 0
    return ldd


def ldd_random_forest(nrow, ncol, seed=0, pit_frac=0.02, missing_frac=0.05):
    """Small irregular test networks: cells drain to a random neighbour with a lower
    random 'elevation' (pit if none); some cells missing."""
    rng = np.random.default_rng(seed)
    z = rng.random((nrow, ncol))
    ldd = np.zeros((nrow, ncol), np.uint8)
    for r in range(nrow):
        for c in range(ncol):
            best, code = z[r, c], 5
            order = [(dr, dc) for dr in (-1, 0, 1) for dc in (-1, 0, 1) if (dr, dc) != (0, 0)]
            rng.shuffle(order)
            for dr, dc in order:
                r2, c2 = r + dr, c + dc
                if 0 <= r2 < nrow and 0 <= c2 < ncol and z[r2, c2] < best:
                    best, code = z[r2, c2], _CODE[dr + 1, dc + 1]
                    if rng.random() < 0.5:
                        break
            if rng.random() < pit_frac:
                code = 5
            ldd[r, c] = code
    miss = rng.random((nrow, ncol)) < missing_frac
    ldd[miss] = 0
    # cells pointing at a missing cell or out of the map become pits (lddrepair)
    for r in range(nrow):
        for c in range(ncol):
            k = ldd[r, c]
            if k in (0, 5):
                continue
            r2, c2 = r + LDD_DR[int(k)], c + LDD_DC[int(k)]
            if not (0 <= r2 < nrow and 0 <= c2 < ncol) or ldd[r2, c2] == 0:
                ldd[r, c] = 5
    return ldd


def upstream_count(ldd):
    """Number of cells draining through each cell (incl. itself) = catchmenttotal(1, ldd).
    Vectorised Kahn sweep; only for moderate grids (tests, synthetic river definition)."""
    nrow, ncol = ldd.shape
    n = nrow * ncol
    flat = ldd.ravel()
    valid = (flat >= 1) & (flat <= 9)
    dr = np.zeros(10, np.int64)
    dc = np.zeros(10, np.int64)
    for k in range(1, 10):
        dr[k], dc[k] = LDD_DR[k], LDD_DC[k]
    idx = np.arange(n, dtype=np.int64)
    down = idx + dr[flat * valid] * ncol + dc[flat * valid]
    down[~valid | (flat == 5)] = -1
    indeg = np.zeros(n, np.int64)
    np.add.at(indeg, down[down >= 0], 1)
    acc = valid.astype(np.int64)
    frontier = idx[valid & (indeg == 0)]
    while frontier.size:
        d = down[frontier]
        ok = d >= 0
        np.add.at(acc, d[ok], acc[frontier[ok]])
        np.subtract.at(indeg, d[ok], 1)
        cand = np.unique(d[ok])
        frontier = cand[indeg[cand] == 0]
    return acc.reshape(nrow, ncol)


def _smooth_noise(rng, nrow, ncol, coarse=64):
    """Bilinear interpolation of coarse white noise (spatially smooth field in [0,1))."""
    gr, gc = max(2, nrow // coarse + 2), max(2, ncol // coarse + 2)
    g = rng.random((gr, gc)).astype(F32)
    y = np.linspace(0, gr - 1.001, nrow, dtype=F32)
    x = np.linspace(0, gc - 1.001, ncol, dtype=F32)
    y0, x0 = y.astype(np.int64), x.astype(np.int64)
    fy, fx = (y - y0)[:, None], (x - x0)[None, :]
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy).astype(F32)


def base_parameters(nrow, ncol, seed=42, block=64, n_layers=4):
    """Per-class parameter values (land use 1..6, soil 1..14 per ``block``² cells) drawn
    from the ranges of examples/wflow_rhine_sbm/intbl/*.tbl."""
    rng = np.random.default_rng(seed)
    br, bc = (nrow + block - 1) // block, (ncol + block - 1) // block
    lu = np.kron(rng.integers(0, 6, (br, bc)), np.ones((block, block), np.int64))[:nrow, :ncol]
    so = np.kron(rng.integers(0, 14, (br, bc)), np.ones((block, block), np.int64))[:nrow, :ncol]

    def by_lu(lo, hi, log=False):
        v = np.exp(rng.uniform(np.log(lo), np.log(hi), 6)) if log else rng.uniform(lo, hi, 6)
        return v.astype(F32)[lu]

    def by_soil(lo, hi, log=False):
        v = np.exp(rng.uniform(np.log(lo), np.log(hi), 14)) if log else rng.uniform(lo, hi, 14)
        return v.astype(F32)[so]

    base = {
        "KsatVer": by_soil(100, 8000, True), "M": by_soil(17, 12880, True), "thetaS": by_soil(0.4, 0.6),
        "thetaR": by_soil(0.01, 0.1), "SoilThickness": by_soil(400, 7100), "RootingDepth": by_lu(2, 750),
        "CanopyGapFraction": by_lu(0.1, 0.8), "Cmax": by_lu(0.0, 1.0), "EoverR": by_lu(0.0, 0.25),
        "PathFrac": by_lu(0.01, 0.9), "InfiltCapSoil": by_soil(100, 600), "InfiltCapPath": by_soil(5, 10),
        "TT": by_lu(-1.42, 0.0), "TTM": by_lu(-1.42, 0.0), "TTI": F32(1.0), "Cfmax": F32(3.76), "WHC": F32(0.1),
        "CapScale": F32(100.0), "rootdistpar": F32(-8000.0), "MaxLeakage": by_soil(0.0, 0.5),
        "KsatHorFrac": by_soil(1.0, 100.0, True), "N": by_lu(0.03, 0.6, True), "N_River": F32(0.036),
        "WaterFrac": (by_lu(0.0, 0.05) * (rng.random(6) < 0.5).astype(F32)[lu]).astype(F32),
    }
    Cm = base["Cmax"]
    Cm[lu == 0] = 0.0  # one open class without canopy (Cmax <= 0 branch of Gash)
    for k in range(n_layers):
        base["c_%d" % k] = by_soil(8.0, 12.0)
        base["KsatVerFrac_%d" % k] = by_soil(0.5, 1.5)
    slope = np.exp(np.log(1e-3) + _smooth_noise(rng, nrow, ncol) * (np.log(0.3) - np.log(1e-3))).astype(F32)
    base["landSlope"] = slope
    base["riverSlope"] = np.maximum(slope * F32(0.3), F32(1e-4)).astype(F32)
    base["elev"] = (_smooth_noise(rng, nrow, ncol, 128) * F32(2000.0)).astype(F32)
    base["TempCor"] = (F32(-0.0065) * base["elev"]).astype(F32)
    return base


class Forcing:
    """Per-step P / PET / TEMP rasters: spatially smooth rain cells over ~40 % of the grid with an
    exponential-like intensity distribution (mean of wet cells ~`rain_mean` mm/day, tail to
    ~10x that), sinusoidal PET and a temperature wave that crosses the snow thresholds."""

    def __init__(self, nrow, ncol, seed=42, timestepsecs=86400, rain_mean=8.0, t_mean=6.0, t_amp=12.0, day0=120.0):
        self.rng = np.random.default_rng(seed + 1)
        self.shape = (nrow, ncol)
        self.scale = F32(timestepsecs / 86400.0)
        self.steps_per_day = 86400.0 / timestepsecs
        self.pattern = _smooth_noise(self.rng, nrow, ncol, 96)
        self.rain_mean, self.t_mean, self.t_amp = rain_mean, t_mean, t_amp
        self.seed = seed
        self.day0 = day0

    def step(self, t):
        rng = np.random.default_rng(1000003 * (t + 1) + 17 + self.seed)
        field = _smooth_noise(rng, *self.shape, coarse=32)
        u = np.clip((field - F32(0.5)) * F32(4.0), F32(0.0), F32(0.999))  # 0 on the dry part
        p = (-np.log1p(-u) * F32(self.rain_mean * 1.5) * F32(rng.uniform(0.2, 2.0))).astype(F32)
        p = (p * self.scale).astype(F32)
        day = self.day0 + t / self.steps_per_day
        pet = np.full(self.shape, (2.0 + 1.5 * np.sin(2 * np.pi * day / 365.0)), F32) * self.scale
        temp = (np.full(self.shape, self.t_mean + self.t_amp * np.sin(2 * np.pi * (day - 100) / 365.0), F32)
                + (self.pattern - F32(0.5)) * F32(8.0) + F32(rng.normal(0, 3)))
        return p.astype(F32), pet.astype(F32), temp.astype(F32)


class _SmoothRng:
    """Drop-in for the few Generator methods random_state uses, returning spatially smooth fields
    (bilinear noise, ~`coarse` cells correlation length) instead of white noise: neighbouring cells
    then have similar water tables, snow packs ... like a real landscape."""

    def __init__(self, seed, coarse):
        self.rng = np.random.default_rng(seed)
        self.coarse = coarse

    def random(self, shape):
        f = _smooth_noise(self.rng, shape[0], shape[1], self.coarse)
        lo, hi = float(f.min()), float(f.max())
        return ((f - lo) / max(hi - lo, 1e-6) * 0.999).astype(np.float64)  # stretched to [0, 1)

    def uniform(self, lo, hi, shape):
        return lo + (hi - lo) * self.random(shape)


def random_state(static, n_layers, layer_thickness, seed=0, smooth=0):
    """A physically consistent but well-mixed initial state (instead of the cold start): water
    tables from the surface to the bottom, partly filled unsaturated layers, snow packs, wet
    canopies, flowing rivers and land.  Exercises the branches a cold start never reaches.
    smooth > 0: spatially correlated fields with that correlation length in cells (benchmarks);
    0: white noise, every cell independent (tests: maximal branch coverage on small grids)."""
    rng = _SmoothRng(seed, smooth) if smooth else np.random.default_rng(seed)
    st = static["SoilThickness"]
    shape = st.shape
    neff = (static["thetaS"] - static["thetaR"]).astype(F32)
    u = rng.random(shape).astype(F32)
    zi = (st * np.where(u < 0.15, F32(0.0) + u * F32(0.2), u)).astype(F32)   # 15 % (nearly) saturated
    zi[rng.random(shape) < 0.03] = 0.0
    state = params.cold_state(static, n_layers)
    state["zi"] = zi
    state["SatWaterDepth"] = (neff * (st - zi)).astype(F32)
    fabs = np.abs(static["f"])
    state["SubsurfaceFlow"] = (((static["KsatHorFrac"] * static["KsatVer"] * static["landSlope"]) / fabs)
                               * (np.exp(-fabs * zi) - np.exp(-fabs * st)) * static["DW"] * F32(1000.0)).astype(F32)
    # unsaturated storage: fill each layer above the water table to a random fraction of capacity
    tops = np.zeros(shape, F32)
    thick_list = list(layer_thickness)
    for k in range(n_layers):
        if k < len(thick_list):
            th = np.minimum(F32(thick_list[k]), np.maximum(st - tops, 0)).astype(F32)
        else:
            th = np.maximum(st - tops, 0).astype(F32)
        unsat = np.clip(zi - tops, 0, th).astype(F32)
        state["UStoreLayerDepth_%d" % k] = (unsat * neff * rng.random(shape).astype(F32) * F32(0.95)).astype(F32)
        tops = (tops + th).astype(F32)
    wet = rng.random(shape) < 0.3
    state["LandRunoffsub"] = np.where(wet, np.exp(rng.uniform(np.log(1e-5), np.log(0.5), shape)), 0).astype(F32)
    state["LandRunoff"] = state["LandRunoffsub"].copy()
    state["WaterLevelL"] = np.where(wet, rng.uniform(0, 0.01, shape), 0).astype(F32)
    river = static["River"] != 0
    state["RiverRunoffsub"] = np.where(river, np.exp(rng.uniform(np.log(1e-3), np.log(300.0), shape)), 0).astype(F32)
    state["RiverRunoff"] = state["RiverRunoffsub"].copy()
    state["WaterLevelR"] = np.where(river, rng.uniform(0.01, 2.0, shape), 0).astype(F32)
    state["WaterLevelRsub"] = state["WaterLevelR"].copy()
    snowy = rng.random(shape) < 0.35
    state["Snow"] = np.where(snowy, rng.uniform(0, 80, shape), 0).astype(F32)
    state["SnowWater"] = (state["Snow"] * rng.uniform(0, 0.1, shape)).astype(F32)
    state["TSoil"] = rng.uniform(-6, 12, shape).astype(F32)
    state["CanopyStorage"] = (static["Cmax"] * rng.uniform(0, 1.3, shape)).astype(F32)
    return state


def make_case(nrow, ncol, *, kind="pits", tile=1024, seed=42, timestepsecs=86400, layer_thickness=(100, 300, 800),
              river_area=500, mask_frac=0.0, init="cold", block=64):
    """Assemble (ldd, river, static, state) for a synthetic case."""
    if kind == "pits":
        ldd = ldd_all_pits(nrow, ncol)
        river = np.zeros((nrow, ncol), np.uint8)
    elif kind == "tiles":
        ldd = ldd_basin_tiles(nrow, ncol, tile=tile, seed=7, mask_frac=mask_frac)
        river = (upstream_count(ldd) >= river_area).astype(np.uint8)
    elif kind == "forest":
        ldd = ldd_random_forest(nrow, ncol, seed=seed)
        river = (upstream_count(ldd) >= river_area).astype(np.uint8)
    else:
        raise ValueError(kind)
    n_layers = len(layer_thickness) + 1 if len(layer_thickness) else 1
    base = base_parameters(nrow, ncol, seed=seed, n_layers=n_layers, block=block)
    up = upstream_count(ldd).astype(F32)
    base["RiverWidth"] = np.where(river != 0, F32(2.0) * np.power(np.maximum(up, 1), F32(0.35)), F32(0.0)).astype(F32)
    base["BankfullDepth"] = np.where(river != 0, F32(1.0), F32(0.0)).astype(F32)
    static = params.derive_static(base, ldd, river, timestepsecs, cellsize=1000.0, n_layers=n_layers)
    if init == "cold":
        state = params.cold_state(static, n_layers)
    else:
        state = random_state(static, n_layers, layer_thickness, seed=seed + 7)
    return ldd, river, static, state


def write_case_dir(path, nrow=24, ncol=24, *, kind="tiles", tile=12, nsteps=6, timestepsecs=86400, seed=3,
                   layer_thickness=(100, 300, 800), model_snow=1, river_area=6):
    """Write a small but complete wflow_sbm case directory (staticmaps/, intbl/, inmaps/ map stacks,
    wflow_sbm.ini) in the reference's on-disk layout, for exercising the framework / BMI mirror."""
    import os

    from . import csf
    from .framework import generate_name_t
    os.makedirs(os.path.join(path, "staticmaps"), exist_ok=True)
    os.makedirs(os.path.join(path, "intbl"), exist_ok=True)
    os.makedirs(os.path.join(path, "inmaps"), exist_ok=True)
    rng = np.random.default_rng(seed)
    ldd = ldd_basin_tiles(nrow, ncol, tile=tile) if kind == "tiles" else ldd_random_forest(nrow, ncol, seed=seed)
    up = upstream_count(ldd)
    river = (up >= river_area).astype(np.int32)
    info = csf.MapInfo(nrow, ncol, x_ul=100000.0, y_ul=500000.0, cellsize=1000.0, projection=1)
    sub = np.where(ldd > 0, 1 + (np.arange(nrow)[:, None] >= nrow // 2) * 1, 0).astype(np.int32)
    w = lambda name, a, vs=None: csf.write_map(os.path.join(path, "staticmaps", name), a, info, value_scale=vs,
                                               mv=-999 if a.dtype.kind != "f" else -999.0)
    w("wflow_ldd.map", np.where(ldd > 0, ldd, -999).astype(np.int32), csf.VS_LDD)
    w("wflow_subcatch.map", np.where(sub > 0, sub, -999).astype(np.int32), csf.VS_ORDINAL)
    w("wflow_river.map", np.where(river > 0, 1, -999).astype(np.int32), csf.VS_ORDINAL)
    w("wflow_landuse.map", np.where(ldd > 0, 1 + rng.integers(0, 3, (nrow, ncol)), -999).astype(np.int32), csf.VS_ORDINAL)
    w("wflow_soil.map", np.where(ldd > 0, 1 + rng.integers(0, 2, (nrow, ncol)), -999).astype(np.int32), csf.VS_ORDINAL)
    gauges = np.full((nrow, ncol), -999, np.int32)
    pits = np.argwhere(ldd == 5)
    for i, (r, c) in enumerate(pits[:8]):
        gauges[r, c] = i + 1
    w("wflow_gauges.map", gauges, csf.VS_ORDINAL)
    dem = (_smooth_noise(rng, nrow, ncol, 8) * F32(500.0) + F32(10.0)).astype(F32)
    w("wflow_dem.map", np.where(ldd > 0, dem, F32(np.nan)).astype(F32))
    w("RiverWidth.map", np.where(river > 0, F32(2.0) * np.power(np.maximum(up, 1).astype(F32), F32(0.35)), F32(0.0)).astype(F32))
    tables = {
        "KsatVer": ["[1,1] <,> [1,1] 600", "[2,2] <,> [1,1] 1500", "[3,> <,> [1,1] 250", "<,> <,> [2,> 900"],
        "M": ["<,> <,> [1,1] 250", "<,> <,> [2,> 900"], "thetaS": ["<,> <,> <,> 0.48"], "thetaR": ["<,> <,> <,> 0.05"],
        "SoilThickness": ["<,> [1,1] <,> 1800", "<,> [2,2] <,> 900"], "SoilMinThickness": ["<,> <,> <,> 400"],
        "RootingDepth": ["[1,1] <,> <,> 300", "[2,> <,> <,> 600"], "MaxCanopyStorage": ["[1,1] <,> <,> 0.0", "[2,> <,> <,> 0.6"],
        "CanopyGapFraction": ["<,2] <,> <,> 0.6", "<2,> <,> <,> 0.2"], "EoverR": ["<,> <,> <,> 0.1"],
        "PathFrac": ["<,> <,> <,> 0.15"], "InfiltCapSoil": ["<,> <,> <,> 200"], "InfiltCapPath": ["<,> <,> <,> 8"],
        "N": ["<,> <,> <,> 0.2"], "N_River": ["<,> <,> <,> 0.04"], "MaxLeakage": ["<,> <,> <,> 0.1"],
        "KsatHorFrac": ["<,> <,> <,> 20"],
        "c": ["<,> <,> <,> [0,0] 9.5", "<,> <,> <,> [1,1] 10.5", "<,> <,> <,> [2,> 11"],
    }
    for name, rows in tables.items():
        with open(os.path.join(path, "intbl", name + ".tbl"), "w") as fh:
            fh.write("\n".join(rows) + "\n")
    forc = Forcing(nrow, ncol, seed=seed, timestepsecs=timestepsecs, rain_mean=12.0)
    for t in range(nsteps):
        P, PET, T = forc.step(t)
        for stem, a in (("P", P), ("PET", PET), ("TEMP", T)):
            csf.write_map(generate_name_t(os.path.join(path, "inmaps", stem), t + 1), a, info)
    import datetime
    start = datetime.datetime(2000, 1, 1)
    end = start + datetime.timedelta(seconds=timestepsecs * (nsteps - 1))
    with open(os.path.join(path, "wflow_sbm.ini"), "w") as fh:
        fh.write("""[inputmapstacks]
Precipitation = /inmaps/P
EvapoTranspiration = /inmaps/PET
Temperature = /inmaps/TEMP
[run]
starttime = %s
endtime = %s
timestepsecs = %d
reinit = 1
[layout]
sizeinmetres = 1
[model]
modeltype = sbm
ModelSnow = %d
UStoreLayerThickness = %s
[API]
P = 0,mm
PET = 0,mm
TEMP = 0,oC
zi = 2,mm
SubsurfaceFlow = 2,mm3
RiverRunoff = 2,m3/s
LandRunoff = 2,m3/s
UStoreLayerDepth_0 = 2,mm
SatWaterDepth = 1,mm
KsatVer = 3,mm
[outputmaps]
self.RiverRunoff = run
[outputcsv_0]
samplemap = staticmaps/wflow_subcatch.map
self.InwaterMM = specrun.csv
self.Precipitation = prec.csv
self.ActEvap+self.Interception = teact.csv
[outputtss_0]
samplemap = staticmaps/wflow_gauges.map
self.RiverRunoff = run.tss
""" % (start.strftime("%Y-%m-%d %H:%M:%S"), end.strftime("%Y-%m-%d %H:%M:%S"), timestepsecs, model_snow,
       ",".join(str(int(t)) for t in layer_thickness) if layer_thickness else "0"))
    return ldd, river


def hbv_case(nrow, ncol, *, kind="tiles", tile=12, seed=3, timestepsecs=86400, mask_frac=0.0, block=4, glacier=0, edge=0):
    """(ldd, static, state) of a synthetic wflow_hbv case: parameter maps in the ranges of the reference's defaults and
    the Rhine tables (wflow/wflow_hbv.py:560-900), a spun-up random state (stateVariables, :131-163).  Rate parameters
    are per timestep like the reference's tables scaled by timestepsecs / basetimestep."""
    if kind == "pits":
        ldd = ldd_all_pits(nrow, ncol)
    elif kind == "tiles":
        ldd = ldd_basin_tiles(nrow, ncol, tile=tile, seed=7, mask_frac=mask_frac)
    else:
        ldd = ldd_random_forest(nrow, ncol, seed=seed)
    rng = np.random.default_rng(seed)
    shape = (nrow, ncol)
    br, bc = (nrow + block - 1) // block, (ncol + block - 1) // block
    lu = np.kron(rng.integers(0, 6, (br, bc)), np.ones((block, block), np.int64))[:nrow, :ncol]
    sc = F32(timestepsecs / 86400.0)

    def by_lu(lo, hi, log=False):
        v = np.exp(rng.uniform(np.log(lo), np.log(hi), 6)) if log else rng.uniform(lo, hi, 6)
        return v.astype(F32)[lu]

    full = lambda v: np.full(shape, v, F32)  # noqa: E731
    fc = by_lu(120.0, 400.0)
    st = {
        "Pcorr": full(1.0), "RFCF": by_lu(0.9, 1.1), "SFCF": by_lu(0.8, 1.2), "TTI": by_lu(0.5, 2.0), "TT": by_lu(-1.5, 0.5),
        "Cfmax": (by_lu(2.5, 4.5) * sc).astype(F32), "CFR": full(0.05), "WHC": full(0.1), "ICF": by_lu(0.0, 3.0),
        "EPF": by_lu(0.0, 0.05), "ECORR": full(1.0), "CEVPF": by_lu(0.8, 1.2), "FieldCapacity": fc,
        "Treshold": (by_lu(0.4, 0.9) * fc).astype(F32), "BetaSeepage": by_lu(1.0, 4.0), "PERC": (by_lu(0.2, 1.5) * sc).astype(F32),
        "Cflux": (by_lu(0.0, 2.5) * sc).astype(F32), "KQuickFlow": (by_lu(0.02, 0.15) * sc).astype(F32),
        "AlphaNL": by_lu(0.3, 1.3), "K4": (by_lu(0.005, 0.05) * sc).astype(F32), "K0": (by_lu(0.1, 0.4) * sc).astype(F32),
        "SUZ": by_lu(20.0, 120.0), "xl": full(1000.0), "yl": full(1000.0),
    }
    elev = (_smooth_noise(rng, nrow, ncol, 128) * F32(2000.0)).astype(F32)
    st["TempCor"] = (F32(-0.0065) * elev).astype(F32)
    slope = np.exp(np.log(1e-3) + _smooth_noise(rng, nrow, ncol) * (np.log(0.3) - np.log(1e-3))).astype(F32)
    st["Slope"] = slope
    up = upstream_count(ldd).astype(F32)
    # every cell carries a channel in wflow_hbv (:1040-1068): width from the upstream area, Manning's alpha
    st["Bw"] = (F32(2.0) * np.power(np.maximum(up, 1), F32(0.35))).astype(F32)
    st["DCL"] = np.where((ldd % 2 == 1) & (ldd != 5), F32(1000.0 * np.sqrt(2.0)), F32(1000.0)).astype(F32)
    N = by_lu(0.03, 0.3, True)
    beta = F32(0.6)
    alp_term = np.power((N / np.sqrt(slope)).astype(np.float64), float(beta)).astype(F32)
    alp_pow = F32(F32(2.0 / 3.0) * beta)
    st["Alpha"] = (alp_term * np.power((st["Bw"] + F32(1.0)).astype(np.float64), float(alp_pow)).astype(F32)).astype(F32)
    if edge:  # edge branches: threshold without interval, no interception store, a soil that cannot hold water
        u = rng.random(shape)
        st["TTI"][u < 0.15] = 0.0
        st["ICF"][(u > 0.15) & (u < 0.3)] = 0.0
        st["Cflux"][(u > 0.3) & (u < 0.4)] = 0.0
    snowy = rng.random(shape) < 0.4
    wet = rng.random(shape) < 0.5
    state = {
        "DrySnow": np.where(snowy, rng.uniform(0, 120, shape), 0).astype(F32),
        "SoilMoisture": (fc * rng.uniform(0.1, 1.0, shape)).astype(F32),
        "UpperZoneStorage": np.where(wet, rng.uniform(0, 60, shape), rng.uniform(0, 0.5, shape)).astype(F32),
        "LowerZoneStorage": rng.uniform(1.0, 80.0, shape).astype(F32),
        "InterceptionStorage": (st["ICF"] * rng.uniform(0, 1.0, shape)).astype(F32),
        "SurfaceRunoff": (np.exp(rng.uniform(np.log(1e-3), np.log(2.0), shape)) * np.maximum(up, 1) ** 0.8).astype(F32),
        "WaterLevel": rng.uniform(0.01, 2.0, shape).astype(F32),
    }
    state["FreeWater"] = (state["DrySnow"] * rng.uniform(0, 0.1, shape)).astype(F32)
    if glacier:
        gf = np.where(rng.random(shape) < 0.35, rng.uniform(0.05, 1.0, shape), 0.0).astype(F32)
        st.update(GlacierFrac=gf, G_TT=full(1.3), G_Cfmax=full(5.3 * float(sc)), G_SIfrac=full(0.002 * float(sc)))
        state["GlacierStore"] = np.where(gf > 0, rng.uniform(0.0, 5500.0, shape), 0.0).astype(F32)
    return ldd, st, state
