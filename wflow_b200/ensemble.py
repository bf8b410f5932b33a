"""Parameter ensembles: M members of one model grid as ONE device model.

The reference runs an ensemble (calibration: wflow/wflow_fit.py:153-158 multiplies M, KsatVer, N, ... of an
area by trial factors; :235-257, :315-347 rerun the whole model per trial) as M serial runs of the same
network.  Members never interact, so here their grids sit side by side in one (nrow, M * ncol) mosaic: every
member is a set of independent basins of the drainage forest, the storage layout packs them into groups like
any other basins, and one step of the model advances all members -- each level of the
sweep carries M times the work.  Members are sharded across GPUs with parallel.shard_members.

Results per member are bit-identical to a stand-alone SbmModel with that member's parameters
(tests/test_gpu_ensemble.py): the same kernels run on the same cells in the same order.
"""
import numpy as np

from . import core

_DCOL = {1: -1, 4: -1, 7: -1, 3: 1, 6: 1, 9: 1}


def _member_ldd(ldd):
    """The member grid's LDD as it may be tiled: a cell whose flow leaves the grid through the west / east edge
    would enter the neighbouring member, so it is taken out -- set_dd never reaches it from a pit anyway, nor
    anything upstream of it (wflow/wflow_funcs.py:79-95).  The cell set_dd drops by its flat-index-0 quirk
    (:88) is flat index 0 of EVERY member's own run, so it is taken out of every tile as well."""
    ldd = np.ascontiguousarray(ldd, dtype=np.uint8).copy()
    ncol = ldd.shape[1]
    for code, dc in _DCOL.items():
        col = 0 if dc < 0 else ncol - 1
        ldd[ldd[:, col] == code, col] = 0
    if ldd[0, 0] not in (0, 5):
        alone = core.Schedule(ldd)
        order, _, _, _ = alone.partition(1)
        if 0 not in set(order[: alone.num_cells].tolist()):
            ldd[0, 0] = 0
    return ldd


class SbmEnsemble:
    """`members` copies of the grid (ldd, river) in one SbmModel.  Rasters are set for all members at once
    (set) or per member (set_members: array [members, nrow, ncol]); get returns [members, nrow, ncol]."""

    def __init__(self, ldd, river, members, **model_kw):
        self.members = int(members)
        if self.members < 1:
            raise ValueError("an ensemble needs at least one member")
        self.grid_shape = tuple(np.asarray(ldd).shape)
        tile = _member_ldd(ldd)
        riv = np.ascontiguousarray(np.asarray(river) != 0, dtype=np.uint8)
        self.ldd = tile
        self.model = core.SbmModel(np.tile(tile, (1, self.members)), np.tile(riv, (1, self.members)), **model_kw)

    # -- rasters ---------------------------------------------------------------------------
    def _mosaic(self, per_member):
        a = np.asarray(per_member, dtype=np.float32)
        if a.shape != (self.members,) + self.grid_shape:
            raise ValueError("expected an array of shape %s" % ((self.members,) + self.grid_shape,))
        return np.ascontiguousarray(np.concatenate(list(a), axis=1))

    def set(self, name, value):
        """The same raster (or scalar) for every member."""
        a = np.asarray(value, dtype=np.float32)
        if a.ndim == 0 or a.size == 1:
            self.model.set(name, a)
        else:
            self.model.set_tiled(name, a)  # one member's raster crosses PCIe; the device repeats it

    def set_members(self, name, per_member):
        self.model.set(name, self._mosaic(per_member))

    def scale(self, name, base, factors):
        """Member m gets base * factors[m] (float32, like a wflow_fit trial multiplying a table parameter)."""
        base = np.broadcast_to(np.asarray(base, dtype=np.float32), self.grid_shape)
        f = np.asarray(factors, dtype=np.float32).reshape(self.members, 1, 1)
        self.set_members(name, base[None] * f)

    def get(self, name, mv=-999.0):
        a = self.model.get(name, mv)
        nc = self.grid_shape[1]
        return np.stack([a[:, m * nc:(m + 1) * nc] for m in range(self.members)])

    # -- stepping --------------------------------------------------------------------------
    def step(self, nsteps=1):
        self.model.step(nsteps)

    def gauges_create(self, flat_index):
        """Gauge set of the given cells (flat indices of the member grid) in every member; sampling returns
        [members * len(flat_index)] values, member-major."""
        idx = np.asarray(flat_index, dtype=np.int64)
        nrow, nc = self.grid_shape
        r, c = idx // nc, idx % nc
        allidx = np.concatenate([r * (nc * self.members) + m * nc + c for m in range(self.members)])
        return self.model.gauges_create(allidx), idx.size

    def gauges_sample(self, handle, field, mv=-999.0):
        h, n = handle
        out = np.zeros(self.members * n, np.float32)
        self.model.gauges_sample(h, field, out.ctypes.data, mv=mv, wait=True)
        return out.reshape(self.members, n)

    def close(self):
        self.model.close()


class HbvEnsemble(SbmEnsemble):
    """`members` copies of a wflow_hbv grid in one HbvModel (hbv.py): wflow_fit's trials of a wflow_hbv case (`-M wflow_hbv`,
    wflow/wflow_fit.py:105-123) as one device model.  Same raster interface as SbmEnsemble; `dynamic` is one timestep for all
    members with the forcing they share."""

    def __init__(self, ldd, members, **model_kw):
        from . import hbv
        self.members = int(members)
        if self.members < 1:
            raise ValueError("an ensemble needs at least one member")
        self.grid_shape = tuple(np.asarray(ldd).shape)
        self.ldd = _member_ldd(ldd)
        self.model = hbv.HbvModel(np.tile(self.ldd, (1, self.members)), **model_kw)

    def set_states(self, state):
        """States shared by all members; the sub-step twins of SurfaceRunoff / WaterLevel start from them (resume(), :1202)."""
        for k, v in state.items():
            self.set(k, v)
        if "SurfaceRunoff" in state and "SurfaceRunoffsub" not in state:
            self.set("SurfaceRunoffsub", state["SurfaceRunoff"])
        if "WaterLevel" in state and "WaterLevelsub" not in state:
            self.set("WaterLevelsub", state["WaterLevel"])

    def dynamic(self, P, PET, TEMP=None, Inflow=None, Seepage=None):
        for name, a in (("P", P), ("PET", PET), ("TEMP", TEMP), ("Inflow", Inflow), ("Seepage", Seepage)):
            if a is not None:
                self.set(name, a)
        self.model.advance()


class CaseEnsemble:
    """M parameter trials of a wflow_sbm CASE (ini, staticmaps, intbl, forcing stacks) advanced together on one GPU.

    The reference calibrates by rerunning the whole model once per trial (wflow/wflow_fit.py:153-158 multiplies
    the maps named in ``[fit] parameter_i`` by the trial's factors, :235-257 and :315-347 run it and collect the
    discharge columns).  Here the trials are the members of one SbmEnsemble: ``factors[name][m]`` multiplies the base map
    ``name`` (a table parameter of initial(): M, KsatVer, N, N_River, SoilThickness ...) of member m before the
    derivations of initial(); forcing and every map the trials share cross PCIe once per step (wf_set_field_tiled).
    ``members`` selects the trials this process holds (parallel.shard_members); results are member-major.
    ``area`` = (id map under the case, code) limits the factors to the cells that hold the code (wflow_fit's areamap / areacode).
    """

    def __init__(self, case_dir, factors, members=None, inifile="wflow_sbm.ini", run_dir="ensemble", device=0, area=None):
        from . import params
        from .framework import WflowSbm
        F32 = np.float32
        fw = WflowSbm(Dir=case_dir, RunDir=run_dir, configfile=inifile, device=device)
        fw.createRunId(NoOverWrite=False)
        fw._runInitial()
        fw._runResume()
        self.fw = fw
        names = sorted(factors)
        nall = len(np.asarray(factors[names[0]]).reshape(-1)) if names else 1
        self.member_ids = list(range(nall)) if members is None else list(members)
        M = len(self.member_ids)
        ldd, river = fw.ldd, fw.River.astype(np.uint8)
        self.shape = ldd.shape
        it_l, it_r = fw._it_fixed
        self.ens = SbmEnsemble(ldd, river, M, timestepsecs=fw.timestepsecs, layer_thickness=fw.layer_thickness,
                               model_snow=fw.modelSnow, soil_inf_redu=fw.soilInfReduction, transfer_method=fw.TransferMethod,
                               ust=fw.UST, it_kin_l=it_l, it_kin_r=it_r, device=device)
        area_mask = None
        if area is not None:
            amap = fw._staticmap(area[0], fail=True)
            area_mask = np.nan_to_num(amap, nan=-1.0).astype(np.int64) == int(area[1])
        # statics: derive per trial, upload what the trials share once
        stat = []
        for m in self.member_ids:
            b = dict(fw._base)
            for n in names:
                scaled = (np.asarray(b[n], F32) * F32(np.asarray(factors[n]).reshape(-1)[m])).astype(F32)
                # area: (file name of an id map under the case, code) -- the factor applies where the map holds the code,
                # like wflow_fit's multVarWithPar (wflow/wflow_fit.py:186-214)
                b[n] = scaled if area_mask is None else np.where(area_mask, scaled, np.broadcast_to(np.asarray(b[n], F32), self.shape)).astype(F32)
            stat.append(params.derive_static(b, ldd, river != 0, fw.timestepsecs, cellsize=fw._cellsize, n_layers=fw.maxLayers))
        self.statics = stat
        self.varied = []
        for k in stat[0]:
            if k in ("River", "Beta"):
                continue
            a0 = np.broadcast_to(np.asarray(stat[0][k], F32), self.shape)
            if all(np.array_equal(a0, np.broadcast_to(np.asarray(s[k], F32), self.shape), equal_nan=True) for s in stat[1:]):
                self.ens.set(k, np.where(np.isnan(a0), F32(0.0), a0))
            else:
                self.varied.append(k)
                self.ens.set_members(k, np.stack([np.nan_to_num(np.broadcast_to(np.asarray(s[k], F32), self.shape), nan=0.0) for s in stat]))
        # states: the case's (instate/ or cold); the water table follows each trial's soil depth (resume(), :2503-2505)
        for k in fw._state_names():
            if k == "SatWaterDepth":
                continue
            self.ens.set(k, np.nan_to_num(np.where(fw._mod.get(k, -999.0) == F32(-999.0), np.nan, fw._mod.get(k, -999.0)), nan=0.0))
        sat = fw._get("SatWaterDepth")
        sat = np.where(sat == F32(-999.0), F32(0.0), sat)
        zi = [np.maximum(F32(0.0), s["SoilThickness"] - sat / (s["thetaS"] - s["thetaR"]).astype(F32)).astype(F32) for s in stat]
        self.ens.set_members("zi", np.stack([np.nan_to_num(z, nan=0.0) for z in zi]))
        self.gauge_cells = None

    @property
    def members(self):
        return len(self.member_ids)

    @property
    def model(self):
        return self.ens.model

    def forcing(self, step):
        """(P, PET[, TEMP]) rasters of timestep `step` (1-based) of the case, as the framework reads them."""
        fw = self.fw
        out = [("P", fw._forcing("P", "Precipitation", "/inmaps/P", step, 0.0)),
               ("PET", fw._forcing("PET", "EvapoTranspiration", "/inmaps/PET", step, 0.0))]
        if fw.modelSnow:
            out.append(("TEMP", fw._forcing("TEMP", "Temperature", "/inmaps/TEMP", step, 10.0)))
        return [(n, np.nan_to_num(np.asarray(a, np.float32), nan=0.0)) for n, a in out]

    def gauges(self, flat_index):
        self.gauge_cells = np.asarray(flat_index, np.int64)
        self._gh = self.ens.gauges_create(self.gauge_cells)
        return self._gh

    def run(self, nsteps, first_step=1, field="RiverRunoff"):
        """Advance all trials `nsteps` timesteps; returns the gauge series [nsteps, members, gauges] (wflow_fit's Q columns)."""
        out = []
        for s in range(nsteps):
            for name, a in self.forcing(first_step + s):
                self.ens.set(name, a)
            self.ens.step()
            if self.gauge_cells is not None:
                out.append(self.ens.gauges_sample(self._gh, field))
        return np.stack(out) if out else None

    def close(self):
        self.ens.close()
        self.fw._mod.close()
