"""`SbmModel`: thin Python owner of a libwflowsbm handle (one model grid on one B200).

This is the raster-level layer the framework mirror (wflow_b200.framework) and the BMI
(wflow_b200.bmi) sit on.  It exposes exactly what the reference's WflowModel does around
its numba kernels: set rasters, step, read rasters -- with the arithmetic on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib

AREA_OPS = {"average": 0, "total": 1, "maximum": 2, "minimum": 3}


def build_schedule(ldd):
    """Host-only level schedule of an LDD raster (no GPU needed): (nodes, nodes_up, lev_off)
    in the reference's set_dd form (wflow/wflow_funcs.py:70-95), flattened."""
    l = _lib.lib()
    ldd = np.ascontiguousarray(ldd, dtype=np.uint8)
    h = C.c_void_p()
    _lib.check(l.wf_schedule_create(ldd.ctypes.data, ldd.shape[0], ldd.shape[1], C.byref(h)))
    try:
        n, nl = int(l.wf_schedule_num_cells(h)), int(l.wf_schedule_num_levels(h))
        nodes = np.zeros(max(n, 1), np.int64)
        up = np.zeros(max(n, 1) * 8, np.int64)
        off = np.zeros(nl + 1, np.int64)
        _lib.check(l.wf_schedule_get(h, nodes.ctypes.data, up.ctypes.data, off.ctypes.data))
    finally:
        l.wf_schedule_destroy(h)
    return nodes[:n], up[: n * 8].reshape(n, 8), off


def _solver_batch(fn, args, nout, device):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in args]
    n = arrs[0].size
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    out = np.zeros(n * nout, np.float64)
    _lib.check(fn(int(device), n, ptrs, out.ctypes.data))
    return out if nout == 1 else out.reshape(n, nout)


def kinematic_wave(Qin, Qold, q, alpha, beta, deltaT, deltaX, device=0):
    """Batched kinematic_wave on the GPU (reference: wflow/wflow_funcs.py:119-152)."""
    return _solver_batch(_lib.lib().wf_kinematic_wave_batch, [Qin, Qold, q, alpha, beta, deltaT, deltaX], 1, device)


def kinematic_wave_ssf(*args, device=0):
    """Batched kinematic_wave_ssf on the GPU (reference: wflow/wflow_funcs.py:210-292);
    returns an (n, 3) array of (ssf, zi, exfilt)."""
    if len(args) != 14:
        raise ValueError("kinematic_wave_ssf takes 14 arrays")
    return _solver_batch(_lib.lib().wf_kinematic_wave_ssf_batch, list(args), 3, device)


class Schedule:
    """Host-only level schedule (no GPU needed).  Can be handed to SbmModel(schedule=...) so the
    network is built once when a quantity derived from it (e.g. the river mask from the upstream
    area) is needed before the model exists."""

    def __init__(self, ldd):
        self._l = _lib.lib()
        self.ldd = np.ascontiguousarray(ldd, dtype=np.uint8)
        self.shape = self.ldd.shape
        h = C.c_void_p()
        _lib.check(self._l.wf_schedule_create(self.ldd.ctypes.data, self.shape[0], self.shape[1], C.byref(h)))
        self._h = h

    @property
    def num_cells(self):
        return int(self._l.wf_schedule_num_cells(self._h))

    @property
    def num_levels(self):
        return int(self._l.wf_schedule_num_levels(self._h))

    def catchment_total(self, values=None):
        out = np.zeros(self.shape, np.float64)
        v = None if values is None else np.ascontiguousarray(values, np.float32)
        _lib.check(self._l.wf_schedule_catchment_total(self._h, None if v is None else v.ctypes.data, out.ctypes.data))
        return out

    def partition(self, ngroups):
        """Storage layout for `ngroups` groups of whole basins: (storage_order, group_off, up_start, nup)."""
        n = self.num_cells
        order = np.zeros(max(n, 1), np.int64)
        goff = np.zeros(int(ngroups) + 1, np.int64)
        ups = np.zeros(max(n, 1), np.int64)
        nup = np.zeros(max(n, 1), np.uint8)
        g = _lib.check(self._l.wf_schedule_partition(self._h, int(ngroups), order.ctypes.data, goff.ctypes.data,
                                                     ups.ctypes.data, nup.ctypes.data))
        return order[:n], goff[: g + 1], ups[:n], nup[:n]

    def _release(self):
        h, self._h = self._h, None
        return h

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.wf_schedule_destroy(self._h)
            self._h = None


class SbmModel:
    def __init__(self, ldd, river=None, *, schedule=None, timestepsecs=86400, layer_thickness=(), model_snow=1, soil_inf_redu=0,
                 transfer_method=0, ust=0, glacier=0, it_kin_l=1, it_kin_r=1, device=0, beta=0.0):
        self._l = _lib.lib()
        ldd = np.ascontiguousarray(ldd, dtype=np.uint8)
        if ldd.ndim != 2:
            raise ValueError("ldd must be a 2-D raster")
        self.shape = ldd.shape
        self._ldd = ldd
        riv = None if river is None else np.ascontiguousarray(np.asarray(river) != 0, dtype=np.uint8)
        cfg = _lib.WfConfig()
        cfg.nrow, cfg.ncol = self.shape
        cfg.n_layer_thickness = len(layer_thickness)
        for i, t in enumerate(layer_thickness):
            cfg.layer_thickness[i] = float(t)
        cfg.model_snow, cfg.soil_inf_redu, cfg.transfer_method = int(model_snow), int(soil_inf_redu), int(transfer_method)
        cfg.ust, cfg.glacier, cfg.it_kin_l, cfg.it_kin_r = int(ust), int(glacier), int(it_kin_l), int(it_kin_r)
        cfg.device, cfg.timestepsecs, cfg.beta = int(device), float(timestepsecs), float(beta)
        self.cfg = cfg
        self.max_layers = len(layer_thickness) + 1 if len(layer_thickness) else 1
        h = C.c_void_p()
        _lib.check(self._l.wf_create_from_schedule(
            C.byref(cfg), None if schedule is None else schedule._release(), ldd.ctypes.data_as(C.c_void_p),
            None if riv is None else riv.ctypes.data_as(C.c_void_p), C.byref(h)))
        self._h = h

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._l.wf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- network ---------------------------------------------------------------------------
    @property
    def num_cells(self):
        return int(self._l.wf_num_cells(self._h))

    @property
    def num_levels(self):
        return int(self._l.wf_num_levels(self._h))

    @property
    def num_river_cells(self):
        return int(self._l.wf_num_river_cells(self._h))

    def network(self, river=False):
        """(nodes, nodes_up, lev_off) in the reference's set_dd form, flattened."""
        n = int(self._l.wf_num_river_cells(self._h) if river else self._l.wf_num_cells(self._h))
        nl = int(self._l.wf_num_river_levels(self._h) if river else self._l.wf_num_levels(self._h))
        nodes = np.zeros(max(n, 1), np.int64)
        up = np.zeros(max(n, 1) * 8, np.int64)
        off = np.zeros(nl + 1, np.int64)
        _lib.check(self._l.wf_get_network(self._h, int(river), nodes.ctypes.data, up.ctypes.data, off.ctypes.data))
        return nodes[:n], up[: n * 8].reshape(n, 8), off

    def cell_order(self):
        o = np.zeros(max(self.num_cells, 1), np.int64)
        _lib.check(self._l.wf_cell_order(self._h, o.ctypes.data))
        return o[: self.num_cells]

    def catchment_total(self, values=None):
        out = np.zeros(self.shape, np.float64)
        v = None if values is None else np.ascontiguousarray(values, np.float32)
        _lib.check(self._l.wf_catchment_total(self._h, None if v is None else v.ctypes.data, out.ctypes.data))
        return out

    # -- fields ----------------------------------------------------------------------------
    def set(self, name, value):
        a = np.asarray(value, dtype=np.float32)
        if a.ndim == 0 or a.size == 1:
            _lib.check(self._l.wf_set_field_uniform(self._h, name.encode(), float(a.reshape(-1)[0])))
            return
        if a.shape != self.shape:
            raise ValueError("raster %s has shape %s, model grid is %s" % (name, a.shape, self.shape))
        a = np.ascontiguousarray(a)
        _lib.check(self._l.wf_set_field(self._h, name.encode(), a.ctypes.data))
        # the copy is asynchronous on the model's stream: keep the host buffer alive until it is consumed
        _lib.check(self._l.wf_synchronize(self._h))

    def set_tiled(self, name, value, wait=True):
        """A raster of tile_ncol = value.shape[1] columns that repeats across the model grid's columns (ensemble members)."""
        a = np.ascontiguousarray(value, dtype=np.float32)
        if a.ndim != 2 or a.shape[0] != self.shape[0] or self.shape[1] % a.shape[1]:
            raise ValueError("raster %s of shape %s does not tile the model grid %s" % (name, a.shape, self.shape))
        _lib.check(self._l.wf_set_field_tiled(self._h, name.encode(), a.ctypes.data, int(a.shape[1])))
        if wait:
            _lib.check(self._l.wf_synchronize(self._h))

    def set_many(self, fields):
        for k, v in fields.items():
            self.set(k, v)

    def set_pinned(self, name, host_ptr):
        """Asynchronous upload from a caller-owned (ideally pinned) nrow*ncol float32 buffer."""
        _lib.check(self._l.wf_set_field(self._h, name.encode(), host_ptr))

    def set_pinned_tiled(self, name, host_ptr, tile_ncol):
        """Asynchronous upload of a caller-owned (pinned) nrow*tile_ncol raster that repeats across the grid's columns."""
        _lib.check(self._l.wf_set_field_tiled(self._h, name.encode(), host_ptr, int(tile_ncol)))

    def get(self, name, mv=-999.0):
        out = np.empty(self.shape, np.float32)
        _lib.check(self._l.wf_get_field(self._h, name.encode(), out.ctypes.data, float(mv)))
        return out

    def request(self, *names, enable=True):
        for n in names:
            _lib.check(self._l.wf_request_output(self._h, n.encode(), int(enable)))

    def bind(self, name, dptr):
        """Zero-copy: read field `name` from a caller-owned device array in network order."""
        _lib.check(self._l.wf_bind_field(self._h, name.encode(), dptr))

    def timing_reset(self):
        _lib.check(self._l.wf_timing_reset(self._h))

    def timing_get(self):
        ms = (C.c_double * 3)()
        n = C.c_int64()
        _lib.check(self._l.wf_timing_get(self._h, ms, C.byref(n)))
        return tuple(ms), int(n.value)

    def device_ptr(self, name):
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self._l.wf_field_device_ptr(self._h, name.encode(), C.byref(p), C.byref(n)))
        return p.value, n.value

    # -- reservoirs / lakes (wflow_sbm.py:2822-2883) ----------------------------------------
    def reservoirs_create(self, locs, areas, ldd_org):
        """Simple reservoirs: id rasters ReserVoirSimpleLocs / ReservoirSimpleAreas and the LDD before the objects were
        turned into pits.  Returns the number of reservoirs."""
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (locs, areas)]
        l = np.ascontiguousarray(ldd_org, dtype=np.uint8)
        return _lib.check(self._l.wf_reservoirs_create(self._h, a[0].ctypes.data, a[1].ctypes.data, l.ctypes.data))

    def irrigation_create(self, areas, intakes):
        """Irrigation areas and their river intakes (id rasters IrrigationAreas / IrrigationSurfaceIntakes,
        wflow_sbm.py:1096-1115, 3014-3031, 3314-3328): the demand of an area becomes the negative Inflow of the intakes
        with its id, the supply comes back as IRSupplymm in the next step.  Returns the number of areas."""
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (areas, intakes)]
        return _lib.check(self._l.wf_irrigation_create(self._h, a[0].ctypes.data, a[1].ctypes.data))

    def paddy_create(self, areas, intakes):
        """Paddy areas and their river intakes (id rasters IrrigationPaddyAreas / IrrigationSurfaceIntakes, 0 = none;
        wflow_sbm.py:762-785, 2811-2815, 3033-3050, 3330-3346): fields h_p / h_min / h_max / CRPST, state PondingDepth.
        Returns the number of areas."""
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (areas, intakes)]
        return _lib.check(self._l.wf_paddy_create(self._h, a[0].ctypes.data, a[1].ctypes.data))

    def lakes_create(self, locs, linked, areas, ldd_org, sh=None, hq=None):
        """Natural lakes: id rasters LakeLocs / LinkedLakeLocs / LakeAreasMap, the original LDD and the curve tables
        sh = {lake id: array[rows, 2] (H, S)}, hq = {lake id: array[rows, 367] (H, Q per day of the year)}."""
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (locs, linked, areas)]
        l = np.ascontiguousarray(ldd_org, dtype=np.uint8)

        def pack(d, ncol):
            d = d or {}
            ids = np.array(sorted(d), np.int32)
            rows = np.array([np.asarray(d[k]).reshape(-1, ncol).shape[0] for k in ids], np.int32)
            data = (np.concatenate([np.asarray(d[k], np.float32).reshape(-1, ncol) for k in ids]) if len(ids)
                    else np.zeros((0, ncol), np.float32))
            return ids, rows, np.ascontiguousarray(data, np.float32)
        si, sr, sd = pack(sh, 2)
        hi, hr, hd = pack(hq, 367)
        return _lib.check(self._l.wf_lakes_create(self._h, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, l.ctypes.data,
                                                  len(si), si.ctypes.data, sr.ctypes.data, sd.ctypes.data,
                                                  len(hi), hi.ctypes.data, hr.ctypes.data, hd.ctypes.data))

    def set_option(self, name, value):
        """[model] options that switch optional parts of dynamic() on: MassWasting, Abstractions, maxitsupply, breakoff."""
        _lib.check(self._l.wf_set_option(self._h, name.encode(), float(value)))

    @property
    def pipeline_mode(self):
        """How the last step_pipelined ran: "skewed" (k_pipeline), "ahead" (per-step kernels, next column on idle SMs) or None."""
        return {1: "skewed", 2: "ahead"}.get(int(self._l.wf_pipeline_mode(self._h)))

    @property
    def supply_iterations(self):
        return int(self._l.wf_supply_iterations(self._h))

    def set_day_of_year(self, jdoy):
        _lib.check(self._l.wf_set_day_of_year(self._h, int(jdoy)))

    # -- kinematic-wave sub-steps ---------------------------------------------------------
    def estimate_iterations(self, which):
        """estimate_iterations_kin_wave (wflow/wflow_funcs.py:103-116); which = "land" | "river"."""
        it = C.c_int32(1)
        _lib.check(self._l.wf_estimate_iterations(self._h, {"land": 0, "river": 1}[which], C.byref(it)))
        return int(it.value)

    def set_substeps(self, it_kin_l, it_kin_r):
        _lib.check(self._l.wf_set_substeps(self._h, int(it_kin_l), int(it_kin_r)))

    # -- stepping --------------------------------------------------------------------------
    def step(self, nsteps=1):
        _lib.check(self._l.wf_step(self._h, int(nsteps)))

    def synchronize(self):
        _lib.check(self._l.wf_synchronize(self._h))

    # -- several timesteps per launch (include/wflow_sbm.h: wf_step_pipelined) ---------------
    def pipeline_reserve(self, max_steps):
        _lib.check(self._l.wf_pipeline_reserve(self._h, int(max_steps)))

    def pipeline_set_forcing(self, name, step, value, wait=True):
        """Forcing raster (P, PET or TEMP) of step `step` of the next pipelined launch.  `value` is a
        numpy raster, or (wait=False) a raw pointer to a caller-owned pinned nrow*ncol float32 buffer."""
        if isinstance(value, int):
            _lib.check(self._l.wf_pipeline_set_forcing(self._h, name.encode(), int(step), value))
            return
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float32), self.shape))
        _lib.check(self._l.wf_pipeline_set_forcing(self._h, name.encode(), int(step), a.ctypes.data))
        _lib.check(self._l.wf_synchronize(self._h))  # the copy is asynchronous: `a` must outlive it

    def pipeline_forcing_ptr(self, name):
        """(device pointer, stride): the [max_steps][stride] float32 forcing buffer in storage order."""
        p, st = C.c_void_p(), C.c_int64()
        _lib.check(self._l.wf_pipeline_forcing_ptr(self._h, name.encode(), C.byref(p), C.byref(st)))
        return p.value, int(st.value)

    def pipeline_capture(self, gauges):
        _lib.check(self._l.wf_pipeline_capture(self._h, -1 if gauges is None else int(gauges)))

    def step_pipelined(self, nsteps, gauge_out=None, mv=-999.0, wait=True):
        """Run `nsteps` timesteps in one launch, step s+1 trailing step s down the network.
        gauge_out: None, a numpy float32 array [nsteps, ngauges] (filled on return), or a raw pointer."""
        if gauge_out is None:
            ptr = None
        elif isinstance(gauge_out, int):
            ptr = gauge_out
        else:
            if gauge_out.dtype != np.float32 or not gauge_out.flags.c_contiguous:
                raise ValueError("gauge_out must be a C-contiguous float32 array")
            ptr, wait = gauge_out.ctypes.data, True
        _lib.check(self._l.wf_step_pipelined(self._h, int(nsteps), ptr, float(mv), 0 if wait else 1))

    def last_step_ms(self):
        ms = (C.c_float * 3)()
        _lib.check(self._l.wf_last_step_ms(self._h, ms))
        return tuple(ms)

    @property
    def lateral_plan(self):
        """How the lateral sweep runs: dict(mode=grid|cluster|cta, cluster_size, groups, blocks)."""
        out = (C.c_int32 * 4)()
        _lib.check(self._l.wf_lateral_plan(self._h, out))
        return {"mode": ("grid", "cluster", "cta", "cluster-wide")[out[0]], "cluster_size": int(out[1]), "groups": int(out[2]),
                "blocks": int(out[3])}

    @property
    def launch_count(self):
        return int(self._l.wf_launch_count(self._h))

    @property
    def stream(self):
        return self._l.wf_stream(self._h)

    def quick_suspend(self):
        _lib.check(self._l.wf_quick_suspend(self._h))

    def quick_resume(self):
        _lib.check(self._l.wf_quick_resume(self._h))

    # -- gauges / areas -------------------------------------------------------------------
    def area_create(self, idmap):
        a = np.ascontiguousarray(idmap, dtype=np.int32)
        if a.shape != self.shape:
            raise ValueError("id map shape mismatch")
        h = _lib.check(self._l.wf_area_create(self._h, a.ctypes.data))
        n = _lib.check(self._l.wf_area_num_ids(self._h, h))
        ids = np.zeros(max(n, 1), np.int32)
        _lib.check(self._l.wf_area_ids(self._h, h, ids.ctypes.data))
        return h, ids[:n]

    def area_reduce(self, area, field, op="average"):
        n = _lib.check(self._l.wf_area_num_ids(self._h, area))
        out = np.zeros(max(n, 1), np.float32)
        _lib.check(self._l.wf_area_reduce(self._h, area, field.encode(), AREA_OPS[op], out.ctypes.data))
        return out[:n]

    def gauges_create(self, flat_index):
        """Persistent gauge set (cells resolved once); returns a handle for gauges_sample."""
        idx = np.ascontiguousarray(flat_index, dtype=np.int64)
        return _lib.check(self._l.wf_gauges_create(self._h, idx.ctypes.data, idx.size))

    def gauges_sample(self, handle, field, out_ptr, mv=-999.0, wait=True):
        """Sample `field` at the gauge set into the caller's float32 buffer (raw pointer; pinned memory
        when wait=False, valid after synchronize())."""
        _lib.check(self._l.wf_gauges_sample(self._h, int(handle), field.encode(), out_ptr, float(mv), 0 if wait else 1))

    def gauges_sample_device(self, handle, field, dev_ptr, mv=-999.0):
        """Sample `field` at the gauge set into a caller-owned DEVICE float32 array (raw pointer), on the model's stream."""
        _lib.check(self._l.wf_gauges_sample_device(self._h, int(handle), field.encode(), dev_ptr, float(mv)))

    def sample(self, field, flat_index, mv=-999.0):
        idx = np.ascontiguousarray(flat_index, dtype=np.int64)
        out = np.zeros(max(idx.size, 1), np.float32)
        _lib.check(self._l.wf_sample(self._h, field.encode(), idx.ctypes.data, idx.size, out.ctypes.data, float(mv)))
        return out[: idx.size]
