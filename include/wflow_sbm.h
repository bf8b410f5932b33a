/*
 * wflow_sbm.h -- C-ABI of libwflowsbm.so: the B200-native per-timestep hot path of
 * wflow_sbm (SBM vertical column + LDD-ordered kinematic wave for subsurface, overland
 * and river flow + gauge/area accumulation).
 *
 * Plain C: opaque handle, raw pointers and sizes, int status codes.  No torch types.
 * Every entry point cites the reference interface it replaces (paths relative to the
 * openstreams/wflow checkout).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to WflowModel to call these instead of sbm_cell()/kin_wave().
 *
 * Conventions
 *   * Rasters are float32, row-major, north-up, nrow*ncol, exactly what
 *     pcr.pcr2numpy(map, mv) hands to the reference's numba kernels
 *     (wflow/wflow_sbm.py:2885-2909).  Cells outside the drainage network are ignored
 *     on input and receive `mv` on output.
 *   * Field names are the reference's attribute names, list-valued
 *     ones as Name_<k> like the reference's state files (wf_DynamicFramework.py:1780-1827).
 *   * All functions return 0 on success, a negative WF_E* code otherwise;
 *     wf_last_error() returns a message for the calling thread's last failure.
 *   * A handle is bound to one CUDA device and is not thread-safe.  There is no CPU
 *     fallback: wf_create fails with WF_ENODEVICE when no sm_100 GPU is usable.
 */
#ifndef WFLOW_SBM_H
#define WFLOW_SBM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WF_MAX_LAYERS 8

enum {
  WF_OK = 0,
  WF_EINVAL = -1,    /* bad argument / unknown field name */
  WF_ENODEVICE = -2, /* no usable CUDA device, or a CUDA call failed */
  WF_ENOMEM = -3,
  WF_ECYCLE = -4,    /* the LDD contains a cycle */
  WF_EUNSUPPORTED = -5
};

/* Model options: the [model]/[run] ini settings that reach the hot path
 * (wflow/wflow_sbm.py:1233-1339, 1757-1766). */
typedef struct wf_config {
  int32_t nrow, ncol;
  int32_t n_layer_thickness;                 /* entries of [model] UStoreLayerThickness (0 = unset -> one layer) */
  double layer_thickness[WF_MAX_LAYERS];     /* mm */
  int32_t model_snow;                        /* [model] ModelSnow */
  int32_t soil_inf_redu;                     /* [model] soilInfRedu */
  int32_t transfer_method;                   /* [model] transfermethod */
  int32_t ust;                               /* [model] Whole_UST_Avail */
  int32_t glacier;                           /* GlacierFrac parameter configured */
  int32_t it_kin_l;                          /* overland sub-steps  (wflow_sbm.py:2911-2923) */
  int32_t it_kin_r;                          /* river sub-steps     (wflow_sbm.py:3159-3171) */
  int32_t device;                            /* CUDA device ordinal */
  double timestepsecs;                       /* [run] timestepsecs */
  double beta;                               /* kinematic-wave exponent; the reference fixes 0.6 as REAL4 (wflow_sbm.py:1643); 0 = that default */
} wf_config;

typedef struct wf_model wf_model;

const char *wf_last_error(void);
const char *wf_version(void);

/* --- network ------------------------------------------------------------------------
 * Replaces set_dd(ldd) (wflow/wflow_funcs.py:70-95), called for the full LDD and for the
 * river-only LDD (wflow/wflow_sbm.py:2385-2389), and owns the level schedule.
 * ldd: PCRaster codes 1..9 (5 = pit), anything else = missing.  river: non-zero = river. */
int wf_create(const wf_config *cfg, const uint8_t *ldd, const uint8_t *river, wf_model **out);
void wf_destroy(wf_model *m);

/* Host-only schedule construction (no GPU needed): the same builder wf_create uses. */
typedef struct wf_schedule wf_schedule;
int wf_schedule_create(const uint8_t *ldd, int nrow, int ncol, wf_schedule **out);
void wf_schedule_destroy(wf_schedule *s);
int64_t wf_schedule_num_cells(const wf_schedule *s);
int64_t wf_schedule_num_levels(const wf_schedule *s);
/* nodes / nodes_up / lev_off as in wf_get_network. */
int wf_schedule_get(const wf_schedule *s, int64_t *nodes, int64_t *nodes_up, int64_t *lev_off);
/* catchmenttotal over a host-only schedule (see wf_catchment_total). */
int wf_schedule_catchment_total(const wf_schedule *s, const float *values, double *out_grid);
/* wf_create reusing an already built full-network schedule (ownership passes to the model; the
 * schedule handle must not be used or destroyed afterwards). */
int wf_create_from_schedule(const wf_config *cfg, wf_schedule *s, const uint8_t *ldd, const uint8_t *river,
                            wf_model **out);

/* Storage layout (host-only, no GPU needed).  The device arrays are stored group-major: a group is
 * a set of whole basins (cells draining to the same pit) that one thread-block cluster sweeps on
 * its own, level by level; set_dd's walk over all basins at once (wflow/wflow_funcs.py:79-95) is
 * thereby split along its independent trees.  Packs the basins of `s` into at most `ngroups` groups
 * and returns the number used.  Outputs (any may be NULL): storage_order[ncells] flat grid index of
 * each storage position; group_off[ngroups+1] storage range of each group; up_start[ncells] /
 * nup[ncells] the contiguous run of upstream neighbours of each cell, in storage positions. */
int wf_schedule_partition(const wf_schedule *s, int ngroups, int64_t *storage_order, int64_t *group_off,
                          int64_t *up_start, uint8_t *nup);

/* --- wflow_hbv ------------------------------------------------------------------------
 * A model whose timestep is wflow_hbv's: replaces one pass of WflowModel.dynamic() of wflow/wflow_hbv.py:1222-1767 -- the
 * HBV-96 column (rain / snow split, interception, snow pack, glacierHBV, soil moisture, upper and lower zone; REAL4 map
 * algebra in the reference, one thread per cell here) and kin_wave (wflow/wflow_funcs.py:155-207) on the FULL network with
 * every cell a channel (wflow_hbv.py:1633-1647), which is the river task of the same level sweep wflow_sbm's river uses.
 * Of cfg: nrow, ncol, timestepsecs, it_kin_r (sub-steps of the wave; wf_set_substeps / wf_estimate_iterations(which = 1)
 * for [model] kinwaveIters), glacier, beta, device.  Fields by the reference's attribute names (wf_set_field / wf_get_field,
 * wf_request_output for the outputs): forcing P PET TEMP Inflow Seepage; parameters Pcorr TempCor TTI TT SFCF RFCF ICF EPF
 * ECORR CEVPF Cfmax CFR WHC FieldCapacity Treshold BetaSeepage PERC Cflux KQuickFlow AlphaNL K4 K0 SUZ xl yl DCL Bw Alpha
 * (+ GlacierFrac G_TT G_Cfmax G_SIfrac); states FreeWater DrySnow SoilMoisture UpperZoneStorage LowerZoneStorage
 * InterceptionStorage SurfaceRunoff SurfaceRunoffsub WaterLevel WaterLevelsub (+ GlacierStore).  wf_set_option:
 * "SetKquickFlow", "ExternalQbase", "MassWasting" (with the Slope raster; one more level sweep between the column and
 * the channel wave, :1351-1364).  Not restated, refused: reservoirs / lakes, updating from observed
 * discharge, multi-step launches.  wf_step, wf_gauges_*, wf_area_*, wf_sample, wf_quick_suspend / resume work as for
 * wflow_sbm models. */
int wf_hbv_create(const wf_config *cfg, const uint8_t *ldd, wf_model **out);

int64_t wf_num_cells(const wf_model *m);        /* cells in the drainage network */
int64_t wf_num_levels(const wf_model *m);
int64_t wf_num_river_cells(const wf_model *m);  /* cells of the river network (rnodes) */
int64_t wf_num_river_levels(const wf_model *m);
/* The schedule in the reference's own form, for bit-exact comparison with set_dd:
 * nodes[ncells] (flat grid indices, level-major, level 0 = farthest from its pit),
 * nodes_up[ncells*8] (-1 padded, neighbour order NW,N,NE,W,E,SW,S,SE), lev_off[nlevels+1].
 * which = 0: full network, 1: river network.  Any output pointer may be NULL. */
int wf_get_network(const wf_model *m, int which, int64_t *nodes, int64_t *nodes_up, int64_t *lev_off);
/* catchmenttotal(values, ldd) on the model's network (PCRaster op used by initial(),
 * wflow/wflow_sbm.py:2070-2073, 3468): out = values accumulated over the upstream area.
 * values == NULL accumulates 1 per cell. */
int wf_catchment_total(const wf_model *m, const float *values, double *out_grid);

/* --- fields -------------------------------------------------------------------------
 * Replace the pcr2numpy / numpy2pcr copies around the kernels (wflow/wflow_sbm.py:2885-2909,
 * 2944-3129, 3190-3201) and wf_supplyMapAsNumpy / wf_setValuesAsNumpy
 * (wflow/wf_DynamicFramework.py:2582, 2203) at the raster level. */
int wf_set_field(wf_model *m, const char *name, const float *grid);        /* host raster -> device */
int wf_set_field_uniform(wf_model *m, const char *name, float value);
/* A raster of nrow x tile_ncol cells that repeats across the columns of the model grid (ncol a multiple of tile_ncol):
 * parameter ensembles hold their members side by side in one grid (wflow/wflow_fit.py:235-257 reruns the model per
 * trial), and everything the members share -- forcing, the parameters a trial does not touch -- crosses PCIe once. */
int wf_set_field_tiled(wf_model *m, const char *name, const float *grid, int tile_ncol);
int wf_get_field(wf_model *m, const char *name, float *grid, float mv);    /* device -> host raster */
/* Optional statics switch parts of the step on by being set: GlacierFrac / G_TT / G_Cfmax / G_SIfrac (with
 * cfg.glacier; glacierHBV, wflow/wflow_funcs.py:549-603) and LAI / Sl / Swood / Kext (LAI-driven Cmax,
 * CanopyGapFraction and EoverR, wflow/wflow_sbm.py:2587-2603; wf_step fails with WF_EINVAL if LAI is set and one of
 * the other three is not). */
/* Enable (1) / disable (0) a diagnostic output so that the kernels materialise it.  With no diagnostic requested
 * the step runs builds of the kernels that have the outputs compiled out. */
int wf_request_output(wf_model *m, const char *name, int enable);
/* Number of known field names and the i-th name / kind (0 forcing, 1 static, 2 state, 3 diagnostic). */
int wf_num_fields(void);
const char *wf_field_name(int i);
int wf_field_kind(int i);
/* Device-resident access (for callers that already hold data on the GPU): pointer to the
 * field's device array in network order (see wf_cell_order) and its length.
 * River-only fields (kind has bit 8 set) are in river-network order. */
int wf_field_device_ptr(wf_model *m, const char *name, void **dptr, int64_t *n);
int wf_cell_order(const wf_model *m, int64_t *flat_index /* [wf_num_cells] */);
/* Same as wf_set_field, but `grid` is a DEVICE pointer to an nrow*ncol raster. */
int wf_set_field_from_device(wf_model *m, const char *name, const float *dgrid);
/* Zero-copy forcing: make the model read field `name` from a caller-owned DEVICE array that is
 * already in network order (river order for river-only fields).  dptr == NULL unbinds. */
int wf_bind_field(wf_model *m, const char *name, float *dptr);

/* --- stepping -----------------------------------------------------------------------
 * Replaces one pass of WflowModel.dynamic() (wflow/wflow_sbm.py:2518-3528): the PCRaster
 * phases (interception, snow, evaporation split), sbm_cell() (:332-900) and kin_wave()
 * (wflow/wflow_funcs.py:155-207), with states re-quantised to float32 at the step boundary
 * like the reference's numpy2pcr copies.  nsteps > 1 repeats with the same forcing. */
int wf_step(wf_model *m, int nsteps);
int wf_synchronize(wf_model *m);

/* --- several timesteps per launch ------------------------------------------------------
 * Replaces a run of consecutive passes of the time loop, wf_DynamicFramework._runDynamic
 * (wflow/wf_DynamicFramework.py:2968-3042) calling WflowModel.dynamic() (wflow/wflow_sbm.py:2518-3528),
 * when the forcing of the coming steps is known (map stacks / netCDF on disk, BMI update_until):
 * timestep s+1 follows timestep s down the drainage network at a distance of 3 + it_kin_r levels
 * inside ONE kernel, instead of waiting for s to reach the outlet.  Results are bit-identical to
 * `nsteps` calls of wf_step with the same forcing.
 *   wf_pipeline_reserve      room for the forcing of `max_steps` steps per launch
 *   wf_pipeline_set_forcing  host raster of P, PET or TEMP for step `step` (0-based) of the next launch;
 *                            asynchronous like wf_set_field (keep the buffer alive until the launch
 *                            has consumed it)
 *   wf_pipeline_forcing_ptr  the device buffer itself, [max_steps][*stride] float32 in storage order
 *                            (wf_cell_order), for forcing produced on the device
 *   wf_pipeline_capture      gauge set (wf_gauges_create; -1 = none) at which every step's RiverRunoff is
 *                            kept -- the one output a step must hand over before the next overwrites it
 *                            (wf_OutputTimeSeriesArea's sampled discharge, wf_DynamicFramework.py:446-499)
 *   wf_step_pipelined        runs `nsteps` (<= 64) steps; gauge_out (may be NULL) receives [nsteps][gauges] values
 *                            (0 at network cells outside the river, mv outside the network); async != 0
 *                            returns without waiting (pinned gauge_out, valid after wf_synchronize).
 * All other fields hold, after the call, what they hold after the last of the steps. */
int wf_pipeline_reserve(wf_model *m, int max_steps);
int wf_pipeline_set_forcing(wf_model *m, const char *name, int step, const float *grid);
int wf_pipeline_forcing_ptr(wf_model *m, const char *name, void **dptr, int64_t *stride);
int wf_pipeline_capture(wf_model *m, int gauges);
int wf_step_pipelined(wf_model *m, int nsteps, float *gauge_out, float mv, int async);
/* How the last wf_step_pipelined ran: 1 = the time-skewed kernel (k_pipeline), 2 = the per-step kernels with the next
 * timestep's column on the SMs a cluster sweep of several groups leaves idle (same results either way); 0 = none yet. */
int wf_pipeline_mode(const wf_model *m);
/* Device time of the kernels of the last wf_step call, milliseconds (CUDA events on the
 * model's stream): [0] vertical column, [1] lateral sweep, [2] whole step. */
int wf_last_step_ms(wf_model *m, float ms[3]);
/* Accumulated device time since wf_timing_reset over all steps (CUDA events on the model's
 * stream around each kernel): ms[0] vertical column, ms[1] lateral sweep, ms[2] whole steps;
 * *nsteps = steps covered.  wf_timing_get synchronises the stream. */
int wf_timing_reset(wf_model *m);
int wf_timing_get(wf_model *m, double ms[3], int64_t *nsteps);
/* How the lateral sweep is launched for this network: out[0] = 0 cooperative grid (one group),
 * 1 thread-block clusters, 2 single blocks, 3 clusters of wider blocks; out[1] = blocks per cluster; out[2] = groups;
 * out[3] = blocks launched (0 before the first step). */
int wf_lateral_plan(const wf_model *m, int32_t out[4]);
/* Number of kernel launches issued by the model since creation. */
int64_t wf_launch_count(const wf_model *m);
/* The CUDA stream (cudaStream_t) the model launches on. */
void *wf_stream(wf_model *m);

/* --- gauge / area accumulation --------------------------------------------------------
 * Replaces wf_OutputTimeSeriesArea.writestep (wflow/wf_DynamicFramework.py:404-499):
 * areaaverage / areatotal / areamaximum / areaminimum of a field over an id map, one
 * value per unique id > 0 (ascending).  Returns an area-set handle >= 0. */
enum { WF_AREA_AVERAGE = 0, WF_AREA_TOTAL = 1, WF_AREA_MAXIMUM = 2, WF_AREA_MINIMUM = 3 };
int wf_area_create(wf_model *m, const int32_t *idmap /* nrow*ncol, <=0 = none */);
int wf_area_num_ids(wf_model *m, int area);
int wf_area_ids(wf_model *m, int area, int32_t *ids);
int wf_area_reduce(wf_model *m, int area, const char *field, int op, float *out /* [num_ids] */);
/* Sample a field at cells (gauges): out[i] = field[flat_index[i]] (mv outside the network). */
int wf_sample(wf_model *m, const char *field, const int64_t *flat_index, int64_t n, float *out, float mv);

/* Persistent gauge set: the cells are resolved once; sampling is then one small kernel and one copy
 * on the model's stream (the per-step read of wf_OutputTimeSeriesArea's sampled cells).  With
 * async != 0 the call returns without waiting: `out` (ideally pinned memory) is valid after
 * wf_synchronize or any later blocking call on the model. */
int wf_gauges_create(wf_model *m, const int64_t *flat_index, int64_t n);
int wf_gauges_sample(wf_model *m, int gauges, const char *field, float *out /* [n] */, float mv, int async);
/* The same sample into a caller-owned DEVICE array (e.g. one row of a [T][n] block that is all-gathered over NCCL every
 * T timesteps: the gauge time series of wf_saveTimeSeries, wflow/wf_DynamicFramework.py:1829-1884, without a host
 * bounce).  Enqueued on the model's stream (wf_stream); returns at once. */
int wf_gauges_sample_device(wf_model *m, int gauges, const char *field, float *dout /* device, [n] */, float mv);

/* --- reservoirs and lakes --------------------------------------------------------------
 * simplereservoir (wflow/wflow_funcs.py:864-949) and naturalLake (:669-861), called by WflowModel.dynamic() before the
 * soil column (wflow/wflow_sbm.py:2822-2883) on the states of the step before; their outflow enters the river cell below
 * the object.  locs / linked / areas: nrow*ncol id rasters (<= 0: none) -- ReserVoirSimpleLocs / ReservoirSimpleAreas,
 * LakeLocs / LinkedLakeLocs / LakeAreasMap; ldd_org: the LDD BEFORE initial() turned the objects into pits (:1899-1907).
 * The parameter rasters (ResSimpleArea, ResMaxVolume, ResTargetFullFrac, ResMaxRelease, ResDemand, ResTargetMinFrac;
 * LakeArea, LakeThreshold, LakeStorFunc, LakeOutflowFunc, Lake_b, Lake_e), the states ReservoirVolume / LakeWaterLevel and
 * filter_P_PET (:1804-1843) are ordinary fields (wf_set_field).  Lake tables: n_sh storage curves (rows of H, S; Lake_SH_<id>.tbl)
 * and n_hq rating curves (rows of H, Q(day 1) .. Q(day 366); Lake_HQ_<id>.tbl), concatenated, keyed by lake id.
 * Return the number of objects found, or a negative status. */
int wf_reservoirs_create(wf_model *m, const int32_t *locs, const int32_t *areas, const uint8_t *ldd_org);
int wf_lakes_create(wf_model *m, const int32_t *locs, const int32_t *linked, const int32_t *areas, const uint8_t *ldd_org,
                    int n_sh, const int32_t *sh_ids, const int32_t *sh_rows, const float *sh_data,
                    int n_hq, const int32_t *hq_ids, const int32_t *hq_rows, const float *hq_data);
/* Options of [model] that switch optional parts of dynamic() on:
 *   "MassWasting"  1: dry snow slides downslope after the snow pack (pcr.accucapacitystate, wflow/wflow_sbm.py:2700-2707)
 *   "Abstractions" 1: a negative Inflow is a demand met from the river, with the reference's supply iteration (:3131-3146,
 *                     3203-3308; reproduced with its re-routing quirk).  Off: Inflow is added as it is.
 *   "maxitsupply"  iterations of that loop at most (default 5, :1241);  "breakoff" its stopping change (default 1e-4, :3206) */
int wf_set_option(wf_model *m, const char *name, double value);
/* Irrigation areas fed from river intakes (self.nrirri > 0; irrigationdemand wflow/wflow_sbm.py:927-945, :3014-3031, 3314-3328;
 * idtoid wflow/wflow_lib.py:81-101).  areas / intakes: the IrrigationAreas and IrrigationSurfaceIntakes id maps
 * (nrow*ncol, <= 0 = none; intakes on river cells).  Every step the demand of an area -- areaaverage(PotTrans -
 * Transpiration) over it, in m3/s -- becomes the negative Inflow of the intakes with its id (the "Abstractions" option is
 * switched on), and what the supply iteration gave them, times (1 - DemandReturnFlowFraction), comes back to the area as
 * IRSupplymm [mm], added to the water available for infiltration in the NEXT step (IRSupplymm is a state: get / set).
 * Not restated: irrigation return points (IrrigationSurfaceReturn), IrriDemandExternal.
 * Returns the number of areas, or a negative wf_status. */
int wf_irrigation_create(wf_model *m, const int32_t *areas, const int32_t *intakes);
/* Paddy areas fed from river intakes (self.nrpaddyirri > 0; wflow/wflow_sbm.py:762-785 ponding inside sbm_cell, :2811-2815
 * pond evaporation, :3033-3050 demand, :3330-3346 supply; idtoid wflow/wflow_lib.py:81-101).  areas / intakes: the
 * IrrigationPaddyAreas and IrrigationSurfaceIntakes id maps (0 or negative = none).  Fields: h_p (bund height; 0 = the cell
 * ponds nothing), h_min / h_max (the pond depth that triggers irrigation / that it fills to), CRPST (crop-stage multiplier,
 * a per-step raster the user's [modelparameters] supplies; 1 when not set), state PondingDepth, outputs ActEvapPond,
 * irr_depth, IrriDemandm3, IRSupplymm.  The step then sweeps twice (subsurface flow and the column tail that ponds; the
 * areas' demand; overland and river flow), with the supply iteration of the abstractions behind it.  Intake cells carry no
 * other Inflow.  Not together with wf_irrigation_create (the reference lets the paddy supply overwrite the other).
 * Returns the number of areas. */
int wf_paddy_create(wf_model *m, const int32_t *areas, const int32_t *intakes);
int wf_supply_iterations(const wf_model *m);   /* self.nrit of the last step */
/* Day of the year (1..366) of the coming steps: column of the lakes' rating tables (wf_supplyJulianDOY). */
int wf_set_day_of_year(wf_model *m, int jdoy);

/* --- kinematic-wave sub-steps --------------------------------------------------------------------
 * estimate_iterations_kin_wave(Q, Beta, alpha, timestepsecs, dx, mv) (wflow/wflow_funcs.py:103-116):
 * it_kin = ceil(1.25 * 95th percentile of the Courant number over the cells with Q > 0), 1 when no cell
 * flows.  which = 0: overland (LandRunoffsub, AlphaL, DL; wflow_sbm.py:2911-2921), 1: river
 * (RiverRunoffsub, AlphaR, DCL; :3159-3169).  A device-wide radix select; synchronises the stream. */
int wf_estimate_iterations(wf_model *m, int which, int32_t *it_kin);
/* Sub-step counts of the following steps ([model] kinwaveIters = 1: per step from the estimate above or
 * ceil(timestepsecs / kinwave*Tstep), wflow_sbm.py:2911-2923, 3159-3171). */
int wf_set_substeps(wf_model *m, int it_kin_l, int it_kin_r);

/* --- batched solvers ----------------------------------------------------------------------
 * The two Newton solvers of the path, exposed on their own so that they can be checked against
 * the reference's scalar functions: kinematic_wave(Qin, Qold, q, alpha, beta, deltaT, deltaX)
 * (wflow/wflow_funcs.py:119-152) and kinematic_wave_ssf(ssf_in, ssf_old, zi_old, r, Ks_hor, Ks,
 * slope, neff, f, D, dt, dx, w, ssf_max) -> (ssf, zi, exfilt) (wflow/wflow_funcs.py:210-292).
 * args: HOST arrays of n float64 each (7 resp. 14 of them); out: HOST array n resp. 3*n. */
int wf_kinematic_wave_batch(int device, int64_t n, const double *const *args, double *out);
int wf_kinematic_wave_ssf_batch(int device, int64_t n, const double *const *args, double *out);

/* --- one-step rollback ------------------------------------------------------------------
 * Replaces wf_QuickSuspend / wf_QuickResume (wflow/wf_DynamicFramework.py:2131-2175), used by
 * BMI update_until() to step back once: a device-side snapshot of all states. */
int wf_quick_suspend(wf_model *m);
int wf_quick_resume(wf_model *m);

#ifdef __cplusplus
}
#endif
#endif /* WFLOW_SBM_H */
