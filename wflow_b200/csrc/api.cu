// C-ABI of libwflowsbm.so (include/wflow_sbm.h): model handle, field staging, stepping.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wflow_sbm.h"
#include "fields.h"
#include "kernels.cuh"
#include "lateral.cuh"
#include "pipeline.cuh"
#include "network.h"
#include "reduce.cuh"
#include "outputs.cuh"
#include "hbv.cuh"
#include "objects.cuh"

namespace wf {

const FieldInfo g_fields[F_COUNT] = {
#define X(n, k) {#n, (int)(k)},
    WF_FIELDS(X)
#undef X
};

int field_index(const char *name) {
  static std::map<std::string, int> idx;
  if (idx.empty())
    for (int i = 0; i < F_COUNT; i++) idx[g_fields[i].name] = i;
  if (idx.size() == (size_t)F_COUNT) {
    // wflow_hbv's names of the channel rasters it shares with wflow_sbm's river (wflow_hbv.py:131-163, 1064-1068)
    idx["SurfaceRunoff"] = F_RiverRunoff;
    idx["SurfaceRunoffsub"] = F_RiverRunoffsub;
    idx["WaterLevel"] = F_WaterLevelR;
    idx["WaterLevelsub"] = F_WaterLevelRsub;
    idx["Alpha"] = F_AlphaR;
  }
  auto it = idx.find(name ? name : "");
  return it == idx.end() ? -1 : it->second;
}

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CUDA_OK(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(WF_ENODEVICE, std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

struct AreaSet {
  std::vector<int32_t> ids;   // ascending unique ids > 0
  int32_t *d_perm = nullptr;  // cells (network order index) sorted by id
  int32_t *d_seg = nullptr;   // ids.size()+1 offsets into perm
  int64_t ncell = 0;
  // two-pass reduction (reduce.cuh): chunks of the ids' runs, their partial results, the result row
  int nchunks = 0;
  int32_t *d_chunk_beg = nullptr, *d_chunk_end = nullptr, *d_id_chunk0 = nullptr;
  AreaPartial *d_partial = nullptr;
  float *d_out = nullptr;
};

}  // namespace wf

using namespace wf;

struct wf_schedule {
  Network net;
};

struct wf_model {
  wf_config cfg;
  Network net, rnet;                 // full and river-only schedules in the reference's (level-major) form
  Network snet, srnet;               // the same forests in storage order (group-major, network.h)
  GroupLayout lay, rlay;
  int lat_mode = 0, cluster_size = 1, lat_blocks = 0;  // how the lateral sweep is launched (plan_lateral)
  size_t lat_smem = 0;
  // the next step's column on the SMs a cluster sweep leaves idle (lateral.cuh column_ahead; built by setup_ahead)
  int32_t *d_ah_progress = nullptr, *d_ah_queue = nullptr, *d_ah_order = nullptr, *d_ah_need = nullptr, *d_ah_rank = nullptr;
  int ah_ntiles = 0, ah_tile = 0, ah_extra = 0, ah_nsweep = 0;
  bool ah_valid = false;
  int last_multistep = 0;             // how the last wf_step_pipelined ran: 1 time-skewed kernel, 2 per-step sweeps + column ahead
  GroupDesc *d_groups = nullptr;
  uint32_t *d_lev_off = nullptr, *d_rlev_off = nullptr, *d_tick_total = nullptr;
  std::vector<int32_t> river_slot;   // network order -> river order (or -1)
  std::vector<int64_t> river_flat;   // river order -> flat grid index (first nrnet: rnet.order)
  int64_t n = 0;
  int32_t nr = 0, nrnet = 0;
  DevModel dm;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_last[3] = {nullptr, nullptr, nullptr};
  int64_t *d_order = nullptr, *d_rorder = nullptr;
  float *d_stage = nullptr;          // nrow*ncol staging raster (wf_get_field)
  // host -> device staging of wf_set_field: a ring of rasters filled on a copy stream, so that the
  // upload of the next step's forcing overlaps the running step's kernels
  static const int kStageRing = 6;
  float *d_ring[kStageRing] = {};
  cudaEvent_t ring_copied[kStageRing] = {}, ring_free[kStageRing] = {};
  int ring_next = 0;
  cudaStream_t copy_stream = nullptr;
  // Forcing rasters set from the host (wf_set_field on P / PET / TEMP / Inflow) are put into storage order ON THE COPY
  // STREAM, into a ring of storage-order arrays per field: the model's stream only waits for the event of the array it is
  // about to read, so upload AND re-ordering of step s+1's forcing overlap the kernels of step s (the three k_gather
  // launches used to sit on the model's stream between the kernels: 0.45 ms of a 13.8 ms step, r1).
  static const int kForcSlots = 3;
  struct ForcRing {
    float *slot[kForcSlots] = {};
    cudaEvent_t ready[kForcSlots] = {}, freed[kForcSlots] = {};
    bool used[kForcSlots] = {};        // a step has been launched that reads the slot since it was filled
    int cur = -1;                      // slot dm.f[id] points at (-1: the field's own array)
    bool pending = false;              // the model's stream has not waited for ready[cur] yet
  };
  std::map<int, ForcRing> forc;
  std::vector<int32_t> pos_of_flat;  // flat grid index -> storage position (-1 outside the network)
  struct GaugeSet {
    int64_t n = 0;
    int32_t *d_pos = nullptr;        // storage position (river order for river fields is resolved per call)
    int32_t *d_rpos = nullptr;       // river-order position or -1
    float *d_out = nullptr;
    std::vector<int32_t> pos, rpos;
  };
  std::vector<GaugeSet> gauges;
  std::vector<float *> snapshot;     // quick-suspend copies of the state arrays
  std::vector<AreaSet> areas;
  int64_t launches = 0;
  int64_t max_tick_items = 1;
  bool timed = false;
  bool derived_dirty = true;          // d_efT / d_efST must be recomputed before the next step
  int ML = 1;
  bool external[F_COUNT] = {};        // field bound to a caller-owned device array
  float *owned[F_COUNT] = {};         // our own allocation while an external one is bound
  std::vector<cudaEvent_t> tev;       // timing window: 3 events per step
  size_t tev_used = 0;
  int64_t tev_extra_steps = 0;        // steps beyond the first of every pipelined launch in the timing window
  // pipelined multi-step sweep (pipeline.cuh)
  float *pipe_forc[3] = {nullptr, nullptr, nullptr};  // P, PET, TEMP of the coming steps: [pipe_cap][npad], storage order
  bool pipe_has[3] = {false, false, false};
  int pipe_cap = 0;
  int pipe_gauges = -1;               // gauge set whose river discharge is captured per step
  int32_t *d_cap_slot = nullptr, *d_cap_col = nullptr;
  float *d_cap_out = nullptr, *d_cap_exp = nullptr;
  int32_t cap_n = 0;
  int pipe_blocks = 0;
  size_t pipe_attr_smem = 0;
  // reservoirs / lakes (objects.cuh)
  struct Objects {
    ObjectSet dev = {};
    std::vector<void *> owned;        // device allocations behind dev
    bool lakes = false;
  };
  Objects reservoirs, lakes;
  IrrigationSet irri = {};            // irrigation areas and their intakes (wf_irrigation_create)
  std::vector<void *> irri_owned;
  IrrigationSet paddy = {};           // paddy areas and their intakes (wf_paddy_create)
  std::vector<void *> paddy_owned;
  std::map<int, uint32_t *> tick_total_of_mask;  // the sweep's tick sizes for a subset of its tasks (paddy: subsurface, then the rest)
  float *d_obj_inflow = nullptr;      // nr: Inflow map + outflow of the objects above each river cell
  int jdoy = 1;                       // day of the year for the lakes' rating tables (wf_set_day_of_year)
  // options of wf_set_option
  bool mass_wasting = false;          // [model] MassWasting
  float *d_mw_flux = nullptr;         // n: snow flux of every cell (k_mass_wasting)
  bool abstractions = false;          // negative Inflow = demand met from the river (supply iteration)
  int maxitsupply = 5;                // [model] maxitsupply
  double breakoff = 0.0001;           // wflow_sbm.py:3206
  SupplyArgs sup = {};
  int last_nrit = 0;                  // iterations of the supply loop in the last step (self.nrit)
  // end-of-dynamic() outputs (outputs.cuh): work arrays, allocated with the first request
  EndArgs end = {};
  float *d_ct1 = nullptr;
  // wflow_hbv (hbv.cuh): a model made by wf_hbv_create runs the HBV column and the river task of the sweep on the whole network
  bool hbv = false;
  HbvOptions hbv_opt = {};
  int hbv_blocks = 0;                 // grid of k_channel_wave (0: not derived yet; setup_sweep resets it)
  bool satwd_valid = false;           // the SatWaterDepth map holds the end of the previous step          // dynamic shared memory limit already raised for the pipelined kernel
};

namespace {

int64_t field_len(const wf_model *m, int id) { return (g_fields[id].kind & K_RIVER) ? m->nr : m->n; }

int ensure_field(wf_model *m, int id) {
  if (m->dm.f[id]) return WF_OK;
  const int64_t len = std::max<int64_t>(field_len(m, id), 1);
  float *p = nullptr;
  CUDA_OK(cudaMalloc(&p, sizeof(float) * len));
  CUDA_OK(cudaMemsetAsync(p, 0, sizeof(float) * len, m->stream));
  m->dm.f[id] = p;
  return WF_OK;
}

// ---- lateral sweep: plan and launch ---------------------------------------------------------------
// The sweep of a group is latency-bound (one barrier + one Newton solve per level), so the plan
// maximises the number of groups swept concurrently while keeping a tick to one or two passes:
//   * many basins          -> one group per thread block (__syncthreads per tick)
//   * a few wide basins    -> one group per thread-block cluster of 2..8 blocks (cluster barrier)
//   * one basin wider than a cluster can hold -> cooperative launch over the whole GPU (grid barrier)
// Cycle figures are measured on B200; WF_LATERAL_MODE / WF_CLUSTER_SIZE /
// WF_GROUPS override the choice (development aid).
struct LateralPlan {
  int mode = WF_GRID, cs = 1, groups = 1;
  double cost = 0;
};

template <int ML>
int resident_clusters(int mode, int cs, size_t smem) {
  if (smem > 48 * 1024) {  // the occupancy queries below refuse more than the default limit otherwise
    cudaFuncSetAttribute(k_lateral_wave<ML, WF_CTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_lateral_wave<ML, WF_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_lateral_wave<ML, WF_CLUSTER_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (mode == WF_CTA) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lateral_wave<ML, WF_CTA>, WF_GROUP_THREADS, smem);
    return sms * per_sm;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cs * 64);
  cfg.blockDim = dim3((unsigned)wf_block_threads(mode));
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nc = 0;
  const cudaError_t e = (mode == WF_CLUSTER_WIDE) ? cudaOccupancyMaxActiveClusters(&nc, k_lateral_wave<ML, WF_CLUSTER_WIDE>, &cfg)
                                                  : cudaOccupancyMaxActiveClusters(&nc, k_lateral_wave<ML, WF_CLUSTER>, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return nc;
}

template <int ML>
LateralPlan plan_lateral(int64_t n, int64_t D, int64_t nbasins, int sms) {
  // Time of one tick, calibrated on B200 (bench workload): a fixed critical
  // path (barrier + one gather + one solve) plus a share that grows with the items an SM holds,
  //   block / cluster sweep:  5.0 us + 9.0 ns per item and SM
  //   grid sweep:             6.0 us + 13.5 ns per item and SM (256-thread blocks, grid barrier)
  const double items = 2.3 * (double)n / (double)std::max<int64_t>(D, 1);  // per tick, whole network
  LateralPlan best;
  best.mode = WF_GRID; best.cs = 1; best.groups = 1;
  best.cost = (double)D * (6.0e3 + 13.5 * items / sms);
  const int sizes[] = {1, 2, 4, 8};
  for (int cs : sizes) {
    const int mode = (cs == 1) ? WF_CTA : WF_CLUSTER;
    const int res = resident_clusters<ML>(mode, cs, StageLayout<ML>::bytes(wf_block_threads(mode)));
    if (res < 1) continue;
    const int64_t G = std::min<int64_t>(nbasins, res);
    if (G < 1) continue;
    const double per_sm = items / (double)G / cs;
    const double cost = (double)D * (5.0e3 + 9.0 * per_sm);
    if (cost < best.cost) { best.mode = mode; best.cs = cs; best.groups = (int)G; best.cost = cost; }
  }
  return best;
}

template <int ML>
int launch_lateral(wf_model *m) {
  if (const char *iv = std::getenv("WF_LAT_INTERLEAVE")) m->dm.lat_interleave = std::atoi(iv);  // development aid
  if (const char *tm = std::getenv("WF_LATERAL_TASKS"))  // development aid: bit0 subsurface, bit1 overland, bit2 river
    if (!m->hbv) m->dm.task_mask = (m->dm.task_mask & 8) | (std::atoi(tm) & 7);
  if (const char *iv = std::getenv("WF_LAT_PREFETCH"))  // development aid: 0 = no L2 warm-up of the coming ticks' operands
    m->dm.task_mask = std::atoi(iv) ? (m->dm.task_mask & ~8) : (m->dm.task_mask | 8);
  void *args[] = {(void *)&m->dm};
  if (m->lat_mode == WF_GRID) {
    if (m->lat_blocks == 0) {
      int dev = 0, sms = 0, per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lateral_wave<ML, WF_GRID>, 256, 0);
      if (per_sm < 1) return fail(WF_ENODEVICE, "k_lateral_wave does not fit on an SM");
      // no more blocks than the busiest tick can use (subsurface + overland + river items)
      m->lat_blocks = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * per_sm, (m->max_tick_items + 255) / 256));
    }
    CUDA_OK(cudaMemsetAsync(m->dm.grid_counter, 0, sizeof(unsigned int), m->stream));
    void *gfn = m->dm.any_diag ? (void *)k_lateral_wave<ML, WF_GRID, true> : (void *)k_lateral_wave<ML, WF_GRID, false>;
    CUDA_OK(cudaLaunchCooperativeKernel(gfn, dim3(m->lat_blocks), dim3(256), args, m->lat_smem, m->stream));
    return WF_OK;
  }
  const bool cta = m->lat_mode == WF_CTA;
  // dynamic shared memory: the group's level offsets (when they fit) + the operand staging columns of the block's threads
  const size_t smem = ((m->lat_smem + 15) & ~(size_t)15) + StageLayout<ML>::bytes(wf_block_threads(m->lat_mode));
  // two builds of every sweep: with the diagnostic outputs (some K_DIAG field requested) and without
  void *fnD = cta ? (void *)k_lateral_wave<ML, WF_CTA, true>
                  : (m->lat_mode == WF_CLUSTER_WIDE ? (void *)k_lateral_wave<ML, WF_CLUSTER_WIDE, true> : (void *)k_lateral_wave<ML, WF_CLUSTER, true>);
  void *fnP = cta ? (void *)k_lateral_wave<ML, WF_CTA, false>
                  : (m->lat_mode == WF_CLUSTER_WIDE ? (void *)k_lateral_wave<ML, WF_CLUSTER_WIDE, false> : (void *)k_lateral_wave<ML, WF_CLUSTER, false>);
  void *fn = m->dm.any_diag ? fnD : fnP;
  if (m->lat_blocks == 0) {
    if (smem > 48 * 1024) {
      CUDA_OK(cudaFuncSetAttribute(fnD, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_OK(cudaFuncSetAttribute(fnP, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int res = resident_clusters<ML>(m->lat_mode, m->cluster_size, smem);
    if (res < 1) return fail(WF_ENODEVICE, "k_lateral_wave does not fit on the device");
    m->lat_blocks = std::min(res, m->dm.ngroups) * m->cluster_size;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)m->lat_blocks);
  cfg.blockDim = dim3((unsigned)wf_block_threads(m->lat_mode));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = m->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)m->cluster_size;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = cta ? 0 : 1;
  CUDA_OK(cudaLaunchKernelExC(&cfg, fn, args));
  return WF_OK;
}

// What every step needs before its kernels: derived statics up to date, the diagnostics flag, the
// source array of every input slot of the column (pointers may have been re-bound since the last step).
template <int ML>
void prepare_step(wf_model *m) {
  const int64_t n = m->n;
  if (m->derived_dirty) {  // statics changed since the derived ones were computed
    k_derive<ML><<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->dm);
    k_derive_lateral<ML><<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->dm);
    m->launches += 2;
    m->derived_dirty = false;
  }
  m->dm.any_diag = 0;
  for (int id = 0; id < F_COUNT; id++)
    if ((g_fields[id].kind & 0xff) == K_DIAG && m->dm.f[id]) m->dm.any_diag = 1;
  if (m->dm.h_wb || m->paddy.n > 0) m->dm.any_diag = 1;  // (the ponds live in the builds with the outputs)
  {  // sources of the staged inputs (VLayout); pointers may have been re-bound since the last step
    using VL = VLayout<ML>;
    static_assert(VL::nslots(true) <= WF_MAX_VSLOTS, "WF_MAX_VSLOTS too small");
    DevModel &dm = m->dm;
    const int nsl = VL::nslots(dm.glacier != 0);
    int cnt = 0;
    for (int s = 0; s < WF_MAX_VSLOTS; s++) {
      const float *src = nullptr;
      if (s < VS_NFIX) src = (s == VS_river_slot) ? reinterpret_cast<const float *>(dm.river_slot) : dm.f[kVSlotField[s]];
      else if (s < VL::S_C) src = dm.f[F_UStoreLayerDepth_0 + (s - VL::S_U)];
      else if (s < VL::S_KVF) src = dm.f[F_c_0 + (s - VL::S_C)];
      else if (s < VL::S_GL) src = dm.f[F_KsatVerFrac_0 + (s - VL::S_KVF)];
      else if (s < nsl) {
        const int g = s - VL::S_GL;
        src = dm.f[g == 0 ? F_GlacierFrac : g == 1 ? F_GlacierStore : g == 2 ? F_G_SIfrac : g == 3 ? F_G_TT : F_G_Cfmax];
      }
      if (s >= nsl) src = nullptr;
      dm.vsrc[s] = src;
      cnt += src != nullptr;
    }
    dm.vsrc_count = cnt;
  }
}

// ---- reservoirs / lakes (objects.cuh): before the column, on the states of the step before (wflow_sbm.py:2822-2883) ----
int launch_objects(wf_model *m) {
  m->dm.obj_inflow = nullptr;
  const int nres = m->reservoirs.dev.n, nlake = m->lakes.dev.n;
  if (nres + nlake == 0) return WF_OK;
  if (!m->d_obj_inflow) CUDA_OK(cudaMalloc(&m->d_obj_inflow, sizeof(float) * (size_t)std::max<int32_t>(m->nr, 1)));
  int first = 1;
  if (nres > 0) {
    for (int id : {F_ResSimpleArea, F_ResMaxVolume, F_ResTargetFullFrac, F_ResMaxRelease, F_ResDemand, F_ResTargetMinFrac})
      if (!m->dm.f[id]) return fail(WF_EINVAL, std::string("reservoir parameter map not set: ") + g_fields[id].name);
    k_object_forcing<<<nres, 128, 0, m->stream>>>(m->dm, m->reservoirs.dev);
    k_reservoirs<<<(nres + 127) / 128, 128, 0, m->stream>>>(m->dm, m->reservoirs.dev);
    k_object_inflow<<<1, 256, 0, m->stream>>>(m->dm, m->reservoirs.dev, m->d_obj_inflow, first);
    first = 0;
    m->launches += 3;
  }
  if (nlake > 0) {
    for (int id : {F_LakeArea, F_LakeThreshold, F_LakeStorFunc, F_LakeOutflowFunc, F_Lake_b, F_Lake_e})
      if (!m->dm.f[id]) return fail(WF_EINVAL, std::string("lake parameter map not set: ") + g_fields[id].name);
    k_object_forcing<<<nlake, 128, 0, m->stream>>>(m->dm, m->lakes.dev);
    k_lakes<<<1, 256, sizeof(float) * 3 * (size_t)nlake, m->stream>>>(m->dm, m->lakes.dev, m->jdoy);
    k_object_inflow<<<1, 256, 0, m->stream>>>(m->dm, m->lakes.dev, m->d_obj_inflow, first);
    m->launches += 3;
  }
  m->dm.obj_inflow = m->d_obj_inflow;
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

// Objects of a location map: one per cell with an id > 0, in raster order.  areas: the map whose class AT the object's cell
// is the object's area (pcr.areaaverage(.., areas) read at the object, wflow_funcs.py:710-713, 882-891); ldd_org: the LDD
// before initial() cut it at the objects (TopoLddOrg), for the cell that receives the outflow.
int build_objects(wf_model *m, const int32_t *locs, const int32_t *areas, const uint8_t *ldd_org, wf_model::Objects &ob,
                  std::vector<int32_t> &ids_out, std::vector<int32_t> &pos_out) {
  static const int dr[10] = {0, 1, 1, 1, 0, 0, 0, -1, -1, -1}, dc[10] = {0, -1, 0, 1, -1, 0, 1, -1, 0, 1};
  const int nrow = m->cfg.nrow, ncol = m->cfg.ncol;
  const int64_t N = (int64_t)nrow * ncol;
  std::vector<int32_t> pos, down, seg(1, 0), perm, areaid;
  for (int64_t fl = 0; fl < N; fl++) {
    if (locs[fl] <= 0) continue;
    const int32_t p = m->pos_of_flat[(size_t)fl];
    if (p < 0) return fail(WF_EINVAL, "a reservoir / lake location lies outside the drainage network");
    int32_t drs = -1;
    const int c = ldd_org ? ldd_org[fl] : 5;
    if (c >= 1 && c <= 9 && c != 5) {
      const int64_t r2 = fl / ncol + dr[c], c2 = fl % ncol + dc[c];
      if (r2 >= 0 && r2 < nrow && c2 >= 0 && c2 < ncol) {
        const int32_t pd = m->pos_of_flat[(size_t)(r2 * ncol + c2)];
        if (pd >= 0) drs = m->river_slot[(size_t)pd];
      }
    }
    pos.push_back(p);
    down.push_back(drs);
    ids_out.push_back(locs[fl]);
    areaid.push_back(areas ? areas[fl] : 0);
  }
  pos_out = pos;
  const int n = (int)pos.size();
  for (int k = 0; k < n; k++) {
    if (areaid[(size_t)k] > 0)
      for (int64_t q = 0; q < m->n; q++)
        if (areas[m->snet.order[(size_t)q]] == areaid[(size_t)k]) perm.push_back((int32_t)q);
    seg.push_back((int32_t)perm.size());
  }
  auto up = [&](const void *src, size_t bytes, const void **dst) -> int {
    void *d = nullptr;
    CUDA_OK(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
    if (bytes) CUDA_OK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    ob.owned.push_back(d);
    *dst = d;
    return WF_OK;
  };
  int rc;
  ob.dev.n = n;
  if ((rc = up(pos.data(), sizeof(int32_t) * n, (const void **)&ob.dev.pos))) return rc;
  if ((rc = up(down.data(), sizeof(int32_t) * n, (const void **)&ob.dev.down_rs))) return rc;
  if ((rc = up(seg.data(), sizeof(int32_t) * seg.size(), (const void **)&ob.dev.area_seg))) return rc;
  if ((rc = up(perm.data(), sizeof(int32_t) * perm.size(), (const void **)&ob.dev.area_perm))) return rc;
  std::vector<float> z((size_t)std::max(n, 1), 0.0f);
  if ((rc = up(z.data(), sizeof(float) * n, (const void **)&ob.dev.prec_av))) return rc;
  if ((rc = up(z.data(), sizeof(float) * n, (const void **)&ob.dev.pet_av))) return rc;
  if ((rc = up(z.data(), sizeof(float) * n, (const void **)&ob.dev.outflow))) return rc;
  return WF_OK;
}

// ---- end-of-dynamic() outputs (outputs.cuh) ----------------------------------------------------------
bool end_outputs_requested(const wf_model *m) {
  for (int id = F_KinWaveVolumeR; id <= F_CumSurfaceWater; id++) {
    if (id >= F_vwc_perc_0 && id <= F_vwc_perc_7) continue;  // written by the sweep's column tail
    if (id == F_SumEvapSat || id == F_SumEvapUstore) continue;  // written by the column kernel
    if (m->dm.f[id]) return true;
  }
  return false;
}
int end_prepare(wf_model *m) {
  const size_t n = (size_t)std::max<int64_t>(m->n, 1), nr = (size_t)std::max<int32_t>(m->nr, 1);
  if (!m->end.preOrg) {
    CUDA_OK(cudaMalloc(&m->end.preWLRsub, sizeof(float) * nr));
    CUDA_OK(cudaMalloc(&m->end.preWLR, sizeof(float) * nr));
    CUDA_OK(cudaMalloc(&m->end.preWLLsub, sizeof(float) * n));
    CUDA_OK(cudaMalloc(&m->end.preOrg, sizeof(float) * n));
    CUDA_OK(cudaMalloc(&m->end.ctR, sizeof(float) * n));
  }
  if (!m->d_ct1) {
    // catchmenttotal(1): the cells of the network upstream of each cell, itself included (static).  Counted on the
    // network set_dd built: a cell that set_dd's flat-index-0 quirk drops (wflow_funcs.py:88) is not on the device at all.
    std::vector<double> grid((size_t)m->cfg.nrow * m->cfg.ncol);
    std::vector<double> acc((size_t)std::max<int64_t>(m->net.ncells, 1));
    std::vector<float> ct(n, 1.0f);
    for (int64_t k = 0; k < m->net.ncells; k++) {
      double v = 1.0;
      for (int j = 0; j < m->net.nup[(size_t)k]; j++) v += acc[(size_t)m->net.up_start[(size_t)k] + j];
      acc[(size_t)k] = v;
      grid[(size_t)m->net.order[(size_t)k]] = v;
    }
    for (int64_t k = 0; k < m->n; k++) ct[(size_t)k] = (float)grid[(size_t)m->snet.order[(size_t)k]];
    CUDA_OK(cudaMalloc(&m->d_ct1, sizeof(float) * n));
    CUDA_OK(cudaMemcpy(m->d_ct1, ct.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
    m->end.ct1 = m->d_ct1;
  }
  return WF_OK;
}

int sweep_mask_totals(wf_model *m, int mask, const uint32_t **out);

template <int ML>
int launch_step(wf_model *m) {
  const int64_t n = m->n;
  if (n == 0) return WF_OK;
  cudaEvent_t e0 = m->ev[0], e1 = m->ev[1], e2 = m->ev[2];
  if (m->tev_used + 3 <= m->tev.size()) {
    e0 = m->tev[m->tev_used]; e1 = m->tev[m->tev_used + 1]; e2 = m->tev[m->tev_used + 2];
    m->tev_used += 3;
  }
  if (m->timed) cudaEventRecord(e0, m->stream);
  const int tb = WF_VERT_THREADS;
  prepare_step<ML>(m);
  const bool end_out = end_outputs_requested(m);
  if (end_out) {
    if (int rc = end_prepare(m)) return rc;
    m->end.satwd_valid = m->satwd_valid ? 1 : 0;
    const int64_t span = std::max<int64_t>(n, m->nr);
    k_end_pre<ML><<<(unsigned)((span + 255) / 256), 256, 0, m->stream>>>(m->dm, m->end);
    m->launches++;
  }
  if (int rc = launch_objects(m)) return rc;
  const size_t vsmem = WF_VERT_STAGE ? VLayout<ML>::bytes(m->dm.glacier != 0) : 0;
  // four builds of the column: with / without the diagnostic outputs, with / without the LAI-driven canopy parameters
  void (*kv)(const DevModel) = nullptr;
  const bool lai = m->dm.f[F_LAI] != nullptr;
  if (m->dm.any_diag) kv = lai ? k_vertical<ML, true, true> : k_vertical<ML, true, false>;
  else kv = lai ? k_vertical<ML, false, true> : k_vertical<ML, false, false>;
  if (vsmem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute((void *)kv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));
  if (VScratch<ML>::enabled)  // WF_VERT_MINBLOCKS blocks with their scratch must fit the shared-memory carve-out
    CUDA_OK(cudaFuncSetAttribute((void *)kv, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  if (m->abstractions && m->dm.f[F_Inflow]) CUDA_OK(cudaMemsetAsync(m->sup.any_neg, 0, sizeof(int32_t), m->stream));
  kv<<<(unsigned)((n + tb - 1) / tb), tb, vsmem, m->stream>>>(m->dm);
  m->launches++;
  if (m->mass_wasting && m->dm.modelSnow) {  // between the snow pack and everything that reads Snow afterwards (:2700-2707)
    if (!m->d_mw_flux) CUDA_OK(cudaMalloc(&m->d_mw_flux, sizeof(float) * (size_t)std::max<int64_t>(n, 1)));
    k_mass_wasting<<<(unsigned)std::max(1, std::min(m->dm.ngroups, 1024)), 1024, 0, m->stream>>>(m->dm, m->d_mw_flux);
    m->launches++;
  }
  if (m->irri.n > 0) {  // the areas' demand of this step -> Inflow of their intakes (wflow_sbm.py:3014-3031)
    k_irrigation_demand<<<(unsigned)m->irri.n, 256, 0, m->stream>>>(m->dm, m->irri, m->dm.obj_inflow ? m->d_obj_inflow : nullptr);
    m->launches++;
  }
  if (m->timed) cudaEventRecord(e1, m->stream);
  if (m->paddy.n > 0 && m->dm.task_mask == 7) {
    // Paddy areas: what a field asks of its intakes depends on its pond AFTER the column tail of this step (ponding,
    // wflow_sbm.py:762-772 inside sbm_cell), and the intakes' demand is met by the overland / river tasks of the same
    // step.  So the sweep runs twice over the same schedule: subsurface flow + column tail, the areas' demand, then the
    // overland and river waves (every task reads its own cell's hand-offs and the level above: same results as one sweep).
    const uint32_t *all = m->dm.tick_total, *tt = nullptr;
    if (int rc = sweep_mask_totals(m, 1, &tt)) return rc;
    m->dm.task_mask = 1; m->dm.tick_total = tt;
    if (int rc = launch_lateral<ML>(m)) return rc;
    k_paddy_demand<<<(unsigned)m->paddy.n, 256, 0, m->stream>>>(m->dm, m->paddy, m->dm.obj_inflow ? m->d_obj_inflow : nullptr);
    if (int rc = sweep_mask_totals(m, 6, &tt)) return rc;
    m->dm.task_mask = 6; m->dm.tick_total = tt;
    const int rc2 = launch_lateral<ML>(m);
    m->dm.task_mask = 7; m->dm.tick_total = all;
    if (rc2) return rc2;
    m->launches += 2;
  } else if (int rc = launch_lateral<ML>(m)) {
    return rc;
  }
  m->launches++;
  if (m->timed) cudaEventRecord(e2, m->stream);
  m->ev_last[0] = e0; m->ev_last[1] = e1; m->ev_last[2] = e2;
  m->last_nrit = 0;
  if (m->abstractions && m->dm.f[F_Inflow] && m->nr > 0) {  // wflow_sbm.py:3203-3308
    int32_t any = 0;
    CUDA_OK(cudaMemcpyAsync(&any, m->sup.any_neg, sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    if (any) {
      for (int k = 1;; k++) {
        m->last_nrit = k;
        CUDA_OK(cudaMemsetAsync(m->sup.delta, 0, sizeof(float), m->stream));
        k_supply_estimate<<<(unsigned)((m->nr + 127) / 128), 128, 0, m->stream>>>(m->dm, m->sup);
        m->launches++;
        if (k >= 2) {  // (the loop's first kin_wave call repeats the sweep's result)
          k_river_again<ML><<<(unsigned)std::max(1, std::min(m->dm.ngroups, 1024)), 512, 0, m->stream>>>(m->dm);
          m->launches++;
        }
        float delta = 0.0f;
        CUDA_OK(cudaMemcpyAsync(&delta, m->sup.delta, sizeof(float), cudaMemcpyDeviceToHost, m->stream));
        CUDA_OK(cudaStreamSynchronize(m->stream));
        if ((double)delta < m->breakoff || k >= m->maxitsupply) break;
      }
    }
  }
  if (m->irri.n > 0) {  // what the intakes were given -> extra water of the areas in the next step (:3314-3328)
    k_irrigation_supply<<<(unsigned)m->irri.n, 256, 0, m->stream>>>(m->dm, m->irri, m->sup.sws);
    m->launches++;
  }
  if (m->paddy.n > 0) {  // :3330-3346 (cover(.., 0): no supply outside the cells that asked)
    CUDA_OK(cudaMemsetAsync(m->dm.f[F_IRSupplymm], 0, sizeof(float) * (size_t)n, m->stream));
    k_paddy_supply<<<(unsigned)m->paddy.n, 256, 0, m->stream>>>(m->dm, m->paddy, m->sup.sws);
    m->launches++;
  }
  if (end_out) {
    if (m->dm.f[F_RunoffCoeff]) {  // (QCatchmentMM alone needs the static upstream cell count only)
      k_catchment_total<<<(unsigned)std::max(1, std::min(m->dm.ngroups, 1024)), 1024, 0, m->stream>>>(m->dm, m->dm.f[F_RainFallPlusMelt], m->end.ctR);
      m->launches++;
    }
    k_end_outputs<ML><<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->dm, m->end);
    m->launches++;
  }
  m->satwd_valid = m->dm.f[F_SatWaterDepth] != nullptr;
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

// One timestep of a wflow_hbv model: the HBV column (hbv.cuh), then kin_wave on the whole network = the river task of the sweep.
int launch_step_hbv(wf_model *m) {
  const int64_t n = m->n;
  if (n == 0) return WF_OK;
  cudaEvent_t e0 = m->ev[0], e1 = m->ev[1], e2 = m->ev[2];
  if (m->tev_used + 3 <= m->tev.size()) {
    e0 = m->tev[m->tev_used]; e1 = m->tev[m->tev_used + 1]; e2 = m->tev[m->tev_used + 2];
    m->tev_used += 3;
  }
  if (m->timed) cudaEventRecord(e0, m->stream);
  m->dm.any_diag = 0;  // (the sweep's river task has no diagnostic build)
  k_hbv_column<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->dm, m->hbv_opt);
  m->launches++;
  if (m->hbv_opt.mass_wasting) {  // between the snow routine and the channel wave (wflow_hbv.py:1351-1364)
    if (!m->d_mw_flux) CUDA_OK(cudaMalloc(&m->d_mw_flux, sizeof(float) * 2 * (size_t)std::max<int64_t>(n, 1)));
    k_hbv_mass_wasting<<<(unsigned)std::max(1, std::min(m->dm.ngroups, 1024)), 1024, 0, m->stream>>>(m->dm, m->d_mw_flux, m->d_mw_flux + n);
    m->launches++;
  }
  if (m->timed) cudaEventRecord(e1, m->stream);
  // the channel wave: a sweep of its own where the plan is one cluster per group (hbv.cuh, k_channel_wave: 1024 threads per
  // block instead of the shared sweep's 640), the river task of the shared sweep otherwise.  WF_HBV_SHARED_SWEEP=1 forces
  // the shared sweep (development aid; same results).
  const char *shared_sweep = std::getenv("WF_HBV_SHARED_SWEEP");
  if (wf_is_cluster(m->lat_mode) && !(shared_sweep && std::atoi(shared_sweep))) {
    const size_t smem = (m->lat_smem + 15) & ~(size_t)15;
    if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute((void *)k_channel_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)m->cluster_size;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(WF_CHANNEL_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = m->stream;
    if (m->hbv_blocks == 0) {
      cfg.gridDim = dim3((unsigned)m->cluster_size * 64);
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, (void *)k_channel_wave, &cfg) != cudaSuccess || nc < 1) {
        cudaGetLastError();
        return fail(WF_ENODEVICE, "k_channel_wave does not fit on the device");
      }
      m->hbv_blocks = std::min(nc, m->dm.ngroups) * m->cluster_size;
    }
    cfg.gridDim = dim3((unsigned)m->hbv_blocks);
    void *args[] = {(void *)&m->dm};
    CUDA_OK(cudaLaunchKernelExC(&cfg, (void *)k_channel_wave, args));
  } else if (int rc = launch_lateral<1>(m)) {
    return rc;
  }
  m->launches++;
  if (m->timed) cudaEventRecord(e2, m->stream);
  m->ev_last[0] = e0; m->ev_last[1] = e1; m->ev_last[2] = e2;
  if (m->dm.f[F_sumlevel] || m->dm.f[F_SurfaceRunoffMM] || (m->hbv_opt.mass_wasting && (m->dm.f[F_storage] || m->dm.f[F_watbal]))) {
    k_hbv_post<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->dm, m->hbv_opt);
    m->launches++;
  }
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

int dispatch_step(wf_model *m) {
  if (m->hbv) return launch_step_hbv(m);
  switch (m->ML) {
#ifndef WF_DEV_ML4
    case 1: return launch_step<1>(m);
    case 2: return launch_step<2>(m);
    case 3: return launch_step<3>(m);
#endif
    case 4: return launch_step<4>(m);
#ifndef WF_DEV_ML4
    case 5: return launch_step<5>(m);
    case 6: return launch_step<6>(m);
    case 7: return launch_step<7>(m);
    case 8: return launch_step<8>(m);
#endif
  }
  return fail(WF_EINVAL, "unsupported number of soil layers");
}

// Everything of the sweep that depends on the sub-step counts: ticks per group, items per tick, the
// busiest tick, shared-memory need.  Called at creation and by wf_set_substeps.
// Items of every tick of every group for the task subset `mask` (mirrors sweep_group's item space, lateral.cuh).
void sweep_tick_totals(const wf_model *m, int mask, std::vector<GroupDesc> *gd_out, std::vector<uint32_t> &tick_total, int64_t *most_out) {
  const DevModel &dm = m->dm;
  const int itR = dm.itR;
  const int G = m->lay.ngroups;
  int64_t most = 0;
  const int Dn = (int)m->net.nlevels(), Dr = (int)m->rnet.nlevels(), shift = Dn - Dr;
  tick_total.clear();
  if (gd_out) gd_out->assign((size_t)G, GroupDesc());
  for (int g = 0; g < G; g++) {
    const uint32_t *lo = m->lay.lev_off.data() + m->lay.lev_idx[(size_t)g];
    const uint32_t *ro = m->rlay.lev_off.data() + m->rlay.lev_idx[(size_t)g];
    const int l0 = m->lay.lev0[(size_t)g], r0 = m->rlay.lev0[(size_t)g];
    const int nl = Dn - l0, nrl = Dr - r0;
    const bool rivers = r0 < Dr && (mask & 4);
    const int nticks = wf_sweep_ticks(nl, r0 < Dr, itR);
    if (gd_out) {
      GroupDesc &d = (*gd_out)[(size_t)g];
      d.lev_idx = (uint32_t)m->lay.lev_idx[(size_t)g];
      d.rlev_idx = (uint32_t)m->rlay.lev_idx[(size_t)g];
      d.lev0 = l0;
      d.rlev0 = r0;
      d.tick_idx = (uint32_t)tick_total.size();
      d.nticks = nticks;
    }
    for (int t = 0; t < nticks; t++) {
      uint32_t items = 0;
      const int t1 = t - WF_OVL_LAG;
      if (t < nl && (mask & 1)) items += ((lo[t + 1] - lo[t]) + 31u) & ~31u;
      if (t1 >= 0 && t1 < nl && (mask & 2)) items += ((lo[t1 + 1] - lo[t1]) + 31u) & ~31u;
      if (rivers)
        for (int v = 0; v < itR; v++) {
          const int q = wf_sweep_q0(l0, shift, r0) + t - v;
          if (q >= 0 && q < nrl) items += ro[q + 1] - ro[q];
        }
      tick_total.push_back(items);
    }
    most = std::max(most, (m->lay.lev_idx[(size_t)g + 1] - m->lay.lev_idx[(size_t)g]) +
                              (m->rlay.lev_idx[(size_t)g + 1] - m->rlay.lev_idx[(size_t)g]) + (int64_t)nticks);
  }
  if (most_out) *most_out = most;
}

// The sweep restricted to a subset of its tasks (bit0 subsurface, bit1 overland, bit2 river): same groups and ticks, other
// tick sizes.  Returns the device array of the subset's tick sizes (cached until setup_sweep runs again).
int sweep_mask_totals(wf_model *m, int mask, const uint32_t **out) {
  auto it = m->tick_total_of_mask.find(mask);
  if (it == m->tick_total_of_mask.end()) {
    std::vector<uint32_t> tt;
    sweep_tick_totals(m, mask, nullptr, tt, nullptr);
    uint32_t *d = nullptr;
    CUDA_OK(cudaMalloc(&d, std::max<size_t>(sizeof(uint32_t) * tt.size(), 16)));
    CUDA_OK(cudaMemcpyAsync(d, tt.data(), sizeof(uint32_t) * tt.size(), cudaMemcpyHostToDevice, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    it = m->tick_total_of_mask.emplace(mask, d).first;
  }
  *out = it->second;
  return WF_OK;
}

int setup_sweep(wf_model *m) {
  DevModel &dm = m->dm;
  const int itR = dm.itR;
  const int G = m->lay.ngroups;
  std::vector<GroupDesc> gd;
  std::vector<uint32_t> tick_total;  // items of every tick of every group (lateral.cuh: tick item space)
  int64_t most = 0;
  sweep_tick_totals(m, dm.task_mask, &gd, tick_total, &most);
  for (auto &kv : m->tick_total_of_mask) cudaFree(kv.second);
  m->tick_total_of_mask.clear();
  cudaFree(m->d_tick_total);
  cudaFree(m->d_groups);
  m->d_tick_total = nullptr;
  m->d_groups = nullptr;
  CUDA_OK(cudaMalloc(&m->d_tick_total, std::max<size_t>(sizeof(uint32_t) * tick_total.size(), 16)));
  CUDA_OK(cudaMemcpy(m->d_tick_total, tick_total.data(), sizeof(uint32_t) * tick_total.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&m->d_groups, std::max<size_t>(sizeof(GroupDesc) * (size_t)G, 16)));
  CUDA_OK(cudaMemcpy(m->d_groups, gd.data(), sizeof(GroupDesc) * (size_t)G, cudaMemcpyHostToDevice));
  dm.tick_total = m->d_tick_total;
  dm.groups = m->d_groups;
  dm.ngroups = G;
  // a group's level offsets and tick sizes live in shared memory during its sweep when they fit
  dm.lev_cache = 0;
  m->lat_smem = 0;
  // (beside the operand staging columns of the block's threads, lateral.cuh: 227 KB of shared memory per block in all)
  const size_t lev_room = (m->lat_mode == WF_GRID) ? (size_t)40 * 1024
                                                   : std::min<size_t>((size_t)96 * 1024, (size_t)226 * 1024 - wf_stage_bytes(m->ML, wf_block_threads(m->lat_mode)));
  if ((size_t)most * 4 <= lev_room) {
    dm.lev_cache = (int32_t)most;
    m->lat_smem = (size_t)most * 4;
  }
  m->lat_blocks = 0;  // re-derived at the next launch
  m->hbv_blocks = 0;
  m->pipe_blocks = 0;
  m->ah_valid = false;
  // busiest tick of the whole network (grid sweep: no more blocks than it can use)
  const int D = dm.nlev;
  int64_t best = 1;
  for (int t = 0; t < D + 3 + itR; t++) {
    int64_t items = 0;
    const int t1 = t - WF_OVL_LAG;
    if (t < D) items += m->net.lev_off[(size_t)t + 1] - m->net.lev_off[(size_t)t];
    if (t1 >= 0 && t1 < D) items += m->net.lev_off[(size_t)t1 + 1] - m->net.lev_off[(size_t)t1];
    for (int v = 0; v < itR; v++) {
      const int lr = (t - 4 - v) - dm.rlev_shift;
      if (lr >= 0 && lr < dm.nrlev) items += m->rnet.lev_off[(size_t)lr + 1] - m->rnet.lev_off[(size_t)lr];
    }
    best = std::max(best, items);
  }
  m->max_tick_items = best + 64;  // two warp-boundary paddings of the item space
  return WF_OK;
}

// Buffers sized by the sub-step counts: outflows of the earlier sub-steps and the shares handed to river receivers.
int setup_substep_buffers(wf_model *m) {
  DevModel &dm = m->dm;
  const size_t n = (size_t)std::max<int64_t>(m->n, 1), nr = (size_t)std::max<int32_t>(m->nr, 1);
  cudaFree(dm.qo_sub); cudaFree(dm.qr_sub); cudaFree(dm.h_qa); cudaFree(dm.h_qb);
  dm.qo_sub = dm.qr_sub = nullptr;
  dm.h_qa = dm.h_qb = nullptr;
  CUDA_OK(cudaMalloc(&dm.h_qa, sizeof(double) * (size_t)dm.itL * n));
  CUDA_OK(cudaMalloc(&dm.h_qb, sizeof(double) * (size_t)dm.itL * n));
  if (dm.itL > 1) CUDA_OK(cudaMalloc(&dm.qo_sub, sizeof(float) * (size_t)(dm.itL - 1) * n));
  if (dm.itR > 1) CUDA_OK(cudaMalloc(&dm.qr_sub, sizeof(float) * (size_t)(dm.itR - 1) * nr));
  return WF_OK;
}

LateralPlan dispatch_plan(int ML, int64_t n, int64_t D, int64_t nbasins, int sms) {
  switch (ML) {
#ifndef WF_DEV_ML4
    case 1: return plan_lateral<1>(n, D, nbasins, sms);
    case 2: return plan_lateral<2>(n, D, nbasins, sms);
    case 3: return plan_lateral<3>(n, D, nbasins, sms);
#endif
    case 4: return plan_lateral<4>(n, D, nbasins, sms);
#ifndef WF_DEV_ML4
    case 5: return plan_lateral<5>(n, D, nbasins, sms);
    case 6: return plan_lateral<6>(n, D, nbasins, sms);
    case 7: return plan_lateral<7>(n, D, nbasins, sms);
    default: return plan_lateral<8>(n, D, nbasins, sms);
#else
    default: return plan_lateral<4>(n, D, nbasins, sms);  // (development build: four layers only)
#endif
  }
}

const int kRequired[] = {F_P, F_PET, F_Cmax, F_EoverR, F_CanopyGapFraction, F_et_RefToPot, F_WaterFrac,
                         F_SoilThickness, F_thetaS, F_thetaR, F_KsatVer, F_f, F_InfiltCapSoil, F_InfiltCapPath,
                         F_PathFrac, F_CapScale, F_rootdistpar, F_RootingDepth, F_AirEntryPressure, F_MaxLeakage,
                         F_xl, F_yl, F_KsatHorFrac, F_landSlope, F_DL, F_DW, F_SW, F_AlphaL};
const int kRequiredSnow[] = {F_TEMP, F_TempCor, F_TTI, F_TT, F_TTM, F_Cfmax, F_WHC, F_w_soil};
const int kRequiredRiver[] = {F_RiverFrac, F_AlphaR, F_DCL, F_Bw};
const int kRequiredGlacier[] = {F_GlacierFrac, F_G_TT, F_G_Cfmax, F_G_SIfrac};
const int kRequiredLAI[] = {F_Sl, F_Swood, F_Kext};  // wflow_sbm.py:1426-1452: needed because LAI is defined

int check_ready(wf_model *m) {
  auto need = [&](int id) -> int {
    if (!m->dm.f[id]) return fail(WF_EINVAL, std::string("field not set: ") + g_fields[id].name);
    return WF_OK;
  };
  if (m->hbv) {  // parameter maps of wflow_hbv's timestep (wflow_hbv.py:560-900)
    for (int id : {F_P, F_PET, F_Pcorr, F_TempCor, F_TTI, F_TT, F_SFCF, F_RFCF, F_ICF, F_EPF, F_ECORR, F_CEVPF, F_Cfmax, F_CFR,
                   F_WHC, F_FieldCapacity, F_Treshold, F_BetaSeepage, F_PERC, F_Cflux, F_KQuickFlow, F_K4, F_xl, F_yl, F_AlphaR,
                   F_DCL, F_Bw})
      if (int rc = need(id)) return rc;
    if (m->hbv_opt.set_kquick) {
      if (int rc = need(F_K0)) return rc;
      if (int rc = need(F_SUZ)) return rc;
    } else if (int rc = need(F_AlphaNL)) {
      return rc;
    }
    if (m->cfg.glacier)
      for (int id : kRequiredGlacier)
        if (int rc = need(id)) return rc;
    if (m->hbv_opt.mass_wasting)
      if (int rc = need(F_Slope)) return rc;
    return WF_OK;
  }
  for (int id : kRequired)
    if (int rc = need(id)) return rc;
  for (int k = 0; k < m->ML; k++) {
    if (int rc = need(F_c_0 + k)) return rc;
    if (int rc = need(F_KsatVerFrac_0 + k)) return rc;
  }
  if (m->cfg.model_snow)
    for (int id : kRequiredSnow)
      if (int rc = need(id)) return rc;
  if (m->cfg.model_snow && m->cfg.soil_inf_redu)
    if (int rc = need(F_cf_soil)) return rc;
  if (m->nr > 0)
    for (int id : kRequiredRiver)
      if (int rc = need(id)) return rc;
  if (m->cfg.glacier)
    for (int id : kRequiredGlacier)
      if (int rc = need(id)) return rc;
  if (m->dm.f[F_LAI])
    for (int id : kRequiredLAI)
      if (int rc = need(id)) return rc;
  if (m->paddy.n > 0)
    for (int id : {F_h_p, F_h_max, F_h_min})
      if (int rc = need(id)) return rc;
  return WF_OK;
}


// ---- pipelined multi-step sweep (pipeline.cuh) -------------------------------------------------------
__global__ void k_expand_capture(const float *__restrict__ cap, const int32_t *__restrict__ col, const int32_t *__restrict__ pos,
                                 float *__restrict__ out, int64_t ngauges, int32_t cap_n, int64_t total, float mv) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  const int64_t s = k / ngauges, g = k - s * ngauges;
  const int32_t c = col[g];
  // river-only field: 0 at the cells of the network that kin_wave does not visit (wflow_funcs.py:172-182), mv outside it
  out[k] = (c >= 0) ? cap[s * cap_n + c] : ((pos[g] >= 0) ? 0.0f : mv);
}

// Which sweep runs the pipelined launch.  With several groups it is the per-step plan's (the storage layout
// is built for it).  A single group (one basin, or a network swept as a whole) has the same layout under
// every plan, and the pipelined sweep wants throughput, not a short barrier: the cooperative grid sweep
// over all SMs, whatever the per-step plan chose (a Rhine-sized basin is a single-block sweep there).
static int pipeline_mode(const wf_model *m) { return (m->dm.ngroups == 1) ? WF_GRID : m->lat_mode; }

template <int ML>
int launch_pipeline(wf_model *m, DevModel &dmp, const PipeArgs &pa) {
  void *args[] = {(void *)&dmp, (void *)&pa};
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int pmode = pipeline_mode(m);
  if (pmode == WF_GRID) {
    void *fn = (void *)k_pipeline<ML, WF_GRID>;
    // level offsets of the one group in shared memory when they fit beside three blocks per SM
    const size_t entries = (size_t)(m->dm.nlev + 1) + (size_t)(m->dm.nrlev + 1);
    const size_t smem = (entries * 4 <= 40 * 1024) ? entries * 4 : 0;
    dmp.lev_cache = smem ? (int32_t)entries : 0;
    if (m->pipe_blocks == 0) {
      int per_sm = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pipeline<ML, WF_GRID>, wf_pipe_threads(WF_GRID), smem);
      if (per_sm < 1) return fail(WF_ENODEVICE, "k_pipeline does not fit on an SM");
      m->pipe_blocks = sms * per_sm;
    }
    CUDA_OK(cudaMemsetAsync(m->dm.grid_counter, 0, sizeof(unsigned int), m->stream));
    CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3((unsigned)m->pipe_blocks), dim3((unsigned)wf_pipe_threads(WF_GRID)), args, smem,
                                        m->stream));
    return WF_OK;
  }
  const bool cta = pmode == WF_CTA;
  void *fn = cta ? (void *)k_pipeline<ML, WF_CTA> : (void *)k_pipeline<ML, WF_CLUSTER>;
  const int threads = cta ? wf_pipe_threads(WF_CTA) : wf_block_threads(pmode);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = m->lat_smem;
  cfg.stream = m->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)m->cluster_size;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = cta ? 0 : 1;
  if (m->pipe_blocks == 0) {
    if (m->lat_smem > 48 * 1024 && m->lat_smem > m->pipe_attr_smem) {
      CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->lat_smem));
      m->pipe_attr_smem = m->lat_smem;
    }
    int res = 0;
    if (cta) {
      int per_sm = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pipeline<ML, WF_CTA>, threads, m->lat_smem);
      res = sms * per_sm;
    } else {
      cfg.gridDim = dim3((unsigned)m->cluster_size * 64);
      if (cudaOccupancyMaxActiveClusters(&res, k_pipeline<ML, WF_CLUSTER>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        res = 0;
      }
    }
    if (res < 1) return fail(WF_ENODEVICE, "k_pipeline does not fit on the device");
    m->pipe_blocks = std::min(res, m->dm.ngroups) * m->cluster_size;
  }
  cfg.gridDim = dim3((unsigned)m->pipe_blocks);
  CUDA_OK(cudaLaunchKernelExC(&cfg, fn, args));
  return WF_OK;
}

template <int ML>
int run_pipeline(wf_model *m, int nsteps) {
  if (m->n == 0) return WF_OK;
  cudaEvent_t e0 = m->ev[0], e1 = m->ev[1], e2 = m->ev[2];
  if (m->tev_used + 3 <= m->tev.size()) {
    e0 = m->tev[m->tev_used]; e1 = m->tev[m->tev_used + 1]; e2 = m->tev[m->tev_used + 2];
    m->tev_used += 3;
    m->tev_extra_steps += nsteps - 1;
  }
  // the column's forcing slots point at the multi-step buffers for this launch
  float *saved[3] = {m->dm.f[F_P], m->dm.f[F_PET], m->dm.f[F_TEMP]};
  m->dm.f[F_P] = m->pipe_forc[0];
  m->dm.f[F_PET] = m->pipe_forc[1];
  if (m->pipe_has[2]) m->dm.f[F_TEMP] = m->pipe_forc[2];
  int rc = check_ready(m);
  if (rc == WF_OK) {
    if (m->timed) { cudaEventRecord(e0, m->stream); cudaEventRecord(e1, m->stream); }
    prepare_step<ML>(m);
    DevModel dmp = m->dm;
    if (m->pipe_gauges >= 0 && m->cap_n > 0) {
      dmp.cap_slot = m->d_cap_slot;
      dmp.cap_out = m->d_cap_out;
      dmp.cap_n = m->cap_n;
    }
    PipeArgs pa;
    pa.nsteps = nsteps;
    pa.lag = 3 + m->dm.itR;
    pa.fstride = m->dm.npad;
    rc = launch_pipeline<ML>(m, dmp, pa);
    if (rc == WF_OK) {
      m->launches++;
      if (m->timed) cudaEventRecord(e2, m->stream);
      m->ev_last[0] = e0; m->ev_last[1] = e1; m->ev_last[2] = e2;
    }
  }
  m->dm.f[F_P] = saved[0]; m->dm.f[F_PET] = saved[1]; m->dm.f[F_TEMP] = saved[2];
  if (rc != WF_OK) return rc;
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

// ---- multi-step launch of the per-step kernels with the next step's column on the idle SMs (lateral.cuh, column_ahead) ----
// Tiles of the column (as many cells as a sweep block has threads) in the order the sweep frees them: a tile whose last
// cell is on level L of group g is free once the barrier of tick L + 3 + it_kinR of g's sweep has completed
// (progress >= L + 4 + it_kinR; the last levels of a group when the group is swept).
template <int ML>
int setup_ahead(wf_model *m) {
  if (m->ah_valid) return WF_OK;
  m->ah_extra = 0;
  m->ah_valid = true;
  if (!wf_is_cluster(m->lat_mode) || m->n == 0) return WF_OK;
  const int threads = wf_block_threads(m->lat_mode);
  const size_t smem = ((m->lat_smem + 15) & ~(size_t)15) + StageLayout<ML>::bytes(threads);
  const int res = resident_clusters<ML>(m->lat_mode, m->cluster_size, smem);
  const int nsweep = std::min(res, m->dm.ngroups);
  int extra = res - nsweep;
  if (const char *e = std::getenv("WF_AHEAD_CLUSTERS")) extra = std::min(extra, std::max(0, std::atoi(e)));  // development aid
  extra = std::min(extra, 8);  // (a handful of clusters is what a sweep of many groups leaves; more would only poll)
  if (nsweep < m->dm.ngroups || extra < 1) return WF_OK;  // every group needs its own resident cluster beside the workers
  const int G = m->dm.ngroups, itR = m->dm.itR;
  const int Dn = (int)m->net.nlevels(), Dr = (int)m->rnet.nlevels();
  const int64_t n = m->n;
  const int ntiles = (int)((n + threads - 1) / threads);
  // group and group-local level of every storage position: groups are consecutive, levels consecutive inside a group
  std::vector<int32_t> need((size_t)ntiles * 4), order((size_t)ntiles), rank((size_t)ntiles);
  auto locate = [&](int64_t pos, int &g, int &need_ticks) {
    g = 0;
    while (g + 1 < G && (int64_t)m->lay.lev_off[(size_t)m->lay.lev_idx[(size_t)g + 1]] <= pos) g++;
    const uint32_t *lo = m->lay.lev_off.data() + m->lay.lev_idx[(size_t)g];
    const int nl = Dn - m->lay.lev0[(size_t)g];
    int L = (int)(std::upper_bound(lo, lo + nl + 1, (uint32_t)pos) - lo) - 1;
    L = std::max(0, std::min(L, nl - 1));
    const bool rivers = m->rlay.lev0[(size_t)g] < Dr;
    const int nticks = wf_sweep_ticks(nl, rivers, itR);
    need_ticks = L + 4 + (rivers ? itR : 0);
    if (need_ticks >= nticks) need_ticks = 0x7fffffff;  // no barrier follows the last tick: free when the group is swept
  };
  std::vector<int64_t> key((size_t)ntiles);
  for (int t = 0; t < ntiles; t++) {
    const int64_t a = (int64_t)t * threads, b = std::min<int64_t>(n, a + threads) - 1;
    int g1, n1, g2, n2;
    locate(b, g2, n2);
    locate(a, g1, n1);
    if (g1 == g2) n1 = n2;
    else n1 = 0x7fffffff;  // the tile holds the last levels of group g1: free when g1 is swept (tiles across more than two
                           // groups are left to k_vertical_rest, below)
    need[(size_t)4 * t + 0] = g1; need[(size_t)4 * t + 1] = n1;
    need[(size_t)4 * t + 2] = g2; need[(size_t)4 * t + 3] = n2;
    key[(size_t)t] = std::max<int64_t>(n1, n2);
    order[(size_t)t] = t;
  }
  // tiles that span more than two groups are never handed to the workers: they sort beyond the queue
  int usable = ntiles;
  for (int t = 0; t < ntiles; t++) {
    const int64_t a = (int64_t)t * threads, b = std::min<int64_t>(n, a + threads) - 1;
    int ga, gb, dummy;
    locate(a, ga, dummy);
    locate(b, gb, dummy);
    if (gb - ga > 1) key[(size_t)t] = (int64_t)1 << 40;
  }
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[(size_t)x] < key[(size_t)y]; });
  while (usable > 0 && key[(size_t)order[(size_t)usable - 1]] >= ((int64_t)1 << 40)) usable--;
  for (int r = 0; r < ntiles; r++) rank[(size_t)order[(size_t)r]] = r;
  cudaFree(m->d_ah_progress); cudaFree(m->d_ah_queue); cudaFree(m->d_ah_order); cudaFree(m->d_ah_need); cudaFree(m->d_ah_rank);
  m->d_ah_progress = m->d_ah_queue = m->d_ah_order = m->d_ah_need = m->d_ah_rank = nullptr;
  CUDA_OK(cudaMalloc(&m->d_ah_progress, sizeof(int32_t) * (size_t)std::max(G, 1)));
  CUDA_OK(cudaMalloc(&m->d_ah_queue, sizeof(int32_t) * 4));
  CUDA_OK(cudaMalloc(&m->d_ah_order, sizeof(int32_t) * (size_t)ntiles));
  CUDA_OK(cudaMalloc(&m->d_ah_need, sizeof(int32_t) * (size_t)ntiles * 4));
  CUDA_OK(cudaMalloc(&m->d_ah_rank, sizeof(int32_t) * (size_t)ntiles));
  CUDA_OK(cudaMemcpy(m->d_ah_order, order.data(), sizeof(int32_t) * (size_t)ntiles, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(m->d_ah_need, need.data(), sizeof(int32_t) * (size_t)ntiles * 4, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(m->d_ah_rank, rank.data(), sizeof(int32_t) * (size_t)ntiles, cudaMemcpyHostToDevice));
  m->ah_ntiles = usable;
  m->ah_tile = threads;
  m->ah_extra = extra;
  m->ah_nsweep = nsweep;
  return WF_OK;
}

template <int ML>
int run_overlap(wf_model *m, int nsteps) {
  if (m->n == 0) return WF_OK;
  cudaEvent_t e0 = m->ev[0], e1 = m->ev[1], e2 = m->ev[2];
  if (m->tev_used + 3 <= m->tev.size()) {
    e0 = m->tev[m->tev_used]; e1 = m->tev[m->tev_used + 1]; e2 = m->tev[m->tev_used + 2];
    m->tev_used += 3;
    m->tev_extra_steps += nsteps - 1;
  }
  float *saved[3] = {m->dm.f[F_P], m->dm.f[F_PET], m->dm.f[F_TEMP]};
  m->dm.f[F_P] = m->pipe_forc[0];
  m->dm.f[F_PET] = m->pipe_forc[1];
  if (m->pipe_has[2]) m->dm.f[F_TEMP] = m->pipe_forc[2];
  int rc = check_ready(m);
  if (rc == WF_OK) {
    if (m->timed) { cudaEventRecord(e0, m->stream); cudaEventRecord(e1, m->stream); }
    prepare_step<ML>(m);
    DevModel dmp = m->dm;
    if (m->pipe_gauges >= 0 && m->cap_n > 0) {
      dmp.cap_slot = m->d_cap_slot;
      dmp.cap_out = m->d_cap_out;
      dmp.cap_n = m->cap_n;
    }
    const int threads = wf_block_threads(m->lat_mode);
    const size_t smem = ((m->lat_smem + 15) & ~(size_t)15) + StageLayout<ML>::bytes(threads);
    void *fn = (m->lat_mode == WF_CLUSTER_WIDE) ? (void *)k_lateral_wave<ML, WF_CLUSTER_WIDE, false, true> : (void *)k_lateral_wave<ML, WF_CLUSTER, false, true>;
    if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned vblocks = (unsigned)((m->n + WF_VERT_THREADS - 1) / WF_VERT_THREADS);
    for (int s = 0; s < nsteps && rc == WF_OK; s++) {
      const int64_t fo = (int64_t)s * m->dm.npad;
      // the column of step s: what the workers of the sweep before did not take (everything for the first step)
      k_vertical_rest<ML, false><<<vblocks, WF_VERT_THREADS, 0, m->stream>>>(dmp, fo, m->d_ah_rank, s > 0 ? m->d_ah_queue : nullptr, m->ah_tile, m->ah_ntiles);
      CUDA_OK(cudaMemsetAsync(m->d_ah_queue, 0, sizeof(int32_t) * 4, m->stream));
      CUDA_OK(cudaMemsetAsync(m->d_ah_progress, 0, sizeof(int32_t) * (size_t)std::max(m->dm.ngroups, 1), m->stream));
      DevModel dms = dmp;
      dms.cap_step = s;
      dms.any_diag = 0;
      const bool last = s + 1 == nsteps;
      dms.ahead.progress = m->d_ah_progress;
      dms.ahead.queue = m->d_ah_queue;
      dms.ahead.order = m->d_ah_order;
      dms.ahead.need = m->d_ah_need;
      dms.ahead.ntiles = last ? 0 : m->ah_ntiles;  // (the last step's sweep has no next column: its workers leave at once)
      dms.ahead.tile = m->ah_tile;
      dms.ahead.nsweep = m->ah_nsweep;
      dms.ahead.fo = fo + m->dm.npad;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)((m->ah_nsweep + m->ah_extra) * m->cluster_size));
      cfg.blockDim = dim3((unsigned)threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = m->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)m->cluster_size;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      void *args[] = {(void *)&dms};
      CUDA_OK(cudaLaunchKernelExC(&cfg, fn, args));
      m->launches += 2;
    }
    if (rc == WF_OK) {
      if (m->timed) cudaEventRecord(e2, m->stream);
      m->ev_last[0] = e0; m->ev_last[1] = e1; m->ev_last[2] = e2;
    }
  }
  m->dm.f[F_P] = saved[0]; m->dm.f[F_PET] = saved[1]; m->dm.f[F_TEMP] = saved[2];
  if (rc != WF_OK) return rc;
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

// Which multi-step launch: the per-step kernels with the next column on the idle SMs where a cluster sweep of several
// groups leaves SMs idle (there the time-skewed kernel is slower than the per-step path), else k_pipeline.
// WF_PIPE_OVERLAP = 0 / 1 forces the choice where both are possible (development aid).
template <int ML>
int run_multistep(wf_model *m, int nsteps) {
  if (int rc = setup_ahead<ML>(m)) return rc;
  bool any_diag = m->dm.h_wb != nullptr;
  for (int id = 0; id < F_COUNT; id++)
    if ((g_fields[id].kind & 0xff) == K_DIAG && m->dm.f[id]) any_diag = true;
  bool overlap = m->ah_extra > 0 && m->dm.ngroups > 1 && !m->dm.f[F_LAI] && !any_diag;
  if (const char *e = std::getenv("WF_PIPE_OVERLAP")) overlap = overlap && std::atoi(e) != 0;
  m->last_multistep = overlap ? 2 : 1;
  return overlap ? run_overlap<ML>(m, nsteps) : run_pipeline<ML>(m, nsteps);
}

int dispatch_pipeline(wf_model *m, int nsteps) {
  switch (m->ML) {
#ifndef WF_DEV_ML4
    case 1: return run_multistep<1>(m, nsteps);
    case 2: return run_multistep<2>(m, nsteps);
    case 3: return run_multistep<3>(m, nsteps);
#endif
    case 4: return run_multistep<4>(m, nsteps);
#ifndef WF_DEV_ML4
    case 5: return run_multistep<5>(m, nsteps);
    case 6: return run_multistep<6>(m, nsteps);
    case 7: return run_multistep<7>(m, nsteps);
    case 8: return run_multistep<8>(m, nsteps);
#endif
  }
  return fail(WF_EINVAL, "unsupported number of soil layers");
}

}  // namespace

static void catchment_total(const Network &net, const float *values, double *out_grid) {
  const int64_t N = (int64_t)net.nrow * net.ncol;
  for (int64_t i = 0; i < N; i++) out_grid[i] = 0.0;
  std::vector<double> acc((size_t)std::max<int64_t>(net.ncells, 1));
  for (int64_t k = 0; k < net.ncells; k++) {  // level order: upstream cells are complete before k
    double v = values ? (double)values[net.order[(size_t)k]] : 1.0;
    for (int j = 0; j < net.nup[(size_t)k]; j++) v += acc[(size_t)net.up_start[(size_t)k] + j];
    acc[(size_t)k] = v;
    out_grid[net.order[(size_t)k]] = v;
  }
}

extern "C" {
static int forcing_acquire(wf_model *m);
static int forcing_release(wf_model *m);
static int create_impl(wf_model *m, const wf_config *cfg, wf_schedule *sched, const uint8_t *ldd, const uint8_t *river);

const char *wf_last_error(void) { return g_err.c_str(); }
const char *wf_version(void) { return "wflow_b200 0.1 (sm_100a)"; }

int wf_num_fields(void) { return F_COUNT; }
const char *wf_field_name(int i) { return (i >= 0 && i < F_COUNT) ? g_fields[i].name : nullptr; }
int wf_field_kind(int i) { return (i >= 0 && i < F_COUNT) ? g_fields[i].kind : -1; }

int wf_create(const wf_config *cfg, const uint8_t *ldd, const uint8_t *river, wf_model **out) {
  return wf_create_from_schedule(cfg, nullptr, ldd, river, out);
}

int wf_schedule_catchment_total(const wf_schedule *s, const float *values, double *out_grid) {
  if (!s || !out_grid) return fail(WF_EINVAL, "null argument");
  catchment_total(s->net, values, out_grid);
  return WF_OK;
}

static int create_common(const wf_config *cfg, wf_schedule *sched, const uint8_t *ldd, const uint8_t *river, bool hbv, wf_model **out);

int wf_create_from_schedule(const wf_config *cfg, wf_schedule *sched, const uint8_t *ldd, const uint8_t *river,
                            wf_model **out) {
  return create_common(cfg, sched, ldd, river, false, out);
}

int wf_hbv_create(const wf_config *cfg, const uint8_t *ldd, wf_model **out) {
  if (!cfg || !ldd || !out) return fail(WF_EINVAL, "null argument");
  if (cfg->nrow <= 0 || cfg->ncol <= 0) return fail(WF_EINVAL, "bad grid size");
  wf_config c = *cfg;
  c.n_layer_thickness = 0;  // no soil layers in HBV; the overland wave does not exist
  c.it_kin_l = 1;
  // every cell with an ldd code carries a channel (self.static["River"] is never consulted by kin_wave, wflow_funcs.py:155-207)
  std::vector<uint8_t> river((size_t)c.nrow * (size_t)c.ncol, 1);
  return create_common(&c, nullptr, ldd, river.data(), true, out);
}

static int create_common(const wf_config *cfg, wf_schedule *sched, const uint8_t *ldd, const uint8_t *river, bool hbv, wf_model **out) {
  if (!cfg || !ldd || !out) return fail(WF_EINVAL, "null argument");
  if (cfg->nrow <= 0 || cfg->ncol <= 0) return fail(WF_EINVAL, "bad grid size");
  if (cfg->n_layer_thickness < 0 || cfg->n_layer_thickness + 1 > WF_MAX_LAYERS)
    return fail(WF_EINVAL, "too many soil layers");
  if (cfg->it_kin_l < 1 || cfg->it_kin_r < 1) return fail(WF_EINVAL, "it_kin_l / it_kin_r must be >= 1");
  if (cfg->it_kin_l > 1024 || cfg->it_kin_r > 1024) return fail(WF_EUNSUPPORTED, "more than 1024 sub-steps are not supported");
  if (!(cfg->timestepsecs > 0)) return fail(WF_EINVAL, "timestepsecs must be > 0");
  const int64_t N = (int64_t)cfg->nrow * cfg->ncol;
  if (N >= (int64_t)1 << 31) return fail(WF_EUNSUPPORTED, "grids of 2^31 cells or more are not supported");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(WF_ENODEVICE, "no CUDA device available (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(WF_ENODEVICE, "bad device ordinal");
  CUDA_OK(cudaSetDevice(cfg->device));

  wf_model *m = new wf_model();
  m->hbv = hbv;
  const int rc = create_impl(m, cfg, sched, ldd, river);
  if (rc != WF_OK) {  // release whatever the failed construction had allocated; keep its message
    const std::string msg = g_err;
    wf_destroy(m);
    g_err = msg;
    return rc;
  }
  *out = m;
  return WF_OK;
}

static int create_impl(wf_model *m, const wf_config *cfg, wf_schedule *sched, const uint8_t *ldd, const uint8_t *river) {
  const int64_t N = (int64_t)cfg->nrow * cfg->ncol;
  m->cfg = *cfg;
  if (m->cfg.beta == 0.0) m->cfg.beta = (double)0.6f;  // REAL4 0.6, wflow_sbm.py:1643
  // wflow_sbm.py:1757-1766
  m->ML = (cfg->n_layer_thickness > 0) ? cfg->n_layer_thickness + 1 : 1;

  if (sched) {
    if (sched->net.nrow != cfg->nrow || sched->net.ncol != cfg->ncol)
      return fail(WF_EINVAL, "schedule and config disagree on the grid size");
    m->net = std::move(sched->net);
    delete sched;
  } else if (build_network(ldd, cfg->nrow, cfg->ncol, m->net) != 0) {
    return fail(WF_ECYCLE, "could not build the drainage network");
  }
  // river-only ldd: non-river cells -> missing (wflow_sbm.py:2387-2389)
  {
    std::vector<uint8_t> lr((size_t)N, 0);
    if (river)
      for (int64_t i = 0; i < N; i++) lr[(size_t)i] = river[i] ? ldd[i] : 0;
    if (build_network(lr.data(), cfg->nrow, cfg->ncol, m->rnet) != 0)
      return fail(WF_ECYCLE, "could not build the river network");
  }
  m->n = m->net.ncells;
  const int64_t n = m->n;
  // ---- storage layout: groups of whole basins, one sweep each (network.h, plan_lateral) ----
  {
    std::vector<int32_t> basin, gob;
    std::vector<int64_t> bcells;
    const int64_t nbasins = label_basins(m->net, basin, bcells);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    LateralPlan plan = dispatch_plan(m->ML, n, m->net.nlevels(), nbasins, sms);
    if (const char *e = std::getenv("WF_LATERAL_MODE")) {
      const std::string v(e);
      if (v == "grid") { plan.mode = WF_GRID; plan.cs = 1; plan.groups = 1; }
      else if (v == "cta") { plan.mode = WF_CTA; plan.cs = 1; plan.groups = (int)std::min<int64_t>(std::max<int64_t>(nbasins, 1), sms); }
      else if (v == "cluster") {
        plan.mode = WF_CLUSTER;
        if (plan.cs < 2) plan.cs = 4;
        plan.groups = (int)std::min<int64_t>(std::max<int64_t>(nbasins, 1), sms / plan.cs);
      }
    }
    if (const char *e = std::getenv("WF_CLUSTER_SIZE")) {
      const int cs = std::atoi(e);
      if (plan.mode == WF_CLUSTER && (cs == 2 || cs == 4 || cs == 8)) {
        plan.cs = cs;
        plan.groups = (int)std::min<int64_t>(std::max<int64_t>(nbasins, 1), sms / cs);
      }
    }
    if (const char *e = std::getenv("WF_GROUPS"))
      if (plan.mode != WF_GRID && std::atoi(e) >= 1) plan.groups = std::atoi(e);
    if (plan.mode == WF_GRID) plan.groups = 1;
    const int G = pack_basins(bcells, plan.groups, gob);
    std::vector<int32_t> group_of((size_t)std::max<int64_t>(n, 1), 0);
    for (int64_t k = 0; k < n; k++) group_of[(size_t)k] = gob.empty() ? 0 : gob[(size_t)basin[(size_t)k]];
    std::vector<int64_t> newpos;
    regroup(m->net, group_of, G, m->snet, m->lay, newpos);
    // river network: a river cell belongs to the group of its basin
    std::vector<int32_t> grp_of_flat;
    {
      std::vector<int32_t> tmp((size_t)N, 0);
      for (int64_t k = 0; k < n; k++) tmp[(size_t)m->net.order[(size_t)k]] = group_of[(size_t)k];
      grp_of_flat.swap(tmp);
    }
    std::vector<int32_t> rgroup_of((size_t)std::max<int64_t>(m->rnet.ncells, 1), 0);
    for (int64_t k = 0; k < m->rnet.ncells; k++) rgroup_of[(size_t)k] = grp_of_flat[(size_t)m->rnet.order[(size_t)k]];
    std::vector<int64_t> rnewpos;
    regroup(m->rnet, rgroup_of, G, m->srnet, m->rlay, rnewpos);
    m->lat_mode = plan.mode;
    m->cluster_size = plan.cs;
    if (plan.mode == WF_CLUSTER) {
      // passes (items / threads of the cluster, rounded up) summed over the ticks of the slowest group, for the
      // regular and the wide block: the wide block has fewer registers per thread (a pass costs ~6 % more)
      double worst_reg = 0, worst_wide = 0;
      const int Dn = (int)m->net.nlevels(), Dr = (int)m->rnet.nlevels(), shift = Dn - Dr;
      for (int g = 0; g < m->lay.ngroups; g++) {
        const uint32_t *lo = m->lay.lev_off.data() + m->lay.lev_idx[(size_t)g];
        const uint32_t *ro = m->rlay.lev_off.data() + m->rlay.lev_idx[(size_t)g];
        const int l0 = m->lay.lev0[(size_t)g], r0 = m->rlay.lev0[(size_t)g];
        const int nl = Dn - l0, nrl = Dr - r0;
        double preg = 0, pwide = 0;
        for (int t = 0; t < wf_sweep_ticks(nl, r0 < Dr, cfg->it_kin_r); t++) {
          int64_t items = 0;
          const int t1 = t - WF_OVL_LAG;
          if (t < nl) items += ((int64_t)(lo[t + 1] - lo[t]) + 31) & ~31LL;
          if (t1 >= 0 && t1 < nl) items += ((int64_t)(lo[t1 + 1] - lo[t1]) + 31) & ~31LL;
          for (int v = 0; v < cfg->it_kin_r; v++) {
            const int q = wf_sweep_q0(l0, shift, r0) + t - v;
            if (q >= 0 && q < nrl) items += ro[q + 1] - ro[q];
          }
          const int64_t treg = (int64_t)plan.cs * WF_GROUP_THREADS, twide = (int64_t)plan.cs * WF_GROUP_THREADS_WIDE;
          preg += (double)std::max<int64_t>(1, (items + treg - 1) / treg);
          pwide += (double)std::max<int64_t>(1, (items + twide - 1) / twide);
        }
        worst_reg = std::max(worst_reg, preg);
        worst_wide = std::max(worst_wide, pwide);
      }
      bool wide = worst_wide * 1.06 < worst_reg;
      if (const char *e = std::getenv("WF_CLUSTER_WIDE")) wide = std::atoi(e) != 0;
      if (wide) m->lat_mode = WF_CLUSTER_WIDE;
    }
  }
  const Network &snet = m->snet, &srnet = m->srnet;
  // river order: river-network cells in storage order, then river-map cells kin_wave never visits
  m->river_slot.assign((size_t)std::max<int64_t>(n, 1), -1);
  {
    std::vector<int32_t> slot_of_flat((size_t)N, -1);
    std::vector<int32_t> &pos_of_flat = m->pos_of_flat;
    pos_of_flat.assign((size_t)N, -1);
    for (int64_t k = 0; k < n; k++) pos_of_flat[(size_t)snet.order[(size_t)k]] = (int32_t)k;
    int32_t s = 0;
    for (int64_t k = 0; k < srnet.ncells; k++) {
      const int64_t fl = srnet.order[(size_t)k];
      slot_of_flat[(size_t)fl] = s++;  // the river network is a sub-forest of the full one
      m->river_flat.push_back(fl);
    }
    m->nrnet = s;
    if (river)
      for (int64_t k = 0; k < n; k++) {
        const int64_t fl = snet.order[(size_t)k];
        if (river[fl] && slot_of_flat[(size_t)fl] < 0) {
          slot_of_flat[(size_t)fl] = s++;
          m->river_flat.push_back(fl);
        }
      }
    m->nr = s;
    for (int64_t k = 0; k < n; k++) m->river_slot[(size_t)k] = slot_of_flat[(size_t)snet.order[(size_t)k]];
  }

  CUDA_OK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  for (auto &e : m->ev) CUDA_OK(cudaEventCreate(&e));

  DevModel &dm = m->dm;
  std::memset(&dm, 0, sizeof(dm));
  dm.n = n;
  dm.nr = m->nr;
  dm.nrnet = m->nrnet;
  dm.nlev = (int32_t)m->net.nlevels();
  dm.nrlev = (int32_t)m->rnet.nlevels();
  dm.rlev_shift = dm.nlev - dm.nrlev;
  dm.ML = m->ML;
  dm.nrLayers = (cfg->n_layer_thickness > 0) ? cfg->n_layer_thickness : 1;
  for (int k = 0; k < WF_MAXL; k++) dm.layerThickness[k] = (k < cfg->n_layer_thickness) ? cfg->layer_thickness[k] : 0.0;
  dm.modelSnow = cfg->model_snow;
  dm.soilInfRedu = cfg->soil_inf_redu;
  dm.transferMethod = cfg->transfer_method;
  dm.ust = cfg->ust;
  dm.glacier = cfg->glacier;
  dm.gash = cfg->timestepsecs >= (23 * 3600);  // wflow_sbm.py:2639
  dm.itL = cfg->it_kin_l;
  dm.itR = cfg->it_kin_r;
  dm.dt = cfg->timestepsecs;
  dm.basetimestep = 86400.0;  // wflow_sbm.py:1222
  dm.beta = m->cfg.beta;
  dm.task_mask = 7;
  if (const char *tm = std::getenv("WF_LATERAL_TASKS")) dm.task_mask = std::atoi(tm);
  if (m->hbv) dm.task_mask = 4;  // the channel of every cell is all the sweep routes (wflow_hbv.py:1633-1647)
  dm.lat_interleave = 1;
#ifdef WF_PROFILE
  if (std::getenv("WF_TICK_PROFILE")) {
    const size_t nt = (size_t)(m->net.nlevels() + 16 + cfg->it_kin_r) * 16;
    CUDA_OK(cudaMallocManaged(&dm.prof, sizeof(unsigned long long) * nt));
    std::memset(dm.prof, 0, sizeof(unsigned long long) * nt);
  }
#endif
  auto upload = [&](const void *src, size_t bytes, void **dst) -> int {
    CUDA_OK(cudaMalloc(dst, (std::max<size_t>(bytes, 16) + 15) & ~(size_t)15));  // (whole 16-byte rows: the byte arrays are also read by aligned words)
    if (bytes) CUDA_OK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return WF_OK;
  };
  int rc;
  std::vector<uint8_t> topo((size_t)std::max<int64_t>(n, 1));
  for (int64_t k = 0; k < n; k++)
    topo[(size_t)k] = (uint8_t)((snet.nup[(size_t)k] << 4) | (ldd[snet.order[(size_t)k]] & 15));
  if ((rc = upload(snet.order.data(), sizeof(int64_t) * n, (void **)&m->d_order))) return rc;
  if ((rc = upload(m->river_flat.data(), sizeof(int64_t) * m->nr, (void **)&m->d_rorder))) return rc;
  if ((rc = upload(m->lay.lev_off.data(), sizeof(uint32_t) * m->lay.lev_off.size(), (void **)&m->d_lev_off))) return rc;
  if ((rc = upload(m->rlay.lev_off.data(), sizeof(uint32_t) * m->rlay.lev_off.size(), (void **)&m->d_rlev_off))) return rc;
  if ((rc = upload(snet.up_start.data(), sizeof(uint32_t) * n, (void **)&dm.up_start))) return rc;
  if ((rc = upload(topo.data(), n, (void **)&dm.topo))) return rc;
  if ((rc = upload(m->river_slot.data(), sizeof(int32_t) * n, (void **)&dm.river_slot))) return rc;
  dm.lev_off = m->d_lev_off;
  dm.rlev_off = m->d_rlev_off;
  if ((rc = setup_sweep(m))) return rc;
  dm.npad = (n + WF_VERT_THREADS - 1) / WF_VERT_THREADS * WF_VERT_THREADS;
  CUDA_OK(cudaMalloc(&dm.d_efT, sizeof(double) * (size_t)m->ML * std::max<int64_t>(dm.npad, 1)));
  CUDA_OK(cudaMalloc(&dm.d_efST, sizeof(double) * std::max<int64_t>(dm.npad, 1)));
  CUDA_OK(cudaMalloc(&dm.d_cpd, sizeof(double) * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMemsetAsync(dm.d_cpd, 0, sizeof(double) * std::max<int64_t>(n, 1), m->stream));
  CUDA_OK(cudaMalloc(&dm.h_r, sizeof(double) * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMalloc(&dm.h_surplus, sizeof(float) * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMalloc(&dm.h_ssf, sizeof(double) * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMalloc(&dm.h_ssf_riv, sizeof(double) * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMalloc(&dm.grid_counter, 256));
  {  // carried powers start invalid (NaN): the first solve computes the full power
    const size_t nb = sizeof(double) * (size_t)std::max<int64_t>(n, 1), rb = sizeof(double) * (size_t)std::max<int32_t>(m->nr, 1);
    CUDA_OK(cudaMalloc(&dm.h_lq, nb));
    CUDA_OK(cudaMalloc(&dm.h_lqb, nb));
    CUDA_OK(cudaMalloc(&dm.h_rq, rb));
    CUDA_OK(cudaMalloc(&dm.h_rqb, rb));
    CUDA_OK(cudaMemsetAsync(dm.h_lq, 0xff, nb, m->stream));
    CUDA_OK(cudaMemsetAsync(dm.h_lqb, 0xff, nb, m->stream));
    CUDA_OK(cudaMemsetAsync(dm.h_rq, 0xff, rb, m->stream));
    CUDA_OK(cudaMemsetAsync(dm.h_rqb, 0xff, rb, m->stream));
  }
  CUDA_OK(cudaMalloc(&dm.h_ulo, sizeof(float) * (size_t)m->ML * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMalloc(&dm.h_inwaterMM, sizeof(float) * std::max<int32_t>(m->nr, 1)));
  CUDA_OK(cudaMalloc(&dm.h_inwO, sizeof(double) * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMalloc(&dm.h_ssf_tr, sizeof(float) * std::max<int32_t>(m->nr, 1)));
  CUDA_OK(cudaMalloc(&dm.h_inwater, sizeof(float) * std::max<int32_t>(m->nr, 1)));
  CUDA_OK(cudaMalloc(&dm.h_racc, sizeof(double) * std::max<int32_t>(m->nr, 1)));
  CUDA_OK(cudaMalloc(&dm.h_rh, sizeof(double) * std::max<int32_t>(m->nr, 1)));
  {
    // river schedule in river order (= storage order of the river network for the first nrnet cells)
    if ((rc = upload(srnet.up_start.data(), sizeof(uint32_t) * srnet.ncells, (void **)&dm.r_up_start))) return rc;
    if ((rc = upload(srnet.nup.data(), srnet.ncells, (void **)&dm.r_nup))) return rc;
    dm.nrlev = (int32_t)m->rnet.nlevels();
    dm.rlev_shift = dm.nlev - dm.nrlev;
  }
  if ((rc = setup_substep_buffers(m))) return rc;
  // states exist from the start (zero = the reference's cold ZeroMap)
  for (int id = 0; id < F_COUNT; id++) {
    if ((g_fields[id].kind & 0xff) != K_STATE) continue;
    const bool channel = id == F_RiverRunoff || id == F_RiverRunoffsub || id == F_WaterLevelR || id == F_WaterLevelRsub;
    if (m->hbv != ((g_fields[id].kind & K_HBV) != 0) && !(m->hbv && (channel || id == F_GlacierStore))) continue;
    if (id >= F_UStoreLayerDepth_0 && id <= F_UStoreLayerDepth_7 && id - F_UStoreLayerDepth_0 >= m->ML) continue;
    if (id == F_GlacierStore && !cfg->glacier) continue;
    if (id == F_ReservoirVolume || id == F_LakeWaterLevel || id == F_PondingDepth) continue;  // created with the objects
    if ((rc = ensure_field(m, id))) return rc;
  }
  CUDA_OK(cudaStreamSynchronize(m->stream));
  return WF_OK;
}

void wf_destroy(wf_model *m) {
  if (!m) return;
  cudaSetDevice(m->cfg.device);
  if (m->stream) cudaStreamSynchronize(m->stream);
#ifdef WF_PROFILE
  if (m->dm.prof) {  // max over the steps run so far, per tick of the first group
    if (FILE *fp = std::fopen(std::getenv("WF_TICK_PROFILE"), "w")) {
      const int nt = m->dm.nlev + 1 + m->dm.itR;
      std::fprintf(fp, "tick,sub_loaded,sub_gathered,sub_solved,sub_done,ovl_loaded,ovl_gathered,ovl_solved,ovl_done,"
                       "riv_loaded,riv_gathered,riv_solved,riv_done,last_arrival,last_release,start_ns\n");
      for (int t = 0; t < nt; t++) {
        std::fprintf(fp, "%d", t);
        for (int k = 0; k < 15; k++) std::fprintf(fp, ",%llu", m->dm.prof[16 * (size_t)t + k]);
        std::fprintf(fp, "\n");
      }
      std::fclose(fp);
    }
    cudaFree(m->dm.prof);
  }
#endif
  if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
  for (int id = 0; id < F_COUNT; id++) {
    if (m->forc.count(id)) continue;  // the ring owns every array the field ever pointed at (below)
    if (m->external[id]) cudaFree(m->owned[id]);
    else cudaFree(m->dm.f[id]);
  }
  for (auto &kv : m->forc)
    for (int k = 0; k < wf_model::kForcSlots; k++) {
      cudaFree(kv.second.slot[k]);
      if (kv.second.ready[k]) cudaEventDestroy(kv.second.ready[k]);
      if (kv.second.freed[k]) cudaEventDestroy(kv.second.freed[k]);
    }
  for (int b = 0; b < wf_model::kStageRing; b++) {
    cudaFree(m->d_ring[b]);
    if (m->ring_copied[b]) cudaEventDestroy(m->ring_copied[b]);
    if (m->ring_free[b]) cudaEventDestroy(m->ring_free[b]);
  }
  for (float *p : m->pipe_forc) cudaFree(p);
  cudaFree(m->d_cap_slot); cudaFree(m->d_cap_col); cudaFree(m->d_cap_out); cudaFree(m->d_cap_exp);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  for (auto &g : m->gauges) { cudaFree(g.d_pos); cudaFree(g.d_rpos); cudaFree(g.d_out); }
  for (cudaEvent_t e : m->tev) cudaEventDestroy(e);
  for (float *p : m->snapshot) cudaFree(p);
  for (auto &a : m->areas) { cudaFree(a.d_perm); cudaFree(a.d_seg); cudaFree(a.d_chunk_beg); cudaFree(a.d_chunk_end); cudaFree(a.d_id_chunk0); cudaFree(a.d_partial); cudaFree(a.d_out); }
  for (void *p : m->reservoirs.owned) cudaFree(p);
  for (void *p : m->lakes.owned) cudaFree(p);
  for (void *p : m->irri_owned) cudaFree(p);
  for (void *p : m->paddy_owned) cudaFree(p);
  for (auto &kv : m->tick_total_of_mask) cudaFree(kv.second);
  cudaFree(m->dm.h_pondsrc);
  cudaFree(m->d_ah_progress); cudaFree(m->d_ah_queue); cudaFree(m->d_ah_order); cudaFree(m->d_ah_need); cudaFree(m->d_ah_rank);
  cudaFree(m->d_obj_inflow); cudaFree(m->d_mw_flux);
  cudaFree(m->sup.sws); cudaFree(m->sup.old_inwater); cudaFree(m->sup.delta); cudaFree(m->sup.any_neg);
  cudaFree(m->end.preWLRsub); cudaFree(m->end.preWLR); cudaFree(m->end.preWLLsub); cudaFree(m->end.preOrg); cudaFree(m->end.ctR); cudaFree(m->d_ct1);
  cudaFree(m->dm.d_efT); cudaFree(m->dm.d_efST); cudaFree(m->dm.d_cpd);
  cudaFree(m->dm.h_lq); cudaFree(m->dm.h_lqb); cudaFree(m->dm.h_rq); cudaFree(m->dm.h_rqb);
  cudaFree(m->dm.h_ssf_riv); cudaFree(m->dm.h_qa); cudaFree(m->dm.h_qb); cudaFree(m->dm.grid_counter);
  cudaFree(m->dm.h_r); cudaFree(m->dm.h_surplus); cudaFree(m->dm.h_ulo); cudaFree(m->dm.h_ssf);
  cudaFree(m->dm.h_inwO); cudaFree(m->dm.h_ssf_tr); cudaFree(m->dm.h_inwater); cudaFree(m->dm.h_racc); cudaFree(m->dm.h_rh);
  cudaFree(m->d_rlev_off); cudaFree(m->d_tick_total); cudaFree(m->d_groups); cudaFree((void *)m->dm.r_up_start); cudaFree((void *)m->dm.r_nup); cudaFree(m->dm.h_inwaterMM); cudaFree(m->dm.h_wb);
  cudaFree(m->dm.qo_sub); cudaFree(m->dm.qr_sub);
  cudaFree((void *)m->dm.up_start); cudaFree((void *)m->dm.topo); cudaFree((void *)m->dm.river_slot);
  cudaFree(m->d_order); cudaFree(m->d_rorder); cudaFree(m->d_lev_off); cudaFree(m->d_stage);
  for (auto &e : m->ev)
    if (e) cudaEventDestroy(e);
  if (m->stream) cudaStreamDestroy(m->stream);
  cudaGetLastError();  // a partially constructed model may have handed null handles to the calls above
  delete m;
}

static void network_to_reference_form(const Network &net, int64_t *nodes, int64_t *nodes_up, int64_t *lev_off) {
  if (nodes) std::copy(net.order.begin(), net.order.end(), nodes);
  if (lev_off) std::copy(net.lev_off.begin(), net.lev_off.end(), lev_off);
  if (nodes_up)
    for (int64_t k = 0; k < net.ncells; k++)
      for (int j = 0; j < 8; j++)
        nodes_up[k * 8 + j] = (j < net.nup[(size_t)k]) ? net.order[(size_t)net.up_start[(size_t)k] + j] : -1;
}

int wf_schedule_create(const uint8_t *ldd, int nrow, int ncol, wf_schedule **out) {
  if (!ldd || !out || nrow <= 0 || ncol <= 0) return fail(WF_EINVAL, "bad argument");
  wf_schedule *s = new wf_schedule();
  if (build_network(ldd, nrow, ncol, s->net) != 0) {
    delete s;
    return fail(WF_ECYCLE, "could not build the drainage network");
  }
  *out = s;
  return WF_OK;
}
void wf_schedule_destroy(wf_schedule *s) { delete s; }
int64_t wf_schedule_num_cells(const wf_schedule *s) { return s ? s->net.ncells : 0; }
int64_t wf_schedule_num_levels(const wf_schedule *s) { return s ? s->net.nlevels() : 0; }
int wf_schedule_get(const wf_schedule *s, int64_t *nodes, int64_t *nodes_up, int64_t *lev_off) {
  if (!s) return fail(WF_EINVAL, "null schedule");
  network_to_reference_form(s->net, nodes, nodes_up, lev_off);
  return WF_OK;
}

int wf_schedule_partition(const wf_schedule *s, int ngroups, int64_t *storage_order, int64_t *group_off,
                          int64_t *up_start, uint8_t *nup) {
  if (!s || ngroups < 1) return fail(WF_EINVAL, "bad argument");
  std::vector<int32_t> basin, gob;
  std::vector<int64_t> bcells, newpos;
  label_basins(s->net, basin, bcells);
  const int G = pack_basins(bcells, ngroups, gob);
  const int64_t n = s->net.ncells;
  std::vector<int32_t> group_of((size_t)std::max<int64_t>(n, 1), 0);
  for (int64_t k = 0; k < n; k++) group_of[(size_t)k] = gob.empty() ? 0 : gob[(size_t)basin[(size_t)k]];
  Network out;
  GroupLayout lay;
  regroup(s->net, group_of, G, out, lay, newpos);
  if (storage_order) std::copy(out.order.begin(), out.order.end(), storage_order);
  if (group_off) std::copy(lay.cell_off.begin(), lay.cell_off.end(), group_off);
  if (up_start)
    for (int64_t k = 0; k < n; k++) up_start[k] = out.up_start[(size_t)k];
  if (nup) std::copy(out.nup.begin(), out.nup.end(), nup);
  return G;
}

int wf_lateral_plan(const wf_model *m, int32_t out[4]) {
  if (!m || !out) return fail(WF_EINVAL, "null argument");
  out[0] = m->lat_mode; out[1] = m->cluster_size; out[2] = m->dm.ngroups; out[3] = m->lat_blocks;
  return WF_OK;
}

int64_t wf_num_cells(const wf_model *m) { return m ? m->n : 0; }
int64_t wf_num_levels(const wf_model *m) { return m ? m->net.nlevels() : 0; }
int64_t wf_num_river_cells(const wf_model *m) { return m ? m->rnet.ncells : 0; }
int64_t wf_num_river_levels(const wf_model *m) { return m ? m->rnet.nlevels() : 0; }

int wf_get_network(const wf_model *m, int which, int64_t *nodes, int64_t *nodes_up, int64_t *lev_off) {
  if (!m) return fail(WF_EINVAL, "null model");
  network_to_reference_form(which ? m->rnet : m->net, nodes, nodes_up, lev_off);
  return WF_OK;
}

int wf_catchment_total(const wf_model *m, const float *values, double *out_grid) {
  if (!m || !out_grid) return fail(WF_EINVAL, "null argument");
  catchment_total(m->net, values, out_grid);
  return WF_OK;
}

static int stage(wf_model *m) {
  if (!m->d_stage) CUDA_OK(cudaMalloc(&m->d_stage, sizeof(float) * (size_t)m->cfg.nrow * m->cfg.ncol));
  return WF_OK;
}

static int gather_into(wf_model *m, int id, const float *dgrid, int tile_ncol = 0) {
  if (int rc = ensure_field(m, id)) return rc;
  if ((g_fields[id].kind & 0xff) == K_STATIC) m->derived_dirty = true;
  if (id == F_zi) m->satwd_valid = false;  // SatWaterDepth follows the water table that was set ...
  if (id == F_SatWaterDepth) m->satwd_valid = true;  // ... unless the caller hands the map over itself (wf_resume reads SatWaterDepth.map)
  const bool riv = g_fields[id].kind & K_RIVER;
  const int64_t len = riv ? m->nr : m->n;
  if (len > 0) {
    if (tile_ncol > 0)
      k_gather_tiled<<<(unsigned)((len + 255) / 256), 256, 0, m->stream>>>(dgrid, riv ? m->d_rorder : m->d_order, m->dm.f[id], len,
                                                                           m->cfg.ncol, tile_ncol);
    else
      k_gather<<<(unsigned)((len + 255) / 256), 256, 0, m->stream>>>(dgrid, riv ? m->d_rorder : m->d_order, m->dm.f[id], len);
    m->launches++;
  }
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

// Host raster -> next slot of the staging ring: the copy runs on the copy stream (it waits only for the
// gather that last read the slot), the model's stream waits for the copy -- so the upload of the next
// step's forcing overlaps the running step's kernels as long as the host enqueues ahead.  The caller
// enqueues its gather from d_ring[*slot] on the model's stream and then calls ring_release.
static int ring_upload(wf_model *m, const float *grid, int *slot, int tile_ncol = 0) {
  const size_t bytes = sizeof(float) * (size_t)m->cfg.nrow * (tile_ncol > 0 ? tile_ncol : m->cfg.ncol);
  const int b = m->ring_next;
  m->ring_next = (b + 1) % wf_model::kStageRing;
  if (!m->copy_stream) CUDA_OK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  if (!m->d_ring[b]) {
    CUDA_OK(cudaMalloc(&m->d_ring[b], sizeof(float) * (size_t)m->cfg.nrow * m->cfg.ncol));
    CUDA_OK(cudaEventCreateWithFlags(&m->ring_copied[b], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&m->ring_free[b], cudaEventDisableTiming));
  } else {
    CUDA_OK(cudaStreamWaitEvent(m->copy_stream, m->ring_free[b], 0));
  }
  CUDA_OK(cudaMemcpyAsync(m->d_ring[b], grid, bytes, cudaMemcpyHostToDevice, m->copy_stream));
  CUDA_OK(cudaEventRecord(m->ring_copied[b], m->copy_stream));
  CUDA_OK(cudaStreamWaitEvent(m->stream, m->ring_copied[b], 0));
  *slot = b;
  return WF_OK;
}
static int ring_release(wf_model *m, int slot) {
  CUDA_OK(cudaEventRecord(m->ring_free[slot], m->stream));
  return WF_OK;
}

// The model's stream waits for the forcing arrays filled on the copy stream that it has not seen yet.
// Called by everything that reads a forcing field on the model's stream.
static int forcing_acquire(wf_model *m) {
  for (auto &kv : m->forc) {
    wf_model::ForcRing &r = kv.second;
    if (r.pending && r.cur >= 0) {
      CUDA_OK(cudaStreamWaitEvent(m->stream, r.ready[r.cur], 0));
      r.pending = false;
    }
  }
  return WF_OK;
}
// After the kernels of a step have been enqueued: the forcing arrays they read may be refilled once they are done.
static int forcing_release(wf_model *m) {
  for (auto &kv : m->forc) {
    wf_model::ForcRing &r = kv.second;
    if (r.cur >= 0 && !m->external[kv.first]) {
      CUDA_OK(cudaEventRecord(r.freed[r.cur], m->stream));
      r.used[r.cur] = true;
    }
  }
  return WF_OK;
}
// wf_set_field of a forcing raster: host -> staging ring -> next storage-order slot, all on the copy stream.
static int set_forcing_async(wf_model *m, int id, const float *grid, int tile_ncol = 0) {
  const bool riv = g_fields[id].kind & K_RIVER;
  const int64_t len = riv ? m->nr : m->n;
  const bool fresh = m->forc.find(id) == m->forc.end();
  wf_model::ForcRing &r = m->forc[id];
  if (fresh && m->dm.f[id]) {
    // the field's own array becomes slot 0 of the ring; steps already enqueued may still read it
    r.slot[0] = m->dm.f[id];
    CUDA_OK(cudaEventCreateWithFlags(&r.ready[0], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&r.freed[0], cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(r.freed[0], m->stream));
    r.used[0] = true;
    r.cur = 0;
  }
  const int k = (r.cur + 1) % wf_model::kForcSlots;
  if (!r.slot[k]) {
    CUDA_OK(cudaMalloc(&r.slot[k], sizeof(float) * (size_t)std::max<int64_t>(len, 1)));
    CUDA_OK(cudaEventCreateWithFlags(&r.ready[k], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&r.freed[k], cudaEventDisableTiming));
  }
  const size_t full = sizeof(float) * (size_t)m->cfg.nrow * m->cfg.ncol;
  const size_t bytes = tile_ncol > 0 ? sizeof(float) * (size_t)m->cfg.nrow * tile_ncol : full;
  const int b = m->ring_next;
  m->ring_next = (b + 1) % wf_model::kStageRing;
  if (!m->copy_stream) CUDA_OK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  if (!m->d_ring[b]) {
    CUDA_OK(cudaMalloc(&m->d_ring[b], full));
    CUDA_OK(cudaEventCreateWithFlags(&m->ring_copied[b], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&m->ring_free[b], cudaEventDisableTiming));
  } else {
    CUDA_OK(cudaStreamWaitEvent(m->copy_stream, m->ring_free[b], 0));  // the gather that last read the staging raster
  }
  if (r.used[k]) CUDA_OK(cudaStreamWaitEvent(m->copy_stream, r.freed[k], 0));  // the step that last read the slot
  r.used[k] = false;
  CUDA_OK(cudaMemcpyAsync(m->d_ring[b], grid, bytes, cudaMemcpyHostToDevice, m->copy_stream));
  if (len > 0) {
    if (tile_ncol > 0)
      k_gather_tiled<<<(unsigned)((len + 255) / 256), 256, 0, m->copy_stream>>>(m->d_ring[b], riv ? m->d_rorder : m->d_order, r.slot[k], len,
                                                                                m->cfg.ncol, tile_ncol);
    else
      k_gather<<<(unsigned)((len + 255) / 256), 256, 0, m->copy_stream>>>(m->d_ring[b], riv ? m->d_rorder : m->d_order, r.slot[k], len);
    m->launches++;
  }
  CUDA_OK(cudaEventRecord(m->ring_free[b], m->copy_stream));
  CUDA_OK(cudaEventRecord(r.ready[k], m->copy_stream));
  r.cur = k;
  r.pending = true;
  m->dm.f[id] = r.slot[k];
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

int wf_set_field(wf_model *m, const char *name, const float *grid) {
  if (!m || !grid) return fail(WF_EINVAL, "null argument");
  const int id = field_index(name);
  if (id < 0) return fail(WF_EINVAL, std::string("unknown field: ") + (name ? name : "(null)"));
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if ((g_fields[id].kind & 0xff) == K_FORCING && !m->external[id]) return set_forcing_async(m, id, grid);
  int b = 0;
  if (int rc = ring_upload(m, grid, &b)) return rc;
  const int rc = gather_into(m, id, m->d_ring[b]);
  if (int rc2 = ring_release(m, b)) return rc2;
  return rc;
}

int wf_set_field_tiled(wf_model *m, const char *name, const float *grid, int tile_ncol) {
  if (!m || !grid) return fail(WF_EINVAL, "null argument");
  if (tile_ncol < 1 || m->cfg.ncol % tile_ncol) return fail(WF_EINVAL, "the grid's columns must be a whole number of tiles");
  const int id = field_index(name);
  if (id < 0) return fail(WF_EINVAL, std::string("unknown field: ") + (name ? name : "(null)"));
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if ((g_fields[id].kind & 0xff) == K_FORCING && !m->external[id]) return set_forcing_async(m, id, grid, tile_ncol);
  int b = 0;
  if (int rc = ring_upload(m, grid, &b, tile_ncol)) return rc;
  const int rc = gather_into(m, id, m->d_ring[b], tile_ncol);
  if (int rc2 = ring_release(m, b)) return rc2;
  return rc;
}

int wf_set_field_from_device(wf_model *m, const char *name, const float *dgrid) {
  if (!m || !dgrid) return fail(WF_EINVAL, "null argument");
  const int id = field_index(name);
  if (id < 0) return fail(WF_EINVAL, std::string("unknown field: ") + (name ? name : "(null)"));
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  return gather_into(m, id, dgrid);
}

int wf_set_field_uniform(wf_model *m, const char *name, float value) {
  if (!m) return fail(WF_EINVAL, "null model");
  const int id = field_index(name);
  if (id < 0) return fail(WF_EINVAL, std::string("unknown field: ") + (name ? name : "(null)"));
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  if (int rc = ensure_field(m, id)) return rc;
  if ((g_fields[id].kind & 0xff) == K_STATIC) m->derived_dirty = true;
  const int64_t len = field_len(m, id);
  if (len > 0) {
    k_fill<<<(unsigned)((len + 255) / 256), 256, 0, m->stream>>>(m->dm.f[id], value, len);
    m->launches++;
  }
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

int wf_get_field(wf_model *m, const char *name, float *grid, float mv) {
  if (!m || !grid) return fail(WF_EINVAL, "null argument");
  const int id = field_index(name);
  if (id < 0) return fail(WF_EINVAL, std::string("unknown field: ") + (name ? name : "(null)"));
  if (!m->dm.f[id]) return fail(WF_EINVAL, std::string("field has no data (not set / not requested): ") + name);
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  if (int rc = stage(m)) return rc;
  const int64_t N = (int64_t)m->cfg.nrow * m->cfg.ncol;
  const bool riv = g_fields[id].kind & K_RIVER;
  k_fill<<<(unsigned)((N + 255) / 256), 256, 0, m->stream>>>(m->d_stage, mv, N);
  m->launches++;
  if (riv && m->n > 0) {
    // river-only arrays: the reference holds 0 at the other cells of the network (np.zeros outputs,
    // wflow_funcs.py:172-182), missing value outside it
    k_scatter_const<<<(unsigned)((m->n + 255) / 256), 256, 0, m->stream>>>(0.0f, m->d_order, m->d_stage, m->n);
    m->launches++;
  }
  const int64_t len = riv ? m->nr : m->n;
  if (len > 0) {
    k_scatter<<<(unsigned)((len + 255) / 256), 256, 0, m->stream>>>(m->dm.f[id], riv ? m->d_rorder : m->d_order, m->d_stage, len);
    m->launches++;
  }
  CUDA_OK(cudaMemcpyAsync(grid, m->d_stage, sizeof(float) * N, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  return WF_OK;
}

int wf_request_output(wf_model *m, const char *name, int enable) {
  if (!m) return fail(WF_EINVAL, "null model");
  const int id = field_index(name);
  if (id < 0) return fail(WF_EINVAL, std::string("unknown field: ") + (name ? name : "(null)"));
  if ((g_fields[id].kind & 0xff) != K_DIAG) return WF_OK;  // forcing/static/state always exist once set
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (enable) {
    if (int rc = ensure_field(m, id)) return rc;
    {  // maps an end-of-dynamic() output is formed from (outputs.cuh)
      static const std::map<int, std::vector<int>> deps = {
          {F_MassBalKinWaveR, {F_Inwater}}, {F_MassBalKinWaveL, {F_InwaterL, F_InflowKinWaveCellLand}},
          {F_RunoffCoeff, {F_RainFallPlusMelt}}, {F_CellStorage, {F_SatWaterDepth}}, {F_DeltaStorage, {F_SatWaterDepth}},
          {F_RootStore, {F_RootStore_unsat}}, {F_vwcRoot, {F_RootStore_unsat}}, {F_vwc_percRoot, {F_RootStore_unsat}},
          {F_ExfiltWaterCubic, {F_ExfiltWater}}, {F_InfiltExcessCubic, {F_InfiltExcess}}, {F_CumActInfilt, {F_ActInfilt}},
          {F_CumCellInFlow, {F_CellInFlow}}, {F_CumPrec, {F_Precipitation}}, {F_CumEvap, {F_ActEvap}},
          {F_CumPotenEvap, {F_PotenEvap}}, {F_CumInt, {F_Interception}}, {F_CumLeakage, {F_ActLeakage}},
          {F_CumExfiltWater, {F_ExfiltWater}}};
      auto it = deps.find(id);
      if (it != deps.end())
        for (int d : it->second)
          if (int rc = ensure_field(m, d)) return rc;
      if (id >= F_vwc_perc_0 && id <= F_vwc_perc_7)
        if (int rc = ensure_field(m, F_vwc_0 + (id - F_vwc_perc_0))) return rc;
    }
    if (m->hbv) {  // wflow_hbv's water balance is formed from its sums (hbv.cuh)
      if (id == F_watbal)
        for (int d : {F_initstorage, F_sumprecip, F_sumsoilevap, F_sumrunoff})
          if (int rc = ensure_field(m, d)) return rc;
      return WF_OK;
    }
    if (id == F_SoilWatbal || id == F_watbal) {
      if (!m->dm.h_wb) CUDA_OK(cudaMalloc(&m->dm.h_wb, sizeof(double) * std::max<int64_t>(m->n, 1)));
      if (int rc = ensure_field(m, F_SoilWatbal)) return rc;
      if (id == F_watbal)
        if (int rc = ensure_field(m, F_SurfaceWatbal)) return rc;
    }
  } else if (m->dm.f[id]) {
    CUDA_OK(cudaStreamSynchronize(m->stream));
    cudaFree(m->dm.f[id]);
    m->dm.f[id] = nullptr;
  }
  return WF_OK;
}

int wf_field_device_ptr(wf_model *m, const char *name, void **dptr, int64_t *n) {
  if (!m || !dptr) return fail(WF_EINVAL, "null argument");
  const int id = field_index(name);
  if (id < 0) return fail(WF_EINVAL, std::string("unknown field: ") + (name ? name : "(null)"));
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  if (int rc = ensure_field(m, id)) return rc;
  if ((g_fields[id].kind & 0xff) == K_STATIC) m->derived_dirty = true;  // the caller may write through the pointer
  *dptr = m->dm.f[id];
  if (n) *n = field_len(m, id);
  return WF_OK;
}

int wf_cell_order(const wf_model *m, int64_t *flat_index) {
  if (!m || !flat_index) return fail(WF_EINVAL, "null argument");
  std::copy(m->snet.order.begin(), m->snet.order.end(), flat_index);
  return WF_OK;
}

int wf_step(wf_model *m, int nsteps) {
  if (!m) return fail(WF_EINVAL, "null model");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  if (int rc = check_ready(m)) return rc;
  m->timed = true;
  for (int s = 0; s < nsteps; s++)
    if (int rc = dispatch_step(m)) return rc;
  return forcing_release(m);
}


// ---- pipelined multi-step stepping ---------------------------------------------------------------------
static int pipe_forcing_index(const char *name) {
  const int id = field_index(name);
  return id == F_P ? 0 : (id == F_PET ? 1 : (id == F_TEMP ? 2 : -1));
}

int wf_pipeline_reserve(wf_model *m, int max_steps) {
  if (!m || max_steps < 1) return fail(WF_EINVAL, "bad argument");
  if (max_steps > 4096) return fail(WF_EUNSUPPORTED, "more than 4096 steps per launch are not supported");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (max_steps <= m->pipe_cap) return WF_OK;
  CUDA_OK(cudaStreamSynchronize(m->stream));
  const size_t per = (size_t)std::max<int64_t>(m->dm.npad, 1);
  for (int k = 0; k < 3; k++) {
    float *p = nullptr;
    CUDA_OK(cudaMalloc(&p, sizeof(float) * per * (size_t)max_steps));
    CUDA_OK(cudaMemsetAsync(p, 0, sizeof(float) * per * (size_t)max_steps, m->stream));
    if (m->pipe_forc[k]) {
      CUDA_OK(cudaMemcpyAsync(p, m->pipe_forc[k], sizeof(float) * per * (size_t)m->pipe_cap, cudaMemcpyDeviceToDevice, m->stream));
      CUDA_OK(cudaStreamSynchronize(m->stream));
      cudaFree(m->pipe_forc[k]);
    }
    m->pipe_forc[k] = p;
  }
  if (m->cap_n > 0) {
    cudaFree(m->d_cap_out);
    m->d_cap_out = nullptr;
    CUDA_OK(cudaMalloc(&m->d_cap_out, sizeof(float) * (size_t)m->cap_n * (size_t)max_steps));
  }
  if (m->pipe_gauges >= 0) {
    cudaFree(m->d_cap_exp);
    m->d_cap_exp = nullptr;
    CUDA_OK(cudaMalloc(&m->d_cap_exp, sizeof(float) * (size_t)std::max<int64_t>(m->gauges[(size_t)m->pipe_gauges].n, 1) * (size_t)max_steps));
  }
  m->pipe_cap = max_steps;
  return WF_OK;
}

int wf_pipeline_set_forcing(wf_model *m, const char *name, int step, const float *grid) {
  if (!m || !grid) return fail(WF_EINVAL, "null argument");
  const int k = pipe_forcing_index(name);
  if (k < 0) return fail(WF_EINVAL, "pipelined forcing is P, PET or TEMP");
  if (step < 0 || step >= m->pipe_cap) return fail(WF_EINVAL, "step outside the reserved window (wf_pipeline_reserve)");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  int b = 0;
  if (int rc = ring_upload(m, grid, &b)) return rc;
  if (m->n > 0) {
    k_gather<<<(unsigned)((m->n + 255) / 256), 256, 0, m->stream>>>(m->d_ring[b], m->d_order, m->pipe_forc[k] + (size_t)step * m->dm.npad, m->n);
    m->launches++;
  }
  m->pipe_has[k] = true;
  if (int rc = ring_release(m, b)) return rc;
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

int wf_pipeline_forcing_ptr(wf_model *m, const char *name, void **dptr, int64_t *stride) {
  if (!m || !dptr) return fail(WF_EINVAL, "null argument");
  const int k = pipe_forcing_index(name);
  if (k < 0) return fail(WF_EINVAL, "pipelined forcing is P, PET or TEMP");
  if (m->pipe_cap < 1) return fail(WF_EINVAL, "call wf_pipeline_reserve first");
  *dptr = m->pipe_forc[k];
  if (stride) *stride = m->dm.npad;
  m->pipe_has[k] = true;  // the caller fills it
  return WF_OK;
}

int wf_pipeline_capture(wf_model *m, int gauges) {
  if (!m || gauges >= (int)m->gauges.size()) return fail(WF_EINVAL, "bad gauge-set handle");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  cudaFree(m->d_cap_slot); cudaFree(m->d_cap_col); cudaFree(m->d_cap_out); cudaFree(m->d_cap_exp);
  m->d_cap_slot = m->d_cap_col = nullptr;
  m->d_cap_out = m->d_cap_exp = nullptr;
  m->cap_n = 0;
  m->pipe_gauges = gauges < 0 ? -1 : gauges;
  if (gauges < 0) return WF_OK;
  const wf_model::GaugeSet &g = m->gauges[(size_t)gauges];
  std::vector<int32_t> slot((size_t)std::max<int32_t>(m->nr, 1), -1), col((size_t)std::max<int64_t>(g.n, 1), -1);
  int32_t ncol = 0;
  for (int64_t i = 0; i < g.n; i++) {
    const int32_t r = g.rpos[(size_t)i];
    if (r < 0 || r >= m->nrnet) continue;  // not a cell of the river network: 0 (inside the network) or mv
    if (slot[(size_t)r] < 0) slot[(size_t)r] = ncol++;
    col[(size_t)i] = slot[(size_t)r];
  }
  m->cap_n = ncol;
  CUDA_OK(cudaMalloc(&m->d_cap_slot, sizeof(int32_t) * slot.size()));
  CUDA_OK(cudaMalloc(&m->d_cap_col, sizeof(int32_t) * col.size()));
  CUDA_OK(cudaMemcpy(m->d_cap_slot, slot.data(), sizeof(int32_t) * slot.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(m->d_cap_col, col.data(), sizeof(int32_t) * col.size(), cudaMemcpyHostToDevice));
  const size_t steps = (size_t)std::max(m->pipe_cap, 1);
  CUDA_OK(cudaMalloc(&m->d_cap_out, sizeof(float) * (size_t)std::max<int32_t>(ncol, 1) * steps));
  CUDA_OK(cudaMalloc(&m->d_cap_exp, sizeof(float) * (size_t)std::max<int64_t>(g.n, 1) * steps));
  return WF_OK;
}

int wf_step_pipelined(wf_model *m, int nsteps, float *gauge_out, float mv, int async) {
  if (!m) return fail(WF_EINVAL, "null model");
  if (nsteps < 1) return fail(WF_EINVAL, "nsteps must be >= 1");
  if (nsteps > m->pipe_cap) return fail(WF_EINVAL, "more steps than reserved (wf_pipeline_reserve)");
  if (nsteps > WF_PIPE_MAX_INFLIGHT) return fail(WF_EUNSUPPORTED, "at most 64 timesteps per pipelined launch");
  if (m->hbv) return fail(WF_EUNSUPPORTED, "wflow_hbv models run the per-step path (wf_step)");
  if (!m->pipe_has[0] || !m->pipe_has[1]) return fail(WF_EINVAL, "pipelined forcing not set: P and PET (wf_pipeline_set_forcing)");
  if (m->cfg.model_snow && !m->pipe_has[2]) return fail(WF_EINVAL, "pipelined forcing not set: TEMP");
  if (m->dm.task_mask != 7) return fail(WF_EUNSUPPORTED, "WF_LATERAL_TASKS is a development aid of the per-step sweep");
  if (gauge_out && m->pipe_gauges < 0) return fail(WF_EINVAL, "no gauge set registered (wf_pipeline_capture)");
  if (m->reservoirs.dev.n + m->lakes.dev.n > 0 || m->mass_wasting || m->abstractions || m->irri.n > 0 || m->paddy.n > 0)
    return fail(WF_EUNSUPPORTED, "reservoirs, lakes, mass wasting, abstractions and irrigation need the per-step path (wf_step)");
  if (end_outputs_requested(m))
    return fail(WF_EUNSUPPORTED, "the end-of-dynamic() outputs (mass balances, cumulative sums ...) need the per-step path (wf_step)");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  m->timed = true;
  if (int rc = dispatch_pipeline(m, nsteps)) return rc;
  if (gauge_out) {
    const wf_model::GaugeSet &g = m->gauges[(size_t)m->pipe_gauges];
    const int64_t total = g.n * (int64_t)nsteps;
    if (total > 0) {
      k_expand_capture<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(m->d_cap_out, m->d_cap_col, g.d_pos, m->d_cap_exp, g.n,
                                                                               m->cap_n, total, mv);
      m->launches++;
      CUDA_OK(cudaMemcpyAsync(gauge_out, m->d_cap_exp, sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost, m->stream));
    }
    if (!async) CUDA_OK(cudaStreamSynchronize(m->stream));
  }
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

int wf_pipeline_mode(const wf_model *m) { return m ? m->last_multistep : 0; }

int wf_synchronize(wf_model *m) {
  if (!m) return fail(WF_EINVAL, "null model");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (m->copy_stream) CUDA_OK(cudaStreamSynchronize(m->copy_stream));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  return WF_OK;
}

int wf_last_step_ms(wf_model *m, float ms[3]) {
  if (!m || !ms) return fail(WF_EINVAL, "null argument");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (!m->ev_last[2]) return fail(WF_EINVAL, "no step has run yet");
  CUDA_OK(cudaEventSynchronize(m->ev_last[2]));
  CUDA_OK(cudaEventElapsedTime(&ms[0], m->ev_last[0], m->ev_last[1]));
  CUDA_OK(cudaEventElapsedTime(&ms[1], m->ev_last[1], m->ev_last[2]));
  CUDA_OK(cudaEventElapsedTime(&ms[2], m->ev_last[0], m->ev_last[2]));
  return WF_OK;
}

int wf_timing_reset(wf_model *m) {
  if (!m) return fail(WF_EINVAL, "null model");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  const size_t want = 3 * 4096;
  while (m->tev.size() < want) {
    cudaEvent_t e;
    CUDA_OK(cudaEventCreate(&e));
    m->tev.push_back(e);
  }
  m->tev_used = 0;
  m->tev_extra_steps = 0;
  return WF_OK;
}

int wf_timing_get(wf_model *m, double ms[3], int64_t *nsteps) {
  if (!m || !ms) return fail(WF_EINVAL, "null argument");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  ms[0] = ms[1] = ms[2] = 0.0;
  for (size_t i = 0; i + 3 <= m->tev_used; i += 3) {
    float a = 0, b = 0;
    CUDA_OK(cudaEventElapsedTime(&a, m->tev[i], m->tev[i + 1]));
    CUDA_OK(cudaEventElapsedTime(&b, m->tev[i + 1], m->tev[i + 2]));
    ms[0] += a; ms[1] += b; ms[2] += a + b;
  }
  if (nsteps) *nsteps = (int64_t)(m->tev_used / 3) + m->tev_extra_steps;
  return WF_OK;
}

int wf_bind_field(wf_model *m, const char *name, float *dptr) {
  if (!m) return fail(WF_EINVAL, "null model");
  const int id = field_index(name);
  if (id < 0) return fail(WF_EINVAL, std::string("unknown field: ") + (name ? name : "(null)"));
  if ((g_fields[id].kind & 0xff) == K_STATIC) m->derived_dirty = true;
  if (dptr) {
    if (!m->external[id]) m->owned[id] = m->dm.f[id];
    m->dm.f[id] = dptr;
    m->external[id] = true;
  } else if (m->external[id]) {
    m->dm.f[id] = m->owned[id];
    m->owned[id] = nullptr;
    m->external[id] = false;
  }
  return WF_OK;
}

int64_t wf_launch_count(const wf_model *m) { return m ? m->launches : 0; }
void *wf_stream(wf_model *m) { return m ? (void *)m->stream : nullptr; }

// ---- gauge / area accumulation ----------------------------------------------------------
int wf_area_create(wf_model *m, const int32_t *idmap) {
  if (!m || !idmap) return fail(WF_EINVAL, "null argument");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  AreaSet a;
  std::vector<std::pair<int32_t, int32_t>> pairs;  // (id, network index)
  pairs.reserve((size_t)m->n);
  for (int64_t k = 0; k < m->n; k++) {
    const int32_t id = idmap[m->snet.order[(size_t)k]];
    if (id > 0) pairs.emplace_back(id, (int32_t)k);
  }
  std::stable_sort(pairs.begin(), pairs.end(),
                   [](const std::pair<int32_t, int32_t> &x, const std::pair<int32_t, int32_t> &y) { return x.first < y.first; });
  std::vector<int32_t> perm(pairs.size()), seg;
  for (size_t i = 0; i < pairs.size(); i++) {
    if (i == 0 || pairs[i].first != pairs[i - 1].first) {
      a.ids.push_back(pairs[i].first);
      seg.push_back((int32_t)i);
    }
    perm[i] = pairs[i].second;
  }
  seg.push_back((int32_t)pairs.size());
  a.ncell = (int64_t)pairs.size();
  {
    std::vector<int32_t> cb, ce, c0(1, 0);
    for (size_t b = 0; b + 1 < seg.size(); b++) {
      for (int32_t p = seg[b]; p < seg[b + 1]; p += WF_AREA_CHUNK) {
        cb.push_back(p);
        ce.push_back(std::min<int32_t>(p + WF_AREA_CHUNK, seg[b + 1]));
      }
      c0.push_back((int32_t)cb.size());
    }
    a.nchunks = (int)cb.size();
    CUDA_OK(cudaMalloc(&a.d_chunk_beg, sizeof(int32_t) * std::max<size_t>(cb.size(), 1)));
    CUDA_OK(cudaMalloc(&a.d_chunk_end, sizeof(int32_t) * std::max<size_t>(ce.size(), 1)));
    CUDA_OK(cudaMalloc(&a.d_id_chunk0, sizeof(int32_t) * c0.size()));
    CUDA_OK(cudaMalloc(&a.d_partial, sizeof(AreaPartial) * std::max<size_t>(cb.size(), 1)));
    CUDA_OK(cudaMalloc(&a.d_out, sizeof(float) * std::max<size_t>(a.ids.size(), 1)));
    if (!cb.empty()) {
      CUDA_OK(cudaMemcpy(a.d_chunk_beg, cb.data(), sizeof(int32_t) * cb.size(), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(a.d_chunk_end, ce.data(), sizeof(int32_t) * ce.size(), cudaMemcpyHostToDevice));
    }
    CUDA_OK(cudaMemcpy(a.d_id_chunk0, c0.data(), sizeof(int32_t) * c0.size(), cudaMemcpyHostToDevice));
  }
  CUDA_OK(cudaMalloc(&a.d_perm, sizeof(int32_t) * std::max<size_t>(perm.size(), 1)));
  CUDA_OK(cudaMalloc(&a.d_seg, sizeof(int32_t) * seg.size()));
  if (!perm.empty()) CUDA_OK(cudaMemcpy(a.d_perm, perm.data(), sizeof(int32_t) * perm.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(a.d_seg, seg.data(), sizeof(int32_t) * seg.size(), cudaMemcpyHostToDevice));
  m->areas.push_back(a);
  return (int)m->areas.size() - 1;
}

int wf_area_num_ids(wf_model *m, int area) {
  if (!m || area < 0 || area >= (int)m->areas.size()) return fail(WF_EINVAL, "bad area handle");
  return (int)m->areas[(size_t)area].ids.size();
}

int wf_area_ids(wf_model *m, int area, int32_t *ids) {
  if (!m || !ids || area < 0 || area >= (int)m->areas.size()) return fail(WF_EINVAL, "bad area handle");
  const auto &v = m->areas[(size_t)area].ids;
  std::copy(v.begin(), v.end(), ids);
  return WF_OK;
}

int wf_area_reduce(wf_model *m, int area, const char *field, int op, float *out) {
  if (!m || !out || area < 0 || area >= (int)m->areas.size()) return fail(WF_EINVAL, "bad area handle");
  const int id = field_index(field);
  if (id < 0 || !m->dm.f[id]) return fail(WF_EINVAL, std::string("field has no data: ") + (field ? field : "(null)"));
  if (op < 0 || op > 3) return fail(WF_EINVAL, "bad reduction op");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  AreaSet &a = m->areas[(size_t)area];
  const int nids = (int)a.ids.size();
  if (nids == 0) return WF_OK;
  const bool riv = g_fields[id].kind & K_RIVER;
  if (a.nchunks > 0) {
    k_area_partial<<<a.nchunks, 256, 0, m->stream>>>(m->dm.f[id], riv ? m->dm.river_slot : nullptr, a.d_perm, a.d_chunk_beg,
                                                     a.d_chunk_end, a.d_partial);
    m->launches++;
  }
  k_area_combine<<<(nids + 127) / 128, 128, 0, m->stream>>>(a.d_partial, a.d_id_chunk0, a.d_seg, nids, op, a.d_out);
  m->launches++;
  CUDA_OK(cudaMemcpyAsync(out, a.d_out, sizeof(float) * nids, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  return WF_OK;
}

int wf_sample(wf_model *m, const char *field, const int64_t *flat_index, int64_t n, float *out, float mv) {
  if (!m || !flat_index || !out) return fail(WF_EINVAL, "null argument");
  const int id = field_index(field);
  if (id < 0 || !m->dm.f[id]) return fail(WF_EINVAL, std::string("field has no data: ") + (field ? field : "(null)"));
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  // translate flat grid indices to array positions on the host (gauge lists are short)
  static thread_local std::vector<int32_t> pos;
  const int64_t N = (int64_t)m->cfg.nrow * m->cfg.ncol;
  const bool riv = g_fields[id].kind & K_RIVER;
  pos.assign((size_t)n, -1);
  std::vector<int> zero((size_t)n, 0);
  for (int64_t i = 0; i < n; i++) {
    const int64_t fl = flat_index[i];
    int32_t k = (fl >= 0 && fl < N) ? m->pos_of_flat[(size_t)fl] : -1;
    if (k >= 0 && riv) {
      const int32_t s = m->river_slot[(size_t)k];
      if (s < 0) zero[(size_t)i] = 1;
      k = s;
    }
    pos[(size_t)i] = k;
  }
  int32_t *d_pos = nullptr;
  float *d_out = nullptr;
  CUDA_OK(cudaMalloc(&d_pos, sizeof(int32_t) * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMalloc(&d_out, sizeof(float) * std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMemcpyAsync(d_pos, pos.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, m->stream));
  if (n > 0) {
    k_sample<<<(unsigned)((n + 127) / 128), 128, 0, m->stream>>>(m->dm.f[id], d_pos, d_out, n, mv);
    m->launches++;
  }
  CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  for (int64_t i = 0; i < n; i++)
    if (zero[(size_t)i]) out[i] = 0.0f;
  cudaFree(d_pos);
  cudaFree(d_out);
  return WF_OK;
}

// Persistent gauge sets: positions resolved once, sampling is one small kernel + one copy.
int wf_gauges_create(wf_model *m, const int64_t *flat_index, int64_t n) {
  if (!m || !flat_index || n < 0) return fail(WF_EINVAL, "bad argument");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  wf_model::GaugeSet g;
  g.n = n;
  const int64_t N = (int64_t)m->cfg.nrow * m->cfg.ncol;
  g.pos.assign((size_t)std::max<int64_t>(n, 1), -1);
  g.rpos.assign((size_t)std::max<int64_t>(n, 1), -1);
  for (int64_t i = 0; i < n; i++) {
    const int64_t fl = flat_index[i];
    const int32_t k = (fl >= 0 && fl < N) ? m->pos_of_flat[(size_t)fl] : -1;
    g.pos[(size_t)i] = k;
    g.rpos[(size_t)i] = (k >= 0) ? m->river_slot[(size_t)k] : -1;
  }
  const size_t nb = sizeof(int32_t) * (size_t)std::max<int64_t>(n, 1);
  CUDA_OK(cudaMalloc(&g.d_pos, nb));
  CUDA_OK(cudaMalloc(&g.d_rpos, nb));
  CUDA_OK(cudaMalloc(&g.d_out, sizeof(float) * (size_t)std::max<int64_t>(n, 1)));
  CUDA_OK(cudaMemcpy(g.d_pos, g.pos.data(), nb, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(g.d_rpos, g.rpos.data(), nb, cudaMemcpyHostToDevice));
  m->gauges.push_back(std::move(g));
  return (int)m->gauges.size() - 1;
}

int wf_gauges_sample(wf_model *m, int gauges, const char *field, float *out, float mv, int async) {
  if (!m || !out || gauges < 0 || gauges >= (int)m->gauges.size()) return fail(WF_EINVAL, "bad gauge-set handle");
  const int id = field_index(field);
  if (id < 0 || !m->dm.f[id]) return fail(WF_EINVAL, std::string("field has no data: ") + (field ? field : "(null)"));
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  wf_model::GaugeSet &g = m->gauges[(size_t)gauges];
  if (g.n == 0) return WF_OK;
  const bool riv = g_fields[id].kind & K_RIVER;
  // river-only fields hold 0 at the other cells of the network (wflow_funcs.py:172-182), mv outside it
  k_sample_gauges<<<(unsigned)((g.n + 127) / 128), 128, 0, m->stream>>>(m->dm.f[id], g.d_pos, riv ? g.d_rpos : nullptr, g.d_out, g.n, mv);
  m->launches++;
  CUDA_OK(cudaMemcpyAsync(out, g.d_out, sizeof(float) * (size_t)g.n, cudaMemcpyDeviceToHost, m->stream));
  if (!async) CUDA_OK(cudaStreamSynchronize(m->stream));
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

int wf_gauges_sample_device(wf_model *m, int gauges, const char *field, float *dout, float mv) {
  if (!m || !dout || gauges < 0 || gauges >= (int)m->gauges.size()) return fail(WF_EINVAL, "bad gauge-set handle");
  const int id = field_index(field);
  if (id < 0 || !m->dm.f[id]) return fail(WF_EINVAL, std::string("field has no data: ") + (field ? field : "(null)"));
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (int rc = forcing_acquire(m)) return rc;
  wf_model::GaugeSet &g = m->gauges[(size_t)gauges];
  if (g.n == 0) return WF_OK;
  const bool riv = g_fields[id].kind & K_RIVER;
  k_sample_gauges<<<(unsigned)((g.n + 127) / 128), 128, 0, m->stream>>>(m->dm.f[id], g.d_pos, riv ? g.d_rpos : nullptr, dout, g.n, mv);
  m->launches++;
  CUDA_OK(cudaGetLastError());
  return WF_OK;
}

// ---- reservoirs / lakes ------------------------------------------------------------------------------
int wf_reservoirs_create(wf_model *m, const int32_t *locs, const int32_t *areas, const uint8_t *ldd_org) {
  if (!m || !locs) return fail(WF_EINVAL, "null argument");
  if (m->hbv) return fail(WF_EUNSUPPORTED, "reservoirs and lakes are not available for wflow_hbv models");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  for (void *p : m->reservoirs.owned) cudaFree(p);
  m->reservoirs = wf_model::Objects();
  std::vector<int32_t> ids, pos;
  if (int rc = build_objects(m, locs, areas, ldd_org, m->reservoirs, ids, pos)) return rc;
  if (m->reservoirs.dev.n > 0)
    if (int rc = ensure_field(m, F_ReservoirVolume)) return rc;
  return m->reservoirs.dev.n;
}

int wf_lakes_create(wf_model *m, const int32_t *locs, const int32_t *linked, const int32_t *areas, const uint8_t *ldd_org,
                    int n_sh, const int32_t *sh_ids, const int32_t *sh_rows, const float *sh_data, int n_hq,
                    const int32_t *hq_ids, const int32_t *hq_rows, const float *hq_data) {
  if (!m || !locs) return fail(WF_EINVAL, "null argument");
  if (m->hbv) return fail(WF_EUNSUPPORTED, "reservoirs and lakes are not available for wflow_hbv models");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  for (void *p : m->lakes.owned) cudaFree(p);
  m->lakes = wf_model::Objects();
  wf_model::Objects &ob = m->lakes;
  std::vector<int32_t> ids, pos;
  if (int rc = build_objects(m, locs, areas, ldd_org, ob, ids, pos)) return rc;
  const int n = ob.dev.n;
  if (n == 0) return 0;
  // linked lakes: LinkedLakeLocs at a lake's cell holds the id of the lake it exchanges water with (wflow_funcs.py:766-777)
  std::vector<int32_t> link((size_t)n, -1), sh_off((size_t)n + 1, 0), hq_off((size_t)n + 1, 0);
  const int ncol = m->cfg.ncol;
  (void)ncol;
  for (int k = 0; k < n; k++) {
    const int64_t fl = m->snet.order[(size_t)pos[(size_t)k]];
    const int32_t want = linked ? linked[fl] : 0;
    if (want > 0)
      for (int j = 0; j < n; j++)
        if (ids[(size_t)j] == want) link[(size_t)k] = j;
  }
  // curve tables, keyed by lake id: storage (rows of H, S) and rating (rows of H, Q(day 1..366))
  std::vector<float> sh, hq;
  for (int k = 0; k < n; k++) {
    size_t off = 0;
    for (int t = 0; t < n_sh; t++) {
      if (sh_ids[t] == ids[(size_t)k]) sh.insert(sh.end(), sh_data + 2 * off, sh_data + 2 * (off + (size_t)sh_rows[t]));
      off += (size_t)sh_rows[t];
    }
    sh_off[(size_t)k + 1] = (int32_t)(sh.size() / 2);
    off = 0;
    for (int t = 0; t < n_hq; t++) {
      if (hq_ids[t] == ids[(size_t)k]) hq.insert(hq.end(), hq_data + 367 * off, hq_data + 367 * (off + (size_t)hq_rows[t]));
      off += (size_t)hq_rows[t];
    }
    hq_off[(size_t)k + 1] = (int32_t)(hq.size() / 367);
  }
  auto up = [&](const void *src, size_t bytes, const void **dst) -> int {
    void *d = nullptr;
    CUDA_OK(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
    if (bytes) CUDA_OK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    ob.owned.push_back(d);
    *dst = d;
    return WF_OK;
  };
  int rc;
  if ((rc = up(link.data(), sizeof(int32_t) * n, (const void **)&ob.dev.linked))) return rc;
  if ((rc = up(sh.data(), sizeof(float) * sh.size(), (const void **)&ob.dev.sh_tab))) return rc;
  if ((rc = up(sh_off.data(), sizeof(int32_t) * sh_off.size(), (const void **)&ob.dev.sh_off))) return rc;
  if ((rc = up(hq.data(), sizeof(float) * hq.size(), (const void **)&ob.dev.hq_tab))) return rc;
  if ((rc = up(hq_off.data(), sizeof(int32_t) * hq_off.size(), (const void **)&ob.dev.hq_off))) return rc;
  ob.lakes = true;
  if (int rc2 = ensure_field(m, F_LakeWaterLevel)) return rc2;
  return n;
}

// id rasters of areas and of the river intakes that carry their ids -> the device form both kinds of irrigation use
static int build_area_intakes(wf_model *m, const int32_t *areas, const int32_t *intakes, IrrigationSet &s, std::vector<void *> &owned) {
  if (!m || !areas || !intakes) return fail(WF_EINVAL, "null argument");
  if (m->hbv) return fail(WF_EUNSUPPORTED, "irrigation areas are not available for wflow_hbv models");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  for (void *p : owned) cudaFree(p);
  owned.clear();
  s = IrrigationSet();
  const int64_t N = (int64_t)m->cfg.nrow * m->cfg.ncol;
  // areas: the ids present on network cells, ascending (np.unique of idtoid); cells of an area in raster order
  std::vector<int32_t> ids;
  for (int64_t fl = 0; fl < N; fl++)
    if (areas[fl] > 0 && m->pos_of_flat[(size_t)fl] >= 0) ids.push_back(areas[fl]);
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  const int n = (int)ids.size();
  if (n == 0) return 0;
  if (m->nr <= 0) return fail(WF_EINVAL, "irrigation areas need a river to take their water from");
  auto area_of = [&](int32_t id) -> int {
    const auto it = std::lower_bound(ids.begin(), ids.end(), id);
    return (it != ids.end() && *it == id) ? (int)(it - ids.begin()) : -1;
  };
  std::vector<std::vector<int32_t>> cells((size_t)n), in_rs((size_t)n), in_pos((size_t)n);
  std::vector<int32_t> orphan_rs, orphan_id;
  for (int64_t fl = 0; fl < N; fl++) {
    const int32_t p = m->pos_of_flat[(size_t)fl];
    if (p < 0) continue;
    if (areas[fl] > 0) cells[(size_t)area_of(areas[fl])].push_back(p);
    if (intakes[fl] > 0) {
      const int32_t rs = m->river_slot[(size_t)p];
      if (rs < 0) return fail(WF_EINVAL, "an irrigation intake lies outside the river");
      const int a = area_of(intakes[fl]);
      if (a >= 0) { in_rs[(size_t)a].push_back(rs); in_pos[(size_t)a].push_back(p); }
      else { orphan_rs.push_back(rs); orphan_id.push_back(intakes[fl]); }
    }
  }
  std::vector<int32_t> seg(1, 0), perm, iseg(1, 0), irs, ipos;
  for (int a = 0; a < n; a++) {
    perm.insert(perm.end(), cells[(size_t)a].begin(), cells[(size_t)a].end());
    seg.push_back((int32_t)perm.size());
    irs.insert(irs.end(), in_rs[(size_t)a].begin(), in_rs[(size_t)a].end());
    ipos.insert(ipos.end(), in_pos[(size_t)a].begin(), in_pos[(size_t)a].end());
    iseg.push_back((int32_t)irs.size());
  }
  auto up = [&](const void *src, size_t bytes, const void **dst) -> int {
    void *d = nullptr;
    CUDA_OK(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
    if (bytes) CUDA_OK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    owned.push_back(d);
    *dst = d;
    return WF_OK;
  };
  int rc;
  if ((rc = up(ids.data(), sizeof(int32_t) * ids.size(), (const void **)&s.id))) return rc;
  if ((rc = up(seg.data(), sizeof(int32_t) * seg.size(), (const void **)&s.seg))) return rc;
  if ((rc = up(perm.data(), sizeof(int32_t) * perm.size(), (const void **)&s.perm))) return rc;
  if ((rc = up(iseg.data(), sizeof(int32_t) * iseg.size(), (const void **)&s.in_seg))) return rc;
  if ((rc = up(irs.data(), sizeof(int32_t) * irs.size(), (const void **)&s.in_rs))) return rc;
  if ((rc = up(ipos.data(), sizeof(int32_t) * ipos.size(), (const void **)&s.in_pos))) return rc;
  if ((rc = up(orphan_rs.data(), sizeof(int32_t) * orphan_rs.size(), (const void **)&s.orphan_rs))) return rc;
  if ((rc = up(orphan_id.data(), sizeof(int32_t) * orphan_id.size(), (const void **)&s.orphan_id))) return rc;
  std::vector<float> z((size_t)n, 0.0f);
  if ((rc = up(z.data(), sizeof(float) * z.size(), (const void **)&s.sqm))) return rc;
  s.n_orphan = (int32_t)orphan_rs.size();
  return n;
}

int wf_irrigation_create(wf_model *m, const int32_t *areas, const int32_t *intakes) {
  if (!m || !areas || !intakes) return fail(WF_EINVAL, "null argument");
  if (m->paddy.n > 0) return fail(WF_EUNSUPPORTED, "irrigation areas together with paddy areas are not supported (the reference lets the paddy supply overwrite the other)");
  const int n = build_area_intakes(m, areas, intakes, m->irri, m->irri_owned);
  if (n <= 0) return n;
  int rc;
  // the demand is an abstraction; its inputs are outputs of the column
  for (int id : {F_Inflow, F_IRSupplymm, F_PotTrans, F_Transpiration})
    if ((rc = ensure_field(m, id))) return rc;
  if ((rc = wf_set_option(m, "Abstractions", 1.0))) return rc;
  m->irri.n = n;  // (last: the step launches the irrigation kernels from here on)
  return n;
}

int wf_paddy_create(wf_model *m, const int32_t *areas, const int32_t *intakes) {
  if (!m || !areas || !intakes) return fail(WF_EINVAL, "null argument");
  if (m->irri.n > 0) return fail(WF_EUNSUPPORTED, "paddy areas together with irrigation areas are not supported (the reference lets the paddy supply overwrite the other)");
  const int n = build_area_intakes(m, areas, intakes, m->paddy, m->paddy_owned);
  if (n <= 0) return n;
  int rc;
  for (int id : {F_Inflow, F_IRSupplymm, F_PondingDepth, F_IrriDemandm3})
    if ((rc = ensure_field(m, id))) return rc;
  if (!m->dm.h_pondsrc) {
    CUDA_OK(cudaMalloc(&m->dm.h_pondsrc, sizeof(float) * (size_t)std::max<int64_t>(m->n, 1)));
    CUDA_OK(cudaMemset(m->dm.h_pondsrc, 0, sizeof(float) * (size_t)std::max<int64_t>(m->n, 1)));
  }
  if ((rc = wf_set_option(m, "Abstractions", 1.0))) return rc;
  m->dm.paddy = 1;
  m->paddy.n = n;  // (last: the step launches the paddy kernels from here on)
  return n;
}

int wf_set_option(wf_model *m, const char *name, double value) {
  if (!m || !name) return fail(WF_EINVAL, "null argument");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  const std::string k(name);
  if (m->hbv) {  // [model] options of wflow_hbv (wflow_hbv.py:383-470)
    if (k == "MassWasting") {
      if (value != 0.0 && m->cfg.glacier)
        return fail(WF_EUNSUPPORTED, "MassWasting together with the glacier module is not supported (the glacier reads the wasted pack)");
      m->hbv_opt.mass_wasting = value != 0.0;
    } else if (k == "SetKquickFlow") m->hbv_opt.set_kquick = value != 0.0;
    else if (k == "ExternalQbase") m->hbv_opt.external_qbase = value != 0.0;
    else if (value != 0.0) return fail(WF_EUNSUPPORTED, "option not available for wflow_hbv models: " + k);
    return WF_OK;
  }
  if (k == "MassWasting") {
    if (value != 0.0 && m->cfg.glacier)
      return fail(WF_EUNSUPPORTED, "MassWasting together with the glacier module is not supported (the glacier reads the wasted pack)");
    m->mass_wasting = value != 0.0;
  } else if (k == "Abstractions") {
    m->abstractions = value != 0.0;
    if (m->abstractions && !m->sup.sws) {
      const size_t nr = (size_t)std::max<int32_t>(m->nr, 1);
      CUDA_OK(cudaMalloc(&m->sup.sws, sizeof(float) * nr));
      CUDA_OK(cudaMalloc(&m->sup.old_inwater, sizeof(float) * nr));
      CUDA_OK(cudaMalloc(&m->sup.delta, sizeof(float)));
      CUDA_OK(cudaMalloc(&m->sup.any_neg, sizeof(int32_t)));
      CUDA_OK(cudaMemset(m->sup.sws, 0, sizeof(float) * nr));
      CUDA_OK(cudaMemset(m->sup.old_inwater, 0, sizeof(float) * nr));
    }
    m->dm.sup_sws = m->abstractions ? m->sup.sws : nullptr;
    m->dm.sup_old_inwater = m->abstractions ? m->sup.old_inwater : nullptr;
    m->dm.sup_any_neg = m->abstractions ? m->sup.any_neg : nullptr;
    m->end.sws = m->dm.sup_sws;
  } else if (k == "maxitsupply") {
    if (value < 1) return fail(WF_EINVAL, "maxitsupply must be >= 1");
    m->maxitsupply = (int)value;
  } else if (k == "breakoff") {
    m->breakoff = value;
  } else {
    return fail(WF_EINVAL, "unknown option: " + k);
  }
  return WF_OK;
}

int wf_supply_iterations(const wf_model *m) { return m ? m->last_nrit : 0; }

int wf_set_day_of_year(wf_model *m, int jdoy) {
  if (!m || jdoy < 1 || jdoy > 366) return fail(WF_EINVAL, "day of the year must be 1..366");
  m->jdoy = jdoy;
  return WF_OK;
}

// ---- kinematic-wave sub-steps ---------------------------------------------------------------------
int wf_set_substeps(wf_model *m, int it_kin_l, int it_kin_r) {
  if (!m) return fail(WF_EINVAL, "null model");
  if (it_kin_l < 1 || it_kin_r < 1) return fail(WF_EINVAL, "it_kin_l / it_kin_r must be >= 1");
  if (it_kin_l > 1024 || it_kin_r > 1024) return fail(WF_EUNSUPPORTED, "more than 1024 sub-steps are not supported");
  if (it_kin_l == m->dm.itL && it_kin_r == m->dm.itR) return WF_OK;
  CUDA_OK(cudaSetDevice(m->cfg.device));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  m->dm.itL = it_kin_l;
  m->dm.itR = it_kin_r;
  m->cfg.it_kin_l = it_kin_l;
  m->cfg.it_kin_r = it_kin_r;
  if (int rc = setup_sweep(m)) return rc;
  return setup_substep_buffers(m);
}

int wf_estimate_iterations(wf_model *m, int which, int32_t *it_kin) {
  if (!m || !it_kin) return fail(WF_EINVAL, "null argument");
  if (which != 0 && which != 1) return fail(WF_EINVAL, "which must be 0 (overland) or 1 (river)");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  *it_kin = 1;
  const int64_t len = which ? m->nr : m->n;
  const float *Q = m->dm.f[which ? F_RiverRunoffsub : F_LandRunoffsub];
  const float *alpha = m->dm.f[which ? F_AlphaR : F_AlphaL];
  const float *dx = m->dm.f[which ? F_DCL : F_DL];
  if (len == 0) return WF_OK;  // no cell with Q > 0: the reference's percentile fails and it falls back to 1
  if (!Q || !alpha || !dx) return fail(WF_EINVAL, "fields of the Courant estimate are not set");
  uint32_t *keys = nullptr;
  unsigned long long *d_cnt = nullptr;  // [0] valid count, [1..256] histogram
  CUDA_OK(cudaMalloc(&keys, sizeof(uint32_t) * (size_t)len));
  CUDA_OK(cudaMalloc(&d_cnt, sizeof(unsigned long long) * 257));
  CUDA_OK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 257, m->stream));
  k_courant_keys<<<(unsigned)((len + 255) / 256), 256, 0, m->stream>>>(Q, alpha, dx, (float)m->dm.beta, (float)m->dm.dt, keys, d_cnt, len);
  m->launches++;
  unsigned long long h[257];
  CUDA_OK(cudaMemcpyAsync(h, d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(cudaStreamSynchronize(m->stream));
  const unsigned long long nvalid = h[0];
  int rc = WF_OK;
  if (nvalid > 0) {
    // k-th smallest valid key, most significant digit first
    auto select = [&](unsigned long long k, uint32_t &out) -> int {
      uint32_t prefix = 0, mask = 0;
      for (int shift = 24; shift >= 0; shift -= 8) {
        CUDA_OK(cudaMemsetAsync(d_cnt + 1, 0, sizeof(unsigned long long) * 256, m->stream));
        k_select_hist<<<(unsigned)std::min<int64_t>((len + 255) / 256, 2048), 256, 0, m->stream>>>(keys, len, mask, prefix, shift, d_cnt + 1);
        m->launches++;
        CUDA_OK(cudaMemcpyAsync(h + 1, d_cnt + 1, sizeof(unsigned long long) * 256, cudaMemcpyDeviceToHost, m->stream));
        CUDA_OK(cudaStreamSynchronize(m->stream));
        int d = 0;
        for (; d < 255; d++) {
          if (k < h[1 + d]) break;
          k -= h[1 + d];
        }
        prefix |= (uint32_t)d << shift;
        mask |= 0xffu << shift;
      }
      out = prefix;
      return WF_OK;
    };
    // np.nanpercentile(a, 95), linear interpolation, float32 like the array (wflow_funcs.py:112)
    const double vidx = 0.95 * (double)(nvalid - 1);
    const unsigned long long lo = (unsigned long long)std::floor(vidx);
    const unsigned long long hi = std::min<unsigned long long>(lo + 1, nvalid - 1);
    const float g = (float)(vidx - (double)lo);
    uint32_t klo = 0, khi = 0;
    rc = select(lo, klo);
    if (rc == WF_OK) rc = (hi == lo) ? (khi = klo, WF_OK) : select(hi, khi);
    if (rc == WF_OK) {
      float a, b;
      std::memcpy(&a, &klo, 4);
      std::memcpy(&b, &khi, 4);
      const float diff = b - a;
      float p95 = (g >= 0.5f) ? b - diff * (1.0f - g) : a + diff * g;
      const float v = std::ceil(1.25f * p95);  // int(np.ceil(1.25 * p95)); a failing conversion means 1 (:113-114)
      if (v == v && v >= 1.0f && v < 2.0e9f) *it_kin = (int32_t)v;
    }
  }
  cudaFree(keys);
  cudaFree(d_cnt);
  return rc;
}

// ---- batched solvers -----------------------------------------------------------------------
static int solver_batch(int device, int64_t n, const double *const *args, int nargs, int nout, double *out, bool ssf) {
  if (!args || !out || n < 0) return fail(WF_EINVAL, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(WF_ENODEVICE, "no CUDA device available (this library has no CPU fallback)");
  if (n == 0) return WF_OK;
  CUDA_OK(cudaSetDevice(device));
  double *d_a = nullptr, *d_o = nullptr;
  CUDA_OK(cudaMalloc(&d_a, sizeof(double) * n * nargs));
  CUDA_OK(cudaMalloc(&d_o, sizeof(double) * n * nout));
  for (int k = 0; k < nargs; k++) CUDA_OK(cudaMemcpy(d_a + (size_t)k * n, args[k], sizeof(double) * n, cudaMemcpyHostToDevice));
  if (ssf) k_kinwave_ssf_batch<<<(unsigned)((n + 127) / 128), 128>>>(n, d_a, d_o);
  else k_kinwave_batch<<<(unsigned)((n + 127) / 128), 128>>>(n, d_a, d_o);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpy(out, d_o, sizeof(double) * n * nout, cudaMemcpyDeviceToHost));
  cudaFree(d_a);
  cudaFree(d_o);
  return WF_OK;
}
int wf_kinematic_wave_batch(int device, int64_t n, const double *const *args, double *out) {
  return solver_batch(device, n, args, 7, 1, out, false);
}
int wf_kinematic_wave_ssf_batch(int device, int64_t n, const double *const *args, double *out) {
  return solver_batch(device, n, args, 14, 3, out, true);
}

// ---- one-step rollback -------------------------------------------------------------------
int wf_quick_suspend(wf_model *m) {
  if (!m) return fail(WF_EINVAL, "null model");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  if (m->snapshot.empty()) m->snapshot.assign(F_COUNT, nullptr);
  for (int id = 0; id < F_COUNT; id++) {
    if ((g_fields[id].kind & 0xff) != K_STATE || !m->dm.f[id]) continue;
    const size_t bytes = sizeof(float) * (size_t)std::max<int64_t>(field_len(m, id), 1);
    if (!m->snapshot[(size_t)id]) CUDA_OK(cudaMalloc(&m->snapshot[(size_t)id], bytes));
    CUDA_OK(cudaMemcpyAsync(m->snapshot[(size_t)id], m->dm.f[id], bytes, cudaMemcpyDeviceToDevice, m->stream));
  }
  return WF_OK;
}

int wf_quick_resume(wf_model *m) {
  if (!m) return fail(WF_EINVAL, "null model");
  if (m->snapshot.empty()) return fail(WF_EINVAL, "no snapshot: call wf_quick_suspend first");
  CUDA_OK(cudaSetDevice(m->cfg.device));
  for (int id = 0; id < F_COUNT; id++) {
    if (!m->snapshot[(size_t)id] || !m->dm.f[id]) continue;
    const size_t bytes = sizeof(float) * (size_t)std::max<int64_t>(field_len(m, id), 1);
    CUDA_OK(cudaMemcpyAsync(m->dm.f[id], m->snapshot[(size_t)id], bytes, cudaMemcpyDeviceToDevice, m->stream));
  }
  return WF_OK;
}

}  // extern "C"
