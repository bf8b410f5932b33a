// Per-object modules of wflow_sbm that sit in front of the routing: simple reservoirs (simplereservoir,
// wflow/wflow_funcs.py:864-949) and natural lakes (naturalLake, :669-861), called from WflowModel.dynamic() at
// wflow/wflow_sbm.py:2822-2883.  Both work on a few hundred objects, each one cell of the grid (its outlet) plus an
// area map for the precipitation / evaporation it receives; their outflow enters the river cell BELOW the object
// (pcr.upstream(TopoLddOrg, outflow), :2849-2852, 2879-2883 -- initial() cuts the LDD at the object, :1899-1907).
// REAL4 map algebra in the reference: float32 here, operation by operation.
//   k_object_forcing   one block per object: area average of Precipitation and PotenEvap (the UNFILTERED maps,
//                      wflow_sbm.py:2621-2626) over the object's area map
//   k_reservoirs       one thread per reservoir: the 6-hourly sub-steps of simplereservoir
//   k_lakes            one thread per lake: modified Puls (LakeOutflowFunc 3) or the sub-stepped rating-curve balance
//                      (1: table by day of year, 2: b (H - H0)^e), storage S = A H or an S-H table, linked lakes
//   k_object_inflow    one thread: outflow of every object added to the inflow of its downstream river cell, in object order
#pragma once
#include "pipeline.cuh"

namespace wf {

struct ObjectSet {
  int32_t n;                 // objects
  const int32_t *pos;        // n   storage position of the object's cell
  const int32_t *down_rs;    // n   river slot of the cell below the object in the ORIGINAL ldd (-1: none)
  const int32_t *area_seg;   // n+1 offsets into area_perm
  const int32_t *area_perm;  //     storage positions of the cells of each object's area map
  float *prec_av, *pet_av;   // n   area averages of the step
  float *outflow;            // n   outflow of the step [m3/s]
  // lakes only
  const int32_t *linked;     // n   index of the lake this one is linked to (its lower neighbour), or -1
  const float *sh_tab;       //     storage curves: rows (H, S), per lake [sh_off[k], sh_off[k+1])
  const int32_t *sh_off;     // n+1
  const float *hq_tab;       //     rating curves: per lake rows of 367 values (H, Q(day 1..366))
  const int32_t *hq_off;     // n+1 (in rows)
};

// float32 value at the object's cell of a cell-order field (0 when the field is absent)
__device__ __forceinline__ float obj_ld(const float *f, int32_t p) { return f ? f[p] : 0.0f; }

__global__ void k_object_forcing(const DevModel m, const ObjectSet o) {
  const int k = blockIdx.x;
  const int beg = o.area_seg[k], end = o.area_seg[k + 1];
  double sp = 0.0, se = 0.0;
  for (int q = beg + threadIdx.x; q < end; q += blockDim.x) {
    const int32_t i = o.area_perm[q];
    sp += (double)fmaxf(0.0f, m.f[F_P][i]);                       // self.Precipitation, :2584
    se += (double)(m.f[F_PET][i] * m.f[F_et_RefToPot][i]);         // self.PotenEvap, :2615
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    sp += __shfl_down_sync(0xffffffffu, sp, s);
    se += __shfl_down_sync(0xffffffffu, se, s);
  }
  __shared__ double a[8], b[8];
  if ((threadIdx.x & 31) == 0) { a[threadIdx.x >> 5] = sp; b[threadIdx.x >> 5] = se; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tp = 0.0, te = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { tp += a[w]; te += b[w]; }
    const int cnt = end - beg;
    // pcr.areaaverage; an object without area cells gets a missing value, covered with 0 for reservoirs (:880-891)
    o.prec_av[k] = cnt > 0 ? (float)(tp / cnt) : 0.0f;
    o.pet_av[k] = cnt > 0 ? (float)(te / cnt) : 0.0f;
  }
}

// inflow of the object's cell: RiverRunoff + LandRunoff + SubsurfaceFlow / 1000 / 1000 / 1000 / timestepsecs, :2833-2836
__device__ __forceinline__ float object_inflow(const DevModel &m, int32_t p) {
  const int32_t rs = m.river_slot[p];
  const float RR = (rs >= 0) ? m.f[F_RiverRunoff][rs] : 0.0f;
  const float dtf = (float)m.dt;
  return (RR + m.f[F_LandRunoff][p]) + ((((m.f[F_SubsurfaceFlow][p] / 1000.0f) / 1000.0f) / 1000.0f) / dtf);
}

// sCurve of wflow_funcs.py:611-629 in REAL4 (b = 1): 1 / (b + exp(-c (X - a)))
__device__ __forceinline__ float scurve_f(float X, float a, float c) {
  return 1.0f / (1.0f + (float)exp((double)((-c) * (X - a))));
}

__global__ void k_reservoirs(const DevModel m, const ObjectSet o) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= o.n) return;
  const int32_t p = o.pos[k];
  const float dtf = (float)m.dt;
  const float inflow = object_inflow(m, p);
  const float prec_av = o.prec_av[k], pet_av = o.pet_av[k];
  const float ResArea = obj_ld(m.f[F_ResSimpleArea], p), maxstorage = obj_ld(m.f[F_ResMaxVolume], p);
  const float target = obj_ld(m.f[F_ResTargetFullFrac], p), maxQ = obj_ld(m.f[F_ResMaxRelease], p);
  const float demand = obj_ld(m.f[F_ResDemand], p), minfull = obj_ld(m.f[F_ResTargetMinFrac], p);
  float storage = m.f[F_ReservoirVolume][p];
  const int nr_loop = max((int)(m.dt / 21600.0), 1);
  const float fl = (float)nr_loop;
  float out_sum = 0.0f, dem_sum = 0.0f, percfull = 0.0f;
  for (int it = 0; it < nr_loop; it++) {
    storage = ((storage + ((inflow * dtf) / fl)) + (((prec_av / fl) / 1000.0f) * ResArea)) - (((pet_av / fl) / 1000.0f) * ResArea);
    percfull = storage / maxstorage;
    const float fac = scurve_f(percfull, minfull, 30.0f);
    const float demandRelease = fminf(((fac * demand) * dtf) / fl, storage);
    storage = storage - demandRelease;
    const float wantrel = fmaxf(0.0f, storage - (maxstorage * target));
    const float overflowQ = fmaxf(storage - maxstorage, 0.0f);
    const float torelease = fminf(wantrel, (overflowQ + ((maxQ * dtf) / fl)) - demandRelease);
    storage = storage - torelease;
    const float outflow = torelease + demandRelease;
    percfull = storage / maxstorage;
    out_sum = out_sum + outflow;
    dem_sum = dem_sum + demandRelease;
  }
  m.f[F_ReservoirVolume][p] = storage;
  const float q = out_sum / dtf;
  o.outflow[k] = q;
  if (m.f[F_OutflowSR]) m.f[F_OutflowSR][p] = q;
  if (m.f[F_ResPercFull]) m.f[F_ResPercFull][p] = percfull;
  if (m.f[F_ResPrecipSR]) m.f[F_ResPrecipSR][p] = prec_av;
  if (m.f[F_ResEvapSR]) m.f[F_ResEvapSR][p] = pet_av;
  if (m.f[F_DemandRelease]) m.f[F_DemandRelease][p] = dem_sum / dtf;
}

// np.interp(x, xp, fp) on float32 tables (float64 arithmetic like numpy, rounded to float32 by numpy2pcr)
__device__ __forceinline__ float interp_tab(float x, const float *tab, int rows, int stride, int xc, int yc) {
  if (rows <= 0) return 0.0f;
  if (x <= tab[xc]) return tab[yc];
  if (x >= tab[(rows - 1) * stride + xc]) return tab[(rows - 1) * stride + yc];
  int lo = 0, hi = rows - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab[mid * stride + xc] <= x) lo = mid; else hi = mid;
  }
  const double x0 = tab[lo * stride + xc], x1 = tab[hi * stride + xc], y0 = tab[lo * stride + yc], y1 = tab[hi * stride + yc];
  return (float)((y1 - y0) / (x1 - x0) * ((double)x - x0) + y0);
}

// x ** y of PCRaster (REAL4; evaluated in float64 and rounded once, see oracle/pcrshim.py); a domain error is a missing value
__device__ __forceinline__ float pcr_pow(float x, float y, bool &mv) {
  if ((x < 0.0f && y != floorf(y)) || (x == 0.0f && y <= 0.0f)) { mv = true; return 0.0f; }
  return (float)pow((double)x, (double)y);
}

// One block for all lakes: linked lakes read each other's level of the running sub-step and hand back-flow to each
// other (the reference does this with whole-map operations per sub-step), so the sub-steps are separated by block barriers.
// Dynamic shared memory: 3 floats per lake (level, outflow of the sub-step, new level).
__global__ void k_lakes(const DevModel m, const ObjectSet o, const int jdoy) {
  extern __shared__ float s_lake[];
  float *const s_wl = s_lake, *const s_out = s_lake + o.n, *const s_new = s_lake + 2 * o.n;
  const float dtf = (float)m.dt;
  const int nr_loop = max((int)(m.dt / 21600.0), 1);
  const float fl = (float)nr_loop;
  for (int k = threadIdx.x; k < o.n; k += blockDim.x) s_wl[k] = m.f[F_LakeWaterLevel][o.pos[k]];
  __syncthreads();
  // per-lake results of the sub-stepped branch are accumulated in global scratch (prec_av/pet_av are inputs; outflow is output)
  for (int it = 0; it < nr_loop; it++) {
    for (int k = threadIdx.x; k < o.n; k += blockDim.x) {
      const int32_t p = o.pos[k];
      const int outfunc = (int)obj_ld(m.f[F_LakeOutflowFunc], p);
      const float H0 = obj_ld(m.f[F_LakeThreshold], p), b = obj_ld(m.f[F_Lake_b], p), e = obj_ld(m.f[F_Lake_e], p);
      const float wl = s_wl[k];
      const float wl_lower = (o.linked[k] >= 0) ? s_wl[o.linked[k]] : wl;  // np_waterlevel_lower, :766-777
      float out_loop;
      bool mv = false;
      if (outfunc == 1) {
        out_loop = interp_tab(wl, o.hq_tab + 367 * o.hq_off[k], o.hq_off[k + 1] - o.hq_off[k], 367, 0, jdoy);
      } else if (wl - wl_lower >= 0.0f) {
        out_loop = fmaxf(b * pcr_pow(wl - H0, e, mv), 0.0f);
      } else {
        out_loop = fminf((-1.0f * b) * pcr_pow(wl_lower - H0, e, mv), 0.0f);
      }
      s_out[k] = mv ? nanf("") : out_loop;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < o.n; k += blockDim.x) {
      const int32_t p = o.pos[k];
      const int storfunc = (int)obj_ld(m.f[F_LakeStorFunc], p);
      const float Area = obj_ld(m.f[F_LakeArea], p);
      const float *sh = o.sh_tab + 2 * o.sh_off[k];
      const int shn = o.sh_off[k + 1] - o.sh_off[k];
      const float wl = s_wl[k];
      const float inflow = object_inflow(m, p);
      const float out_loop = s_out[k];
      const float out_cov = (out_loop == out_loop) ? out_loop : 0.0f;  // pcr.cover(outflow_loop, 0.0)
      // back-flow: a lake whose upper neighbour has a negative outflow loses that water to it (outflow_linked, :804-815)
      float out_linked = 0.0f;
      for (int u = 0; u < o.n; u++)
        if (o.linked[u] == k && s_out[u] < 0.0f) out_linked = s_out[u];
      const float storage_start = (storfunc == 1) ? Area * wl : interp_tab(wl, sh, shn, 2, 0, 1);
      float st = storage_start + ((inflow * dtf) / fl);                      // :817-825
      st = st + (((o.prec_av[k] / fl) / 1000.0f) * Area);
      st = st - (((o.pet_av[k] / fl) / 1000.0f) * Area);
      st = st - ((out_cov * dtf) / fl);
      st = st + ((out_linked * dtf) / fl);
      s_new[k] = (storfunc == 1) ? wl + ((st - storage_start) / Area) : interp_tab(st, sh, shn, 2, 1, 0);  // :827-831
      // running results: the mean of the positive outflows (float64 in numpy, :833-839), the last storage
      const float pos = (out_loop > 0.0f) ? out_loop : 0.0f;  // NaN compares false: 0
      o.outflow[k] = (it == 0) ? pos : o.outflow[k] + pos;     // (float32 running sum of <= 4 terms; averaged below)
      if (m.f[F_LakeVolume]) m.f[F_LakeVolume][p] = st;  // storage_loop of the last sub-step (:852)
    }
    __syncthreads();
    for (int k = threadIdx.x; k < o.n; k += blockDim.x) s_wl[k] = s_new[k];
    __syncthreads();
  }
  for (int k = threadIdx.x; k < o.n; k += blockDim.x) {
    const int32_t p = o.pos[k];
    const int storfunc = (int)obj_ld(m.f[F_LakeStorFunc], p), outfunc = (int)obj_ld(m.f[F_LakeOutflowFunc], p);
    const float Area = obj_ld(m.f[F_LakeArea], p), H0 = obj_ld(m.f[F_LakeThreshold], p), b = obj_ld(m.f[F_Lake_b], p);
    const float prec_av = o.prec_av[k], pet_av = o.pet_av[k];
    float waterlevel = s_wl[k], outflow = o.outflow[k] / fl;
    if (outfunc == 3) {
      // modified Puls approach (LISFLOOD), :718-752, on the level of the start of the step
      const float wl_start = m.f[F_LakeWaterLevel][p];
      const float inflow = object_inflow(m, p);
      const float LakeFactor = Area / (dtf * (float)sqrt((double)b));
      const float storage_start = (storfunc == 1) ? Area * wl_start : interp_tab(wl_start, o.sh_tab + 2 * o.sh_off[k], o.sh_off[k + 1] - o.sh_off[k], 2, 0, 1);
      const float SI = ((storage_start + (((prec_av - pet_av) * Area) / 1000.0f)) / dtf) + inflow;
      const float SIadj = SI - ((Area * H0) / dtf);
      if (SIadj > 0.0f) {
        const float r = (float)sqrt((double)((LakeFactor * LakeFactor) + (2.0f * SIadj)));
        const float d = (-LakeFactor) + r;
        outflow = d * d;
      } else {
        outflow = 0.0f;
      }
      const float storage = (SI - outflow) * dtf;
      waterlevel = storage / Area;
      if (m.f[F_LakeVolume]) m.f[F_LakeVolume][p] = storage;
    }
    m.f[F_LakeWaterLevel][p] = waterlevel;
    o.outflow[k] = outflow;
    if (m.f[F_LakeOutflow]) m.f[F_LakeOutflow][p] = outflow;
    if (m.f[F_LakePrecip]) m.f[F_LakePrecip][p] = prec_av;
    if (m.f[F_LakeEvap]) m.f[F_LakeEvap][p] = pet_av;
  }
}

// Inflow of the river cells below the objects: self.Inflow = self.OutflowDwn + self.Inflow, once per module (:2852, 2883).
// first = 1: the array starts from the external Inflow map (or 0).  One thread: object order, deterministic.
__global__ void k_object_inflow(const DevModel m, const ObjectSet o, float *__restrict__ inflow, const int first) {
  if (first) {
    for (int r = threadIdx.x; r < m.nr; r += blockDim.x) inflow[r] = m.f[F_Inflow] ? m.f[F_Inflow][r] : 0.0f;
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  // pcr.upstream sums the outflows of the objects that drain into the same cell first, then adds the map
  for (int k = 0; k < o.n; k++) {
    const int32_t r = o.down_rs[k];
    if (r < 0) continue;
    float up = 0.0f;
    bool firstk = true;
    for (int j = 0; j < k; j++)
      if (o.down_rs[j] == r) firstk = false;
    if (!firstk) continue;
    for (int j = k; j < o.n; j++)
      if (o.down_rs[j] == r) up = up + o.outflow[j];
    inflow[r] = up + inflow[r];
  }
}

// ---- snow mass wasting (wflow_sbm.py:2700-2707) -------------------------------------------------------
// self.Snow = pcr.accucapacitystate(TopoLdd, Snow, MaxFlux): dry snow slides to the cell below, at most MaxFlux per cell
// (a slope- and pack-dependent fraction of the cell's own pack); what leaves a pit leaves the map.  One more sweep over
// the levels: one block per group of the storage layout, a block barrier per level, float32 like the map algebra.
__global__ void k_mass_wasting(const DevModel m, float *__restrict__ flux) {
  for (int g = blockIdx.x; g < m.ngroups; g += gridDim.x) {
    const GroupDesc gd = m.groups[g];
    const uint32_t *lo = m.lev_off + gd.lev_idx;
    const int nl = m.nlev - gd.lev0;
    for (int l = 0; l < nl; l++) {
      const uint32_t b = lo[l], e = lo[l + 1];
      for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        const float snow = m.f[F_Snow][i];
        const float frac = fminf(0.5f, m.f[F_landSlope][i] / 5.67f) * fminf(1.0f, snow / 10000.0f);  // 5.67 = tan 80 degrees
        const float cap = frac * snow;
        float tot = snow;
        const int nup = m.topo[i] >> 4;
        const uint32_t us = m.up_start[i];
        for (int j = 0; j < nup; j++) tot = tot + flux[us + j];
        const float out = fminf(tot, cap);
        flux[i] = out;
        m.f[F_Snow][i] = tot - out;
      }
      __syncthreads();
    }
  }
}

// ---- irrigation areas (irrigationdemand, wflow_sbm.py:927-945; :3014-3031, 3314-3328; idtoid, wflow_lib.py:81-101) ---------
// An irrigation area asks for what its crops lacked this step: IrriDemand = areaaverage(PotTrans - Transpiration) [mm],
// IrriDemandm3 = IrriDemand * areatotal(xl * yl) / 1000 / timestepsecs [m3/s].  idtoid hands that figure to the river cells of the
// intake map that carry the area's id, as a NEGATIVE Inflow -- from there on it is an abstraction like any other (the
// overland task's first estimate, the supply iteration above).  What the intakes of an area were given,
// SurfaceWaterSupply * (1 - DemandReturnFlowFraction), averaged over them (np.mean of float32 values, in raster order),
// is spread over the area as IRSupplymm = supply / (area / 1000 / timestepsecs) [mm], which the NEXT step's column adds
// to the water available for infiltration.  idtoid's left-over values are the reference's and kept: an intake whose
// id is no area's asks for its own id in m3/s, an area without an intake is "supplied" its own id.
// All REAL4 map algebra: float32, operation by operation; the area sums in float64 like the area statistics.
struct IrrigationSet {
  int32_t n;                 // areas, ascending id
  const int32_t *id;         // n    id of the area
  const int32_t *seg;        // n+1  offsets into perm
  const int32_t *perm;       //      storage positions of the cells of each area
  const int32_t *in_seg;     // n+1  offsets into in_rs / in_pos: the intakes of each area, raster order
  const int32_t *in_rs;      //      river slot of the intake cell
  const int32_t *in_pos;     //      storage position of the intake cell
  int32_t n_orphan;          // intakes whose id is no area's
  const int32_t *orphan_rs;
  const int32_t *orphan_id;
  float *sqm;                // n    areatotal(xl * yl) of the step
};

// block-wide sum of float64 partials in a fixed order (thread-strided partials, then a tree over the threads)
__device__ __forceinline__ double irri_block_sum(double v, double *sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// after the column, ahead of the sweep: the demand of every area becomes the (negative) Inflow of its intakes
__global__ void __launch_bounds__(256) k_irrigation_demand(const DevModel m, const IrrigationSet s, float *obj_inflow) {
  __shared__ double sh[256];
  const int a = blockIdx.x;
  const int beg = s.seg[a], end = s.seg[a + 1];
  double se = 0.0, sa = 0.0;
  for (int p = beg + threadIdx.x; p < end; p += blockDim.x) {
    const int i = s.perm[p];
    se += (double)(m.f[F_PotTrans][i] - m.f[F_Transpiration][i]);
    sa += (double)(m.f[F_xl][i] * m.f[F_yl][i]);
  }
  se = irri_block_sum(se, sh);
  sa = irri_block_sum(sa, sh);
  const int cnt = end - beg;
  const float et = (cnt > 0) ? (float)(se / (double)cnt) : 0.0f;   // pcr.areaaverage
  const float sqm = (float)sa;                                     // pcr.areatotal
  const float m3 = ((et * sqm) / 1000.0f) / (float)m.dt;           // :942
  for (int p = beg + threadIdx.x; p < end; p += blockDim.x) {
    const int i = s.perm[p];
    if (m.f[F_IrriDemand]) m.f[F_IrriDemand][i] = et;
    if (m.f[F_IrriDemandm3]) m.f[F_IrriDemandm3][i] = m3;
  }
  if (threadIdx.x == 0) {
    s.sqm[a] = sqm;
    // idtoid(IrrigationAreas, IrrigationSurfaceIntakes, IrriDemandm3) * -1: the mean over the area of a uniform value
    // (np.mean in float32 may differ from the value in its last bit; not reproduced)
    const float v = m3 * -1.0f;
    for (int k = s.in_seg[a]; k < s.in_seg[a + 1]; k++) {
      const int rs = s.in_rs[k];
      m.f[F_Inflow][rs] = v;
      if (obj_inflow) obj_inflow[rs] = v;
    }
    if (a == 0)
      for (int k = 0; k < s.n_orphan; k++) {  // an intake without an area keeps idtoid's initial value: its own id
        const float v0 = (float)s.orphan_id[k] * -1.0f;
        m.f[F_Inflow][s.orphan_rs[k]] = v0;
        if (obj_inflow) obj_inflow[s.orphan_rs[k]] = v0;
      }
  }
}

// np.mean of k float32 values (numpy's pairwise sum: sequential below 8 values, eight partial sums up to 128)
__device__ __forceinline__ float irri_np_mean(const float *v, int k) {
  float res;
  if (k < 8) {
    res = -0.0f;
    for (int i = 0; i < k; i++) res = res + v[i];
  } else {
    float r[8];
    for (int j = 0; j < 8; j++) r[j] = v[j];
    int i = 8;
    for (; i < k - (k % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] = r[j] + v[i + j];
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < k; i++) res = res + v[i];
  }
  return res / (float)k;
}

// after the supply iteration: what the intakes were given, spread over their areas for the next step
__global__ void __launch_bounds__(256) k_irrigation_supply(const DevModel m, const IrrigationSet s, const float *sws) {
  __shared__ float sv;
  const int a = blockIdx.x;
  if (threadIdx.x == 0) {
    const int kb = s.in_seg[a], k = s.in_seg[a + 1] - kb;
    float val;
    if (k == 0) {
      val = (float)s.id[a];  // idtoid: no source point with this id, the area keeps its id as the value
    } else {
      float tmp[128];
      const int kk = k < 128 ? k : 128;  // (more than 128 intakes of one area: the first 128)
      for (int j = 0; j < kk; j++) {
        const float drf = m.f[F_DemandReturnFlowFraction] ? m.f[F_DemandReturnFlowFraction][s.in_pos[kb + j]] : 0.0f;
        tmp[j] = sws[s.in_rs[kb + j]] * (1.0f - drf);  // :3320
      }
      val = irri_np_mean(tmp, kk);
    }
    // numpy2pcr(Scalar, ., 0.0): a value of exactly 0 is missing, covered with 0 (:3326-3328)
    sv = (val == 0.0f) ? 0.0f : val / ((s.sqm[a] / 1000.0f) / (float)m.dt);
  }
  __syncthreads();
  const float out = sv;
  for (int p = s.seg[a] + threadIdx.x; p < s.seg[a + 1]; p += blockDim.x) m.f[F_IRSupplymm][s.perm[p]] = out;
}

// ---- paddy areas (wflow_sbm.py:3033-3050, 3330-3346; idtoid, wflow_lib.py:81-101) -----------------------------------------
// A paddy cell whose pond has fallen below h_min asks for the water that fills it to h_max, times the crop stage:
// irr_depth = ifthenelse(PondingDepth < h_min, h_max - PondingDepth, 0) * CRPST [mm], and the reference turns that into
// IrriDemandm3 = irr_depth / 1000 * areatotal(xl * yl, area) PER CELL (the area's whole surface, not the cell's -- kept).
// idtoid hands the intakes of the area np.mean of that map over the area's cells, times -1 / timestepsecs, as their Inflow;
// a mean of exactly 0 is a missing value there and leaves the intake alone.  After the supply iteration the cells that asked
// (IrriDemandm3 > 0) share what their intakes were given: IRSupplymm = mean(SurfaceWaterSupply of the intakes) * timestepsecs
// * 1000 / areatotal(xl * yl, the cells that asked), 0 everywhere else.  idtoid's left-overs as for the irrigation areas.
// Runs between the subsurface sweep (which ponds) and the overland / river sweep (which meets the demand).
__global__ void __launch_bounds__(256) k_paddy_demand(const DevModel m, const IrrigationSet s, float *obj_inflow) {
  __shared__ double sh[256];
  __shared__ float s_sqm;
  const int a = blockIdx.x;
  const int beg = s.seg[a], end = s.seg[a + 1];
  double sa = 0.0;
  for (int p = beg + threadIdx.x; p < end; p += blockDim.x) {
    const int i = s.perm[p];
    sa += (double)(m.f[F_xl][i] * m.f[F_yl][i]);
  }
  sa = irri_block_sum(sa, sh);
  if (threadIdx.x == 0) s_sqm = (float)sa;  // pcr.areatotal
  __syncthreads();
  const float sqm = s_sqm;
  double sm = 0.0, sq = 0.0;
  for (int p = beg + threadIdx.x; p < end; p += blockDim.x) {
    const int i = s.perm[p];
    const float pond = m.f[F_PondingDepth][i];
    const float crpst = m.f[F_CRPST] ? m.f[F_CRPST][i] : 1.0f;
    const float irr = ((pond < m.f[F_h_min][i]) ? (m.f[F_h_max][i] - pond) : 0.0f) * crpst;  // :3034-3039
    const float m3 = (irr / 1000.0f) * sqm;                                                  // :3041
    if (m.f[F_irr_depth]) m.f[F_irr_depth][i] = irr;
    m.f[F_IrriDemandm3][i] = m3;
    sm += (double)m3;
    if (m3 > 0.0f) sq += (double)(m.f[F_xl][i] * m.f[F_yl][i]);
  }
  sm = irri_block_sum(sm, sh);
  sq = irri_block_sum(sq, sh);
  if (threadIdx.x == 0) {
    s.sqm[a] = (float)sq;  // areatotal over the cells that asked (:3337-3342), for k_paddy_supply
    const int cnt = end - beg;
    float mean;
    if (cnt <= 128) {  // np.mean of float32 values in raster order (numpy's pairwise sum up to 128 values)
      float tmp[128];
      for (int j = 0; j < cnt; j++) tmp[j] = m.f[F_IrriDemandm3][s.perm[beg + j]];
      mean = irri_np_mean(tmp, cnt);
    } else {
      mean = (float)(sm / (double)cnt);  // (numpy adds blocks of 128 pairwise in float32; within its rounding of this)
    }
    const float scale = (float)(-1.0 / m.dt);
    const float v = (mean == 0.0f) ? 0.0f : mean * scale;  // exactly 0: missing in idtoid's result, the intake keeps its Inflow (none)
    for (int k = s.in_seg[a]; k < s.in_seg[a + 1]; k++) {
      const int rs = s.in_rs[k];
      m.f[F_Inflow][rs] = v;
      if (obj_inflow) obj_inflow[rs] = v;
    }
    if (a == 0)
      for (int k = 0; k < s.n_orphan; k++) {  // an intake without an area keeps idtoid's initial value: its own id
        const float v0 = (float)s.orphan_id[k] * scale;
        m.f[F_Inflow][s.orphan_rs[k]] = v0;
        if (obj_inflow) obj_inflow[s.orphan_rs[k]] = v0;
      }
  }
}

// after the supply iteration (IRSupplymm zeroed before the launch)
__global__ void __launch_bounds__(256) k_paddy_supply(const DevModel m, const IrrigationSet s, const float *sws) {
  __shared__ float sv;
  const int a = blockIdx.x;
  if (threadIdx.x == 0) {
    const int kb = s.in_seg[a], k = s.in_seg[a + 1] - kb;
    float val;
    if (k == 0) {
      val = (float)s.id[a];  // idtoid: no source point with this id, the area keeps its id as the value
    } else {
      float tmp[128];
      const int kk = k < 128 ? k : 128;
      for (int j = 0; j < kk; j++) tmp[j] = sws[s.in_rs[kb + j]];
      val = irri_np_mean(tmp, kk);
    }
    sv = (val == 0.0f) ? 0.0f : ((val * (float)m.dt) * 1000.0f) / s.sqm[a];  // :3344-3346
  }
  __syncthreads();
  const float out = sv;
  for (int p = s.seg[a] + threadIdx.x; p < s.seg[a + 1]; p += blockDim.x) {
    const int i = s.perm[p];
    if (m.f[F_IrriDemandm3][i] > 0.0f) m.f[F_IRSupplymm][i] = out;
  }
}

// ---- abstractions: Inflow < 0 (wflow_sbm.py:3131-3146, 3203-3308) ---------------------------------------
// The demand of a river cell is met from what reaches it: SurfaceWaterSupply = min(upstream(RiverRunoff) + Inwater, -Inflow).
// The first estimate is made inside the sweep (task_overland); the reference then repeats { new estimate from the routed
// discharge; kin_wave } until the estimate stops changing.  Its kin_wave calls of the loop are handed the UNCHANGED lateral
// inflow of the first call and a discharge array that the previous call of the loop has overwritten (:3242-3274, SURVEY.md
// appendix A.4): the first call of the loop repeats the sweep's result, every further one routes the river once more with
// the same lateral inflow.  Reproduced as such (k_supply_estimate + k_river_again), not "fixed".
struct SupplyArgs {
  float *sws;          // nr  SurfaceWaterSupply
  float *old_inwater;  // nr  Inwater before the abstraction (self.OldInwater)
  float *delta;        // 1   max |change of the estimate| (float bits, >= 0)
  int32_t *any_neg;    // 1   some Inflow < 0
};

__global__ void k_supply_estimate(const DevModel m, const SupplyArgs a) {
  const int rs = blockIdx.x * blockDim.x + threadIdx.x;
  if (rs >= m.nr) return;
  const float inflow = m.f[F_Inflow][rs];
  float sws = 0.0f;
  if (inflow < 0.0f) {
    // pcr.upstream(TopoLdd, RiverRunoff): the river cells that drain into this one (river order: contiguous run)
    float up = 0.0f;
    if (rs < m.nrnet) {
      const int nup = m.r_nup[rs];
      const uint32_t us = m.r_up_start[rs];
      for (int j = 0; j < nup; j++) up = up + m.f[F_RiverRunoff][us + j];
    }
    sws = fminf(up + a.old_inwater[rs], -1.0f * inflow);
  }
  const float d = fabsf(a.sws[rs] - sws);
  if (d > 0.0f) atomicMax(reinterpret_cast<int *>(a.delta), __float_as_int(d));
  a.sws[rs] = sws;
  if (m.f[F_Inwater]) m.f[F_Inwater][rs] = a.old_inwater[rs] + ((sws > 0.0f) ? -1.0f * sws : inflow);
}

// the river kinematic wave once more (all sub-steps) from the current discharge, lateral inflow as the sweep left it
template <int ML>
__global__ void __launch_bounds__(512, 1) k_river_again(const DevModel m) {
  for (int g = blockIdx.x; g < m.ngroups; g += gridDim.x) {
    const GroupDesc gd = m.groups[g];
    if (gd.rlev0 >= m.nrlev) continue;
    const uint32_t *ro = m.rlev_off + gd.rlev_idx;
    const int nrl = m.nrlev - gd.rlev0;
    for (int v = 0; v < m.itR; v++) {
      for (int q = 0; q < nrl; q++) {
        const uint32_t b = ro[q], e = ro[q + 1];
        for (uint32_t rs = b + threadIdx.x; rs < e; rs += blockDim.x) {
          RivOps riv;
          load_riv(m, (int32_t)rs, riv);
          task_river<WF_CTA>(m, (int32_t)rs, v, riv, 0 WF_PNONE);
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace wf
