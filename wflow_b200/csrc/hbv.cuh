// wflow_hbv on the device: the HBV-96 column of one timestep, ahead of the kinematic wave.
//
// Reference: WflowModel.dynamic(), wflow/wflow_hbv.py:1222-1767.  Everything up to the lateral inflow of the channel
// (:1264-1612) is cell-local REAL4 map algebra: rain / snow split, interception, snow pack with refreezing and liquid
// water, glacierHBV (wflow/wflow_funcs.py:549-603), soil moisture with direct runoff and seepage, upper zone (percolation,
// capillary flux, quick flow), lower zone (base flow).  It runs here as one thread per cell in float32, operation by
// operation in the reference's order (the library is compiled without FMA contraction); pcr.exp and ** are evaluated in
// float64 and rounded to REAL4 once, like the PCRaster stand-in the goldens were made with (oracle/pcrshim.py).
// The kinematic wave that follows (kin_wave, wflow/wflow_funcs.py:155-207, called on the FULL network with every cell a
// channel, wflow_hbv.py:1633-1647) is the river task of the sweep in lateral.cuh: a wf_hbv_create model is a model whose
// river network is its whole network and whose sweep runs the river task only; this kernel hands it the lateral inflow
// (h_inwater, river order) where the SBM model's overland task does.
//
// Snow mass wasting (:1351-1364) is one more level sweep between the column and the channel wave (k_hbv_mass_wasting).
// Not restated (refused for these models): reservoirs / lakes (:1533-1594), state updating from observed discharge (:1684-1726).
#pragma once
#include "kernels.cuh"
#include "lateral.cuh"

namespace wf {

struct HbvOptions {
  int32_t mass_wasting;    // [model] MassWasting: the snow pack slides downslope between the column and the channel wave (:1351-1364)
  int32_t set_kquick;      // [model] SetKquickFlow: K0 / SUZ quick flow instead of the non-linear upper zone (:1466-1500)
  int32_t external_qbase;  // [model] ExternalQbase: seepage of an external model instead of the lower zone's base flow (:1271, :1512)
};

#define HF(name) (__ldg(m.f[F_##name] + i))
#define HOUT(name, v)                            \
  do {                                           \
    if (m.f[F_##name]) m.f[F_##name][i] = (v);   \
  } while (0)

__device__ __forceinline__ float hbv_pow(float x, float y) { return (float)pow((double)x, (double)y); }

#ifndef WF_HBV_MINBLOCKS
#define WF_HBV_MINBLOCKS 8  // resident blocks of 256 threads per SM (32 registers: the loads of 37 rasters wait on DRAM, occupancy pays for the spills: 1.78 ms at 4 blocks, 1.48 at 6, 1.38 at 8)
#endif
__global__ void __launch_bounds__(256, WF_HBV_MINBLOCKS) k_hbv_column(const DevModel m, const HbvOptions opt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.n) return;
  const int32_t rs = __ldg(m.river_slot + i);
  const float dtf = (float)m.dt;
  // ---- forcing, rain / snow split (:1264-1300) ----
  float Precipitation = fmaxf(0.0f, HF(P)) * HF(Pcorr);
  float PotEvaporation = HF(PET);
  const float Temperature = (m.f[F_TEMP] ? HF(TEMP) : 10.0f) + HF(TempCor);
  const float TT = HF(TT), TTI = HF(TTI);
  float RainFrac;
  if (1.0f * TTI == 0.0f) RainFrac = (Temperature <= TT) ? 0.0f : 1.0f;
  else RainFrac = fminf((Temperature - (TT - TTI / 2.0f)) / TTI, 1.0f);
  RainFrac = fmaxf(RainFrac, 0.0f);
  const float SnowFrac = 1.0f - RainFrac;
  Precipitation = HF(SFCF) * SnowFrac * Precipitation + HF(RFCF) * RainFrac * Precipitation;
  // ---- interception (:1302-1322) ----
  float IS = m.f[F_InterceptionStorage][i];
  const float Interception = fminf(Precipitation, HF(ICF) - IS);
  IS = IS + Interception;
  Precipitation = Precipitation - Interception;
  PotEvaporation = (float)exp((double)(-HF(EPF) * Precipitation)) * HF(ECORR) * PotEvaporation;
  PotEvaporation = HF(CEVPF) * PotEvaporation;
  const float IntEvap = fminf(IS, PotEvaporation);
  IS = IS - IntEvap;
  const float RestEvap = fmaxf(0.0f, PotEvaporation - IntEvap);
  // ---- snow pack (:1332-1362) ----
  const float SnowFall = SnowFrac * Precipitation;
  const float RainFall = RainFrac * Precipitation;
  const float Cfmax = HF(Cfmax);
  const float PotSnowMelt = (Temperature > TT) ? Cfmax * (Temperature - TT) : 0.0f;
  const float PotRefreezing = (Temperature < TT) ? Cfmax * HF(CFR) * (TT - Temperature) : 0.0f;
  float FreeWater = m.f[F_FreeWater][i];
  float DrySnow = m.f[F_DrySnow][i];
  const float Refreezing = (Temperature < TT) ? fminf(PotRefreezing, FreeWater) : 0.0f;
  const float SnowMelt = fminf(PotSnowMelt, DrySnow);
  DrySnow = DrySnow + SnowFall + Refreezing - SnowMelt;
  FreeWater = FreeWater - Refreezing;
  const float MaxFreeWater = DrySnow * HF(WHC);
  FreeWater = FreeWater + SnowMelt + RainFall;
  const float InSoil = fmaxf(FreeWater - MaxFreeWater, 0.0f);
  FreeWater = FreeWater - InSoil;
  const float RainAndSnowmelt = RainFall + SnowMelt;
  const float SnowCover = (DrySnow > 0.0f) ? 1.0f : 0.0f;
  // ---- soil moisture: direct runoff (:1370-1378) ----
  float NetInSoil = InSoil;
  const float FC = HF(FieldCapacity);
  float SM = m.f[F_SoilMoisture][i] + NetInSoil;
  float DirectRunoff = fmaxf(SM - FC, 0.0f);
  SM = SM - DirectRunoff;
  NetInSoil = NetInSoil - DirectRunoff;
  // ---- glacier (:1378-1404; glacierHBV, wflow_funcs.py:549-603) ----
  if (m.glacier) {
    const float ts = (float)(m.dt / m.basetimestep);
    const float gf = HF(GlacierFrac), gtt = HF(G_TT);
    float S2G = HF(G_SIfrac) * ts * DrySnow;
    S2G = (gf > 0.0f) ? S2G : 0.0f;
    S2G = fminf(S2G, (float)(8.0 * (m.dt / m.basetimestep)));
    DrySnow = DrySnow - (S2G * gf);
    float GS = m.f[F_GlacierStore][i] + S2G;
    const float PotMelt = (Temperature > gtt) ? HF(G_Cfmax) * ts * (Temperature - gtt) : 0.0f;
    float GM = (DrySnow < 10.0f) ? fminf(PotMelt, GS) : 0.0f;
    GS = GS - GM;
    GM = GM * gf;  // :1402
    FreeWater = FreeWater + GM;
    m.f[F_GlacierStore][i] = GS;
    HOUT(Snow2Glacier, S2G);
    HOUT(GlacierMelt, GM);
  }
  // ---- soil evaporation, seepage (:1412-1442) ----
  const float Tr = HF(Treshold);
  const float SoilEvap = (SM > Tr) ? fminf(SM, RestEvap) : fminf(SM, fminf(RestEvap, PotEvaporation * (SM / Tr)));
  SM = SM - SoilEvap;
  const float ActEvap = IntEvap + SoilEvap;
  const float HBVSeepage = hbv_pow(fminf(SM / FC, 1.0f), HF(BetaSeepage)) * NetInSoil;
  SM = SM - HBVSeepage;
  const float Backtosoil = fminf(FC - SM, DirectRunoff);
  DirectRunoff = DirectRunoff - Backtosoil;
  SM = SM + Backtosoil;
  const float InUpperZone = DirectRunoff + HBVSeepage;
  // ---- upper zone (:1448-1506) ----
  const float PERC = HF(PERC);
  float UZ = m.f[F_UpperZoneStorage][i] + InUpperZone;
  const float Percolation = fminf(PERC, UZ - InUpperZone / 2.0f);
  UZ = UZ - Percolation;
  float CapFlux = HF(Cflux) * ((FC - SM) / FC);
  CapFlux = fminf(UZ, CapFlux);
  CapFlux = fminf(FC - SM, CapFlux);
  UZ = UZ - CapFlux;
  SM = SM + CapFlux;
  float QuickFlow, RealQuickFlow;
  if (!opt.set_kquick) {
    const bool low = Percolation < PERC;
    const float qf = low ? 0.0f : HF(KQuickFlow) * hbv_pow(UZ - fminf(InUpperZone / 2.0f, UZ), 1.0f + HF(AlphaNL));
    QuickFlow = fminf(qf, UZ);
    UZ = fmaxf(low ? UZ : UZ - QuickFlow, 0.0f);
    RealQuickFlow = 0.0f;
  } else {
    QuickFlow = HF(KQuickFlow) * UZ;
    RealQuickFlow = fmaxf(0.0f, HF(K0) * (UZ - HF(SUZ)));
    UZ = UZ - QuickFlow - RealQuickFlow;
  }
  // ---- lower zone, runoff generation (:1508-1528) ----
  float LZ = m.f[F_LowerZoneStorage][i] + Percolation;
  const float BaseFlow = fminf(LZ, HF(K4) * LZ);
  LZ = LZ - BaseFlow;
  float DirectRunoffStorage;
  if (opt.external_qbase) DirectRunoffStorage = QuickFlow + (m.f[F_Seepage] ? HF(Seepage) : 0.0f) + RealQuickFlow;
  else DirectRunoffStorage = QuickFlow + BaseFlow + RealQuickFlow;
  const float InwaterMM = fmaxf(0.0f, DirectRunoffStorage);
  const float ToCubic = ((HF(xl) * HF(yl)) * 0.001f) / dtf;  // :1014
  float Inwater = InwaterMM * ToCubic;
  // ---- sources and demands (:1599-1606), lateral inflow of the channel (:1612) ----
  const float Inflow = (m.f[F_Inflow] && rs >= 0) ? __ldg(m.f[F_Inflow] + rs) : 0.0f;
  const float SurfaceRunoff = (rs >= 0) ? m.f[F_RiverRunoff][rs] : 0.0f;
  const float SWS = (Inflow < 0.0f) ? fmaxf(-1.0f * Inwater, SurfaceRunoff) : 0.0f;
  Inwater = Inwater + ((SWS > 0.0f) ? -1.0f * SWS : Inflow);
  if (rs >= 0) {
    m.h_inwater[rs] = Inwater;
    if (m.f[F_Inwater]) m.f[F_Inwater][rs] = Inwater;
  }
  // ---- states ----
  m.f[F_InterceptionStorage][i] = IS;
  m.f[F_DrySnow][i] = DrySnow;
  m.f[F_FreeWater][i] = FreeWater;
  m.f[F_SoilMoisture][i] = SM;
  m.f[F_UpperZoneStorage][i] = UZ;
  m.f[F_LowerZoneStorage][i] = LZ;
  // ---- outputs on request ----
  HOUT(Precipitation, Precipitation); HOUT(PotEvaporation, PotEvaporation); HOUT(Temperature, Temperature);
  HOUT(IntEvap, IntEvap); HOUT(SnowMelt, SnowMelt); HOUT(SnowCover, SnowCover); HOUT(SoilEvap, SoilEvap);
  HOUT(ActEvap, ActEvap); HOUT(HBVSeepage, HBVSeepage); HOUT(DirectRunoff, DirectRunoff); HOUT(InUpperZone, InUpperZone);
  HOUT(Percolation, Percolation); HOUT(CapFlux, CapFlux); HOUT(QuickFlow, QuickFlow); HOUT(RealQuickFlow, RealQuickFlow);
  HOUT(BaseFlow, BaseFlow); HOUT(InSoil, InSoil); HOUT(RainAndSnowmelt, RainAndSnowmelt); HOUT(NetInSoil, NetInSoil);
  HOUT(InwaterMM, InwaterMM); HOUT(QuickFlowCubic, (QuickFlow + RealQuickFlow) * ToCubic); HOUT(BaseFlowCubic, BaseFlow * ToCubic);
  HOUT(SurfaceWaterSupply, SWS);
  // ---- water balance (:1731-1766): storage of the cell, sums since the start of the run ----
  const float storage = (((FreeWater + DrySnow) + SM) + UZ) + LZ;
  HOUT(storage, storage);
#define HSUM(name, v)                                          \
  do {                                                         \
    if (m.f[F_##name]) m.f[F_##name][i] = m.f[F_##name][i] + (v); \
  } while (0)
  HSUM(sumprecip, Precipitation); HSUM(sumevap, ActEvap); HSUM(sumsoilevap, SoilEvap); HSUM(sumpotevap, PotEvaporation);
  HSUM(sumtemp, Temperature); HSUM(sumrunoff, InwaterMM); HSUM(suminflow, Inflow);
#undef HSUM
  if (m.f[F_watbal])  // wf_request_output made the four maps it is formed from
    m.f[F_watbal][i] = (((m.f[F_initstorage][i] - storage) + m.f[F_sumprecip][i]) - m.f[F_sumsoilevap][i]) - m.f[F_sumrunoff][i];
}

// Snow mass wasting (wflow_hbv.py:1351-1364): DrySnow = accucapacitystate(TopoLdd, DrySnow, SnowFluxFrac * DrySnow) and the
// same for the liquid water in the pack, SnowFluxFrac = min(0.5, Slope / 5.67) * min(1, DrySnow / 10000) formed from the pack
// BEFORE it slides.  One more sweep over the levels, both stores at once: one block per group, a block barrier per level,
// float32 like the map algebra (the SBM model's k_mass_wasting, objects.cuh, on wflow_hbv's rasters).  Nothing of the column
// reads the pack after this point except the glacier module (refused together) and the water balance (k_hbv_post).
__global__ void k_hbv_mass_wasting(const DevModel m, float *__restrict__ flux, float *__restrict__ flux_w) {
  for (int g = blockIdx.x; g < m.ngroups; g += gridDim.x) {
    const GroupDesc gd = m.groups[g];
    const uint32_t *lo = m.lev_off + gd.lev_idx;
    const int nl = m.nlev - gd.lev0;
    for (int l = 0; l < nl; l++) {
      const uint32_t b = lo[l], e = lo[l + 1];
      for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        const float snow = m.f[F_DrySnow][i], fw = m.f[F_FreeWater][i];
        const float frac = fminf(0.5f, m.f[F_Slope][i] / 5.67f) * fminf(1.0f, snow / 10000.0f);  // 5.67 = tan 80 degrees
        float tot = snow, totw = fw;
        const int nup = m.topo[i] >> 4;
        const uint32_t us = m.up_start[i];
        for (int j = 0; j < nup; j++) {
          tot = tot + flux[us + j];
          totw = totw + flux_w[us + j];
        }
        const float out = fminf(tot, frac * snow), outw = fminf(totw, frac * fw);
        flux[i] = out;
        flux_w[i] = outw;
        m.f[F_DrySnow][i] = tot - out;
        m.f[F_FreeWater][i] = totw - outw;
      }
      __syncthreads();
    }
  }
}

// behind the sweep: what the water balance and the outputs read of the routed channel (:1650-1668, 1745)
__global__ void __launch_bounds__(256) k_hbv_post(const DevModel m, const HbvOptions opt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.n) return;
  if (opt.mass_wasting && (m.f[F_storage] || m.f[F_watbal])) {  // the pack has moved since the column formed these
    const float storage = (((m.f[F_FreeWater][i] + m.f[F_DrySnow][i]) + m.f[F_SoilMoisture][i]) + m.f[F_UpperZoneStorage][i]) + m.f[F_LowerZoneStorage][i];
    if (m.f[F_storage]) m.f[F_storage][i] = storage;
    if (m.f[F_watbal])
      m.f[F_watbal][i] = (((m.f[F_initstorage][i] - storage) + m.f[F_sumprecip][i]) - m.f[F_sumsoilevap][i]) - m.f[F_sumrunoff][i];
  }
  const int32_t rs = __ldg(m.river_slot + i);
  if (rs < 0) return;
  if (m.f[F_sumlevel]) m.f[F_sumlevel][i] = m.f[F_sumlevel][i] + m.f[F_WaterLevelR][rs];
  if (m.f[F_SurfaceRunoffMM]) {
    const float QMMConv = (float)m.dt / ((m.f[F_xl][i] * m.f[F_yl][i]) * 0.001f);  // :1013
    m.f[F_SurfaceRunoffMM][i] = m.f[F_RiverRunoff][rs] * QMMConv;
  }
}

#undef HF
#undef HOUT

// ---- the channel wave of every cell as a sweep of its own ------------------------------------------------------------
// kin_wave on the full network (wflow_hbv.py:1633-1647; wflow/wflow_funcs.py:155-207).  The shared sweep of lateral.cuh
// can run it (river task only), but its blocks are sized for the subsurface task: 640 threads at 96 registers, 2560
// threads per cluster of four for the ~4 x 1024 channel items a tick of the bench shard holds -- two rounds per tick.  The
// channel task alone needs far fewer registers, so this kernel runs one cluster per group with 1024 threads per block
// (64 registers): a tick's items fit one round.  Same schedule as the river task of the shared sweep, re-based: at tick t
// sub-step v solves level t - v of the group, so a cell's sub-step v follows its own sub-step v - 1 and its upstream
// neighbours' sub-step v by exactly one tick, one cluster barrier per tick, levels + it - 1 ticks.  task_river is the
// shared device function: same arithmetic, same hand-off arrays, bit-identical results (tests/test_gpu_hbv.py).
#define WF_CHANNEL_THREADS 1024
__global__ void __launch_bounds__(WF_CHANNEL_THREADS, 1) k_channel_wave(const __grid_constant__ DevModel m) {
  extern __shared__ __align__(16) uint32_t s_ro[];
#ifdef WF_PROFILE
  unsigned long long *pbase = nullptr;
  const unsigned long long c0 = 0;
#endif
  cg::cluster_group cl = cg::this_cluster();
  const int crank = (int)cl.block_rank(), csize = (int)cl.num_blocks();
  const int nclusters = (int)gridDim.x / csize, cid = (int)blockIdx.x / csize;
  // item warps dealt round robin to the blocks of the cluster (every SM gets the same mix of sub-steps)
  const int tid = (((int)threadIdx.x >> 5) * csize + crank) * 32 + ((int)threadIdx.x & 31);
  const int nthreads = csize * (int)blockDim.x;
  const int itR = m.itR;
  for (int g = cid; g < m.ngroups; g += nclusters) {
    const GroupDesc gd = m.groups[g];
    const int nrl = (gd.rlev0 < m.nrlev) ? m.nrlev - gd.rlev0 : 0;
    const uint32_t *ro = m.rlev_off + gd.rlev_idx;
    if (nrl + 1 <= m.lev_cache) {  // the group's level offsets into shared memory
      __syncthreads();
      for (int k = threadIdx.x; k <= nrl; k += blockDim.x) s_ro[k] = ro[k];
      __syncthreads();
      ro = s_ro;
    }
    const int nticks = (nrl > 0) ? nrl + itR - 1 : 0;
    for (int t = 0; t < nticks; t++) {
      // items of the tick: [sub-step 0: level t | sub-step 1: level t - 1 | ...]
      const int v0 = max(0, t - (nrl - 1)), v1 = min(itR - 1, t);
      int j = tid;
      for (int v = v0; v <= v1; v++) {
        const int q = t - v;
        const uint32_t rb = ro[q];
        const int rc = (int)(ro[q + 1] - rb);
        while (j < rc) {
          const int32_t rs = (int32_t)(rb + (uint32_t)j);
          RivOps o;
          load_riv(m, rs, o);
          task_river<WF_CLUSTER>(m, rs, v, o, 0 WF_PPASS(2));
          j += nthreads;
        }
        j -= rc;
      }
      // (warming L2 with the level that enters the wave next tick measured slower, 7.04 against 6.70 ms: with a tick's items
      //  filling every thread the kernel is bound by the solves' instructions, not by the cold operands of a quarter of them)
      __syncwarp();
      if (t + 1 < nticks) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      }
    }
    if (g + nclusters < m.ngroups) {  // the next group re-uses the shared-memory offsets
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  }
}

}  // namespace wf
