// Outputs of the END of WflowModel.dynamic() (wflow/wflow_sbm.py:3310-3312, 3348-3370, 3429-3526; updateKinWaveVolume
// :946-953): kinematic-wave volumes and mass balances, unit conversions, the runoff coefficient over the upstream area,
// storage change, root-zone water content, cumulative sums.  REAL4 map algebra on the float32 maps the step has just
// produced -- nothing of it feeds the next step, so it is materialised only when one of these outputs is requested
// (wf_request_output): one small kernel before the step (what the outputs need from BEFORE it), one after, and for the
// runoff coefficient one level sweep (pcr.catchmenttotal of RainFallPlusMelt, :3465-3470).
#pragma once
#include "kernels.cuh"

namespace wf {

struct EndArgs {
  float *preWLRsub;  // nr   WaterLevelRsub before the step (OldKinWaveVolumeR)
  float *preWLR;     // nr   WaterLevelR before the step (CumSurfaceWater, :3058)
  float *preWLLsub;  // n    WaterLevelLsub before the step
  float *preOrg;     // n    OrgStorage (:2627-2629)
  float *ctR;        // n    catchmenttotal(RainFallPlusMelt)
  const float *ct1;  // n    catchmenttotal(1): upstream cell count incl. the cell itself (static)
  const float *sws;  // nr   SurfaceWaterSupply of the step (abstractions enabled), else nullptr
  int32_t satwd_valid;  // the SatWaterDepth map holds the end of the previous step (else: derived from zi, like resume())
};

#define WF_MVF (-999.0f)

template <int ML>
__global__ void k_end_pre(const DevModel m, const EndArgs e) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m.nr) {
    e.preWLRsub[i] = m.f[F_WaterLevelRsub][i];
    e.preWLR[i] = m.f[F_WaterLevelR][i];
  }
  if (i >= m.n) return;
  e.preWLLsub[i] = m.f[F_WaterLevelLsub][i];
  float su = 0.0f;  // sum_list_cover, wflow_lib.py:65-78
#pragma unroll
  for (int k = 0; k < ML; k++) su = su + m.f[F_UStoreLayerDepth_0 + k][i];
  const float sat = (e.satwd_valid && m.f[F_SatWaterDepth])
                        ? m.f[F_SatWaterDepth][i]
                        : (LD(thetaS) - LD(thetaR)) * (LD(SoilThickness) - m.f[F_zi][i]);  // :2503-2505 / :2751
  e.preOrg[i] = su + sat;
}

// pcr.catchmenttotal(values, ldd) on the device: one block per group of the storage layout, level after level (the cells of a
// level only read the totals of the level before: contiguous runs, network.h).  float32 sums like the map algebra.
__global__ void k_catchment_total(const DevModel m, const float *__restrict__ values, float *__restrict__ acc) {
  for (int g = blockIdx.x; g < m.ngroups; g += gridDim.x) {
    const GroupDesc gd = m.groups[g];
    const uint32_t *lo = m.lev_off + gd.lev_idx;
    const int nl = m.nlev - gd.lev0;
    for (int l = 0; l < nl; l++) {
      const uint32_t b = lo[l], e = lo[l + 1];
      for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        float v = values[i];
        const int nup = m.topo[i] >> 4;
        const uint32_t us = m.up_start[i];
        for (int j = 0; j < nup; j++) v = v + acc[us + j];
        acc[i] = v;
      }
      __syncthreads();
    }
  }
}

template <int ML>
__global__ void k_end_outputs(const DevModel m, const EndArgs e) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.n) return;
  const float dtf = (float)m.dt;
  const float xl = LD(xl), yl = LD(yl), thS = LD(thetaS), thR = LD(thetaR);
  const float ToCubic = ((xl * yl) * 0.001f) / dtf;  // :1998
  const float QMMConv = dtf / ((xl * yl) * 0.001f);  // :1985
  const int32_t rs = m.river_slot[i];
  const bool riv = rs >= 0;
  const float RR = riv ? m.f[F_RiverRunoff][rs] : 0.0f, LR = m.f[F_LandRunoff][i];
  // kinematic-wave volumes, :946-953
  float KWVR = 0.0f, OldKWVR = 0.0f, Inwater = 0.0f;
  if (riv) {
    const float Bw = m.f[F_Bw][rs], DCL = m.f[F_DCL][rs];
    KWVR = (m.f[F_WaterLevelRsub][rs] * Bw) * DCL;
    OldKWVR = (e.preWLRsub[rs] * Bw) * DCL;
    if (m.f[F_Inwater]) Inwater = m.f[F_Inwater][rs];
  }
  const float KWVL = (m.f[F_WaterLevelLsub][i] * LD(SW)) * LD(DL);
  const float OldKWVL = (e.preWLLsub[i] * LD(SW)) * LD(DL);
  OUTF(KinWaveVolumeR, KWVR); OUTF(KinWaveVolumeL, KWVL); OUTF(OldKinWaveVolumeR, OldKWVR); OUTF(OldKinWaveVolumeL, OldKWVL);
  // pcr.upstream(TopoLdd, RiverRunoff), :3349: the discharge of the cells that drain into this one, neighbour order of _up_nb
  float inflowKW = 0.0f;
  if (m.f[F_InflowKinWaveCell] || m.f[F_MassBalKinWaveR]) {
    const int nup = m.topo[i] >> 4;
    const uint32_t us = m.up_start[i];
    for (int j = 0; j < nup; j++) {
      const int32_t r = m.river_slot[us + j];
      if (r >= 0) inflowKW = inflowKW + m.f[F_RiverRunoff][r];
    }
  }
  OUTF(InflowKinWaveCell, inflowKW);
  OUTF(MassBalKinWaveR, (((((-KWVR) + OldKWVR) / dtf) + inflowKW) + Inwater) - RR);  // :3351-3356
  if (m.f[F_MassBalKinWaveL])                                                        // :3365-3370
    m.f[F_MassBalKinWaveL][i] = (((((-KWVL) + OldKWVL) / dtf) + m.f[F_InflowKinWaveCellLand][i]) + m.f[F_InwaterL][i]) - LR;
  OUTF(RiverRunoffMM, RR * QMMConv); OUTF(LandRunoffMM, LR * QMMConv);               // :3305-3312
  if (m.f[F_QCatchmentMM] || m.f[F_RunoffCoeff]) {
    const float ct1 = e.ct1[i];
    const float temp = (((ct1 * xl) * 0.001f) * 0.001f) * yl;  // :1989-1997
    const float QMMConvUp = (temp != 0.0f) ? (float)(m.dt * 0.001) / temp : WF_MVF;  // x / 0 is a missing value in PCRaster
    const float QC = (QMMConvUp != WF_MVF) ? RR * QMMConvUp : WF_MVF;                // :3465
    OUTF(QCatchmentMM, QC);
    if (m.f[F_RunoffCoeff]) m.f[F_RunoffCoeff][i] = (QC != WF_MVF && e.ctR[i] != 0.0f) ? (QC / e.ctR[i]) / ct1 : WF_MVF;  // :3466-3470
  }
  float su = 0.0f;
#pragma unroll
  for (int k = 0; k < ML; k++) su = su + m.f[F_UStoreLayerDepth_0 + k][i];
  const float zi = m.f[F_zi][i];
  const float Sat = m.f[F_SatWaterDepth] ? m.f[F_SatWaterDepth][i] : (thS - thR) * (LD(SoilThickness) - zi);
  const float CellStorage = su + Sat;  // :3473-3475
  OUTF(sumUstore, su); OUTF(CellStorage, CellStorage); OUTF(DeltaStorage, CellStorage - e.preOrg[i]);
  const float SatWaterFlux = m.f[F_SubsurfaceFlow][i] / (((xl * 1000.0f) * yl) * 1000.0f);  // :3480
  OUTF(SatWaterFlux, SatWaterFlux);
  if (m.f[F_SnowCover]) m.f[F_SnowCover][i] = (m.f[F_Snow] && m.f[F_Snow][i] > 0.0f) ? 1.0f : 0.0f;  // :3499-3501
  const float ActRoot = fminf(LD(SoilThickness) * 0.99f, LD(RootingDepth));  // :2783
  const float RSsat = fmaxf(0.0f, ActRoot - zi) * (thS - thR);               // :3431-3433
  OUTF(RootStore_sat, RSsat);
  if (m.f[F_RootStore] || m.f[F_vwcRoot] || m.f[F_vwc_percRoot]) {
    const float RS = RSsat + m.f[F_RootStore_unsat][i];  // :3456
    const float vwcRoot = (ActRoot != 0.0f) ? (RS / ActRoot) + thR : WF_MVF;
    OUTF(RootStore, RS); OUTF(vwcRoot, vwcRoot);
    OUTF(vwc_percRoot, (vwcRoot != WF_MVF && thS != 0.0f) ? (vwcRoot / thS) * 100.0f : WF_MVF);
  }
  if (m.f[F_ExfiltWaterCubic]) m.f[F_ExfiltWaterCubic][i] = m.f[F_ExfiltWater][i] * ToCubic;  // :3128-3129
  if (m.f[F_InfiltExcessCubic]) m.f[F_InfiltExcessCubic][i] = m.f[F_InfiltExcess][i] * ToCubic;
  OUTF(SurfaceWaterSupply, (e.sws && riv) ? e.sws[rs] : 0.0f);  // :3138-3140, 3217-3221
  // cumulative sums, :3487-3504 and :3058
#define WF_CUM(dst, add)                                    \
  do {                                                      \
    if (m.f[F_##dst]) m.f[F_##dst][i] = m.f[F_##dst][i] + (add); \
  } while (0)
  WF_CUM(CumOutFlow, SatWaterFlux);
  WF_CUM(CumActInfilt, m.f[F_ActInfilt][i]);
  WF_CUM(CumCellInFlow, m.f[F_CellInFlow][i]);
  WF_CUM(CumPrec, m.f[F_Precipitation][i]);
  WF_CUM(CumEvap, m.f[F_ActEvap][i]);
  WF_CUM(CumPotenEvap, m.f[F_PotenEvap][i]);
  WF_CUM(CumInt, m.f[F_Interception][i]);
  WF_CUM(CumLeakage, m.f[F_ActLeakage][i]);
  WF_CUM(CumExfiltWater, m.f[F_ExfiltWater][i]);
  WF_CUM(CumSurfaceWater, riv ? e.preWLR[rs] / 1000.0f : 0.0f);
#undef WF_CUM
}

}  // namespace wf
