// Field table of libwflowsbm: one row per raster the hot path reads or writes.
// Names are the reference's attribute names (wflow/wflow_sbm.py).
#pragma once
#include <cstdint>

namespace wf {

enum FieldKind : int { K_FORCING = 0, K_STATIC = 1, K_STATE = 2, K_DIAG = 3, K_RIVER = 0x100, K_HBV = 0x200 };

// X(name, kind)   -- K_RIVER marks arrays that exist only for river cells (river order),
//                    K_HBV the rasters of the wflow_hbv column (hbv.cuh; states created for wf_hbv_create models only)
#define WF_FIELDS(X)                                                                               \
  /* forcing (wf_updateparameters, wflow_sbm.py:2583) */                                           \
  X(P, K_FORCING) X(PET, K_FORCING) X(TEMP, K_FORCING) X(Inflow, K_FORCING | K_RIVER)              \
  /* statics of the PCRaster phases (wflow_sbm.py:2636-2819) */                                    \
  X(Cmax, K_STATIC) X(EoverR, K_STATIC) X(CanopyGapFraction, K_STATIC) X(et_RefToPot, K_STATIC)    \
  X(TempCor, K_STATIC) X(TTI, K_STATIC) X(TT, K_STATIC) X(TTM, K_STATIC) X(Cfmax, K_STATIC)        \
  X(WHC, K_STATIC) X(w_soil, K_STATIC) X(cf_soil, K_STATIC) X(WaterFrac, K_STATIC)                 \
  X(RiverFrac, K_STATIC | K_RIVER)                                                                 \
  /* statics of the soil column (static/layer dtypes, wflow_sbm.py:2176-2299) */                   \
  X(SoilThickness, K_STATIC) X(thetaS, K_STATIC) X(thetaR, K_STATIC) X(KsatVer, K_STATIC)          \
  X(f, K_STATIC) X(InfiltCapSoil, K_STATIC) X(InfiltCapPath, K_STATIC) X(PathFrac, K_STATIC)       \
  X(CapScale, K_STATIC) X(rootdistpar, K_STATIC) X(RootingDepth, K_STATIC)                         \
  X(AirEntryPressure, K_STATIC) X(MaxLeakage, K_STATIC) X(xl, K_STATIC) X(yl, K_STATIC)            \
  X(c_0, K_STATIC) X(c_1, K_STATIC) X(c_2, K_STATIC) X(c_3, K_STATIC)                              \
  X(c_4, K_STATIC) X(c_5, K_STATIC) X(c_6, K_STATIC) X(c_7, K_STATIC)                              \
  X(KsatVerFrac_0, K_STATIC) X(KsatVerFrac_1, K_STATIC) X(KsatVerFrac_2, K_STATIC)                 \
  X(KsatVerFrac_3, K_STATIC) X(KsatVerFrac_4, K_STATIC) X(KsatVerFrac_5, K_STATIC)                 \
  X(KsatVerFrac_6, K_STATIC) X(KsatVerFrac_7, K_STATIC)                                            \
  /* lateral statics */                                                                            \
  X(KsatHorFrac, K_STATIC) X(landSlope, K_STATIC) X(DL, K_STATIC) X(DW, K_STATIC)                  \
  X(SW, K_STATIC) X(AlphaL, K_STATIC)                                                              \
  X(AlphaR, K_STATIC | K_RIVER) X(DCL, K_STATIC | K_RIVER) X(Bw, K_STATIC | K_RIVER)               \
  /* glacier */                                                                                    \
  X(GlacierFrac, K_STATIC) X(G_TT, K_STATIC) X(G_Cfmax, K_STATIC) X(G_SIfrac, K_STATIC)            \
  /* reservoirs and lakes (wflow_sbm.py:1804-1907, 2822-2883; objects.cuh): rasters read at the objects' cells */   \
  X(filter_P_PET, K_STATIC) X(ResSimpleArea, K_STATIC) X(ResMaxVolume, K_STATIC) X(ResTargetFullFrac, K_STATIC)     \
  X(ResMaxRelease, K_STATIC) X(ResDemand, K_STATIC) X(ResTargetMinFrac, K_STATIC)                  \
  X(LakeArea, K_STATIC) X(LakeThreshold, K_STATIC) X(LakeStorFunc, K_STATIC) X(LakeOutflowFunc, K_STATIC)          \
  X(Lake_b, K_STATIC) X(Lake_e, K_STATIC) X(DemandReturnFlowFraction, K_STATIC)                    \
  /* dynamic vegetation (wflow_sbm.py:2587-2603): present when LAI is */                           \
  X(LAI, K_STATIC) X(Sl, K_STATIC) X(Swood, K_STATIC) X(Kext, K_STATIC)                            \
  /* states (stateVariables(), wflow_sbm.py:955-1010; zi is the live saturated-zone state) */      \
  X(CanopyStorage, K_STATE) X(Snow, K_STATE) X(SnowWater, K_STATE) X(TSoil, K_STATE)               \
  X(zi, K_STATE)                                                                                   \
  X(UStoreLayerDepth_0, K_STATE) X(UStoreLayerDepth_1, K_STATE) X(UStoreLayerDepth_2, K_STATE)     \
  X(UStoreLayerDepth_3, K_STATE) X(UStoreLayerDepth_4, K_STATE) X(UStoreLayerDepth_5, K_STATE)     \
  X(UStoreLayerDepth_6, K_STATE) X(UStoreLayerDepth_7, K_STATE)                                    \
  X(SubsurfaceFlow, K_STATE) X(LandRunoff, K_STATE) X(LandRunoffsub, K_STATE)                      \
  X(WaterLevelL, K_STATE) X(WaterLevelLsub, K_STATE)                                               \
  X(RiverRunoff, K_STATE | K_RIVER) X(RiverRunoffsub, K_STATE | K_RIVER)                           \
  X(WaterLevelR, K_STATE | K_RIVER) X(WaterLevelRsub, K_STATE | K_RIVER)                           \
  X(GlacierStore, K_STATE) X(ReservoirVolume, K_STATE) X(LakeWaterLevel, K_STATE)                  \
  /* irrigation (wflow_sbm.py:2034, 2755, 3314-3328): the supply of a step is extra water of the next one */ \
  X(IRSupplymm, K_STATE)                                                                           \
  /* paddy areas (wflow_sbm.py:762-785, 2812-2815, 3033-3050, 3330-3346; wf_paddy_create): the pond on the field, its bund   \
   * height and the depths that trigger and size the irrigation; CRPST is the crop-stage multiplier of the user's ini */  \
  X(PondingDepth, K_STATE) X(h_p, K_STATIC) X(h_max, K_STATIC) X(h_min, K_STATIC) X(CRPST, K_FORCING)                     \
  /* diagnostics, materialised only on request */                                                  \
  X(SatWaterDepth, K_DIAG) X(Precipitation, K_DIAG) X(PotenEvap, K_DIAG) X(Temperature, K_DIAG)    \
  X(ThroughFall, K_DIAG) X(StemFlow, K_DIAG) X(Interception, K_DIAG) X(PotTransSoil, K_DIAG)       \
  X(SnowMelt, K_DIAG) X(SnowFall, K_DIAG) X(RainFallPlusMelt, K_DIAG)                              \
  X(AvailableForInfiltration, K_DIAG) X(RunoffRiverCells, K_DIAG) X(RunoffLandCells, K_DIAG)       \
  X(ActEvapOpenWaterRiver, K_DIAG) X(ActEvapOpenWaterLand, K_DIAG) X(RestEvap, K_DIAG)             \
  X(PotSoilEvap, K_DIAG) X(PotTrans, K_DIAG) X(UStoreDepth, K_DIAG) X(CapFlux, K_DIAG)             \
  X(Transfer, K_DIAG) X(ActEvapUStore, K_DIAG) X(ActEvapSat, K_DIAG) X(Transpiration, K_DIAG)      \
  X(soilevap, K_DIAG) X(soilevapsat, K_DIAG) X(ActEvap, K_DIAG) X(Recharge, K_DIAG)                \
  X(InwaterMM, K_DIAG) X(qo_toriver, K_DIAG) X(ssf_toriver, K_DIAG) X(Inwater, K_DIAG | K_RIVER)   \
  X(ExfiltWater, K_DIAG) X(InfiltExcess, K_DIAG) X(ExcessWater, K_DIAG) X(ActInfilt, K_DIAG)       \
  X(ActLeakage, K_DIAG) X(SoilInf, K_DIAG) X(PathInf, K_DIAG) X(InfiltExcessSoil, K_DIAG)          \
  X(InfiltExcessPath, K_DIAG) X(ActInfiltSoil, K_DIAG) X(ActInfiltPath, K_DIAG)                    \
  X(ExfiltFromUstore, K_DIAG) X(ExfiltFromSat, K_DIAG) X(InwaterL, K_DIAG)                         \
  X(InflowKinWaveCellLand, K_DIAG) X(CellInFlow, K_DIAG) X(SoilWatbal, K_DIAG)                     \
  X(InterceptionWatBal, K_DIAG) X(SurfaceWatbal, K_DIAG) X(watbal, K_DIAG)                         \
  X(RootStore_unsat, K_DIAG)                                                                       \
  X(vwc_0, K_DIAG) X(vwc_1, K_DIAG) X(vwc_2, K_DIAG) X(vwc_3, K_DIAG)                              \
  X(vwc_4, K_DIAG) X(vwc_5, K_DIAG) X(vwc_6, K_DIAG) X(vwc_7, K_DIAG)                              \
  X(Snow2Glacier, K_DIAG) X(GlacierMelt, K_DIAG)                                                   \
  /* end of dynamic(), wflow_sbm.py:3310-3312, 3348-3370, 3429-3526 (outputs.cuh; the Cum* maps accumulate) */ \
  X(KinWaveVolumeR, K_DIAG) X(KinWaveVolumeL, K_DIAG) X(OldKinWaveVolumeR, K_DIAG)                 \
  X(OldKinWaveVolumeL, K_DIAG) X(MassBalKinWaveR, K_DIAG) X(MassBalKinWaveL, K_DIAG)               \
  X(InflowKinWaveCell, K_DIAG) X(RiverRunoffMM, K_DIAG) X(LandRunoffMM, K_DIAG)                    \
  X(QCatchmentMM, K_DIAG) X(RunoffCoeff, K_DIAG) X(CellStorage, K_DIAG) X(DeltaStorage, K_DIAG)    \
  X(SatWaterFlux, K_DIAG) X(sumUstore, K_DIAG) X(SnowCover, K_DIAG) X(RootStore, K_DIAG)           \
  X(RootStore_sat, K_DIAG) X(vwcRoot, K_DIAG) X(vwc_percRoot, K_DIAG)                              \
  X(vwc_perc_0, K_DIAG) X(vwc_perc_1, K_DIAG) X(vwc_perc_2, K_DIAG) X(vwc_perc_3, K_DIAG)          \
  X(vwc_perc_4, K_DIAG) X(vwc_perc_5, K_DIAG) X(vwc_perc_6, K_DIAG) X(vwc_perc_7, K_DIAG)          \
  X(SumEvapSat, K_DIAG) X(SumEvapUstore, K_DIAG) X(ExfiltWaterCubic, K_DIAG)                       \
  X(InfiltExcessCubic, K_DIAG) X(SurfaceWaterSupply, K_DIAG)                                       \
  X(CumOutFlow, K_DIAG) X(CumActInfilt, K_DIAG) X(CumCellInFlow, K_DIAG) X(CumPrec, K_DIAG)        \
  X(CumEvap, K_DIAG) X(CumPotenEvap, K_DIAG) X(CumInt, K_DIAG) X(CumLeakage, K_DIAG)               \
  X(CumExfiltWater, K_DIAG) X(CumSurfaceWater, K_DIAG)                                             \
  X(OutflowSR, K_DIAG) X(ResPercFull, K_DIAG) X(ResPrecipSR, K_DIAG) X(ResEvapSR, K_DIAG)          \
  X(DemandRelease, K_DIAG) X(LakeOutflow, K_DIAG) X(LakeVolume, K_DIAG) X(LakePrecip, K_DIAG)      \
  X(LakeEvap, K_DIAG) X(IrriDemand, K_DIAG) X(IrriDemandm3, K_DIAG) X(ActEvapPond, K_DIAG) X(irr_depth, K_DIAG)      \
  /* wflow_hbv (wflow/wflow_hbv.py:560-900 parameters, :131-163 states, :1222-1767 dynamic(); hbv.cuh).  Shared with the   \
   * rows above: P PET TEMP Inflow, TempCor TTI TT Cfmax WHC xl yl, the glacier rows, DCL Bw and AlphaR (= Alpha), the       \
   * channel states RiverRunoff(sub) / WaterLevelR(sub) (= SurfaceRunoff(sub) / WaterLevel(sub)) and some outputs */        \
  X(Seepage, K_FORCING | K_HBV)                                                                    \
  X(Pcorr, K_STATIC | K_HBV) X(SFCF, K_STATIC | K_HBV) X(RFCF, K_STATIC | K_HBV) X(ICF, K_STATIC | K_HBV)                  \
  X(EPF, K_STATIC | K_HBV) X(ECORR, K_STATIC | K_HBV) X(CEVPF, K_STATIC | K_HBV) X(CFR, K_STATIC | K_HBV)                  \
  X(FieldCapacity, K_STATIC | K_HBV) X(Treshold, K_STATIC | K_HBV) X(BetaSeepage, K_STATIC | K_HBV)                        \
  X(PERC, K_STATIC | K_HBV) X(Cflux, K_STATIC | K_HBV) X(KQuickFlow, K_STATIC | K_HBV) X(AlphaNL, K_STATIC | K_HBV)        \
  X(K4, K_STATIC | K_HBV) X(K0, K_STATIC | K_HBV) X(SUZ, K_STATIC | K_HBV) X(Slope, K_STATIC | K_HBV)                                                 \
  X(FreeWater, K_STATE | K_HBV) X(DrySnow, K_STATE | K_HBV) X(SoilMoisture, K_STATE | K_HBV)                               \
  X(UpperZoneStorage, K_STATE | K_HBV) X(LowerZoneStorage, K_STATE | K_HBV) X(InterceptionStorage, K_STATE | K_HBV)        \
  X(PotEvaporation, K_DIAG | K_HBV) X(IntEvap, K_DIAG | K_HBV) X(SoilEvap, K_DIAG | K_HBV) X(HBVSeepage, K_DIAG | K_HBV)   \
  X(DirectRunoff, K_DIAG | K_HBV) X(InUpperZone, K_DIAG | K_HBV) X(Percolation, K_DIAG | K_HBV)                            \
  X(QuickFlow, K_DIAG | K_HBV) X(RealQuickFlow, K_DIAG | K_HBV) X(BaseFlow, K_DIAG | K_HBV) X(InSoil, K_DIAG | K_HBV)      \
  X(NetInSoil, K_DIAG | K_HBV) X(RainAndSnowmelt, K_DIAG | K_HBV) X(QuickFlowCubic, K_DIAG | K_HBV)                        \
  X(BaseFlowCubic, K_DIAG | K_HBV)                                                                                        \
  /* water balance of wflow_hbv (:1208-1215, 1731-1766): the sums accumulate from the start of the run */                  \
  X(initstorage, K_STATIC | K_HBV) X(storage, K_DIAG | K_HBV) X(sumprecip, K_DIAG | K_HBV) X(sumevap, K_DIAG | K_HBV)     \
  X(sumrunoff, K_DIAG | K_HBV) X(sumlevel, K_DIAG | K_HBV) X(sumpotevap, K_DIAG | K_HBV) X(sumsoilevap, K_DIAG | K_HBV)   \
  X(sumtemp, K_DIAG | K_HBV) X(suminflow, K_DIAG | K_HBV) X(SurfaceRunoffMM, K_DIAG | K_HBV)

enum FieldId : int {
#define X(n, k) F_##n,
  WF_FIELDS(X)
#undef X
      F_COUNT
};

struct FieldInfo {
  const char *name;
  int kind;
};

extern const FieldInfo g_fields[F_COUNT];
int field_index(const char *name);

}  // namespace wf
