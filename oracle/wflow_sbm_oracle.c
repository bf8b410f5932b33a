/*
 * wflow_sbm_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's per-timestep hot path of wflow_sbm, used
 * ONLY as the checker for the CUDA path (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).  Nothing under wflow_b200/ may
 * link, import or call it.
 *
 * Parity status
 *   * numba half (set_dd, sbm_cell, kinematic_wave, kinematic_wave_ssf, kin_wave):
 *     PINNED -- validated element-wise against the reference's own numba kernels run
 *     in the build container (oracle/gen_golden.py -> tests/golden/*.npz,
 *     tests/test_oracle_vs_reference.py).
 *   * PCRaster half (interception, snow, glacier, LAI-driven canopy parameters, evap split, river
 *     inflow glue of WflowModel.dynamic): PINNED -- the reference's whole, UNMODIFIED dynamic() is
 *     executed in the build container with `pcraster` = the float32-numpy shim of oracle/pcrshim.py
 *     (oracle/ref_dynamic.py), and every state and output of this file is bit-identical to it on nine
 *     trajectories that put the edge branches on the grid (oracle/gen_golden_dynamic.py ->
 *     tests/golden/dyn_*.npz, tests/test_oracle_dynamic_golden.py; live re-check in
 *     tests/test_oracle_vs_reference.py).  What the shim cannot verify about PCRaster itself is listed
 *     in its header (float32 libm vs. rounded float64 for ln / exp / pow).
 *     The first run of that pin found one deviation, fixed since: kin_wave receives a REAL4 array and
 *     writes into it, so the river outflow is rounded to float32 BETWEEN sub-steps (see below).
 *     The reference's acceptance numbers are reproduced as well (tests/test_reference_cases.py).
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference checkout).  Written from the behaviour of that code, not copied from it.
 *
 * Precision contract (reference behaviour): PCRaster maps are REAL4, so every state is
 * float32 at step boundaries and the map-algebra phases round to float after each
 * operation; sbm_cell / kin_wave compute in float64 on float32-valued inputs.
 * Compile with -O2 -ffp-contract=off (no FMA contraction, no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define WFO_MAXL 8

/* ------------------------------------------------------------------------------------
 * Field table.  All fields are full-grid float32 rasters (nrow*ncol, row-major); the
 * Python wrapper passes a pointer table indexed by this enum (NULL = not supplied /
 * not requested).  Names are the reference's attribute names.
 * ---------------------------------------------------------------------------------- */
#define WFO_FIELDS(X) \
  /* forcing */ X(P) X(PET) X(TEMP) X(Inflow) \
  /* static, PCRaster phases */ \
  X(Cmax) X(EoverR) X(CanopyGapFraction) X(et_RefToPot) X(TempCor) X(TTI) X(TT) X(TTM) \
  X(Cfmax) X(WHC) X(w_soil) X(cf_soil) X(RiverFrac) X(WaterFrac) \
  /* static, soil column */ \
  X(SoilThickness) X(thetaS) X(thetaR) X(KsatVer) X(f) X(InfiltCapSoil) X(InfiltCapPath) \
  X(PathFrac) X(CapScale) X(rootdistpar) X(RootingDepth) X(AirEntryPressure) X(MaxLeakage) \
  X(xl) X(yl) \
  X(c_0) X(c_1) X(c_2) X(c_3) X(c_4) X(c_5) X(c_6) X(c_7) \
  X(KsatVerFrac_0) X(KsatVerFrac_1) X(KsatVerFrac_2) X(KsatVerFrac_3) \
  X(KsatVerFrac_4) X(KsatVerFrac_5) X(KsatVerFrac_6) X(KsatVerFrac_7) \
  /* static, lateral */ \
  X(KsatHorFrac) X(landSlope) X(DL) X(DW) X(SW) X(AlphaL) X(Beta) X(River) X(AlphaR) X(DCL) X(Bw) \
  /* glacier (optional) */ X(GlacierFrac) X(G_TT) X(G_Cfmax) X(G_SIfrac) \
  /* dynamic vegetation (optional: present when LAI is) */ X(LAI) X(Sl) X(Swood) X(Kext) \
  /* states (in/out) */ \
  X(CanopyStorage) X(Snow) X(SnowWater) X(TSoil) X(zi) X(SatWaterDepth) \
  X(UStoreLayerDepth_0) X(UStoreLayerDepth_1) X(UStoreLayerDepth_2) X(UStoreLayerDepth_3) \
  X(UStoreLayerDepth_4) X(UStoreLayerDepth_5) X(UStoreLayerDepth_6) X(UStoreLayerDepth_7) \
  X(SubsurfaceFlow) X(LandRunoff) X(LandRunoffsub) X(WaterLevelL) X(WaterLevelLsub) \
  X(RiverRunoff) X(RiverRunoffsub) X(WaterLevelR) X(WaterLevelRsub) X(GlacierStore) \
  /* diagnostics (out, optional) */ \
  X(Precipitation) X(PotenEvap) X(Temperature) X(ThroughFall) X(StemFlow) X(Interception) \
  X(PotTransSoil) X(SnowMelt) X(SnowFall) X(RainFallPlusMelt) X(AvailableForInfiltration) \
  X(RunoffRiverCells) X(RunoffLandCells) X(ActEvapOpenWaterRiver) X(ActEvapOpenWaterLand) \
  X(RestEvap) X(PotSoilEvap) X(PotTrans) X(UStoreDepth) X(CapFlux) X(Transfer) \
  X(ActEvapUStore) X(ActEvapSat) X(Transpiration) X(soilevap) X(soilevapsat) X(ActEvap) \
  X(Recharge) X(InwaterMM) X(qo_toriver) X(ssf_toriver) X(Inwater) X(ExfiltWater) \
  X(InfiltExcess) X(ExcessWater) X(ActInfilt) X(ActLeakage) X(SoilInf) X(PathInf) \
  X(InfiltExcessSoil) X(InfiltExcessPath) X(ActInfiltSoil) X(ActInfiltPath) \
  X(ExfiltFromUstore) X(ExfiltFromSat) X(InwaterL) X(InflowKinWaveCellLand) X(CellInFlow) \
  X(SoilWatbal) X(InterceptionWatBal) X(SurfaceWatbal) X(watbal) X(RootStore_unsat) \
  X(vwc_0) X(vwc_1) X(vwc_2) X(vwc_3) X(vwc_4) X(vwc_5) X(vwc_6) X(vwc_7) \
  X(Snow2Glacier) X(GlacierMelt) \
  /* end of dynamic(), wflow_sbm.py:3310-3312, 3348-3370, 3429-3526 (out, optional; the Cum* maps accumulate in place) */ \
  X(KinWaveVolumeR) X(KinWaveVolumeL) X(OldKinWaveVolumeR) X(OldKinWaveVolumeL) X(MassBalKinWaveR) X(MassBalKinWaveL) \
  X(InflowKinWaveCell) X(RiverRunoffMM) X(LandRunoffMM) X(QCatchmentMM) X(RunoffCoeff) X(CellStorage) X(DeltaStorage) \
  X(SatWaterFlux) X(sumUstore) X(SnowCover) X(RootStore) X(RootStore_sat) X(vwcRoot) X(vwc_percRoot) \
  X(vwc_perc_0) X(vwc_perc_1) X(vwc_perc_2) X(vwc_perc_3) X(vwc_perc_4) X(vwc_perc_5) X(vwc_perc_6) X(vwc_perc_7) \
  X(SumEvapSat) X(SumEvapUstore) X(ExfiltWaterCubic) X(InfiltExcessCubic) X(SurfaceWaterSupply) \
  X(CumOutFlow) X(CumActInfilt) X(CumCellInFlow) X(CumPrec) X(CumEvap) X(CumPotenEvap) X(CumInt) X(CumLeakage) \
  X(CumExfiltWater) X(CumSurfaceWater)

enum {
#define X(n) WFO_##n,
  WFO_FIELDS(X)
#undef X
      WFO_NFIELDS
};

static const char *wfo_names[] = {
#define X(n) #n,
    WFO_FIELDS(X)
#undef X
};

/* Threads for wfo_step (cells of one level are independent, so the level loops are
 * parallelised with OpenMP for the CPU-baseline timing; results do not depend on it). */
static int g_threads = 1;
void wfo_set_threads(int n) { g_threads = n > 0 ? n : 1; }
int wfo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int wfo_has_openmp(void) {
#ifdef _OPENMP
  return 1;
#else
  return 0;
#endif
}

int wfo_nfields(void) { return WFO_NFIELDS; }
const char *wfo_field_name(int i) { return (i >= 0 && i < WFO_NFIELDS) ? wfo_names[i] : 0; }

typedef struct {
  int32_t nrow, ncol;
  int32_t maxLayers;      /* wflow_sbm.py:1757-1766 */
  int32_t nrLayers;
  int32_t modelSnow, soilInfRedu, transferMethod, ust;
  int32_t it_kinL, it_kinR;
  int32_t glacier;        /* hasattr(self,"GlacierFrac") */
  int32_t reserved;
  double timestepsecs, basetimestep;
  double layerThickness[WFO_MAXL]; /* [model] UStoreLayerThickness entries */
} wfo_params;

/* network in the reference's list-of-levels form, flattened */
typedef struct {
  int64_t nlev;
  int64_t ncells;
  int64_t *lev_off;  /* nlev+1 */
  int64_t *nodes;    /* ncells, flat grid index */
  int64_t *nodes_up; /* ncells*8, -1 padded */
} wfo_network;

/* ------------------------------------------------------------------------------------
 * set_dd  (wflow/wflow_funcs.py:43-95)
 * BFS from the pits upstream; result reversed so level 0 is the farthest from its pit.
 * Keeps the reference quirk at :88 (np.any(idx_up) is a VALUE test: an upstream set
 * consisting only of flat index 0 is dropped).
 * ---------------------------------------------------------------------------------- */
static int up_nb(const uint8_t *ldd, int64_t idx0, int nrow, int ncol, int64_t out[8]) {
  /* wflow_funcs.py:48-67; _ldd_us = flipped [[7,8,9],[4,5,6],[1,2,3]] = [[3,2,1],[6,5,4],[9,8,7]] */
  static const uint8_t us[3][3] = {{3, 2, 1}, {6, 5, 4}, {9, 8, 7}};
  int64_t r = idx0 / ncol, c = idx0 % ncol;
  int n = 0;
  for (int dr = -1; dr <= 1; dr++) {
    int64_t row = r + dr;
    for (int dc = -1; dc <= 1; dc++) {
      int64_t col = c + dc;
      if (dr == 0 && dc == 0) continue;
      if (row < 0 || row >= nrow || col < 0 || col >= ncol) continue;
      int64_t idx = idx0 + dc + (int64_t)dr * ncol;
      if (ldd[idx] == us[dr + 1][dc + 1]) out[n++] = idx;
    }
  }
  return n;
}

int wfo_set_dd(const uint8_t *ldd, int nrow, int ncol, wfo_network *net) {
  int64_t N = (int64_t)nrow * ncol;
  int64_t cap = 0;
  for (int64_t i = 0; i < N; i++)
    if (ldd[i] >= 1 && ldd[i] <= 9) cap++;
  /* downstream-first storage, reversed at the end */
  int64_t *nodes = (int64_t *)malloc(sizeof(int64_t) * (cap + 1));
  int64_t *ups = (int64_t *)malloc(sizeof(int64_t) * 8 * (cap + 1));
  int64_t *loff = (int64_t *)malloc(sizeof(int64_t) * (cap + 2));
  int64_t nn = 0, nlev = 0;
  loff[0] = 0;
  for (int64_t i = 0; i < N; i++)
    if (ldd[i] == 5) nodes[nn++] = i; /* :79 */
  while (1) {
    int64_t beg = loff[nlev], end = nn;
    loff[++nlev] = end;
    int64_t nnext = 0;
    for (int64_t k = beg; k < end; k++) {
      int64_t up[8];
      int n = up_nb(ldd, nodes[k], nrow, ncol, up);
      int any = 0;
      for (int j = 0; j < n; j++) any |= (up[j] != 0);
      for (int j = 0; j < 8; j++) ups[k * 8 + j] = -1;
      if (any) { /* :88 */
        for (int j = 0; j < n; j++) {
          ups[k * 8 + j] = up[j];
          if (nn + nnext >= cap + 1) { free(nodes); free(ups); free(loff); return -1; } /* cycle guard */
          nodes[nn + nnext++] = up[j];
        }
      }
    }
    if (nnext == 0) break;
    nn += nnext;
  }
  net->nlev = nlev;
  net->ncells = nn;
  net->lev_off = (int64_t *)malloc(sizeof(int64_t) * (nlev + 1));
  net->nodes = (int64_t *)malloc(sizeof(int64_t) * (nn > 0 ? nn : 1));
  net->nodes_up = (int64_t *)malloc(sizeof(int64_t) * 8 * (nn > 0 ? nn : 1));
  int64_t o = 0;
  for (int64_t l = nlev - 1; l >= 0; l--) { /* :95 reversed */
    net->lev_off[nlev - 1 - l] = o;
    int64_t w = loff[l + 1] - loff[l];
    memcpy(net->nodes + o, nodes + loff[l], sizeof(int64_t) * w);
    memcpy(net->nodes_up + 8 * o, ups + 8 * loff[l], sizeof(int64_t) * 8 * w);
    o += w;
  }
  net->lev_off[nlev] = o;
  free(nodes); free(ups); free(loff);
  return 0;
}

void wfo_free_network(wfo_network *net) {
  free(net->lev_off); free(net->nodes); free(net->nodes_up);
  net->lev_off = net->nodes = net->nodes_up = 0;
}

/* ------------------------------------------------------------------------------------
 * kinematic_wave  (wflow/wflow_funcs.py:119-152)
 * ---------------------------------------------------------------------------------- */
double wfo_kinematic_wave(double Qin, double Qold, double q, double alpha, double beta,
                          double deltaT, double deltaX) {
  const double epsilon = 1e-12;
  const int MAX_ITERS = 3000;
  if ((Qin + Qold + q) == 0.0) return 0.0;
  double ab_pQ = alpha * beta * pow(((Qold + Qin) / 2.0), beta - 1.0);
  double deltaTX = deltaT / deltaX;
  double C = deltaTX * Qin + alpha * pow(Qold, beta) + deltaT * q;
  double Qkx = (deltaTX * Qin + Qold * ab_pQ + deltaT * q) / (deltaTX + ab_pQ);
  if (isnan(Qkx)) Qkx = 0.0;
  Qkx = fmax(Qkx, 1e-30);
  double fQkx = deltaTX * Qkx + alpha * pow(Qkx, beta) - C;
  double dfQkx = deltaTX + alpha * beta * pow(Qkx, beta - 1.0);
  Qkx = Qkx - fQkx / dfQkx;
  Qkx = fmax(Qkx, 1e-30);
  int count = 0;
  while (fabs(fQkx) > epsilon && count < MAX_ITERS) {
    fQkx = deltaTX * Qkx + alpha * pow(Qkx, beta) - C;
    dfQkx = deltaTX + alpha * beta * pow(Qkx, beta - 1.0);
    Qkx = Qkx - fQkx / dfQkx;
    Qkx = fmax(Qkx, 1e-30);
    count++;
  }
  return Qkx;
}

/* python max(a, b): returns b only if b > a (NaN-propagation like the reference) */
static inline double pymax(double a, double b) { return (b > a) ? b : a; }
static inline double pymin(double a, double b) { return (b < a) ? b : a; }

/* ------------------------------------------------------------------------------------
 * kinematic_wave_ssf  (wflow/wflow_funcs.py:210-292)
 * ---------------------------------------------------------------------------------- */
void wfo_kinematic_wave_ssf(double ssf_in, double ssf_old, double zi_old, double r,
                            double Ks_hor, double Ks, double slope, double neff, double f,
                            double D, double dt, double dx, double w, double ssf_max,
                            double *ssf_out, double *zi_out, double *exfilt_out) {
  const double epsilon = 1e-3;
  const int MAX_ITERS = 3000;
  if ((ssf_in + ssf_old) == 0.0 && (r <= 0)) {
    *ssf_out = 0.0; *zi_out = D; *exfilt_out = 0.0;
    return;
  }
  double ssf_n = (ssf_old + ssf_in) / 2.0;
  int count = 0;
  double zi = log((f * ssf_n) / (w * Ks_hor * Ks * slope) + exp(-f * D)) / -f;
  double Cn = (Ks_hor * Ks * exp(-f * zi) * slope) / neff;
  double c = (dt / dx) * ssf_in + (1.0 / Cn) * ssf_n + dt * (r - (zi_old - zi) * neff * w);
  double fQ = (dt / dx) * ssf_n + (1.0 / Cn) * ssf_n - c;
  double dfQ = (dt / dx) + 1 / Cn;
  ssf_n = ssf_n - (fQ / dfQ);
  if (isnan(ssf_n)) ssf_n = 0.0;
  ssf_n = pymax(ssf_n, 1e-30);
  while (fabs(fQ) > epsilon && count < MAX_ITERS) {
    zi = log((f * ssf_n) / (w * Ks_hor * Ks * slope) + exp(-f * D)) / -f;
    Cn = (Ks_hor * Ks * exp(-f * zi) * slope) / neff;
    c = (dt / dx) * ssf_in + (1.0 / Cn) * ssf_n + dt * (r - (zi_old - zi) * neff * w);
    fQ = (dt / dx) * ssf_n + (1.0 / Cn) * ssf_n - c;
    dfQ = (dt / dx) + 1.0 / Cn;
    ssf_n = ssf_n - (fQ / dfQ);
    if (isnan(ssf_n)) ssf_n = 0.0;
    ssf_n = pymax(ssf_n, 1e-30);
    count++;
  }
  ssf_n = pymin(ssf_n, (ssf_max * w));
  double rest = zi_old - (ssf_in + r * dx - ssf_n) / (w * dx) / neff;
  *exfilt_out = pymin(rest, 0.0) * -neff;
  *zi_out = pymax(0.0, zi);
  *ssf_out = ssf_n;
}

/* _sCurve (wflow_sbm.py:106-123) */
static inline double sCurve(double X, double a, double b, double c) {
  return 1.0 / (b + exp(-c * (X - a)));
}

/* infiltration (wflow_sbm.py:212-250) */
static void infiltration(double Avail, double PathFrac, double cf_soil, double TSoil,
                         double InfiltCapSoil, double InfiltCapPath, double UStoreCapacity,
                         int modelSnow, int soilInfReduction, double *InfiltSoilPath,
                         double *InfiltSoil, double *InfiltPath, double *SoilInf, double *PathInf,
                         double *InfiltExcess) {
  *SoilInf = Avail * (1 - PathFrac);
  *PathInf = Avail * PathFrac;
  double soilInfRedu = 1.0;
  if (modelSnow & soilInfReduction) {
    double bb = 1.0 / (1.0 - cf_soil);
    soilInfRedu = sCurve(TSoil, 0.0, bb, 8.0) + cf_soil;
  }
  double MaxInfiltSoil = pymin(InfiltCapSoil * soilInfRedu, *SoilInf);
  double MaxInfiltPath = pymin(InfiltCapPath * soilInfRedu, *PathInf);
  double cap0 = pymax(0.0, UStoreCapacity);
  *InfiltSoilPath = pymin(MaxInfiltPath + MaxInfiltSoil, cap0);
  if ((MaxInfiltPath + MaxInfiltSoil) > 0.0) {
    *InfiltSoil = MaxInfiltSoil * pymin(1.0, cap0 / (MaxInfiltPath + MaxInfiltSoil));
    *InfiltPath = MaxInfiltPath * pymin(1.0, cap0 / (MaxInfiltPath + MaxInfiltSoil));
  } else {
    *InfiltSoil = 0.0;
    *InfiltPath = 0.0;
  }
  *InfiltExcess = (*SoilInf - MaxInfiltSoil) + (*PathInf - MaxInfiltPath);
}

/* unsatzone_flow_layer (wflow_sbm.py:253-272) */
static void unsatzone_flow_layer(double *USd_io, double KsatVerFrac, double KsatVer, double f,
                                 double z, double L_sat, double c, double *sum_ast_out) {
  double USd = *USd_io;
  double sum_ast = 0;
  double st = KsatVerFrac * KsatVer * exp(-f * z) * pymin(pow(USd / L_sat, c), 1.0);
  double st_sat = pymax(0, USd - L_sat);
  double m = pymin(st, st_sat);
  USd = USd - m;
  sum_ast = sum_ast + m;
  double ast = pymin(st - m, USd);
  double itsd = ceil(ast / (L_sat * 0.02));
  long its = (long)itsd;
  if (its < 1) its = 1;
  double k = KsatVerFrac * KsatVer / (double)its * exp(-f * z);
  for (long i = 0; i < its; i++) {
    st = k * pymin(pow(USd / L_sat, c), 1.0);
    ast = pymin(st, USd);
    USd = USd - ast;
    sum_ast = sum_ast + ast;
  }
  *USd_io = USd;
  *sum_ast_out = sum_ast;
}

/* actTransp_unsat_SBM (wflow_sbm.py:126-209) */
static void actTransp_unsat_SBM(double RootingDepth, double *USd_io, double sumLayer,
                                double *RestPotEvap_io, double *sumActEvapUStore_io, double c,
                                double L, double thetaS, double thetaR, double hb, int ust) {
  double USd = *USd_io;
  double AvailCap;
  if (ust >= 1) AvailCap = USd * 0.99;
  else if (L > 0) AvailCap = pymin(1.0, pymax(0.0, (RootingDepth - sumLayer) / L));
  else AvailCap = 0.0;
  double MaxExtr = AvailCap * USd;
  double h1 = hb, h2 = 100, h3 = 400, h4 = 15849;
  double par_lambda = 2 / (c - 3);
  double vwc = (L > 0.0) ? USd / L : 0.0;
  vwc = pymax(vwc, 0.0000001);
  double head = hb / pow(vwc / (thetaS - thetaR), 1 / par_lambda);
  head = pymax(head, hb);
  double alpha;
  if (head <= h1) alpha = 1;
  else if (head >= h4) alpha = 0;
  else if ((head < h2) & (head > h1)) alpha = 1;
  else if ((head > h3) & (head < h4)) alpha = 1 - (head - h3) / (h4 - h3);
  else alpha = 1;
  double mn = pymin(pymin(MaxExtr, *RestPotEvap_io), USd); /* python min(a,b,c) */
  double ActEvapUStore = mn * alpha;
  *USd_io = USd - ActEvapUStore;
  *RestPotEvap_io = *RestPotEvap_io - ActEvapUStore;
  *sumActEvapUStore_io = ActEvapUStore + *sumActEvapUStore_io;
}

/* ------------------------------------------------------------------------------------
 * One full timestep = WflowModel.dynamic() (wflow/wflow_sbm.py:2518-3528) for the
 * configuration space covered by the build: no reservoirs/lakes, no irrigation/paddy,
 * no state updating, no LAI, Inflow >= 0 (no abstraction iteration), MassWasting off.
 * fields[]: pointer table (see WFO_FIELDS); states are updated in place.
 * `active` marks cells that are not MV in the clone (PCRaster phases run there);
 * sbm_cell / kin_wave visit the cells of net / rnet.
 * Returns 0, or a negative code for unsupported input.
 * ---------------------------------------------------------------------------------- */
#define F(name) (fields[WFO_##name])
#define OUT(name, i, v) do { if (fields[WFO_##name]) fields[WFO_##name][i] = (float)(v); } while (0)

int wfo_step(const wfo_params *p, const wfo_network *net, const wfo_network *rnet,
             const uint8_t *ldd, const uint8_t *active, float **fields) {
  const int64_t N = (int64_t)p->nrow * p->ncol;
  const int ML = p->maxLayers;
  if (ML < 1 || ML > WFO_MAXL) return -2;
  const float dtf = (float)p->timestepsecs;
  const int gash = p->timestepsecs >= (23 * 3600); /* wflow_sbm.py:2639 */

  float *Uf[WFO_MAXL];
  const float *cf[WFO_MAXL], *kvf[WFO_MAXL];
  for (int k = 0; k < ML; k++) {
    Uf[k] = fields[WFO_UStoreLayerDepth_0 + k];
    cf[k] = fields[WFO_c_0 + k];
    kvf[k] = fields[WFO_KsatVerFrac_0 + k];
    if (!Uf[k] || !cf[k] || !kvf[k]) return -3;
  }

  /* work arrays (float64 = the reference's dyn/layer structured arrays) */
  double *W = (double *)calloc((size_t)N * 24 + 8, sizeof(double));
  double *Ul = (double *)calloc((size_t)N * ML, sizeof(double));
  if (!W || !Ul) { free(W); free(Ul); return -4; }
  double *ssf_new = W;               /* N+1 (sentinel) */
  double *qo_new = W + (N + 1);      /* N+1 */
  double *AvailInf = W + 2 * (N + 1);
  double *PotSoilEvapD = AvailInf + N, *PotTransD = PotSoilEvapD + N, *RunoffLandD = PotTransD + N,
         *AEOWLandD = RunoffLandD + N, *ziD = AEOWLandD + N, *SatD = ziD + N, *InwaterO = SatD + N,
         *acc_flow = InwaterO + N, *acc_h = acc_flow + N, *qo_toriver_acc = acc_h + N,
         *Qo_in = qo_toriver_acc + N, *ssf_toriverD = Qo_in + N, *LandRunoffsubD = ssf_toriverD + N,
         *WLLsubD = LandRunoffsubD + N, *InwaterMMf = WLLsubD + N, *RainPlusMeltD = InwaterMMf + N,
         *OldCanopy = RainPlusMeltD + N, *RunoffRiverD = OldCanopy + N;

  /* values of BEFORE the step that the end-of-dynamic() outputs need (only when one of them is requested) */
  int want_end = 0;
  for (int k = WFO_KinWaveVolumeR; k <= WFO_CumSurfaceWater; k++) want_end |= fields[k] != 0;
  float *pre = want_end ? (float *)calloc((size_t)N * 4, sizeof(float)) : 0;
  if (want_end && !pre) { free(W); free(Ul); return -4; }
  float *preWLR = pre, *preWLRsub = pre ? pre + N : 0, *preWLLsub = pre ? pre + 2 * N : 0, *preOrg = pre ? pre + 3 * N : 0;

  /* ---------------- P0-P3: PCRaster REAL4 phases (wflow_sbm.py:2583-2819) ---------------- */
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int64_t i = 0; i < N; i++) {
    if (!active[i]) continue;
    if (pre) {
      preWLR[i] = F(WaterLevelR)[i]; preWLRsub[i] = F(WaterLevelRsub)[i]; preWLLsub[i] = F(WaterLevelLsub)[i];
      float su = 0.0f;                                                /* sum_list_cover, wflow_lib.py:65-78 */
      for (int k = 0; k < ML; k++) su = su + Uf[k][i];
      /* OrgStorage :2627-2629: the SatWaterDepth MAP of the end of the previous step (or of resume()) */
      preOrg[i] = su + (F(SatWaterDepth) ? F(SatWaterDepth)[i]
                                         : (F(thetaS)[i] - F(thetaR)[i]) * (F(SoilThickness)[i] - F(zi)[i]));
    }
    float Precipitation = fmaxf(0.0f, F(P)[i]);                       /* :2584 */
    float PotenEvap = F(PET)[i] * F(et_RefToPot)[i];                  /* :2615 */
    float Temperature = F(TEMP) ? F(TEMP)[i] : 0.0f;
    if (p->modelSnow) Temperature = Temperature + F(TempCor)[i];      /* :2617 */
    float CanopyStorage = F(CanopyStorage)[i];
    OldCanopy[i] = CanopyStorage;
    float PotEvap = PotenEvap;
    float CGF = F(CanopyGapFraction)[i], Cmax = F(Cmax)[i];
    float EoverR = F(EoverR) ? F(EoverR)[i] : 0.0f;
    if (F(LAI)) { /* LAI-driven canopy, wflow_sbm.py:2587-2603 (PotenEvap is still the raw forcing there) */
      const float LAI = F(LAI)[i], Kext = F(Kext)[i];
      Cmax = F(Sl)[i] * LAI + F(Swood)[i];                            /* :2590 */
      CGF = (float)exp((double)((-Kext) * LAI));                      /* :2591 */
      const float Ewet = (1.0f - CGF) * F(PET)[i];                    /* :2595 */
      EoverR = (Precipitation > 0.0f) ? fminf(0.25f, Ewet / fmaxf(0.0001f, Precipitation)) : 0.0f; /* :2596-2603 */
    }
    float ThroughFall, Interception, StemFlow, PotTransSoil;
    if (gash) {
      /* rainfall_interception_gash, wflow_funcs.py:321-373 */
      float pt = 0.1f * CGF;
      float one_m = (1.0f - CGF) - pt;
      float P_sat = 0.0f;
      if (EoverR != 0.0f && one_m != 0.0f) { /* PCRaster: x/0 -> MV -> cover -> 0 */
        float arg = 1.0f - (EoverR / one_m);
        if (arg > 0.0f) {                    /* ln(x<=0) -> MV -> cover -> 0 */
          float lnv = (float)log((double)arg);
          P_sat = ((-Cmax) / EoverR) * lnv;
        }
      }
      P_sat = fmaxf(0.0f, P_sat);
      int large = Precipitation > P_sat;
      float Iwet = large ? (one_m * P_sat) - Cmax : Precipitation * one_m;
      float Isat = large ? EoverR * (Precipitation - P_sat) : 0.0f;
      float Idry = large ? Cmax : 0.0f;
      float Itrunc = 0.0f;
      StemFlow = pt * Precipitation;
      ThroughFall = ((((Precipitation - Iwet) - Idry) - Isat) - Itrunc) - StemFlow;
      Interception = ((Iwet + Idry) + Isat) + Itrunc;
      if (Cmax <= 0.0f) { ThroughFall = Precipitation; Interception = 0.0f; StemFlow = 0.0f; }
      float OverEstimate = (Interception > PotEvap) ? Interception - PotEvap : 0.0f;
      Interception = fminf(Interception, PotEvap);
      ThroughFall = ThroughFall + OverEstimate;
      PotTransSoil = fmaxf(0.0f, PotEvap - Interception);             /* :2654 */
    } else {
      /* rainfall_interception_modrut, wflow_funcs.py:376-429 */
      float pp = CGF, pt = 0.1f * pp;
      float Pfrac = fmaxf(((1.0f - pp) - pt), 0.0f) * Precipitation;
      float DD = (CanopyStorage > Cmax) ? CanopyStorage - Cmax : 0.0f;
      CanopyStorage = CanopyStorage - DD;
      CanopyStorage = CanopyStorage + Pfrac;
      float dC = -1.0f * fminf(CanopyStorage, PotEvap);
      CanopyStorage = CanopyStorage + dC;
      float LeftOver = PotEvap + dC;
      float D = (CanopyStorage > Cmax) ? CanopyStorage - Cmax : 0.0f;
      CanopyStorage = CanopyStorage - D;
      ThroughFall = (DD + D) + pp * Precipitation;
      StemFlow = Precipitation * pt;
      float NetInterception = (Precipitation - ThroughFall) - StemFlow;
      PotTransSoil = fmaxf(0.0f, LeftOver);                           /* :2673 */
      Interception = NetInterception;                                 /* :2674 */
    }
    float EffectivePrecipitation = ThroughFall + StemFlow;            /* :2676 */
    float RainFallPlusMelt, SnowMelt = 0.0f, SnowFall = 0.0f;
    float TSoil = F(TSoil) ? F(TSoil)[i] : 0.0f;
    if (p->modelSnow) {
      /* :2678-2698 and SnowPackHBV, wflow_funcs.py:490-546 */
      TSoil = TSoil + F(w_soil)[i] * (Temperature - TSoil);
      float Snow = F(Snow)[i], SnowWater = F(SnowWater)[i];
      float TTI = F(TTI)[i], TT = F(TT)[i], TTM = F(TTM)[i], Cfmax = F(Cfmax)[i], WHC = F(WHC)[i];
      const float CFR = 0.05f;
      float RainFrac;
      if (1.0f * TTI == 0.0f) RainFrac = (Temperature <= TT) ? 0.0f : 1.0f;
      else RainFrac = fminf((Temperature - (TT - TTI / 2.0f)) / TTI, 1.0f);
      RainFrac = fmaxf(RainFrac, 0.0f);
      float SnowFrac = 1.0f - RainFrac;
      float Pc = (1.0f * SnowFrac) * EffectivePrecipitation + (1.0f * RainFrac) * EffectivePrecipitation;
      SnowFall = SnowFrac * Pc;
      float RainFall = RainFrac * Pc;
      float PotSnowMelt = (Temperature > TTM) ? Cfmax * (Temperature - TTM) : 0.0f;
      float PotRefreezing = (Temperature < TTM) ? (Cfmax * CFR) * (TTM - Temperature) : 0.0f;
      float Refreezing = (Temperature < TTM) ? fminf(PotRefreezing, SnowWater) : 0.0f;
      SnowMelt = fminf(PotSnowMelt, Snow);
      Snow = ((Snow + SnowFall) + Refreezing) - SnowMelt;
      SnowWater = SnowWater - Refreezing;
      float MaxSnowWater = Snow * WHC;
      SnowWater = (SnowWater + SnowMelt) + RainFall;
      RainFall = fmaxf(SnowWater - MaxSnowWater, 0.0f);
      SnowWater = SnowWater - RainFall;
      RainFallPlusMelt = RainFall;
      if (p->glacier) {
        /* glacierHBV, wflow_funcs.py:549-603; wflow_sbm.py:2715-2743 */
        float GF = F(GlacierFrac)[i], GS = F(GlacierStore)[i];
        float ratio = (float)(p->timestepsecs / p->basetimestep);
        float Snow2Glacier = (F(G_SIfrac)[i] * ratio) * Snow;
        Snow2Glacier = (GF > 0.0f) ? Snow2Glacier : 0.0f;
        Snow2Glacier = fminf(Snow2Glacier, 8.0f * ratio);
        Snow = Snow - (Snow2Glacier * GF);
        GS = GS + Snow2Glacier;
        float PotMelt = (Temperature > F(G_TT)[i]) ? (F(G_Cfmax)[i] * ratio) * (Temperature - F(G_TT)[i]) : 0.0f;
        float GlacierMelt = (Snow < 10.0f) ? fminf(PotMelt, GS) : 0.0f;
        GS = GS - GlacierMelt;
        F(GlacierStore)[i] = GS;
        GlacierMelt = GlacierMelt * GF;
        RainFallPlusMelt = RainFallPlusMelt + GlacierMelt;
        OUT(Snow2Glacier, i, Snow2Glacier);
        OUT(GlacierMelt, i, GlacierMelt);
      }
      F(Snow)[i] = Snow;
      F(SnowWater)[i] = SnowWater;
      F(TSoil)[i] = TSoil;
    } else {
      RainFallPlusMelt = EffectivePrecipitation;                      /* :2745 */
    }
    float thetaS = F(thetaS)[i], thetaR = F(thetaR)[i], ST = F(SoilThickness)[i];
    float zi = F(zi)[i];
    float SatWaterDepth = (thetaS - thetaR) * (ST - zi);              /* :2751 */
    float Avail = RainFallPlusMelt + 0.0f;                            /* :2755 IRSupplymm = 0 */
    float UStoreDepth = 0.0f;
    for (int k = 0; k < ML; k++) UStoreDepth = UStoreDepth + Uf[k][i]; /* :2758 */
    float RiverFrac = F(RiverFrac)[i], WaterFrac = F(WaterFrac)[i];
    float RunoffRiverCells = fminf(1.0f, RiverFrac) * Avail;          /* :2763 */
    float RunoffLandCells = fminf(1.0f, WaterFrac) * Avail;           /* :2766 */
    Avail = fmaxf((Avail - RunoffRiverCells) - RunoffLandCells, 0.0f); /* :2775 */
    float AEOWRiver = fminf((F(WaterLevelR)[i] * 1000.0f) * RiverFrac, RiverFrac * PotTransSoil); /* :2790 */
    float AEOWLand = fminf((F(WaterLevelL)[i] * 1000.0f) * WaterFrac, WaterFrac * PotTransSoil);  /* :2795 */
    float RestEvap = (PotTransSoil - AEOWRiver) - AEOWLand;           /* :2807 */
    float PotSoilEvap = RestEvap * CGF;                               /* :2818 */
    float PotTrans = RestEvap * (1.0f - CGF);                         /* :2819 */

    F(CanopyStorage)[i] = CanopyStorage;
    AvailInf[i] = Avail; PotSoilEvapD[i] = PotSoilEvap; PotTransD[i] = PotTrans;
    RunoffLandD[i] = RunoffLandCells; AEOWLandD[i] = AEOWLand; ziD[i] = zi; SatD[i] = SatWaterDepth;
    RainPlusMeltD[i] = RainFallPlusMelt; RunoffRiverD[i] = RunoffRiverCells;
    InwaterMMf[i] = (double)(RunoffRiverCells - AEOWRiver);           /* :3055 (REAL4) */
    LandRunoffsubD[i] = F(LandRunoffsub)[i];
    for (int k = 0; k < ML; k++) Ul[(size_t)k * N + i] = Uf[k][i];
    /* dyn["WaterLevelLsub"] is never refreshed from the map (only written where SW>0,
       wflow_sbm.py:873-878) and starts at 0: equivalent to 0 wherever SW<=0. */
    WLLsubD[i] = 0.0;

    OUT(Precipitation, i, Precipitation); OUT(PotenEvap, i, PotenEvap); OUT(Temperature, i, Temperature);
    OUT(ThroughFall, i, ThroughFall); OUT(StemFlow, i, StemFlow); OUT(Interception, i, Interception);
    OUT(PotTransSoil, i, PotTransSoil); OUT(SnowMelt, i, SnowMelt); OUT(SnowFall, i, SnowFall);
    OUT(RainFallPlusMelt, i, RainFallPlusMelt); OUT(AvailableForInfiltration, i, Avail);
    OUT(RunoffRiverCells, i, RunoffRiverCells); OUT(RunoffLandCells, i, RunoffLandCells);
    OUT(ActEvapOpenWaterRiver, i, AEOWRiver); OUT(ActEvapOpenWaterLand, i, AEOWLand);
    OUT(RestEvap, i, RestEvap); OUT(PotSoilEvap, i, PotSoilEvap); OUT(PotTrans, i, PotTrans);
    OUT(InwaterMM, i, RunoffRiverCells - AEOWRiver);
    /* InterceptionWatBal :3507 */
    OUT(InterceptionWatBal, i,
        (((Precipitation - Interception) - StemFlow) - ThroughFall) - ((float)OldCanopy[i] - CanopyStorage));
  }

  /* ---------------- P5: sbm_cell (wflow_sbm.py:332-900), float64 ---------------- */
  const double dt = p->timestepsecs;
  for (int64_t lev = 0; lev < net->nlev; lev++) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(g_threads)
    for (int64_t kk = net->lev_off[lev]; kk < net->lev_off[lev + 1]; kk++) {
      const int64_t idx = net->nodes[kk];
      const int64_t *nbs = net->nodes_up + 8 * kk;
      const double ST = F(SoilThickness)[idx], thetaS = F(thetaS)[idx], thetaR = F(thetaR)[idx];
      const double neff = (double)(float)(F(thetaS)[idx] - F(thetaR)[idx]);          /* :2158 REAL4 */
      const double SWC = (double)(float)(F(SoilThickness)[idx] * (F(thetaS)[idx] - F(thetaR)[idx])); /* :1935 */
      const double KsatVer = F(KsatVer)[idx], fpar = F(f)[idx];
      const double ActRoot = (double)fminf(F(SoilThickness)[idx] * 0.99f, F(RootingDepth)[idx]); /* :2783 */

      /* layer thickness, :2301-2326 (float64 from REAL4 SoilThickness) */
      double thick[WFO_MAXL], tops[WFO_MAXL + 1];
      {
        double sumT = 0.0;
        for (int n = 0; n < ML; n++) {
          double sumL = sumT, tmp, th;
          if (p->nrLayers > 1 && n < p->nrLayers) {
            tmp = p->layerThickness[n] + 0.0;
            th = fmin(tmp, fmax(ST - sumL, 0.0));
          } else {
            tmp = fmax(ST - sumL, 0.0);
            th = tmp;
          }
          sumT = tmp + sumT;
          thick[n] = th;
        }
        tops[0] = 0.0;
        double cs = 0.0;
        for (int n = 0; n < ML; n++) { cs = cs + thick[n]; tops[n + 1] = cs; } /* cumsum :375 */
      }
      double *U[WFO_MAXL];
      for (int k = 0; k < ML; k++) U[k] = &Ul[(size_t)k * N + idx];

      double zi = ziD[idx];
      const double SWDold = SatD[idx];
      double sumUSold = 0.0;
      for (int k = 0; k < ML; k++) sumUSold += *U[k];
      /* n = where(zi > sumlayer_0): duplicates removed by np.unique sit at SoilThickness >= zi */
      int m = 0;
      for (int k = 0; k < ML; k++) {
        if (k > 0 && tops[k] == tops[k - 1]) break; /* np.unique */
        if (zi > tops[k]) m = k + 1;
      }
      const int nL = (m > 1) ? m : 1;
      double L[WFO_MAXL], z[WFO_MAXL];
      for (int k = 0; k < nL - 1; k++) L[k] = thick[k];
      L[nL - 1] = (m > 1) ? zi - tops[nL - 1] : zi;
      { double cs = 0.0; for (int k = 0; k < nL; k++) { cs = cs + L[k]; z[k] = cs; } }

      double ActEvapUStore = 0.0;
      double ssf_in, ssf_toriver = 0.0;
      const int isRiver = F(River)[idx] != 0.0f;
      const double slope_i = F(landSlope)[idx];
      double chanperc[8];
      if (isRiver) { /* :395-405 */
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < 8; j++) {
          int64_t nb = nbs[j];
          double sv = (nb < 0) ? 0.0 : ssf_new[nb];
          uint8_t lnb = (nb < 0) ? 0 : ldd[nb];
          double slnb = (nb < 0) ? 0.0 : (double)F(landSlope)[nb];
          chanperc[j] = (lnb != ldd[idx]) ? slnb / (slope_i + slnb) : 0.0;
          s1 = s1 + (1 - chanperc[j]) * sv;
          s2 = s2 + chanperc[j] * sv;
        }
        ssf_in = s1;
        ssf_toriver = s2 / (1000 * 1000 * 1000) / dt;
      } else {
        double s1 = 0.0;
        for (int j = 0; j < 8; j++) s1 = s1 + ((nbs[j] < 0) ? 0.0 : ssf_new[nbs[j]]);
        ssf_in = s1;
      }
      ssf_toriverD[idx] = ssf_toriver;
      const double CellInFlow = ssf_in;

      double sumUn = 0.0;
      for (int k = 0; k < m; k++) sumUn += *U[k];
      double UStoreCapacity = SWC - SatD[idx] - sumUn;                /* :413 */
      double InfiltSoilPath, InfiltSoil, InfiltPath, SoilInf, PathInf, InfiltExcess;
      infiltration(AvailInf[idx], F(PathFrac)[idx], F(cf_soil) ? F(cf_soil)[idx] : 0.0, F(TSoil) ? F(TSoil)[idx] : 0.0,
                   F(InfiltCapSoil)[idx], F(InfiltCapPath)[idx], UStoreCapacity, p->modelSnow,
                   p->soilInfRedu, &InfiltSoilPath, &InfiltSoil, &InfiltPath, &SoilInf, &PathInf,
                   &InfiltExcess);

      /* unsatzone_flow, :275-329 */
      double ast = 0.0;
      {
        *U[0] = *U[0] + InfiltSoilPath;
        if (L[0] > 0.0) {
          if (p->transferMethod == 1 && ML == 1) {
            double Sd = SWC - SWDold;
            if (Sd <= 0.00001) {
              ast = 0.0; /* reference leaves `ast` undefined here when maxLayers==1; see note */
            } else {
              double st = kvf[0][idx] * KsatVer * exp(-fpar * z[0]) *
                          (pymin(*U[0], L[0] * (thetaS - thetaR)) / Sd);
              ast = pymin(st, *U[0]);
              *U[0] = *U[0] - ast;
            }
          } else {
            double L_sat = L[0] * (thetaS - thetaR);
            unsatzone_flow_layer(U[0], kvf[0][idx], KsatVer, fpar, z[0], L_sat, cf[0][idx], &ast);
          }
        } else {
          ast = 0.0;
        }
        for (int mm = 1; mm < nL; mm++) {
          *U[mm] = *U[mm] + ast;
          if (L[mm] > 0.0) {
            double L_sat = L[mm] * (thetaS - thetaR);
            unsatzone_flow_layer(U[mm], kvf[mm][idx], KsatVer, fpar, z[mm], L_sat, cf[mm][idx], &ast);
          } else {
            ast = 0.0;
          }
        }
      }
      const double Transfer = ast;

      /* evapotranspiration from layers, :465-591 */
      double soilevapunsat = 0.0, soilevapsat = 0.0, ActEvapSat = 0.0, RestPotTrans = 0.0;
      double PotSoilEvap = PotSoilEvapD[idx];
      double Sat = SatD[idx];
      for (int k = 0; k < nL; k++) {
        if (k == 0) {
          double SaturationDeficit = SWC - Sat;
          if (ML == 1) {
            soilevapunsat = PotSoilEvap * pymin(1.0, SaturationDeficit / SWC);
          } else if (nL == 1) {
            if (zi > 0) soilevapunsat = PotSoilEvap * pymin(1.0, *U[0] / (zi * (thetaS - thetaR)));
            else soilevapunsat = 0.0;
          } else {
            soilevapunsat = PotSoilEvap * pymin(1.0, *U[0] / (thick[0] * (thetaS - thetaR)));
          }
          soilevapunsat = pymin(soilevapunsat, *U[0]);
          PotSoilEvap = PotSoilEvap - soilevapunsat;
          *U[0] = *U[0] - soilevapunsat;
          if (ML == 1) {
            soilevapsat = 0.0;
          } else if (nL == 1) {
            soilevapsat = PotSoilEvap * pymin(1.0, (thick[0] - zi) / thick[0]);
            soilevapsat = pymin(soilevapsat, (thick[0] - zi) * (thetaS - thetaR));
          } else {
            soilevapsat = 0.0;
          }
          Sat = Sat - soilevapsat;
          double wetroots = sCurve(zi, ActRoot, 1.0, F(rootdistpar)[idx]);
          ActEvapSat = pymin(PotTransD[idx] * wetroots, Sat);
          Sat = Sat - ActEvapSat;
          RestPotTrans = PotTransD[idx] - ActEvapSat;
        }
        actTransp_unsat_SBM(ActRoot, U[k], tops[k], &RestPotTrans, &ActEvapUStore, cf[k][idx], L[k],
                            thetaS, thetaR, F(AirEntryPressure)[idx], p->ust);
      }
      const double soilevap = soilevapunsat + soilevapsat;

      /* layer balance, :593-607 */
      double du = 0.0;
      for (int k = nL - 1; k >= 0; k--) {
        du = pymax(0, *U[k] - L[k] * (thetaS - thetaR));
        *U[k] = *U[k] - du;
        if (k > 0) *U[k - 1] = *U[k - 1] + du;
      }
      double Ksat = kvf[nL - 1][idx] * KsatVer * exp(-fpar * zi);     /* :609 */
      sumUn = 0.0;
      for (int k = 0; k < m; k++) sumUn += *U[k];
      UStoreCapacity = SWC - Sat - sumUn;                              /* :615 */
      double MaxCapFlux = pymax(0.0, pymin(pymin(pymin(Ksat, ActEvapUStore), UStoreCapacity), Sat));
      double CapFluxScale = 0.0;
      if (zi > ActRoot)
        CapFluxScale = F(CapScale)[idx] / (F(CapScale)[idx] + zi - ActRoot) * dt / p->basetimestep;
      double CapFlux = MaxCapFlux * CapFluxScale;
      double netCapflux = CapFlux, actCapFlux = 0.0;
      for (int k = nL - 1; k >= 0; k--) {
        double toadd = pymin(netCapflux, pymax(L[k] * (thetaS - thetaR) - *U[k], 0.0));
        *U[k] = *U[k] + toadd;
        netCapflux = netCapflux - toadd;
        actCapFlux = actCapFlux + toadd;
      }
      double DeepKsat = KsatVer * exp(-fpar * ST);
      double DeepTransfer = pymin(Sat, DeepKsat);
      double ActLeakage = pymax(0.0, pymin(F(MaxLeakage)[idx], DeepTransfer));
      const double DW = F(DW)[idx], DLd = F(DL)[idx];
      double r = (ast - actCapFlux - ActLeakage - ActEvapSat - soilevapsat) * DW * 1000;

      /* ssfmax :2291-2299 */
      const double Kh = F(KsatHorFrac)[idx];
      double ssfmax = ((Kh * KsatVer * slope_i) / fpar) * (exp(0) - exp(-fpar * ST));
      double ssf_n, zi_n, ExfiltSatWater;
      wfo_kinematic_wave_ssf(ssf_in, (double)F(SubsurfaceFlow)[idx], zi, r, Kh, KsatVer, slope_i, neff,
                             fpar, ST, 1.0, DLd * 1000, DW * 1000, ssfmax, &ssf_n, &zi_n, &ExfiltSatWater);
      ssf_new[idx] = ssf_n;
      zi = pymin(zi_n, ST);                                           /* :702 */
      Sat = (ST - zi) * (thetaS - thetaR);

      int m_new = 0;
      for (int k = 0; k < ML; k++) {
        if (k > 0 && tops[k] == tops[k - 1]) break;
        if (zi > tops[k]) m_new = k + 1;
      }
      const int nLn = (m_new > 1) ? m_new : 1;
      double L_new[WFO_MAXL];
      for (int k = 0; k < nLn - 1; k++) L_new[k] = thick[k];
      L_new[nLn - 1] = (m_new > 1) ? zi - tops[nLn - 1] : zi;

      double ExfiltFromUstore = 0.0;
      for (int k = nL - 1; k >= 0; k--) { /* :718-734 */
        if (k < m_new) ExfiltFromUstore = pymax(0, *U[k] - L_new[k] * (thetaS - thetaR));
        else ExfiltFromUstore = *U[k];
        *U[k] = *U[k] - ExfiltFromUstore;
        if (k > 0) *U[k - 1] = *U[k - 1] + ExfiltFromUstore;
      }
      double ExfiltWater = ExfiltSatWater + ExfiltFromUstore;
      double ExcessWater = AvailInf[idx] - InfiltSoilPath - InfiltExcess + du;
      double ActInfilt = InfiltSoilPath - du;
      double ActInfiltSoil = 0.0, ActInfiltPath = 0.0;
      if ((InfiltSoil + InfiltPath) > 0.0) {
        ActInfiltSoil = InfiltSoil - du * InfiltSoil / (InfiltPath + InfiltSoil);
        ActInfiltPath = InfiltPath - du * InfiltPath / (InfiltPath + InfiltSoil);
      }
      double ExcessWaterSoil = pymax(SoilInf - ActInfiltSoil, 0.0);
      double ExcessWaterPath = pymax(PathInf - ActInfiltPath, 0.0);
      InwaterO[idx] = pymax(ExfiltWater + ExcessWater + InfiltExcess + RunoffLandD[idx] - AEOWLandD[idx] - 0, 0.0) *
                      ((double)F(xl)[idx] * (double)F(yl)[idx]) * 0.001 / dt; /* :774 */
      double sumU = 0.0;
      for (int k = 0; k < ML; k++) sumU += *U[k];

      /* vwc :792-807 */
      for (int k = 0; k < ML; k++) {
        float *vw = fields[WFO_vwc_0 + k];
        if (!vw) continue;
        double vwc = vw[idx];  /* layer["vwc"] keeps its value where neither branch writes (:792-803) */
        if (k < m_new) {
          if (thick[k] > 0) vwc = (*U[k] + (thick[k] - L_new[k]) * (thetaS - thetaR)) / thick[k] + thetaR;
        } else {
          vwc = thetaS;
        }
        vw[idx] = (float)vwc;
        if (fields[WFO_vwc_perc_0 + k]) fields[WFO_vwc_perc_0 + k][idx] = (float)((vwc / thetaS) * 100.0); /* :805-807 */
      }
      double rootStore_unsat = 0;
      for (int k = 0; k < nLn; k++)
        if (L_new[k] > 0) rootStore_unsat = rootStore_unsat + (pymax(0.0, ActRoot - tops[k]) / L_new[k]) * *U[k];

      ziD[idx] = zi; SatD[idx] = Sat;
      /* SoilWatbal :885-898 */
      double SoilWatbal = ActInfilt - ((Sat + sumU) - (sumUSold + SWDold)) +
                          (CellInFlow - ssf_n) / (DW * DLd * 1000 * 1000) - ExfiltWater - soilevap -
                          ActEvapUStore - ActEvapSat - ActLeakage;

      /* numpy2pcr back to REAL4, :2944-3129 */
      F(SubsurfaceFlow)[idx] = (float)ssf_n;
      F(zi)[idx] = (float)zi;
      if (F(SatWaterDepth)) F(SatWaterDepth)[idx] = (float)Sat;
      for (int k = 0; k < ML; k++) Uf[k][idx] = (float)*U[k];
      OUT(UStoreDepth, idx, sumU); OUT(CapFlux, idx, actCapFlux); OUT(Transfer, idx, Transfer);
      OUT(ActEvapUStore, idx, ActEvapUStore); OUT(ActEvapSat, idx, ActEvapSat);
      OUT(Transpiration, idx, (float)ActEvapUStore + (float)ActEvapSat);
      OUT(soilevap, idx, soilevap); OUT(soilevapsat, idx, soilevapsat);
      if (F(ActEvap)) {
        float aer = F(ActEvapOpenWaterRiver) ? F(ActEvapOpenWaterRiver)[idx]
                                             : (float)(RunoffRiverD[idx] - InwaterMMf[idx]);
        float tr = (float)ActEvapUStore + (float)ActEvapSat;
        F(ActEvap)[idx] = ((((float)soilevap + tr) + aer) + (float)AEOWLandD[idx]) + 0.0f; /* :3002 */
      }
      OUT(Recharge, idx, (((float)Transfer - (float)actCapFlux) - (float)ActEvapSat) - (float)soilevapsat);
      OUT(ssf_toriver, idx, ssf_toriver); OUT(ExfiltWater, idx, ExfiltWater);
      OUT(InfiltExcess, idx, InfiltExcess); OUT(ExcessWater, idx, ExcessWater); OUT(ActInfilt, idx, ActInfilt);
      OUT(ActLeakage, idx, ActLeakage); OUT(SoilInf, idx, SoilInf); OUT(PathInf, idx, PathInf);
      OUT(InfiltExcessSoil, idx, ExcessWaterSoil); OUT(InfiltExcessPath, idx, ExcessWaterPath);
      OUT(ActInfiltSoil, idx, ActInfiltSoil); OUT(ActInfiltPath, idx, ActInfiltPath);
      OUT(ExfiltFromUstore, idx, ExfiltFromUstore); OUT(ExfiltFromSat, idx, ExfiltSatWater);
      OUT(InwaterL, idx, InwaterO[idx]); OUT(CellInFlow, idx, CellInFlow);
      OUT(SoilWatbal, idx, SoilWatbal); OUT(RootStore_unsat, idx, rootStore_unsat);
      OUT(SumEvapSat, idx, (float)soilevapsat + (float)ActEvapSat);          /* :3118-3122 */
      OUT(SumEvapUstore, idx, (float)soilevapunsat + (float)ActEvapUStore);  /* :3123-3127 */
      if (F(SurfaceWatbal) || F(watbal)) { /* :3515-3526 (REAL4) */
        float sw = (float)RainPlusMeltD[idx] + 0.0f;
        sw = sw - (float)InfiltExcess; sw = sw - (float)ExcessWater; sw = sw - (float)RunoffRiverD[idx];
        sw = sw - (float)RunoffLandD[idx]; sw = sw - (float)ActInfilt;
        sw = sw - ((float)OldCanopy[idx] - F(CanopyStorage)[idx]);
        OUT(SurfaceWatbal, idx, sw);
        OUT(watbal, idx, (float)SoilWatbal + sw);
      }
    }
  }

  /* overland flow, wflow_sbm.py:823-883 */
  const int itL = p->it_kinL;
  for (int v = 0; v < itL; v++) {
    memset(qo_new, 0, sizeof(double) * (N + 1));
    for (int64_t lev = 0; lev < net->nlev; lev++) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(g_threads)
      for (int64_t kk = net->lev_off[lev]; kk < net->lev_off[lev + 1]; kk++) {
        const int64_t idx = net->nodes[kk];
        const int64_t *nbs = net->nodes_up + 8 * kk;
        double qo_in, qo_toriver_vol;
        const double SW = F(SW)[idx];
        if (F(River)[idx] != 0.0f) {
          const double slope_i = F(landSlope)[idx];
          double s1 = 0.0, s2 = 0.0, s3 = 0.0;
          for (int j = 0; j < 8; j++) {
            int64_t nb = nbs[j];
            double qv = (nb < 0) ? 0.0 : qo_new[nb];
            uint8_t lnb = (nb < 0) ? 0 : ldd[nb];
            double slnb = (nb < 0) ? 0.0 : (double)F(landSlope)[nb];
            double cp = (lnb != ldd[idx]) ? slnb / (slope_i + slnb) : 0.0;
            s1 = s1 + (1 - cp) * qv; s2 = s2 + cp * qv; s3 = s3 + qv;
          }
          if (SW > 0.0) { qo_in = s1; qo_toriver_vol = s2 * (dt / itL); }
          else { qo_in = 0.0; qo_toriver_vol = s3 * (dt / itL); }
        } else {
          double s1 = 0.0;
          for (int j = 0; j < 8; j++) s1 = s1 + ((nbs[j] < 0) ? 0.0 : qo_new[nbs[j]]);
          qo_in = s1; qo_toriver_vol = 0.0;
        }
        const double DLd = F(DL)[idx], AlphaL = F(AlphaL)[idx], Beta = F(Beta)[idx];
        double q = InwaterO[idx] / DLd;                                /* :830 */
        double qn = wfo_kinematic_wave(qo_in, LandRunoffsubD[idx], q, AlphaL, Beta, dt / itL, DLd);
        qo_new[idx] = qn;
        acc_flow[idx] = acc_flow[idx] + qn * (dt / itL);
        Qo_in[idx] = Qo_in[idx] + qo_in * (dt / itL);
        qo_toriver_acc[idx] = qo_toriver_acc[idx] + qo_toriver_vol;
        if (SW > 0) WLLsubD[idx] = (AlphaL * pow(qn, Beta)) / SW;
        acc_h[idx] = acc_h[idx] + WLLsubD[idx];
        LandRunoffsubD[idx] = qn;
      }
    }
  }
  for (int64_t kk = 0; kk < net->ncells; kk++) {
    const int64_t idx = net->nodes[kk];
    F(LandRunoffsub)[idx] = (float)LandRunoffsubD[idx];
    F(WaterLevelLsub)[idx] = (float)WLLsubD[idx];
    F(LandRunoff)[idx] = (float)(acc_flow[idx] / dt);
    F(WaterLevelL)[idx] = (float)(acc_h[idx] / itL);
    OUT(qo_toriver, idx, qo_toriver_acc[idx] / dt);
    OUT(InflowKinWaveCellLand, idx, Qo_in[idx] / dt);
  }

  /* ---------------- P7-P8: river (wflow_sbm.py:3052-3201) ---------------- */
  {
    double *Qold = acc_flow;  /* reuse work arrays */
    double *qr = acc_h;
    double *racc = Qo_in, *rwl = InwaterO, *racch = AvailInf;
    double *Qnew = qo_new;
    for (int64_t i = 0; i < N; i++) { racc[i] = 0; rwl[i] = 0; racch[i] = 0; Qold[i] = 0; qr[i] = 0; }
    for (int64_t i = 0; i < N; i++) {
      if (!active[i]) continue;
      float ToCubic = ((F(xl)[i] * F(yl)[i]) * 0.001f) / dtf;          /* :1998 */
      float qtr = (float)(qo_toriver_acc[i] / dt);
      float str_ = (float)ssf_toriverD[i];
      float Inwater = (((float)InwaterMMf[i] * ToCubic) + qtr) + str_;  /* :3063 */
      float Inflow = F(Inflow) ? F(Inflow)[i] : 0.0f;
      if (Inflow < 0.0f) { free(W); free(Ul); free(pre); return -5; }             /* abstraction loop not restated */
      Inwater = Inwater + Inflow;                                       /* :3144 (SurfaceWaterSupply == 0) */
      OUT(Inwater, i, Inwater);
      qr[i] = (double)(Inwater / F(DCL)[i]);                            /* :3152 */
      Qold[i] = (double)F(RiverRunoffsub)[i];
    }
    const int itR = p->it_kinR;
    for (int v = 0; v < itR; v++) { /* kin_wave, wflow_funcs.py:155-207 */
      memset(Qnew, 0, sizeof(double) * (N + 1));
      for (int64_t lev = 0; lev < rnet->nlev; lev++) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(g_threads)
        for (int64_t kk = rnet->lev_off[lev]; kk < rnet->lev_off[lev + 1]; kk++) {
          const int64_t idx = rnet->nodes[kk];
          const int64_t *nbs = rnet->nodes_up + 8 * kk;
          double Qin = 0.0;
          for (int j = 0; j < 8; j++) Qin = Qin + ((nbs[j] < 0) ? 0.0 : Qnew[nbs[j]]);
          const double A = F(AlphaR)[idx], B = F(Beta)[idx];
          double qn = wfo_kinematic_wave(Qin, Qold[idx], qr[idx], A, B, dt / itR, (double)F(DCL)[idx]);
          Qnew[idx] = qn;
          racc[idx] = racc[idx] + qn * (dt / itR);
          rwl[idx] = (A * pow(qn, B)) / (double)F(Bw)[idx];
          racch[idx] = racch[idx] + rwl[idx];
          /* the caller hands kin_wave a REAL4 array (pcr2numpy of the RiverRunoffsub map, wflow_sbm.py:3157) and
             kin_wave stores into it (wflow_funcs.py:199): the outflow is rounded to float32 between sub-steps */
          Qold[idx] = (double)(float)qn;
        }
      }
    }
    for (int64_t i = 0; i < N; i++) {
      if (!active[i]) continue;
      F(RiverRunoff)[i] = (float)(racc[i] / dt);                        /* :3189 */
      F(WaterLevelR)[i] = (float)(racch[i] / itR);
      F(RiverRunoffsub)[i] = (float)Qnew[i];
      F(WaterLevelRsub)[i] = (float)rwl[i];
    }
  }
  /* ---------------- end of dynamic(): kinematic-wave volumes and mass balances, unit conversions, storage change,
   * root-zone water content, cumulative sums (wflow_sbm.py:3310-3312, 3348-3370, 3429-3526; updateKinWaveVolume :946-953).
   * REAL4 map algebra on the float32 maps of this step; the maps it reads must be requested with the outputs
   * (oracle.py does that).  A division by zero is a missing value in PCRaster: -999 here. ---------------- */
  if (want_end) {
    const float MV = -999.0f;
    float *ct1 = (float *)calloc((size_t)N * 3, sizeof(float));
    if (!ct1) { free(W); free(Ul); free(pre); return -4; }
    float *ctR = ct1 + N, *inflowKW = ct1 + 2 * N;
    const int need_ct = F(QCatchmentMM) || F(RunoffCoeff);
    for (int64_t kk = 0; kk < net->ncells; kk++) { /* pcr.upstream(TopoLdd, RiverRunoff) :3349, neighbour order of _up_nb */
      const int64_t idx = net->nodes[kk];
      const int64_t *nbs = net->nodes_up + 8 * kk;
      float up = 0.0f;
      for (int j = 0; j < 8; j++)
        if (nbs[j] >= 0) up = up + F(RiverRunoff)[nbs[j]];
      inflowKW[idx] = up;
    }
    if (need_ct) {
      /* pcr.catchmenttotal over the LDD MAP (:3467-3469, :1989): every cell with a valid ldd code counts, also the cells
       * that set_dd's flat-index-0 quirk leaves out of the network (wflow_funcs.py:88).  Kahn order, float32 sums. */
      static const int dr[10] = {0, 1, 1, 1, 0, 0, 0, -1, -1, -1}, dc[10] = {0, -1, 0, 1, -1, 0, 1, -1, 0, 1};
      int64_t *ds = (int64_t *)malloc(sizeof(int64_t) * (size_t)N * 2);
      int32_t *indeg = (int32_t *)calloc((size_t)N, sizeof(int32_t));
      if (!ds || !indeg) { free(ds); free(indeg); free(ct1); free(W); free(Ul); free(pre); return -4; }
      int64_t *queue = ds + N;
      for (int64_t i = 0; i < N; i++) {
        ds[i] = -1;
        ct1[i] = 1.0f;
        ctR[i] = F(RainFallPlusMelt) ? F(RainFallPlusMelt)[i] : 0.0f;
        const int c = ldd[i];
        if (c < 1 || c > 9 || c == 5) continue;
        const int64_t r2 = i / p->ncol + dr[c], c2 = i % p->ncol + dc[c];
        if (r2 < 0 || r2 >= p->nrow || c2 < 0 || c2 >= p->ncol) continue;
        const int64_t d = r2 * p->ncol + c2;
        if (ldd[d] < 1 || ldd[d] > 9) continue;
        ds[i] = d;
        indeg[d]++;
      }
      int64_t qh = 0, qt = 0;
      for (int64_t i = 0; i < N; i++)
        if (ldd[i] >= 1 && ldd[i] <= 9 && indeg[i] == 0) queue[qt++] = i;
      while (qh < qt) {
        const int64_t c = queue[qh++], d = ds[c];
        if (d < 0) continue;
        ct1[d] = ct1[d] + ct1[c];
        ctR[d] = ctR[d] + ctR[c];
        if (--indeg[d] == 0) queue[qt++] = d;
      }
      free(ds); free(indeg);
    }
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int64_t i = 0; i < N; i++) {
      if (!active[i]) continue;
      const float xl = F(xl)[i], yl = F(yl)[i], thS = F(thetaS)[i], thR = F(thetaR)[i];
      const float ToCubic = ((xl * yl) * 0.001f) / dtf;                 /* :1998 */
      const float QMMConv = dtf / ((xl * yl) * 0.001f);                 /* :1985 */
      const float RR = F(RiverRunoff)[i], LR = F(LandRunoff)[i];
      const float KWVR = (F(WaterLevelRsub)[i] * F(Bw)[i]) * F(DCL)[i]; /* :950-953 */
      const float KWVL = (F(WaterLevelLsub)[i] * F(SW)[i]) * F(DL)[i];
      const float OldKWVR = (preWLRsub[i] * F(Bw)[i]) * F(DCL)[i], OldKWVL = (preWLLsub[i] * F(SW)[i]) * F(DL)[i];
      OUT(KinWaveVolumeR, i, KWVR); OUT(KinWaveVolumeL, i, KWVL);
      OUT(OldKinWaveVolumeR, i, OldKWVR); OUT(OldKinWaveVolumeL, i, OldKWVL);
      OUT(InflowKinWaveCell, i, inflowKW[i]);
      if (F(MassBalKinWaveR))                                           /* :3351-3356 */
        F(MassBalKinWaveR)[i] = (((((-KWVR) + OldKWVR) / dtf) + inflowKW[i]) + F(Inwater)[i]) - RR;
      if (F(MassBalKinWaveL))                                           /* :3365-3370 */
        F(MassBalKinWaveL)[i] = (((((-KWVL) + OldKWVL) / dtf) + F(InflowKinWaveCellLand)[i]) + F(InwaterL)[i]) - LR;
      OUT(RiverRunoffMM, i, RR * QMMConv); OUT(LandRunoffMM, i, LR * QMMConv);   /* :3305-3312 */
      if (need_ct) {
        const float temp = (((ct1[i] * xl) * 0.001f) * 0.001f) * yl;    /* :1989-1997 */
        const float QMMConvUp = (temp != 0.0f) ? (float)(p->timestepsecs * 0.001) / temp : MV;
        const float QC = (QMMConvUp != MV) ? RR * QMMConvUp : MV;       /* :3465 */
        OUT(QCatchmentMM, i, QC);
        OUT(RunoffCoeff, i, (QC != MV && ctR[i] != 0.0f) ? (QC / ctR[i]) / ct1[i] : MV); /* :3466-3470 */
      }
      float su = 0.0f;
      for (int k = 0; k < ML; k++) su = su + Uf[k][i];
      const float Sat = F(SatWaterDepth) ? F(SatWaterDepth)[i] : (float)SatD[i];
      const float CellStorage = su + Sat;                               /* :3473-3475 */
      OUT(sumUstore, i, su); OUT(CellStorage, i, CellStorage); OUT(DeltaStorage, i, CellStorage - preOrg[i]);
      const float SatWaterFlux = F(SubsurfaceFlow)[i] / (((xl * 1000.0f) * yl) * 1000.0f); /* :3480 */
      OUT(SatWaterFlux, i, SatWaterFlux);
      if (F(SnowCover)) F(SnowCover)[i] = (F(Snow) && F(Snow)[i] > 0.0f) ? 1.0f : 0.0f;  /* :3499-3501 */
      const float ActRoot = fminf(F(SoilThickness)[i] * 0.99f, F(RootingDepth)[i]);      /* :2783 */
      const float RSsat = fmaxf(0.0f, ActRoot - F(zi)[i]) * (thS - thR);                 /* :3431-3433 */
      OUT(RootStore_sat, i, RSsat);
      if (F(RootStore) || F(vwcRoot) || F(vwc_percRoot)) {
        const float RS = RSsat + F(RootStore_unsat)[i];                 /* :3456 */
        const float vwcRoot = (ActRoot != 0.0f) ? (RS / ActRoot) + thR : MV;
        OUT(RootStore, i, RS); OUT(vwcRoot, i, vwcRoot);
        OUT(vwc_percRoot, i, (vwcRoot != MV && thS != 0.0f) ? (vwcRoot / thS) * 100.0f : MV);
      }
      if (F(ExfiltWaterCubic)) F(ExfiltWaterCubic)[i] = F(ExfiltWater)[i] * ToCubic;     /* :3128-3129 */
      if (F(InfiltExcessCubic)) F(InfiltExcessCubic)[i] = F(InfiltExcess)[i] * ToCubic;
      OUT(SurfaceWaterSupply, i, 0.0f);                                 /* no abstraction: Inflow >= 0 (:3138-3140) */
      /* cumulative sums, :3487-3504 and :3058 */
      if (F(CumOutFlow)) F(CumOutFlow)[i] = F(CumOutFlow)[i] + SatWaterFlux;
      if (F(CumActInfilt)) F(CumActInfilt)[i] = F(CumActInfilt)[i] + F(ActInfilt)[i];
      if (F(CumCellInFlow)) F(CumCellInFlow)[i] = F(CumCellInFlow)[i] + F(CellInFlow)[i];
      if (F(CumPrec)) F(CumPrec)[i] = F(CumPrec)[i] + F(Precipitation)[i];
      if (F(CumEvap)) F(CumEvap)[i] = F(CumEvap)[i] + F(ActEvap)[i];
      if (F(CumPotenEvap)) F(CumPotenEvap)[i] = F(CumPotenEvap)[i] + F(PotenEvap)[i];
      if (F(CumInt)) F(CumInt)[i] = F(CumInt)[i] + F(Interception)[i];
      if (F(CumLeakage)) F(CumLeakage)[i] = F(CumLeakage)[i] + F(ActLeakage)[i];
      if (F(CumExfiltWater)) F(CumExfiltWater)[i] = F(CumExfiltWater)[i] + F(ExfiltWater)[i];
      if (F(CumSurfaceWater)) F(CumSurfaceWater)[i] = F(CumSurfaceWater)[i] + (preWLR[i] / 1000.0f);
    }
    free(ct1);
  }
  free(W); free(Ul); free(pre);
  return 0;
}

/* kin_wave, wflow/wflow_funcs.py:155-207, as a function of its own (wflow_hbv.py:1633-1647 calls it on the full network with
 * every cell a channel).  Qold: the REAL4 array the caller hands over (pcr2numpy of the SurfaceRunoffsub map), which
 * kin_wave stores into between sub-steps (:199): the outflow is rounded to float32 there.  q REAL4; Alpha, Beta, DCL, Bw
 * float64 (the `static` mirror).  Outputs (float64, N values each, untouched outside the network): acc_flow [m3],
 * Qnew [m3/s], waterlevel_av, waterlevel [m]. */
int wfo_kin_wave(const wfo_network *net, int64_t N, float *Qold, const float *q, const double *Alpha, const double *Beta,
                 const double *DCL, const double *Bw, double deltaT, int it, double *acc_flow, double *Qnew_out,
                 double *wl_av, double *wl) {
  double *Qnew = (double *)calloc((size_t)N + 1, sizeof(double));
  double *acc_h = (double *)calloc((size_t)N, sizeof(double));
  if (!Qnew || !acc_h) { free(Qnew); free(acc_h); return -4; }
  memset(acc_flow, 0, sizeof(double) * (size_t)N);
  memset(wl, 0, sizeof(double) * (size_t)N);
  for (int v = 0; v < it; v++) {
    memset(Qnew, 0, sizeof(double) * ((size_t)N + 1));
    for (int64_t lev = 0; lev < net->nlev; lev++) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(g_threads)
      for (int64_t kk = net->lev_off[lev]; kk < net->lev_off[lev + 1]; kk++) {
        const int64_t idx = net->nodes[kk];
        const int64_t *nbs = net->nodes_up + 8 * kk;
        double Qin = 0.0;
        for (int j = 0; j < 8; j++) Qin = Qin + ((nbs[j] < 0) ? 0.0 : Qnew[nbs[j]]);
        const double qn = wfo_kinematic_wave(Qin, (double)Qold[idx], (double)q[idx], Alpha[idx], Beta[idx], deltaT / it, DCL[idx]);
        Qnew[idx] = qn;
        acc_flow[idx] = acc_flow[idx] + qn * (deltaT / it);
        wl[idx] = (Alpha[idx] * pow(qn, Beta[idx])) / Bw[idx];
        acc_h[idx] = acc_h[idx] + wl[idx];
        Qold[idx] = (float)qn;
      }
    }
  }
  for (int64_t i = 0; i < N; i++) {
    Qnew_out[i] = Qnew[i];
    wl_av[i] = acc_h[i] / it;
  }
  free(Qnew);
  free(acc_h);
  return 0;
}
