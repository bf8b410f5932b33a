// Lateral flow: the level-ordered kinematic waves as a skewed wavefront.
//
// The reference walks the network three times per timestep, serially, upstream to downstream:
// subsurface flow inside sbm_cell (wflow/wflow_sbm.py:370-821), overland flow (:823-883) and the
// river (kin_wave, wflow/wflow_funcs.py:155-207, once per sub-step).  The dependencies are
//     subsurface(level l)      <- subsurface(l-1)
//     overland(l)              <- overland(l-1), subsurface(l)            (InwaterO of the cell)
//     river(l, sub-step v)     <- river(l-1, v), river(l, v-1), overland(l-1), subsurface(l-1)
// so one persistent cooperative kernel advances a wavefront: at tick t it solves
//     subsurface for level t, overland for level t-1, river sub-step v for level t-2-v,
// all concurrently in different warps, one grid barrier per tick: D + 1 + it_kinR ticks whose
// critical path is ONE Newton solve instead of the reference's (2 + it_kinR) * D serial solves.
//
// Inside a level the cells are contiguous in network order and so are each cell's upstream
// neighbours (network.h): every load of a level is coalesced, the inflow gather is a short
// contiguous run read with ld.global.cg (written by other SMs one tick earlier).
#pragma once
#include "kernels.cuh"

// build switches of this file (measured variants)
#ifndef WF_OVL_TAIL_PREFETCH
#define WF_OVL_TAIL_PREFETCH 0
#endif
#ifndef WF_SSF_INLINE
#define WF_SSF_INLINE 0
#endif
#ifndef WF_STAGE_AHEAD
#define WF_STAGE_AHEAD 1  // 1: the staged sweeps fetch the operands of a thread's NEXT item while it works on the current one
                          //    (sweep_group_ahead); 0: after the arrival at the tick barrier (sweep_group)
#endif

namespace wf {

// x^b for x >= 0 as exp(b*log(x)): the two CUDA double routines are each accurate to ~1 ulp, the
// result to ~1e-15 relative -- three times cheaper than pow()'s correctly rounded path, far inside
// the 1e-9 the float64 quantities are held to.  x = 0 gives 0 (b > 0) or inf (b < 0) like pow.
__device__ __forceinline__ double pow_el(double x, double b) { return wf_exp(b * wf_log(x)); }
// float32 x^b on the special-function unit (ex2.approx / lg2.approx), ~1e-6 relative: first guesses only.
__device__ __forceinline__ float pow_fast(float x, float b) { return exp2f(b * __log2f(x)); }

// Development aid (-DWF_PROFILE, a separate build): per tick, the latest time since the tick began at
// which any thread passed four points of each task (operands loaded, inflow gathered, solved, done).
#ifdef WF_PROFILE
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define WF_PARGS , unsigned long long *pslot, unsigned long long pc0
#define WF_PPASS(type) , (pbase ? pbase + 4 * (type) : nullptr), c0
__device__ __forceinline__ void prof_stamp(unsigned long long *slot, unsigned long long c0, long long dep) {
  const unsigned mask = __activemask();
  const unsigned d = (unsigned)(gtimer() - c0) + (unsigned)(dep & 0ll);
  const unsigned mx = __reduce_max_sync(mask, d);
  if ((threadIdx.x & 31) == (__ffs(mask) - 1)) atomicMax(slot, (unsigned long long)mx);
}
#define WF_STAMP(k, dep)                                                          \
  do {                                                                            \
    if (pslot) prof_stamp(pslot + (k), pc0, __double_as_longlong((double)(dep))); \
  } while (0)
#else
#define WF_PARGS
#define WF_PPASS(type)
#define WF_STAMP(k, dep)
#endif

// Load of a value another thread of the sweep wrote one tick earlier.  Across blocks (grid or
// cluster barrier) it is read from L2 (ld.global.cg); inside one block (__syncthreads) an ordinary
// load is coherent.
template <int MODE, typename T>
__device__ __forceinline__ T ldx(const T *p) {
  if (MODE == 2) return *(const volatile T *)p;
  return __ldcg(p);
}

// Qold^beta for the constant term of the kinematic wave.  Qold is the float32-rounded outflow of the solve
// one (sub-)step earlier, which left its float64 solution q64 and q64^beta behind: when the state still
// is that value, Qold^beta = q64^beta * (1 + d)^beta with d = (Qold - q64)/q64 ~ 1e-8 (two series terms
// are exact to rounding); otherwise (cold start, a state set from outside) the full power.
__device__ __forceinline__ double carried_pow(double Qold, double q64, double qb64, double beta) {
  if (Qold == 0.0 && beta > 0.0) return 0.0;
  if (q64 > 0.0 && (float)q64 == (float)Qold) {
    const double d = (Qold - q64) * qrcp(q64);
    return qb64 * fma(d, fma(d, 0.5 * beta * (beta - 1.0), beta), 1.0);  // (series of this file: explicit FMAs)
  }
  return pow_el(Qold, beta);
}

// kinematic_wave, wflow/wflow_funcs.py:119-152.
// The reference's Newton iteration converges to the root of  dT/dX*Q + alpha*Q^beta = C  to
// rounding (it only stops on |f| <= 1e-12 or its 3000-iteration cap), so the result is a property
// of the equation, not of the path.  Here:
//   1. a float32 Newton run from the reference's own first guess, to ~1e-6 (cheap; its constant
//      term is formed in float32 too, so it does not wait for the float64 one);
//   2. float64 Newton steps from there.  Q^beta is evaluated once in full (exp(beta*log Q)) and
//      then carried along the tiny steps by its Taylor series; a relative step below 1e-8 proves
//      the next iterate is converged to rounding (quadratic convergence), so 2 steps are typical.
// Agreement with the reference: a few ulp (tests/test_gpu_parity.py asserts 1e-9 on its golden set).
// Also returns Q^beta at the solution (the water level needs it).
// QoldBeta = Qold^beta (the caller usually carries it over from the solve that produced Qold, carried_pow).
__device__ __noinline__ double kinematic_wave(double Qin, double Qold, double QoldBeta, double q, double alpha, double beta,
                                              double deltaT, double deltaX, double &Qbeta) {
  Qbeta = 0.0;
  if ((Qin + Qold + q) == 0.0) return 0.0;
  const double deltaTX = qdiv(deltaT, deltaX);
  const double C = deltaTX * Qin + alpha * QoldBeta + deltaT * q;
  // float32 phase
  const float a32 = (float)alpha, b32 = (float)beta, dtx32 = (float)deltaTX;
  const float qin32 = (float)Qin, qold32 = (float)Qold, lat32 = (float)(deltaT * q);
  // (this phase only has to land within ~1e-5 of the root: hardware exp2/log2 and fast division)
  // (own iteration from here to the float64 Newton loop's end: explicit FMAs, it converges to the same root)
  const float C32 = fmaf(dtx32, qin32, fmaf(a32, pow_fast(qold32, b32), lat32));
  const float ab_pQ = a32 * b32 * pow_fast((qold32 + qin32) * 0.5f, b32 - 1.0f);
  float Qf = __fdividef(fmaf(dtx32, qin32, fmaf(qold32, ab_pQ, lat32)), dtx32 + ab_pQ);
  if (isnan(Qf)) Qf = 0.0f;
  {
    // The root lies in (ub/4, ub], ub = min(C/dtx, (C/alpha)^(1/beta)) (each term alone reaches C at its
    // bound).  The reference's linearised guess collapses towards 0 when Qold + Qin is tiny, and Newton
    // then creeps up the steep side of Q^beta for ~9 iterations; clamped into the bracket it needs ~3.
    const float ub = fminf(__fdividef(C32, dtx32), pow_fast(__fdividef(C32, a32), __fdividef(1.0f, b32)));
    if (ub > 0.0f && ub < 3.0e38f) Qf = fminf(fmaxf(Qf, 0.25f * ub), ub);
  }
  Qf = fmaxf(Qf, 1e-30f);
#pragma unroll 1
  for (int it = 0; it < 12; it++) {
    const float qb = pow_fast(Qf, b32);
    const float f = fmaf(dtx32, Qf, fmaf(a32, qb, -C32));
    const float df = fmaf(a32 * b32, __fdividef(qb, Qf), dtx32);
    const float Qn = fmaxf(Qf - __fdividef(f, df), 1e-30f);
    const bool done = fabsf(Qn - Qf) <= 1e-5f * Qn;
    Qf = Qn;
    if (done) break;
  }
  // float64 phase (Q >= 1e-30 and df > 0 are ordinary numbers: reciprocal-based quotients)
  double Qkx = (double)Qf;
  double qb = pow_el(Qkx, beta);
  const double t2 = 0.5 * beta * (beta - 1.0), t3 = beta * (beta - 1.0) * (beta - 2.0) * (1.0 / 6.0);
#pragma unroll 1
  for (int it = 0; it < 60; it++) {
    const double rQ = qrcp(Qkx);
    const double f = fma(deltaTX, Qkx, fma(alpha, qb, -C));
    const double dqb = beta * (qb * rQ);  // d(Q^beta)/dQ
    const double df = fma(alpha, dqb, deltaTX);
    const double Qn = fmax(Qkx - qdiv(f, df), 1e-30);
    const double d = (Qn - Qkx) * rQ;
    const double ad = fabs(d);
    // Q^beta at the new iterate
    if (ad < 1e-3) qb = qb * fma(d, fma(d, fma(d, t3, t2), beta), 1.0);
    else qb = pow_el(Qn, beta);
    Qkx = Qn;
    // Newton on this concave f converges quadratically, relative error after a step ~ 0.2 * (step)^2:
    // a relative step below 3e-5 leaves the new iterate within 2e-10 of the root
    if (fabs(f) <= 1e-12 || ad <= 3e-5) break;
  }
  Qbeta = qb;
  return Qkx;
}

// kinematic_wave_ssf, wflow/wflow_funcs.py:210-292.
// The reference runs Newton's method on the continuity residual
//     F(s) = dt/dx*(s - ssf_in) - dt*(r - (zi_old - zi(s))*neff*w),   zi(s) = -log(f*s/den + exp(-f*D))/f,
// from the poor first guess (ssf_old+ssf_in)/2 (up to ~23 iterations on the bench workload, and a
// level waits for its slowest cell), stops once |F| <= 1e-3 at the previous iterate and returns that
// iterate's zi.  Because |F| <= 1e-3 bounds the distance to the root, the result is pinned to the
// root itself: zi to 1e-3/(neff*w) ~ 4e-9 mm and s to 1e-3/F' mm^3, whatever the path.  So the same
// loop is started from a much better point: the root of F written in u = log(x) = -f*zi,
//     F(u) = A*e^u + B*u + C0   (convex, increasing),
// found by a few Halley steps from the previous water table u = -f*zi_old.  The reference loop then
// needs one or two iterations; exit tests, NaN guard, 1e-30 floor and the returned (stale) zi are the
// reference's.  exp(-f*zi) inside the loop is the log's own argument x (saves an exp; ~1e-14 rel).
#if WF_SSF_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void kinematic_wave_ssf(double ssf_in, double ssf_old, double zi_old, double r,
                                                double Ks_hor, double Ks, double slope, double neff, double f,
                                                double D, double efD, double dt, double dx, double w, double ssf_max,
                                                double &ssf_out, double &zi_out, double &exfilt_out) {
  const double epsilon = 1e-3;
  const int MAX_ITERS = 3000;
  if ((ssf_in + ssf_old) == 0.0 && (r <= 0)) {
    ssf_out = 0.0; zi_out = D; exfilt_out = 0.0;
    return;
  }
  if (zi_old >= D) {
    // A dry column.  The reference evaluates its residual at the first guess (ssf_old + ssf_in)/2 BEFORE the loop
    // (:224-244) and, when that is already within epsilon -- no inflow, (almost) no recharge: the guess sits at
    // zi = D like the state -- returns after that single update, with the zi of the guess.  The start from the
    // root below cannot reproduce this exit (with f*D ~ 100 the 1e-30 floor on ssf alone stands for a water table
    // metres above D), so this one evaluation is done literally, in the reference's own arithmetic.
    double s0 = (ssf_old + ssf_in) / 2.0;
    const double z0 = wf_div(wf_log(wf_div(f * s0, w * Ks_hor * Ks * slope) + efD), -f);
    const double Cn0 = wf_div(Ks_hor * Ks * wf_exp(-f * z0) * slope, neff);
    const double iC = wf_div(1.0, Cn0);
    const double c0 = wf_div(dt, dx) * ssf_in + iC * s0 + dt * (r - (zi_old - z0) * neff * w);
    const double fQ0 = wf_div(dt, dx) * s0 + iC * s0 - c0;
    if (!(fabs(fQ0) > epsilon)) {
      s0 = s0 - wf_div(fQ0, wf_div(dt, dx) + iC);
      if (isnan(s0)) s0 = 0.0;
      s0 = pymax(s0, 1e-30);
      s0 = pymin(s0, (ssf_max * w));
      const double rest0 = zi_old - wf_div(wf_div(ssf_in + r * dx - s0, w * dx), neff);
      exfilt_out = pymin(rest0, 0.0) * -neff;
      zi_out = pymax(0.0, z0);
      ssf_out = s0;
      return;
    }
  }
  // quotients of positive parameters (slope >= 1e-5, f > 0, Ks > 0, dx > 0): ordinary numbers
  const double den = w * Ks_hor * Ks * slope;
  const double dtdx = qdiv(dt, dx);
  const double fden = qdiv(f, den);
  const double nf = -qrcp(f);
  const double K1 = qdiv(neff, Ks_hor * Ks * slope);  // 1/Cn = K1 / exp(-f*zi)
  double ssf_n;
  {
    const double A = qdiv(dtdx, fden);
    const double B = qdiv(dt * neff * w, f);
    const double C0 = -A * efD - dtdx * ssf_in - dt * (r - zi_old * neff * w);
    double u = -f * fmin(fmax(zi_old, 0.0), D);
    double e = wf_exp(u);
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      const double d2 = A * e;  // (own first-guess iteration: explicit FMAs)
      const double Fu = fma(B, u, d2 + C0);
      const double d1 = d2 + B;
      const double du = qdiv(2.0 * Fu * d1, fma(2.0 * d1, d1, -Fu * d2));
      u = u - du;
      // e^u follows a small step by its series (|du| < 1/64: truncation below 1e-19)
      const double t = -du;
      if (fabs(du) < 0.015625)
        e = e * fma(t, fma(t, fma(t, fma(t, fma(t, fma(t, fma(t, 1.0 / 5040, 1.0 / 720), 1.0 / 120), 1.0 / 24), 1.0 / 6), 0.5), 1.0), 1.0);
      else
        e = wf_exp(u);
      if (!(fabs(du) > 1e-6)) break;  // cubic convergence: the step after one below 1e-6 would be below 1e-17
    }
    ssf_n = qdiv(e - efD, fden);
    if (!(ssf_n == ssf_n) || isinf(ssf_n)) ssf_n = (ssf_old + ssf_in) / 2.0;  // the reference's first guess
    ssf_n = pymax(ssf_n, 1e-30);
  }
  int count = 0;
  double zi, fQ;
  double x = ssf_n * fden + efD;  // exp(-f*zi) at the iterate (saves the reference's exp; ~1e-14 rel)
  double lx = wf_log(x);
  for (;;) {
    zi = lx * nf;
    const double invCn = qdiv(K1, x);
    const double c = dtdx * ssf_in + invCn * ssf_n + dt * (r - (zi_old - zi) * neff * w);
    fQ = dtdx * ssf_n + invCn * ssf_n - c;
    const double dfQ = dtdx + invCn;
    double sn = ssf_n - qdiv(fQ, dfQ);
    if (isnan(sn)) sn = 0.0;
    sn = pymax(sn, 1e-30);
    const bool stalled = fabs(sn - ssf_n) <= 4.440892098500626e-16 * sn;
    ssf_n = sn;
    count++;
    if (!(fabs(fQ) > epsilon && count <= MAX_ITERS && !stalled)) break;
    // log of the next iterate's x from this one's: log1p(d) by its series for a small relative step
    const double xn = sn * fden + efD;
    const double d = qdiv(xn - x, x);
    if (fabs(d) < 1e-4) lx = fma(d, fma(d, fma(d, fma(d, -0.25, 1.0 / 3), -0.5), 1.0), lx);
    else lx = wf_log(xn);
    x = xn;
  }
  ssf_n = pymin(ssf_n, (ssf_max * w));
  const double rest = zi_old - qdiv(qdiv(ssf_in + r * dx - ssf_n, w * dx), neff);
  exfilt_out = pymin(rest, 0.0) * -neff;
  zi_out = pymax(0.0, zi);
  ssf_out = ssf_n;
}

// Fraction of an upstream neighbour's outflow that a river cell takes straight into the channel
// (wflow_sbm.py:395-398, 843-846): slope-weighted for neighbours that drain in another direction.
// It depends on statics only and every cell has exactly one receiver, so k_derive_lateral stores it
// per UPSTREAM cell (d_cpd; -1 when the receiver is not a river cell) and the SENDER splits its
// outflow into the land share (1-cp)*Q and the channel share cp*Q -- the very products the reference
// forms in the receiver -- so the receiver's gather is two plain contiguous sums.
template <int ML>
__global__ void k_derive_lateral(const DevModel m) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.n) return;
  const uint8_t tp = __ldg(m.topo + i);
  const int nup = tp >> 4;
  const int myldd = tp & 15;
  const int64_t us = __ldg(m.up_start + i);
  const double slope_i = LD(landSlope);
  const bool riverMap = __ldg(m.river_slot + i) >= 0;
  for (int j = 0; j < nup; j++) {
    const double slnb = (double)__ldg(m.f[F_landSlope] + us + j);
    const int nbldd = __ldg(m.topo + us + j) & 15;
    m.d_cpd[us + j] = riverMap ? ((nbldd != myldd) ? slnb / (slope_i + slnb) : 0.0) : -1.0;
  }
  if (myldd == 5) m.d_cpd[i] = -1.0;  // a pit has no receiver
}

// ---- operands of the three tasks --------------------------------------------------------------
// Everything a task reads from its OWN cell (statics, its states, the column kernel's hand-offs) is
// known before the tick starts: the thread fetches it for its next item BEFORE the barrier of the
// running tick, so that after the barrier only the neighbours' fresh outflows remain to be read
// (one L2 round trip instead of a DRAM-deep dependent chain).  None of these values is written by
// another thread during the sweep.
#ifndef WF_GATHER_NEAR
#define WF_GATHER_NEAR 3  // upstream neighbours gathered without a branch
#endif

template <int ML>
struct SubOps {
  uint32_t us;
  int32_t rs;
  int nup;
  float slope, ST, thetaS, thetaR, KsatVer, f, Kh, DW, DL, zi, ssf_old, surplus, xl, yl;
  double rsum, efD, cpd;
  float Uhi[ML], Ulo[ML];
};
struct OvlOps {
  uint32_t us;
  int32_t rs;
  int nup;
  float SW, AlphaL, DL, Qold;
  double cpd, q64, qb64;
};
struct RivOps {
  uint32_t us;
  int nup;
  float DCL, A, Bw;
};

template <int ML>
__device__ __forceinline__ void load_sub(const DevModel &m, const int64_t i, SubOps<ML> &o) {
  const uint8_t tp = __ldg(m.topo + i);
  o.nup = tp >> 4;
  o.us = __ldg(m.up_start + i);
  o.rs = __ldg(m.river_slot + i);
  o.slope = LD(landSlope); o.ST = LD(SoilThickness); o.thetaS = LD(thetaS); o.thetaR = LD(thetaR);
  o.KsatVer = LD(KsatVer); o.f = LD(f); o.Kh = LD(KsatHorFrac); o.DW = LD(DW); o.DL = LD(DL);
  o.xl = LD(xl); o.yl = LD(yl);
  o.zi = m.f[F_zi][i];
  o.ssf_old = m.f[F_SubsurfaceFlow][i];
  o.surplus = m.h_surplus[i];
  o.rsum = m.h_r[i];
  o.efD = __ldg(m.d_efST + i);  // exp(-f * SoilThickness), derived static
  o.cpd = __ldg(m.d_cpd + i);
#pragma unroll
  for (int k = 0; k < ML; k++) {
    o.Uhi[k] = m.f[F_UStoreLayerDepth_0 + k][i];
    o.Ulo[k] = m.h_ulo[(int64_t)k * m.n + i];
  }
}
__device__ __forceinline__ void load_ovl(const DevModel &m, const int64_t i, OvlOps &o) {
  const uint8_t tp = __ldg(m.topo + i);
  o.nup = tp >> 4;
  o.us = __ldg(m.up_start + i);
  o.rs = __ldg(m.river_slot + i);
  o.SW = LD(SW); o.AlphaL = LD(AlphaL); o.DL = LD(DL);
  o.Qold = m.f[F_LandRunoffsub][i];
  o.cpd = __ldg(m.d_cpd + i);
  o.q64 = m.h_lq[i];
  o.qb64 = m.h_lqb[i];
}
__device__ __forceinline__ void load_riv(const DevModel &m, const int32_t rs, RivOps &o) {
  o.nup = __ldg(m.r_nup + rs);
  o.us = __ldg(m.r_up_start + rs);
  o.DCL = __ldg(m.f[F_DCL] + rs);
  o.A = __ldg(m.f[F_AlphaR] + rs);
  o.Bw = __ldg(m.f[F_Bw] + rs);
}

// ---- operand staging through shared memory (cluster / single-block sweeps) ------------------------
// The operand sets above are DRAM-cold (every array is touched once per timestep) and the 96/128-register sweeps
// cannot hold them across the barrier.  So the thread that will run an item next tick has the item's operands
// copied into its own shared-memory column by cp.async -- no registers, no waiting -- BEFORE it waits at the
// barrier; after the barrier they are a shared-memory read away and the tick's critical path starts with the
// gather of the neighbours' fresh outflows instead of a DRAM round trip (r1 tick profile: 1.2 us of 7.9).
// Layout: word w of thread x at sw[w * T + x] (conflict-free), doubles likewise behind the words.
// Per-thread bytes of one staging column: 16 + 2 ML operand words, the word of the topology byte array that holds the
// item's upstream count, two words for the item itself (sweep_group_ahead), three float64 operands.
__host__ __device__ constexpr int wf_stage_words(int ML) { return 16 + 2 * ML + 3; }
// columns per thread: two where they fit beside the level cache (the next item's operands land while the current
// item's are still being read); with one, an item's operands move to registers before the next item's are requested
__host__ __device__ constexpr int wf_stage_nbuf(int ML) { return (WF_STAGE_AHEAD && ML <= 5) ? 2 : 1; }
__host__ __device__ constexpr size_t wf_stage_bytes(int ML, int threads) {
  return (size_t)threads * (size_t)wf_stage_nbuf(ML) * (size_t)(wf_stage_words(ML) * 4 + 3 * 8);
}
template <int ML>
struct StageLayout {
  static constexpr int NW = wf_stage_words(ML), ND = 3, NBUF = wf_stage_nbuf(ML);  // 4-byte words, 8-byte words per thread and column
  static constexpr int W_TOPO = NW - 3, W_IDX = NW - 2, W_TYPE = NW - 1;
  __host__ __device__ static constexpr size_t bytes(int threads) { return wf_stage_bytes(ML, threads); }
};
__device__ __forceinline__ void cp_async4(void *sdst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *sdst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int ML>
__device__ __forceinline__ void stage_sub(const DevModel &m, const int64_t i, float *sw, double *sd, const int T) {
  const int x = threadIdx.x;
  int w = 0;
#define WF_ST4(src) cp_async4(sw + (w++) * T + x, (src))
  WF_ST4(m.up_start + i); WF_ST4(m.river_slot + i);
  WF_ST4(m.f[F_landSlope] + i); WF_ST4(m.f[F_SoilThickness] + i); WF_ST4(m.f[F_thetaS] + i); WF_ST4(m.f[F_thetaR] + i);
  WF_ST4(m.f[F_KsatVer] + i); WF_ST4(m.f[F_f] + i); WF_ST4(m.f[F_KsatHorFrac] + i); WF_ST4(m.f[F_DW] + i);
  WF_ST4(m.f[F_DL] + i); WF_ST4(m.f[F_xl] + i); WF_ST4(m.f[F_yl] + i); WF_ST4(m.f[F_zi] + i);
  WF_ST4(m.f[F_SubsurfaceFlow] + i); WF_ST4(m.h_surplus + i);
#pragma unroll
  for (int k = 0; k < ML; k++) { WF_ST4(m.f[F_UStoreLayerDepth_0 + k] + i); WF_ST4(m.h_ulo + (int64_t)k * m.n + i); }
  cp_async8(sd + 0 * T + x, m.h_r + i); cp_async8(sd + 1 * T + x, m.d_efST + i); cp_async8(sd + 2 * T + x, m.d_cpd + i);
}
template <int ML>
__device__ __forceinline__ void unstage_sub(const DevModel &m, const int64_t i, const float *sw, const double *sd, const int T,
                                            SubOps<ML> &o) {
  const int x = threadIdx.x;
  int w = 0;
#define WF_UF (sw[(w++) * T + x])
  o.us = __float_as_uint(WF_UF); o.rs = __float_as_int(WF_UF);
  o.slope = WF_UF; o.ST = WF_UF; o.thetaS = WF_UF; o.thetaR = WF_UF; o.KsatVer = WF_UF; o.f = WF_UF; o.Kh = WF_UF;
  o.DW = WF_UF; o.DL = WF_UF; o.xl = WF_UF; o.yl = WF_UF; o.zi = WF_UF; o.ssf_old = WF_UF; o.surplus = WF_UF;
#pragma unroll
  for (int k = 0; k < ML; k++) { o.Uhi[k] = WF_UF; o.Ulo[k] = WF_UF; }
  o.rsum = sd[0 * T + x]; o.efD = sd[1 * T + x]; o.cpd = sd[2 * T + x];
}
__device__ __forceinline__ void stage_ovl(const DevModel &m, const int64_t i, float *sw, double *sd, const int T) {
  const int x = threadIdx.x;
  int w = 0;
  WF_ST4(m.up_start + i); WF_ST4(m.river_slot + i); WF_ST4(m.f[F_SW] + i); WF_ST4(m.f[F_AlphaL] + i); WF_ST4(m.f[F_DL] + i);
  WF_ST4(m.f[F_LandRunoffsub] + i);
  cp_async8(sd + 0 * T + x, m.d_cpd + i); cp_async8(sd + 1 * T + x, m.h_lq + i); cp_async8(sd + 2 * T + x, m.h_lqb + i);
}
__device__ __forceinline__ void unstage_ovl(const DevModel &m, const int64_t i, const float *sw, const double *sd, const int T, OvlOps &o) {
  const int x = threadIdx.x;
  int w = 0;
  o.us = __float_as_uint(WF_UF); o.rs = __float_as_int(WF_UF);
  o.SW = WF_UF; o.AlphaL = WF_UF; o.DL = WF_UF; o.Qold = WF_UF;
  o.cpd = sd[0 * T + x]; o.q64 = sd[1 * T + x]; o.qb64 = sd[2 * T + x];
}
__device__ __forceinline__ void stage_riv(const DevModel &m, const int32_t rs, float *sw, const int T) {
  const int x = threadIdx.x;
  int w = 0;
  WF_ST4(m.r_up_start + rs); WF_ST4(m.f[F_DCL] + rs); WF_ST4(m.f[F_AlphaR] + rs); WF_ST4(m.f[F_Bw] + rs);
}
__device__ __forceinline__ void unstage_riv(const DevModel &m, const int32_t rs, const float *sw, const int T, RivOps &o) {
  const int x = threadIdx.x;
  int w = 0;
  o.us = __float_as_uint(WF_UF); o.DCL = WF_UF; o.A = WF_UF; o.Bw = WF_UF;
}
#undef WF_ST4
#undef WF_UF

// Arrival at the cluster barrier from inside a task (divergent code: the per-thread form of the instruction).
// A task of the cluster sweeps arrives as soon as what the NEXT tick reads is stored -- the outflow its
// downstream neighbour gathers -- and does the rest of its work (column tail, step means, the channel's lateral
// inflow) behind the arrival, while the barrier collects the other warps: r1 profile, 28 % of all stall cycles were
// warps waiting at the barrier for the subsurface tails, and the barrier itself takes 1.7 us from the last arrival.
// What a task stores behind its arrival is released by the thread's NEXT arrival, so its readers run two ticks
// later (sweep_group's schedule).
__device__ __forceinline__ void wave_arrive_early() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }

// ---- task 1: subsurface flow + post-ssf column tail of one cell ------------------------------
// DIAG = false compiles the diagnostic outputs out (and with them the values they keep alive): the production sweep
// EARLY (cluster sweeps): `signal` = this is the thread's last item of the tick -> arrive at the barrier once the outflow is stored
template <int ML, int MODE, bool DIAG = true, bool EARLY = false>
__device__ __forceinline__ void task_subsurface(const DevModel &m, const int64_t i, const SubOps<ML> &o, const bool signal WF_PARGS) {
  const int nup = o.nup;
  const uint32_t us = o.us;
  const int32_t rs = o.rs;
  const bool riverMap = rs >= 0;
  const double slope_i = o.slope;
  const double ST = o.ST, thetaS = o.thetaS, thetaR = o.thetaR;
  const float neff_f = o.thetaS - o.thetaR;
  const double KsatVer = o.KsatVer, fpar = o.f, Kh = o.Kh;
  const double DW = o.DW, DLd = o.DL;
  const double zi_old = o.zi;
  const double rsum = o.rsum;
  float *const ssfp = m.f[F_SubsurfaceFlow];
  const double ssf_old = o.ssf_old;
  const double surplus = o.surplus;
  const double xlyl = (double)o.xl * (double)o.yl;
  double U[ML];
#pragma unroll
  for (int k = 0; k < ML; k++) U[k] = (double)o.Uhi[k] + (double)o.Ulo[k];

  WF_STAMP(0, U[0] + xlyl + ssf_old + rsum + surplus + DW + Kh + zi_old);
  // inflow gather, reference summation order (wflow_sbm.py:395-410)
  double ssf_in = 0.0, ssf_tr = 0.0;
  // All loads of a cell in flight at once.  Most cells have 0-3 upstream neighbours, so the loads of neighbours 3..7 sit
  // behind a branch that a warp only takes when one of its cells needs them (predicated-off loads still cost an issue
  // slot each: the 8-wide gathers were 10 % of the sweep's instructions).
  double sv[8], tv[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { sv[j] = 0.0; tv[j] = 0.0; }
#pragma unroll
  for (int j = 0; j < WF_GATHER_NEAR; j++) {
    if (j < nup) {
      sv[j] = ldx<MODE>(m.h_ssf + (int64_t)us + j);
      if (riverMap) tv[j] = ldx<MODE>(m.h_ssf_riv + (int64_t)us + j);
    }
  }
  if (nup > WF_GATHER_NEAR) {
#pragma unroll
    for (int j = WF_GATHER_NEAR; j < 8; j++) {
      if (j < nup) {
        sv[j] = ldx<MODE>(m.h_ssf + (int64_t)us + j);
        if (riverMap) tv[j] = ldx<MODE>(m.h_ssf_riv + (int64_t)us + j);
      }
    }
  }
  // reference summation order (wflow_sbm.py:395-410); the absent neighbours add +0.0, which changes nothing
#pragma unroll
  for (int j = 0; j < WF_GATHER_NEAR; j++) {
    ssf_in = ssf_in + sv[j];  // (1 - chanperc) * ssf of the neighbour (the whole of it for a land receiver)
    ssf_tr = ssf_tr + tv[j];  // chanperc * ssf
  }
  if (nup > WF_GATHER_NEAR) {
#pragma unroll
    for (int j = WF_GATHER_NEAR; j < 8; j++) {
      ssf_in = ssf_in + sv[j];
      ssf_tr = ssf_tr + tv[j];
    }
  }
  if (riverMap) m.h_ssf_tr[rs] = (float)qdiv(ssf_tr / (1000 * 1000 * 1000), m.dt);

  WF_STAMP(1, ssf_in);
  const double neff = (double)neff_f;
  const double dneff = thetaS - thetaR;
  const double r = rsum * DW * 1000;                                   // :674
  const double efD = o.efD;
  const double ssfmax = qdiv(Kh * KsatVer * slope_i, fpar) * (1.0 - efD);  // :2291-2299
  double ssf_n, zi_n, ExfiltSatWater;
  kinematic_wave_ssf(ssf_in, ssf_old, zi_old, r, Kh, KsatVer, slope_i, neff, fpar, ST, efD, 1.0, DLd * 1000,
                     DW * 1000, ssfmax, ssf_n, zi_n, ExfiltSatWater);
  WF_STAMP(2, ssf_n);
  const double zi = pymin(zi_n, ST);  // :702
  const double Sat = (ST - zi) * dneff;
  ssfp[i] = (float)ssf_n;
  // gathered by the downstream cell one tick later; a river receiver gets the two shares
  if (o.cpd >= 0.0) {
    m.h_ssf[i] = (1 - o.cpd) * ssf_n;
    m.h_ssf_riv[i] = o.cpd * ssf_n;
  } else {
    m.h_ssf[i] = ssf_n;
  }
  if (EARLY && signal) wave_arrive_early();  // the downstream cell's gather is all the next tick needs
  m.f[F_zi][i] = (float)zi;

  // post-ssf column tail, :707-821
  double thick[ML], tops[ML + 1];
  layer_geometry<ML>(m, ST, thick, tops);
  const int mc_old = unsat_layers<ML>(zi_old, tops);
  const int nL = (mc_old > 1) ? mc_old : 1;
  const int mc_new = unsat_layers<ML>(zi, tops);
  const int nLn = (mc_new > 1) ? mc_new : 1;
  double L_new[ML];
#pragma unroll
  for (int k = 0; k < ML; k++)
    L_new[k] = (k < nLn - 1) ? thick[k] : ((k == nLn - 1) ? ((mc_new > 1) ? zi - tops[k] : zi) : 0.0);
  double ExfiltFromUstore = 0.0;
  bool changed = false;
#pragma unroll
  for (int k = ML - 1; k >= 0; k--) {
    if (k < nL) {
      if (k < mc_new) ExfiltFromUstore = pymax(0.0, U[k] - L_new[k] * dneff);
      else ExfiltFromUstore = U[k];
      changed |= (ExfiltFromUstore != 0.0);
      U[k] = U[k] - ExfiltFromUstore;
      if (k > 0) U[k - 1] = U[k - 1] + ExfiltFromUstore;
    }
  }
  if (changed) {
#pragma unroll
    for (int k = 0; k < ML; k++) m.f[F_UStoreLayerDepth_0 + k][i] = (float)U[k];
  }
  const double ExfiltWater = ExfiltSatWater + ExfiltFromUstore;
  // a paddy field keeps what would run off, up to its bund (:762-772): the pond takes it before the overland wave sees it
  double ponding_add = 0.0;
  if (DIAG && m.paddy) {
    const double hp = (double)__ldg(m.f[F_h_p] + i);
    if (hp > 0) {
      const double pond = (double)m.f[F_PondingDepth][i];
      ponding_add = pymin(ExfiltWater + (double)m.h_pondsrc[i], hp - pond);
      m.f[F_PondingDepth][i] = (float)(pond + ponding_add);
    }
  }
  // InwaterO :774 -- ExfiltWater + [ExcessWater + InfiltExcess + RunoffLandCells - ActEvapOpenWaterLand] - ponding_add
  const double InwaterO = qdiv(pymax(ExfiltWater + surplus - ponding_add, 0.0) * xlyl * 0.001, m.dt);
  m.h_inwO[i] = InwaterO;

  double sumU = 0.0;
#pragma unroll
  for (int k = 0; k < ML; k++) sumU += U[k];
  if (DIAG && m.any_diag) {
  OUTF(SatWaterDepth, Sat); OUTF(UStoreDepth, sumU);
  OUTF(ssf_toriver, riverMap ? qdiv(ssf_tr / (1000 * 1000 * 1000), m.dt) : 0.0);
  OUTF(ExfiltWater, ExfiltWater); OUTF(ExfiltFromUstore, ExfiltFromUstore); OUTF(ExfiltFromSat, ExfiltSatWater);
  OUTF(InwaterL, InwaterO); OUTF(CellInFlow, ssf_in);
#pragma unroll
  for (int k = 0; k < ML; k++) {
    float *vw = m.f[F_vwc_0 + k];
    if (vw) {  // :792-807 (layer["vwc"] keeps its value where neither branch writes)
      double vwc = vw[i];
      if (k < mc_new) {
        if (thick[k] > 0) vwc = qdiv(U[k] + (thick[k] - L_new[k]) * dneff, thick[k]) + thetaR;
      } else {
        vwc = thetaS;
      }
      vw[i] = (float)vwc;
      if (m.f[F_vwc_perc_0 + k]) m.f[F_vwc_perc_0 + k][i] = (float)(qdiv(vwc, thetaS) * 100.0);
    }
  }
  if (m.f[F_RootStore_unsat]) {  // :809-821
    const double ActRoot = (double)fminf(LD(SoilThickness) * 0.99f, LD(RootingDepth));
    double rsu = 0.0;
#pragma unroll
    for (int k = 0; k < ML; k++)
      if (k < nLn && L_new[k] > 0) rsu = rsu + qdiv(pymax(0.0, ActRoot - tops[k]), L_new[k]) * U[k];
    m.f[F_RootStore_unsat][i] = (float)rsu;
  }
  if (m.f[F_SoilWatbal] && m.h_wb) {  // :885-898; h_wb holds the cell-local terms (vertical kernel)
    const double wb = m.h_wb[i] - (Sat + sumU) + qdiv(ssf_in - ssf_n, DW * DLd * 1000 * 1000) - ExfiltWater;
    m.f[F_SoilWatbal][i] = (float)wb;
    if (m.f[F_watbal] && m.f[F_SurfaceWatbal]) m.f[F_watbal][i] = (float)wb + m.f[F_SurfaceWatbal][i];
  }
  }
  WF_STAMP(3, InwaterO);
}

// ---- task 2: overland flow of one cell, all sub-steps (wflow_sbm.py:830-879) -----------------
template <int MODE, bool DIAG = true, bool EARLY = false>
__device__ __forceinline__ void task_overland(const DevModel &m, const int64_t i, const OvlOps &o, const bool signal WF_PARGS) {
  const int nup = o.nup;
  const uint32_t us = o.us;
  const int32_t rs = o.rs;
  const bool riverMap = rs >= 0;
  const int itL = m.itL;
  const double SW = o.SW, AlphaL = o.AlphaL, Beta = m.beta, DLd = o.DL;
  const double dts = qdiv(m.dt, (double)itL);
#if WF_OVL_TAIL_PREFETCH
  // What the channel's lateral inflow (behind the solve) reads of a river cell is DRAM-cold: requested now, it arrives
  // while the wave is solved (nine in ten warps of a level hold a river cell and would wait for it in the tail).
  if (riverMap) {
    prefetch_l1(m.f[F_xl] + i); prefetch_l1(m.f[F_yl] + i); prefetch_l1(m.h_inwaterMM + rs);
    if (m.obj_inflow) prefetch_l1(m.obj_inflow + rs);
    else if (m.f[F_Inflow]) prefetch_l1(m.f[F_Inflow] + rs);
  }
#endif
  const double q = qdiv(ldx<MODE>(m.h_inwO + i), DLd);
  double Qold = o.Qold;
  double QoldBeta = carried_pow(Qold, o.q64, o.qb64, Beta);
  double acc_flow = 0.0, acc_h = 0.0, Qo_in_acc = 0.0, qo_toriver_acc = 0.0;
  double wlsub = 0.0;  // dyn["WaterLevelLsub"] starts at 0 and is only written where SW > 0 (:873-878)
#pragma unroll 1
  for (int v = 0; v < itL; v++) {
    float *qbuf = (v == itL - 1) ? m.f[F_LandRunoffsub] : m.qo_sub + (int64_t)v * m.n;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    float qv[8];
    double av[8], bv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { qv[j] = 0.0f; av[j] = 0.0; bv[j] = 0.0; }
    const double *const qa = m.h_qa + (int64_t)v * m.n + us, *const qb = m.h_qb + (int64_t)v * m.n + us;
#pragma unroll
    for (int j = 0; j < WF_GATHER_NEAR; j++) {
      if (j < nup) {
        qv[j] = ldx<MODE>(qbuf + (int64_t)us + j);
        if (riverMap) { av[j] = ldx<MODE>(qa + j); bv[j] = ldx<MODE>(qb + j); }
      }
    }
    if (nup > WF_GATHER_NEAR) {
#pragma unroll
      for (int j = WF_GATHER_NEAR; j < 8; j++) {
        if (j < nup) {
          qv[j] = ldx<MODE>(qbuf + (int64_t)us + j);
          if (riverMap) { av[j] = ldx<MODE>(qa + j); bv[j] = ldx<MODE>(qb + j); }
        }
      }
    }
    // sums in the reference's neighbour order; absent neighbours add +0.0
    if (riverMap) {
#pragma unroll
      for (int j = 0; j < WF_GATHER_NEAR; j++) {
        s1 = s1 + av[j];  // (1 - chanperc) * qo of the neighbour
        s2 = s2 + bv[j];  // chanperc * qo
        s3 = s3 + (double)qv[j];
      }
      if (nup > WF_GATHER_NEAR) {
#pragma unroll
        for (int j = WF_GATHER_NEAR; j < 8; j++) { s1 = s1 + av[j]; s2 = s2 + bv[j]; s3 = s3 + (double)qv[j]; }
      }
    } else {
#pragma unroll
      for (int j = 0; j < WF_GATHER_NEAR; j++) s1 = s1 + (double)qv[j];
      if (nup > WF_GATHER_NEAR) {
#pragma unroll
        for (int j = WF_GATHER_NEAR; j < 8; j++) s1 = s1 + (double)qv[j];
      }
    }
    double qo_in, qo_toriver_vol;
    if (riverMap) {
      if (SW > 0.0) { qo_in = s1; qo_toriver_vol = s2 * dts; }
      else { qo_in = 0.0; qo_toriver_vol = s3 * dts; }
    } else {
      qo_in = s1; qo_toriver_vol = 0.0;
    }
    double qbeta;
    WF_STAMP(0, q + Qold + AlphaL + SW);
    WF_STAMP(1, qo_in);
    const double qn = kinematic_wave(qo_in, Qold, QoldBeta, q, AlphaL, Beta, dts, DLd, qbeta);
    WF_STAMP(2, qn);
    acc_flow = acc_flow + qn * dts;
    Qo_in_acc = Qo_in_acc + qo_in * dts;
    qo_toriver_acc = qo_toriver_acc + qo_toriver_vol;
    if (SW > 0) wlsub = qdiv(AlphaL * qbeta, SW);
    acc_h = acc_h + wlsub;
    const float qn_f = (float)qn;
    qbuf[i] = qn_f;
    if (v == itL - 1) {  // for the next timestep's constant term
      m.h_lq[i] = qn;
      m.h_lqb[i] = qbeta;
    }
    Qold = qn;          // (the reference keeps float64 between sub-steps, wflow_sbm.py:879)
    QoldBeta = qbeta;
    if (o.cpd >= 0.0) {  // the receiver is a river cell: hand it the two shares
      m.h_qa[(int64_t)v * m.n + i] = (1 - o.cpd) * (double)qn_f;
      m.h_qb[(int64_t)v * m.n + i] = o.cpd * (double)qn_f;
    }
  }
  if (EARLY && signal) wave_arrive_early();  // every sub-step's outflow is stored: the step means and the channel's inflow follow
  m.f[F_WaterLevelLsub][i] = (float)wlsub;
  m.f[F_LandRunoff][i] = (float)qdiv(acc_flow, m.dt);
  m.f[F_WaterLevelL][i] = (float)qdiv(acc_h, (double)itL);
  const float qo_toriver_f = (float)qdiv(qo_toriver_acc, m.dt);
  if (DIAG && m.any_diag) {
    OUTF(qo_toriver, qo_toriver_f);
    OUTF(InflowKinWaveCellLand, qdiv(Qo_in_acc, m.dt));
  }

  if (riverMap) {  // lateral inflow of the channel, wflow_sbm.py:3052-3065, 3144
    const float dtf = (float)m.dt;
    const float ToCubic = ((LD(xl) * LD(yl)) * 0.001f) / dtf;  // :1998
    float Inwater = ((m.h_inwaterMM[rs] * ToCubic) + qo_toriver_f) + ldx<MODE>(m.h_ssf_tr + rs);
    // self.Inflow: the external map, plus the outflow of the reservoirs / lakes above the cell (:2852, 2883, 3144)
    if (m.obj_inflow || m.f[F_Inflow]) {
      const float inflow = m.obj_inflow ? __ldg(m.obj_inflow + rs) : m.f[F_Inflow][rs];
      float take = inflow;
      if (m.sup_sws) {  // abstractions enabled (some Inflow may be negative): first estimate of the supply, :3131-3146
        float sws = 0.0f;
        if (inflow < 0.0f) {
          // pcr.upstream(TopoLdd, RiverRunoff) of the step before: the river cells draining into this one are solved
          // for this step ticks later, so their discharge still is the old one here
          float up = 0.0f;
          if (rs < m.nrnet) {
            const int rn = __ldg(m.r_nup + rs);
            const uint32_t ru = __ldg(m.r_up_start + rs);
            for (int j = 0; j < rn; j++) up = up + m.f[F_RiverRunoff][ru + j];
          }
          sws = fminf(up + Inwater, -1.0f * inflow);
          if (sws > 0.0f) take = -1.0f * sws;
        }
        m.sup_sws[rs] = sws;
        m.sup_old_inwater[rs] = Inwater;
        if (inflow < 0.0f) *m.sup_any_neg = 1;
      }
      Inwater = Inwater + take;
    }
    m.h_inwater[rs] = Inwater;
    if (m.f[F_Inwater]) m.f[F_Inwater][rs] = Inwater;
    if (rs >= m.nrnet) {  // river-map cell outside the river network: kin_wave never visits it (zeros)
      m.f[F_RiverRunoff][rs] = 0.0f; m.f[F_WaterLevelR][rs] = 0.0f;
      m.f[F_RiverRunoffsub][rs] = 0.0f; m.f[F_WaterLevelRsub][rs] = 0.0f;
    }
  }
  WF_STAMP(3, acc_flow);
}

// ---- task 3: one sub-step of the river kinematic wave for one river cell (river order) -------
// kin_wave, wflow/wflow_funcs.py:155-207; caller glue wflow_sbm.py:3152-3201.
template <int MODE>
__device__ __forceinline__ void task_river(const DevModel &m, const int32_t rs, const int v, const RivOps &o, const int step WF_PARGS) {
  const int itR = m.itR;
  const int nup = o.nup;
  const uint32_t us = o.us;
  float *qbuf = (v == itR - 1) ? m.f[F_RiverRunoffsub] : m.qr_sub + (int64_t)v * m.nr;
  const float DCLf = o.DCL;
  const float inw = ldx<MODE>(m.h_inwater + rs);
  const double A = o.A, Bw = o.Bw;
  const double Qo = (v == 0) ? (double)m.f[F_RiverRunoffsub][rs] : (double)ldx<MODE>(m.qr_sub + (int64_t)(v - 1) * m.nr + rs);
  double racc0 = 0.0, rh0 = 0.0;
  if (v > 0) { racc0 = ldx<MODE>(m.h_racc + rs); rh0 = ldx<MODE>(m.h_rh + rs); }
  double Qin = 0.0;
  {
    float qv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) qv[j] = 0.0f;
#pragma unroll
    for (int j = 0; j < WF_GATHER_NEAR; j++)
      if (j < nup) qv[j] = ldx<MODE>(qbuf + us + j);
    if (nup > WF_GATHER_NEAR) {
#pragma unroll
      for (int j = WF_GATHER_NEAR; j < 8; j++)
        if (j < nup) qv[j] = ldx<MODE>(qbuf + us + j);
    }
#pragma unroll
    for (int j = 0; j < WF_GATHER_NEAR; j++) Qin = Qin + (double)qv[j];
    if (nup > WF_GATHER_NEAR) {
#pragma unroll
      for (int j = WF_GATHER_NEAR; j < 8; j++) Qin = Qin + (double)qv[j];
    }
  }
  const double qr = (double)(inw / DCLf);  // :3152 (REAL4 division)
  const double dtr = qdiv(m.dt, (double)itR);
  // Qo: the previous sub-step's outflow of this cell, or the state at the first sub-step (:199)
  double qbeta;
  WF_STAMP(0, qr + A + Bw + Qo + racc0);
  WF_STAMP(1, Qin);
  const double QoBeta = carried_pow(Qo, ldx<MODE>(m.h_rq + rs), ldx<MODE>(m.h_rqb + rs), m.beta);
  const double qn = kinematic_wave(Qin, Qo, QoBeta, qr, A, m.beta, dtr, (double)DCLf, qbeta);
  m.h_rq[rs] = qn;
  m.h_rqb[rs] = qbeta;
  WF_STAMP(2, qn);
  const double wl = qdiv(A * qbeta, Bw);
  const double racc = racc0 + qn * dtr, rh = rh0 + wl;
  qbuf[rs] = (float)qn;
  if (v == itR - 1) {
    const float rr = (float)qdiv(racc, m.dt);
    m.f[F_RiverRunoff][rs] = rr;
    if (m.cap_slot) {  // pipelined sweep: the step's discharge at a gauge, before the next step overwrites it
      const int32_t g = __ldg(m.cap_slot + rs);
      if (g >= 0) m.cap_out[(int64_t)(step + m.cap_step) * m.cap_n + g] = rr;
    }
    m.f[F_WaterLevelR][rs] = (float)qdiv(rh, (double)itR);
    m.f[F_WaterLevelRsub][rs] = (float)wl;
  } else {
    m.h_racc[rs] = racc;
    m.h_rh[rs] = rh;
  }
  WF_STAMP(3, racc);
}

// The first sub-step reads its own cell's Qold from the state array that the LAST sub-step
// overwrites: with it_kinR sub-steps in flight on different levels that is safe, because a cell's
// sub-step v runs at tick level+2+v, strictly after its sub-step v-1 and before anyone reads v.

// ---- the wavefront -------------------------------------------------------------------------------

// Pull the operands of the cell this thread will most likely solve next tick into L2 while the
// current tick is still waiting at the barrier (they are cold DRAM lines otherwise: every array is
// touched once per timestep).
__device__ __forceinline__ void prefetch_subsurface(const DevModel &m, const int64_t i) {
  prefetch_l2(m.topo + i); prefetch_l2(m.up_start + i); prefetch_l2(m.river_slot + i);
  prefetch_l2(m.h_r + i); prefetch_l2(m.h_surplus + i); prefetch_l2(m.f[F_zi] + i); prefetch_l2(m.f[F_SubsurfaceFlow] + i);
  prefetch_l2(m.f[F_landSlope] + i); prefetch_l2(m.f[F_SoilThickness] + i); prefetch_l2(m.f[F_thetaS] + i);
  prefetch_l2(m.f[F_thetaR] + i); prefetch_l2(m.f[F_KsatVer] + i); prefetch_l2(m.f[F_f] + i);
  prefetch_l2(m.f[F_KsatHorFrac] + i); prefetch_l2(m.f[F_DW] + i); prefetch_l2(m.f[F_DL] + i);
  prefetch_l2(m.f[F_xl] + i); prefetch_l2(m.f[F_yl] + i);
  for (int k = 0; k < m.ML; k++) { prefetch_l2(m.f[F_UStoreLayerDepth_0 + k] + i); prefetch_l2(m.h_ulo + (int64_t)k * m.n + i); }
  prefetch_l2(m.d_efST + i); prefetch_l2(m.d_cpd + i);
}
__device__ __forceinline__ void prefetch_overland(const DevModel &m, const int64_t i) {
  prefetch_l2(m.topo + i); prefetch_l2(m.up_start + i); prefetch_l2(m.river_slot + i); prefetch_l2(m.d_cpd + i);
  prefetch_l2(m.f[F_SW] + i); prefetch_l2(m.f[F_AlphaL] + i); prefetch_l2(m.f[F_LandRunoffsub] + i); prefetch_l2(m.f[F_DL] + i);
  prefetch_l2(m.h_lq + i); prefetch_l2(m.h_lqb + i);
}
// the same for a river cell that the first river sub-step reaches soon (the later sub-steps find the lines in L2)
__device__ __forceinline__ void prefetch_river(const DevModel &m, const int64_t rs) {
  prefetch_l2(m.r_up_start + rs); prefetch_l2(m.r_nup + rs); prefetch_l2(m.f[F_DCL] + rs); prefetch_l2(m.f[F_AlphaR] + rs);
  prefetch_l2(m.f[F_Bw] + rs); prefetch_l2(m.f[F_RiverRunoffsub] + rs); prefetch_l2(m.h_rq + rs); prefetch_l2(m.h_rqb + rs);
}

// Who sweeps a group (a set of whole basins, GroupLayout in network.h), and with which barrier:
//   WF_GRID     all blocks of a cooperative launch, grid barrier     (one wide group)
//   WF_CLUSTER  one thread-block cluster per group, cluster barrier  (several basins, wide levels)
//   WF_CTA      one thread block per group, __syncthreads            (many small basins / a narrow network)
// Groups never exchange data, so clusters / blocks of different groups never wait for each other.
enum { WF_GRID = 0, WF_CLUSTER = 1, WF_CTA = 2, WF_CLUSTER_WIDE = 3 };
// WF_CLUSTER_WIDE: the cluster sweep with WF_GROUP_THREADS_WIDE threads per block (fewer registers per
// thread), chosen when it lets every item of a tick have its own thread
#ifndef WF_GROUP_THREADS_WIDE
#define WF_GROUP_THREADS_WIDE 640
#endif
__host__ __device__ constexpr bool wf_is_cluster(int mode) { return mode == WF_CLUSTER || mode == WF_CLUSTER_WIDE; }
__host__ __device__ constexpr int wf_block_threads(int mode) {
  return mode == WF_GRID ? 256 : (mode == WF_CLUSTER_WIDE ? WF_GROUP_THREADS_WIDE : WF_GROUP_THREADS);
}

// Split barrier of a sweep: arrive once the tick's outflows are stored, fetch the next item's own
// operands while the others arrive, then wait.
//   grid     per-launch counter in global memory (zeroed before the launch): one atomic per block
//   cluster  hardware cluster barrier, arrive.release / wait.acquire
//   block    __syncthreads at the wait (loads in flight cross it)
struct WaveSync {
  unsigned int *counter;
  unsigned int nblocks, target;
};
template <int MODE>
__device__ __forceinline__ void wave_arrive(WaveSync &ws) {
  if (MODE == WF_GRID) {
    __syncthreads();
    ws.target += ws.nblocks;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ws.counter, 1u);
    }
  } else if (wf_is_cluster(MODE)) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  }
}
template <int MODE>
__device__ __forceinline__ void wave_wait(WaveSync &ws) {
  if (MODE == WF_GRID) {
    if (threadIdx.x == 0) {
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ws.counter) : "memory");
      } while ((int)(seen - ws.target) < 0);
    }
    __syncthreads();
  } else if (wf_is_cluster(MODE)) {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

// One group's wavefront.  Tick t (level L = lev0 + t) runs, concurrently on different warps,
//   [subsurface cells of level L][overland cells of level L-1][river cells of level L-2-v, sub-step v = 0..itR-1]
// as one flat item space: consecutive threads take consecutive items, so a warp almost always runs
// a single task type.  lo / ro: the group's level offsets (shared memory when they fit).
struct WaveItem {
  int type;      // -1 none, 0 subsurface, 1 overland, 2 river
  uint32_t idx;  // storage position (river order for type 2)
  int v;         // river sub-step
  int nup;       // staged sweeps: number of upstream neighbours, fetched ahead of the barrier with the operands (a byte array:
                 // not a cp.async size, and the gather of the next tick starts with it)
};

// Schedule of a group's sweep (local tick t; level = group-local level index):
//     subsurface flow of level t, overland flow of level t-2, river sub-step v of the river level that is full level t-4-v.
// The two-tick distances are what the early arrival needs (wave_arrive_early): InwaterO leaves the subsurface
// task behind its arrival and reaches the overland task of the same cell two ticks later; the channel's lateral
// inflow likewise from the overland task to the first river sub-step.  D + 3 + it_kinR ticks.
__host__ __device__ constexpr int wf_sweep_ticks(int nl, bool rivers, int itR) { return rivers ? nl + 3 + itR : nl + 2; }
// river level (group-local) that sub-step 0 solves at tick 0; sub-step v lags v levels behind
__host__ __device__ constexpr int wf_sweep_q0(int lev0, int rlev_shift, int rlev0) { return (lev0 - 4) - rlev_shift - rlev0; }
#define WF_OVL_LAG 2

template <int ML, int MODE, bool DIAG>
__device__ __forceinline__ void sweep_group(const DevModel &m, const GroupDesc gd, const uint32_t *lo, const uint32_t *ro,
                                            const uint32_t *tt, const int tid, const int nthreads, WaveSync &ws, float *stage) {
  constexpr bool EARLY = wf_is_cluster(MODE);  // arrival as soon as the hand-off to the next tick is stored
  constexpr bool STAGE = MODE != WF_GRID;      // operands of the coming tick's item through shared memory
  const int D = m.nlev, itR = m.itR;
  const int nl = D - gd.lev0;                        // levels of the group
  const bool rivers = gd.rlev0 < m.nrlev && (m.task_mask & 4);
  const int nrl = rivers ? m.nrlev - gd.rlev0 : 0;   // river levels of the group
  const int nticks = gd.nticks;
  const int q0 = wf_sweep_q0(gd.lev0, m.rlev_shift, gd.rlev0);
  const int T = (int)blockDim.x;
  float *const sw = stage;
  double *const sd = reinterpret_cast<double *>(stage + (size_t)StageLayout<ML>::NW * T);
  // Item space of tick t, every task type starting on a warp boundary (a warp that straddled two
  // types would run both solves one after the other and the whole tick would wait for it):
  //   [subsurface cells of level t | pad][overland cells of level t-2 | pad][river cells, sub-step 0..itR-1]
  auto tick_items = [&](int t) -> int { return (int)tt[t]; };  // precomputed on the host (api.cu)
  auto decode_item = [&](int j, int t) -> WaveItem {
    WaveItem it;
    it.type = -1; it.idx = 0; it.v = 0; it.nup = 0;
    int cnt0 = 0, cnt1 = 0;
    const int t1 = t - WF_OVL_LAG;
    if (t < nl && (m.task_mask & 1)) cnt0 = (int)(lo[t + 1] - lo[t]);
    if (t1 >= 0 && t1 < nl && (m.task_mask & 2)) cnt1 = (int)(lo[t1 + 1] - lo[t1]);
    const int c0p = (cnt0 + 31) & ~31, c1p = (cnt1 + 31) & ~31;
    if (j < c0p) {
      if (j < cnt0) { it.type = 0; it.idx = lo[t] + (uint32_t)j; }
    } else if (j < c0p + c1p) {
      if (j - c0p < cnt1) { it.type = 1; it.idx = lo[t1] + (uint32_t)(j - c0p); }
    } else if (rivers) {
      int jj = j - c0p - c1p;
#pragma unroll 1
      for (int v = 0; v < itR; v++) {
        const int q = q0 + t - v;
        if (q < 0 || q >= nrl) continue;
        const uint32_t rb = ro[q];
        const int rc = (int)(ro[q + 1] - rb);
        if (jj < rc) { it.type = 2; it.v = v; it.idx = rb + (uint32_t)jj; break; }
        jj -= rc;
      }
    }
    return it;
  };
  // Grid sweep (256 threads per block, registers to spare): the operands of the thread's next item are fetched
  // into registers ahead of the barrier.  Cluster / block sweeps: into the thread's shared-memory column.
  constexpr bool PRELOAD = (MODE == WF_GRID);
  SubOps<ML> sub;
  OvlOps ovl;
  RivOps riv;
  // first item of this thread in the coming tick, with its operands on their way
  auto prepare = [&](int t) -> WaveItem {
    WaveItem it;
    it.type = -1; it.idx = 0; it.v = 0; it.nup = 0;
    if (t < nticks && tid < tick_items(t)) {
      it = decode_item(tid, t);
      if (PRELOAD) {
        if (it.type == 0) load_sub<ML>(m, (int64_t)it.idx, sub);
        else if (it.type == 1) load_ovl(m, (int64_t)it.idx, ovl);
        else if (it.type == 2) load_riv(m, (int32_t)it.idx, riv);
      } else if (STAGE) {
        if (it.type == 0) { stage_sub<ML>(m, (int64_t)it.idx, sw, sd, T); it.nup = __ldg(m.topo + it.idx) >> 4; }
        else if (it.type == 1) { stage_ovl(m, (int64_t)it.idx, sw, sd, T); it.nup = __ldg(m.topo + it.idx) >> 4; }
        else if (it.type == 2) { stage_riv(m, (int32_t)it.idx, sw, T); it.nup = __ldg(m.r_nup + it.idx); }
      }
    }
    return it;
  };
  WaveItem nxt = prepare(0);
  for (int t = 0; t < nticks; t++) {
#ifdef WF_PROFILE
    unsigned long long *pbase = (m.prof && gd.lev_idx == 0) ? m.prof + 16 * (size_t)t : nullptr;
    const unsigned long long c0 = gtimer();
    if (pbase && tid == 0) pbase[14] = c0;
#endif
    const int total = tick_items(t);
    const bool more = t + 1 < nticks;
    bool first = true, arrived = false;
    if (STAGE) cp_async_wait_all();
    for (int j = tid; j < total; j += nthreads, first = false) {
      WaveItem it;
      if (first) it = nxt;                // decoded, operands fetched, before the barrier
      else it = decode_item(j, t);
      const bool signal = EARLY && more && (j + nthreads >= total);  // the thread's last item of the tick
      if (it.type == 0) {
        if (first && STAGE) { unstage_sub<ML>(m, (int64_t)it.idx, sw, sd, T, sub); sub.nup = it.nup; }
        else if (!first || !PRELOAD) load_sub<ML>(m, (int64_t)it.idx, sub);
        task_subsurface<ML, MODE, DIAG, EARLY>(m, (int64_t)it.idx, sub, signal WF_PPASS(0));
        arrived = signal;
      } else if (it.type == 1) {
        if (first && STAGE) { unstage_ovl(m, (int64_t)it.idx, sw, sd, T, ovl); ovl.nup = it.nup; }
        else if (!first || !PRELOAD) load_ovl(m, (int64_t)it.idx, ovl);
        task_overland<MODE, DIAG, EARLY>(m, (int64_t)it.idx, ovl, signal WF_PPASS(1));
        arrived = signal;
      } else if (it.type == 2) {
        if (first && STAGE) { unstage_riv(m, (int32_t)it.idx, sw, T, riv); riv.nup = it.nup; }
        else if (!first || !PRELOAD) load_riv(m, (int32_t)it.idx, riv);
        task_river<MODE>(m, (int32_t)it.idx, it.v, riv, 0 WF_PPASS(2));
      }
    }
    if (more) {
#ifdef WF_PROFILE
      if (pbase) prof_stamp(pbase + 12, c0, 0);
#endif
      if (EARLY) {
        if (!arrived) wave_arrive_early();
      } else {
        wave_arrive<MODE>(ws);
      }
      // while the others arrive: this thread's next item and its operands ...
      nxt = prepare(t + 1);
      if (STAGE) {
        // ... (staged sweeps) nothing else for the thread's FIRST item of a tick: its cp.async copies are issued a barrier
        // wait ahead of their use, which covers the DRAM latency (measured: the L2 warm-up of the coming ticks' operands
        // made no difference once the operands were staged, WF_LAT_PREFETCH, and cost 9 % of the instructions).  Items
        // beyond the first (levels wider than the sweep has threads) are not staged: their lines are pulled into L2.
        const int t2 = t + 2;
        if (t2 < nticks && !(m.task_mask & 8)) {
          const int total2 = tick_items(t2);
          for (int j = tid + nthreads; j < total2; j += nthreads) {
            const WaveItem it = decode_item(j, t2);
            if (it.type == 0) prefetch_subsurface(m, (int64_t)it.idx);
            else if (it.type == 1) prefetch_overland(m, (int64_t)it.idx);
            else if (it.type == 2 && it.v == 0) prefetch_river(m, (int64_t)it.idx);
          }
        }
      } else {
        // ... (grid sweep) and the operands of the tick after into L2
        const int pd = 2;
        if (t + pd < nl) {
          const uint32_t nb = lo[t + pd], ne = lo[t + pd + 1];
          if (nb + (uint32_t)tid < ne && !(m.task_mask & 8)) prefetch_subsurface(m, (int64_t)nb + tid);
        }
        const int q = q0 + t + pd;  // river level of the first sub-step in two ticks
        if (q >= 0 && q < nrl) {
          const uint32_t rb = ro[q], re = ro[q + 1];
          if (rb + (uint32_t)tid < re && !(m.task_mask & 8)) prefetch_river(m, (int64_t)rb + tid);
        }
      }
      wave_wait<MODE>(ws);
#ifdef WF_PROFILE
      if (pbase) prof_stamp(pbase + 13, c0, 0);
#endif
    }
  }
}

// The staged sweeps (cluster / single block), software-pipelined over a thread's items.
//
// r2 tick profile of sweep_group on the bench shard: a thread's first instructions of a tick waited ~1.1 us for its
// staged operands.  They were requested behind the thread's arrival at the barrier, and the thread that arrives LAST --
// the one the tick waits for -- finds the barrier open: its copies had no time to land.  And the items beyond a
// thread's first of a tick (levels wider than the sweep has threads: every tick of a single-block sweep of several
// basins) loaded their operands where they were used.  Here every item's operands are requested one whole item
// earlier: while item k is solved, cp.async brings item k+1's operands (the thread's next item of this tick or its first
// of the next tick) into the thread's other staging column.  Nothing that is staged is written by another thread
// during the sweep (statics, the column kernel's hand-offs, the cell's own states), so the early copy reads the same
// values; what the neighbours produce is read behind the barrier as before.  The item itself (position, type,
// sub-step) waits in the column too, so it takes no registers while the current item is solved.
template <int ML, int MODE, bool DIAG, bool AHEAD = false>
__device__ __forceinline__ void sweep_group_ahead(const DevModel &m, const GroupDesc gd, const uint32_t *lo, const uint32_t *ro,
                                                  const uint32_t *tt, const int tid, const int nthreads, WaveSync &ws, float *stage, int32_t *progress) {
  constexpr bool EARLY = wf_is_cluster(MODE);
  using SL = StageLayout<ML>;
  constexpr int NW = SL::NW, ND = SL::ND, NBUF = SL::NBUF;
  const int D = m.nlev, itR = m.itR;
  const int nl = D - gd.lev0;
  const bool rivers = gd.rlev0 < m.nrlev && (m.task_mask & 4);
  const int nrl = rivers ? m.nrlev - gd.rlev0 : 0;
  const int nticks = gd.nticks;
  const int q0 = wf_sweep_q0(gd.lev0, m.rlev_shift, gd.rlev0);
  const int T = (int)blockDim.x, x = (int)threadIdx.x;
  double *const sdbase = reinterpret_cast<double *>(stage + (size_t)NBUF * NW * T);
  auto decode_item = [&](int j, int t) -> WaveItem {  // the item space of sweep_group
    WaveItem it;
    it.type = -1; it.idx = 0; it.v = 0; it.nup = 0;
    int cnt0 = 0, cnt1 = 0;
    const int t1 = t - WF_OVL_LAG;
    if (t < nl && (m.task_mask & 1)) cnt0 = (int)(lo[t + 1] - lo[t]);
    if (t1 >= 0 && t1 < nl && (m.task_mask & 2)) cnt1 = (int)(lo[t1 + 1] - lo[t1]);
    const int c0p = (cnt0 + 31) & ~31, c1p = (cnt1 + 31) & ~31;
    if (j < c0p) {
      if (j < cnt0) { it.type = 0; it.idx = lo[t] + (uint32_t)j; }
    } else if (j < c0p + c1p) {
      if (j - c0p < cnt1) { it.type = 1; it.idx = lo[t1] + (uint32_t)(j - c0p); }
    } else if (rivers) {
      int jj = j - c0p - c1p;
#pragma unroll 1
      for (int v = 0; v < itR; v++) {
        const int q = q0 + t - v;
        if (q < 0 || q >= nrl) continue;
        const uint32_t rb = ro[q];
        const int rc = (int)(ro[q + 1] - rb);
        if (jj < rc) { it.type = 2; it.v = v; it.idx = rb + (uint32_t)jj; break; }
        jj -= rc;
      }
    }
    return it;
  };
  // request an item's operands into column b (and park the item there)
  auto stage_item = [&](const WaveItem &it, const int b) {
    float *const sw = stage + (size_t)b * NW * T;
    double *const sd = sdbase + (size_t)b * ND * T;
    reinterpret_cast<uint32_t *>(sw)[SL::W_IDX * T + x] = it.idx;
    reinterpret_cast<uint32_t *>(sw)[SL::W_TYPE * T + x] = (uint32_t)(it.type & 0xff) | ((uint32_t)it.v << 8);
    if (it.type == 0) {
      stage_sub<ML>(m, (int64_t)it.idx, sw, sd, T);
      cp_async4(sw + SL::W_TOPO * T + x, m.topo + (it.idx & ~3u));
    } else if (it.type == 1) {
      stage_ovl(m, (int64_t)it.idx, sw, sd, T);
      cp_async4(sw + SL::W_TOPO * T + x, m.topo + (it.idx & ~3u));
    } else if (it.type == 2) {
      stage_riv(m, (int32_t)it.idx, sw, T);
      cp_async4(sw + SL::W_TOPO * T + x, m.r_nup + (it.idx & ~3u));
    }
  };
  auto staged_item = [&](const int b) -> WaveItem {
    const uint32_t *const sw = reinterpret_cast<const uint32_t *>(stage + (size_t)b * NW * T);
    WaveItem it;
    it.idx = sw[SL::W_IDX * T + x];
    const uint32_t tv = sw[SL::W_TYPE * T + x];
    it.type = (int)(int8_t)(tv & 0xff);
    it.v = (int)(tv >> 8);
    it.nup = 0;
    return it;
  };
  // the byte of the item's entry in the staged word of the topology array
  auto staged_byte = [&](const int b, const uint32_t idx) -> int {
    const uint32_t w = reinterpret_cast<const uint32_t *>(stage + (size_t)b * NW * T)[SL::W_TOPO * T + x];
    return (int)((w >> (8 * (idx & 3u))) & 0xffu);
  };
  SubOps<ML> sub;
  OvlOps ovl;
  RivOps riv;
  int p = 0;  // column of the current item
  WaveItem cur;
  cur.type = -1; cur.idx = 0; cur.v = 0; cur.nup = 0;
  if (tid < (int)tt[0]) cur = decode_item(tid, 0);
  stage_item(cur, 0);
  for (int t = 0; t < nticks; t++) {
#ifdef WF_PROFILE
    unsigned long long *pbase = (m.prof && gd.lev_idx == 0) ? m.prof + 16 * (size_t)t : nullptr;
    const unsigned long long c0 = gtimer();
    if (pbase && tid == 0) pbase[14] = c0;
#endif
    const int total = (int)tt[t];
    const bool more = t + 1 < nticks;
    bool arrived = false;
    if (tid >= total) {
      // nothing to do this tick: the thread's first item of the next tick and its operands
      cur.type = -1;
      if (more && tid < (int)tt[t + 1]) cur = decode_item(tid, t + 1);
      stage_item(cur, p);
    }
    for (int j = tid; j < total; j += nthreads) {
      // cur = item j of this tick, its operands in (or on their way into) column p
      WaveItem nx;
      nx.type = -1; nx.idx = 0; nx.v = 0; nx.nup = 0;
      if (j + nthreads < total) nx = decode_item(j + nthreads, t);
      else if (more && tid < (int)tt[t + 1]) nx = decode_item(tid, t + 1);
      cp_async_wait_all();  // (requested one item ago)
      const int q = (NBUF == 2) ? (p ^ 1) : p;
      const float *const sw = stage + (size_t)p * NW * T;
      const double *const sd = sdbase + (size_t)p * ND * T;
      if (NBUF == 2) stage_item(nx, q);  // two columns: the next item's copies start now, this item's operands are read where they are used
      const bool signal = EARLY && more && (j + nthreads >= total);  // the thread's last item of the tick
      const int type = cur.type;
      const uint32_t idx = cur.idx;
      if (type == 0) {
        unstage_sub<ML>(m, (int64_t)idx, sw, sd, T, sub);
        sub.nup = staged_byte(p, idx) >> 4;
        if (NBUF == 1) stage_item(nx, q);  // one column: its operands are in registers now
        task_subsurface<ML, MODE, DIAG, EARLY>(m, (int64_t)idx, sub, signal WF_PPASS(0));
        arrived = signal;
      } else if (type == 1) {
        unstage_ovl(m, (int64_t)idx, sw, sd, T, ovl);
        ovl.nup = staged_byte(p, idx) >> 4;
        if (NBUF == 1) stage_item(nx, q);
        task_overland<MODE, DIAG, EARLY>(m, (int64_t)idx, ovl, signal WF_PPASS(1));
        arrived = signal;
      } else {
        const int v = cur.v;
        if (type == 2) {
          unstage_riv(m, (int32_t)idx, sw, T, riv);
          riv.nup = staged_byte(p, idx);
        }
        if (NBUF == 1) stage_item(nx, q);
        if (type == 2) task_river<MODE>(m, (int32_t)idx, v, riv, 0 WF_PPASS(2));
      }
      cur = staged_item(q);
      p = q;
    }
    if (more) {
#ifdef WF_PROFILE
      if (pbase) prof_stamp(pbase + 12, c0, 0);
#endif
      if (EARLY) {
        if (!arrived) wave_arrive_early();
      } else {
        wave_arrive<MODE>(ws);
      }
      wave_wait<MODE>(ws);
      // the barrier of tick t has completed: what the tick's tasks stored ahead of their arrival, and the tick before's
      // behind it, is visible to whoever reads the group's progress with an acquire (column_ahead)
      // (a relaxed store: the data it announces was fenced by its writers' arrivals, MEMBAR.ALL.GPU ahead of UCGABAR_ARV,
      //  and this thread has passed the acquiring wait; a release here would put another fence on the tick's critical path)
      if (AHEAD && progress && tid == 0) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(progress), "r"(t + 1) : "memory");
#ifdef WF_PROFILE
      if (pbase) prof_stamp(pbase + 13, c0, 0);
#endif
    }
  }
  cp_async_wait_all();  // (nothing is pending after the last item; the columns are re-used by the cluster's next group)
}

// ---- the column of the next timestep on the SMs the sweep leaves idle ---------------------------------------------
// A cluster sweep holds one cluster per group: 32 groups of 4 blocks are 128 of the 148 SMs, and the column kernel of the
// next timestep starts when the last group has reached its outlet.  But the column of a cell only needs the cell's OWN
// lateral tasks of this step to be done -- the wavefront has passed level L once the barrier of tick L + 3 + it_kinR has
// completed -- so the clusters beyond the groups (as many as fit the GPU beside them) run column tiles of the next
// timestep meanwhile, in the order the tiles become ready, behind the groups' progress counters.  They take a new tile
// only while some group is still sweeping; k_vertical_rest does what is left.  Multi-step launches only (between two
// launches the states would be mid-step): wf_step_pipelined with the per-step sweeps, bit-identical to wf_step.
template <int ML>
__device__ __noinline__ void column_ahead(const DevModel &m) {  // (its own function: the sweep's registers are not its business)
  __shared__ int s_tile;
  const DevModel::WorkAhead &w = m.ahead;
  for (;;) {
    if (threadIdx.x == 0) {
      int t = -1, done;
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(w.queue + 1) : "memory");
      if (done < w.nsweep) t = atomicAdd(w.queue, 1);
      if (t >= w.ntiles) t = -1;
      if (t >= 0) {
        const int tile = w.order[t];
        for (int k = 0; k < 2; k++) {  // the tile's group(s) must have swept past its last level
          const int32_t *p = w.progress + w.need[4 * tile + 2 * k];
          const int need = w.need[4 * tile + 2 * k + 1];
          int have, spins = 0;
          for (;;) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(p) : "memory");
            if (have >= need) break;
            __nanosleep(400);
            if (++spins > (1 << 23)) __trap();  // (seconds: the sweep is not advancing)
          }
        }
        t = tile;
      }
      s_tile = t;
    }
    __syncthreads();
    const int tile = s_tile;
    __syncthreads();
    if (tile < 0) break;
    const int64_t i = (int64_t)tile * w.tile + threadIdx.x;
    if (i < m.n) column_cell<ML, 1, false, false, false>(m, i, nullptr, nullptr, (int)threadIdx.x, w.fo, nullptr);
  }
}

// AHEAD: the build of the multi-step launch whose spare clusters run the next timestep's column (column_ahead); the
// per-step sweep is compiled without it (the progress stores and the worker branch cost the per-step sweep 4 %).
template <int ML, int MODE, bool DIAG = true, bool AHEAD = false>
__global__ void __launch_bounds__(wf_block_threads(MODE), 1) k_lateral_wave(const __grid_constant__ DevModel m) {
  extern __shared__ __align__(16) uint32_t s_lev[];
  int crank = 0, csize = 1;
  if (wf_is_cluster(MODE)) {
    cg::cluster_group cl = cg::this_cluster();
    crank = (int)cl.block_rank();
    csize = (int)cl.num_blocks();
  }
  // grid: every block sweeps the one group; otherwise cluster (or block) c takes groups c, c + nclusters, ...
  int nclusters = (MODE == WF_GRID) ? 1 : (int)gridDim.x / csize;
  const int cid = (MODE == WF_GRID) ? 0 : (int)blockIdx.x / csize;
  const bool ahead = AHEAD && wf_is_cluster(MODE) && !DIAG && m.ahead.queue != nullptr;
  if (ahead) {
    if (cid >= m.ahead.nsweep) {  // a cluster beyond the groups: next timestep's column tiles
      column_ahead<ML>(m);
      return;
    }
    nclusters = m.ahead.nsweep;
  }
  // Position in the flat item space of a tick.  The item space is type-major ([subsurface | overland | river], the
  // subsurface items the longest), so with block-major positions the first block of a cluster ran nothing but
  // subsurface solves and the last one a few river cells: consecutive item warps are dealt round robin to the blocks
  // instead, and every SM of the cluster gets the same mix.
  int tid = (MODE == WF_GRID) ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : crank * (int)blockDim.x + (int)threadIdx.x;
  if (wf_is_cluster(MODE) && m.lat_interleave) tid = (((int)threadIdx.x >> 5) * csize + crank) * 32 + ((int)threadIdx.x & 31);
  const int nthreads = (MODE == WF_GRID) ? (int)(gridDim.x * blockDim.x) : csize * (int)blockDim.x;
  WaveSync ws;
  ws.counter = m.grid_counter;
  ws.nblocks = gridDim.x;
  ws.target = 0;
  // operand staging columns of the block's threads behind the level cache (16-byte aligned)
  float *const stage = reinterpret_cast<float *>(s_lev + (((size_t)m.lev_cache + 3) & ~(size_t)3));
  for (int g = cid; g < m.ngroups; g += nclusters) {
    const GroupDesc gd = m.groups[g];
    const int nlo = m.nlev - gd.lev0 + 1;
    const int nro = (gd.rlev0 < m.nrlev) ? m.nrlev - gd.rlev0 + 1 : 0;
    const uint32_t *lo = m.lev_off + gd.lev_idx, *ro = m.rlev_off + gd.rlev_idx, *tt = m.tick_total + gd.tick_idx;
    if (nlo + nro + gd.nticks <= m.lev_cache) {  // level offsets and tick sizes of the group into shared memory
      __syncthreads();
      for (int k = threadIdx.x; k < nlo; k += blockDim.x) s_lev[k] = lo[k];
      for (int k = threadIdx.x; k < nro; k += blockDim.x) s_lev[nlo + k] = ro[k];
      for (int k = threadIdx.x; k < gd.nticks; k += blockDim.x) s_lev[nlo + nro + k] = tt[k];
      __syncthreads();
      lo = s_lev;
      ro = s_lev + nlo;
      tt = s_lev + nlo + nro;
    }
#if WF_STAGE_AHEAD
    if (MODE != WF_GRID) sweep_group_ahead<ML, MODE, DIAG, AHEAD>(m, gd, lo, ro, tt, tid, nthreads, ws, stage, ahead ? m.ahead.progress + g : nullptr);
    else
#endif
    sweep_group<ML, MODE, DIAG>(m, gd, lo, ro, tt, tid, nthreads, ws, stage);
    // The next group of this cluster touches other cells only, but a thread's early arrival may still be pending:
    // the cluster sweeps close a group with one full barrier so that the next group starts on a fresh phase.
    if (wf_is_cluster(MODE) && (g + nclusters < m.ngroups || ahead)) {
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (ahead && tid == 0)  // everything of the group is stored and released: its column tiles of the next step are free
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(m.ahead.progress + g), "r"(0x7fffffff) : "memory");
    if (MODE == WF_GRID) break;
  }
  if (ahead && tid == 0) {
    __threadfence();
    atomicAdd(m.ahead.queue + 1, 1);  // this cluster has swept its groups
  }
}

// ---- batched solver kernels (wf_kinematic_wave_batch / wf_kinematic_wave_ssf_batch) -----------
__global__ void k_kinwave_batch(int64_t n, const double *a, double *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double qb;
  out[i] = kinematic_wave(a[i], a[n + i], pow_el(a[n + i], a[4 * n + i]), a[2 * n + i], a[3 * n + i], a[4 * n + i], a[5 * n + i],
                          a[6 * n + i], qb);
}
__global__ void k_kinwave_ssf_batch(int64_t n, const double *a, double *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double f = a[8 * n + i], D = a[9 * n + i];
  double s, z, e;
  kinematic_wave_ssf(a[i], a[n + i], a[2 * n + i], a[3 * n + i], a[4 * n + i], a[5 * n + i], a[6 * n + i], a[7 * n + i], f, D,
                     exp(-f * D), a[10 * n + i], a[11 * n + i], a[12 * n + i], a[13 * n + i], s, z, e);
  out[3 * i] = s; out[3 * i + 1] = z; out[3 * i + 2] = e;
}

}  // namespace wf
