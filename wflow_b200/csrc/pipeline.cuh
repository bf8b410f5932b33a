// Several timesteps in flight: the wavefront skewed in TIME as well as in space.
//
// The reference runs timestep s+1 only after timestep s has walked the whole network
// (wf_DynamicFramework._runDynamic, wflow/wf_DynamicFramework.py:2968-3042), and the per-step sweep
// of lateral.cuh inherits that: D + 1 + it_kinR barrier-separated ticks per step, each one Newton
// solve deep, with only the cells of two or three levels to work on -- latency-bound (a step of the
// 15 754-cell Rhine case takes as long as one of a million cells).  But nothing of step s+1 needs
// the whole of step s.  Per cell the order of work is
//     column(s) -> subsurface(s) -> overland(s) -> river(s, 0 .. it_kinR-1) -> column(s+1) -> ...
// and every task reads, besides its own cell, only what the cells ONE level upstream produced one
// tick earlier.  So step s+1 may follow step s down the network at a distance of
//     lag = 3 + it_kinR ticks:
// at tick t the kernel runs, for every step s in flight (tau = t - s*lag),
//     the column of level tau, subsurface flow of level tau-1, overland flow of level tau-2 and
//     river sub-step v of level tau-3-v,
// all in one flat item space, one barrier per tick.  K steps cost D + 2 + it_kinR + (K-1)*lag ticks
// instead of K*(D + 1 + it_kinR): the Rhine case advances 29 days in 470 ticks instead of 7 800, and
// on wide networks a tick carries the work of ~D/lag steps, so the sweep is bound by instruction
// throughput like the column kernel, not by barrier + solve latency.
//
// Every state array is updated in place in step order per cell; a hand-off written at tick t is read
// at tick t+1 and next overwritten at tick t+lag (lag >= 4), so one copy of every array serves all
// steps in flight.  Results are bit-identical to K calls of the per-step path (same device functions,
// same operand values, same order: tests/test_gpu_pipeline.py).  What differs per step is only the
// forcing, read from a [K][stride] buffer, and what leaves a step before the next overwrites it: the
// river discharge at the gauge cells (DevModel::cap_*).
//
// Memory visibility: a cell's tasks run on different SMs from tick to tick.  Every barrier used here
// is an acquire that invalidates the SM's L1 (CCTL.IVALL after UCGABAR_WAIT / ld.acquire.gpu in SASS),
// all stores are write-through, and the column reads its states from L2 (ldmut<1>), never through the
// non-coherent path.
#pragma once
#include "lateral.cuh"

namespace wf {

#ifdef WF_PROFILE
#define WF_PNONE , nullptr, 0ull
#else
#define WF_PNONE
#endif

// block size and launch bounds of the pipelined sweep per barrier kind.  The sweep is throughput-bound,
// so the grid sweep asks for three 256-thread blocks per SM (<= 85 registers) instead of one.
__host__ __device__ constexpr int wf_pipe_threads(int mode) {
  return mode == WF_GRID ? 256 : (mode == WF_CTA ? WF_GROUP_THREADS : WF_GROUP_THREADS_WIDE);
}
__host__ __device__ constexpr int wf_pipe_minblocks(int mode) { return mode == WF_GRID ? 3 : 1; }

struct PipeArgs {
  int32_t nsteps;   // timesteps of this launch
  int32_t lag;      // ticks between consecutive steps on the same level (3 + it_kinR)
  int64_t fstride;  // elements between the forcing rasters of consecutive steps
};

// item j of the lateral part of local tick t (the item space of lateral.cuh's sweep_group)
__device__ __forceinline__ WaveItem decode_lateral(int j, const int t, const uint32_t *lo, const uint32_t *ro, const int nl,
                                                   const int nrl, const int q0, const int itR) {
  WaveItem it;
  it.type = -1; it.idx = 0; it.v = 0;
  int cnt0 = 0, cnt1 = 0;
  if (t < nl) cnt0 = (int)(lo[t + 1] - lo[t]);
  if (t >= 1 && t - 1 < nl) cnt1 = (int)(lo[t] - lo[t - 1]);
  const int c0p = (cnt0 + 31) & ~31, c1p = (cnt1 + 31) & ~31;
  if (j < c0p) {
    if (j < cnt0) { it.type = 0; it.idx = lo[t] + (uint32_t)j; }
  } else if (j < c0p + c1p) {
    if (j - c0p < cnt1) { it.type = 1; it.idx = lo[t - 1] + (uint32_t)(j - c0p); }
  } else if (nrl > 0) {
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
}

template <int ML, int MODE>
__device__ __forceinline__ void pipe_group(const DevModel &m, const PipeArgs pa, const GroupDesc gd, const uint32_t *lo,
                                           const uint32_t *ro, const uint32_t *tt, const int tid, const int nthreads,
                                           WaveSync &ws) {
  const int itR = m.itR;
  const int nl = m.nlev - gd.lev0;                                      // levels of the group
  const int nrl = (gd.rlev0 < m.nrlev) ? m.nrlev - gd.rlev0 : 0;        // river levels of the group
  const int nt1 = gd.nticks + 1;                                        // local ticks of one step: the column's + the lateral sweep's
  const int q0 = (gd.lev0 - 2) - m.rlev_shift - gd.rlev0;
  const int lag = pa.lag;
  const int total = nt1 + (pa.nsteps - 1) * lag;
  SubOps<ML> sub;
  OvlOps ovl;
  RivOps riv;
  for (int t = 0; t < total; t++) {
    // steps in flight: 0 <= t - s*lag < nt1
    int s_hi = t / lag;
    if (s_hi > pa.nsteps - 1) s_hi = pa.nsteps - 1;
    int s_lo = t - nt1 + 1;
    s_lo = (s_lo <= 0) ? 0 : (s_lo + lag - 1) / lag;
    int base = 0;  // items of the steps already laid out in this tick, modulo the thread count
    for (int s = s_lo; s <= s_hi; s++) {
      const int tau = t - s * lag;
      int cntV = 0;
      uint32_t v0 = 0;
      if (tau < nl) { v0 = lo[tau]; cntV = (int)(lo[tau + 1] - v0); }
      const int cVp = (cntV + 31) & ~31;  // every task type starts on a warp boundary
      const int cntL = (tau >= 1) ? (int)tt[tau - 1] : 0;
      const int cnt = cVp + ((cntL + 31) & ~31);
      int j = tid - base;
      if (j < 0) j += nthreads;
      for (; j < cnt; j += nthreads) {
        if (j < cVp) {
          if (j < cntV) column_cell<ML, 1>(m, (int64_t)v0 + j, nullptr, nullptr, 0, (int64_t)s * pa.fstride);
          continue;
        }
        const WaveItem it = decode_lateral(j - cVp, tau - 1, lo, ro, nl, nrl, q0, itR);
        if (it.type == 0) {
          load_sub<ML>(m, (int64_t)it.idx, sub);
          task_subsurface<ML, MODE>(m, (int64_t)it.idx, sub WF_PNONE);
        } else if (it.type == 1) {
          load_ovl(m, (int64_t)it.idx, ovl);
          task_overland<MODE>(m, (int64_t)it.idx, ovl WF_PNONE);
        } else if (it.type == 2) {
          load_riv(m, (int32_t)it.idx, riv);
          task_river<MODE>(m, (int32_t)it.idx, it.v, riv, s WF_PNONE);
        }
      }
      base = (base + cnt) % nthreads;
    }
    if (t + 1 < total) {
      wave_arrive<MODE>(ws);
      wave_wait<MODE>(ws);
    }
  }
}

template <int ML, int MODE>
__global__ void __launch_bounds__(wf_pipe_threads(MODE), wf_pipe_minblocks(MODE)) k_pipeline(const DevModel m, const PipeArgs pa) {
  extern __shared__ uint32_t s_lev[];
  int crank = 0, csize = 1;
  if (wf_is_cluster(MODE)) {
    cg::cluster_group cl = cg::this_cluster();
    crank = (int)cl.block_rank();
    csize = (int)cl.num_blocks();
  }
  const int nclusters = (MODE == WF_GRID) ? 1 : (int)gridDim.x / csize;
  const int cid = (MODE == WF_GRID) ? 0 : (int)blockIdx.x / csize;
  const int tid = (MODE == WF_GRID) ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : crank * (int)blockDim.x + (int)threadIdx.x;
  const int nthreads = (MODE == WF_GRID) ? (int)(gridDim.x * blockDim.x) : csize * (int)blockDim.x;
  WaveSync ws;
  ws.counter = m.grid_counter;
  ws.nblocks = gridDim.x;
  ws.target = 0;
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
    pipe_group<ML, MODE>(m, pa, gd, lo, ro, tt, tid, nthreads, ws);
    if (MODE == WF_GRID) break;
  }
}

}  // namespace wf
