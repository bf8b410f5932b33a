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
// all in one flat item space (type-major, see pipe_group), one barrier per tick.  K steps cost D + 2 + it_kinR + (K-1)*lag ticks
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
#ifndef WF_PIPE_GRID_MINBLOCKS
#define WF_PIPE_GRID_MINBLOCKS 3
#endif
__host__ __device__ constexpr int wf_pipe_minblocks(int mode) { return mode == WF_GRID ? WF_PIPE_GRID_MINBLOCKS : 1; }

struct PipeArgs {
  int32_t nsteps;   // timesteps of this launch
  int32_t lag;      // ticks between consecutive steps on the same level (3 + it_kinR)
  int64_t fstride;  // elements between the forcing rasters of consecutive steps
};

// river item j of local lateral tick t: sub-step v = 0 .. itR-1 works on river level q0 + t - v
__device__ __forceinline__ WaveItem decode_river(int jj, const int t, const uint32_t *ro, const int nrl, const int q0,
                                                 const int itR) {
  WaveItem it;
  it.type = -1; it.idx = 0; it.v = 0;
#pragma unroll 1
  for (int v = 0; v < itR; v++) {
    const int q = q0 + t - v;
    if (q < 0 || q >= nrl) continue;
    const uint32_t rb = ro[q];
    const int rc = (int)(ro[q + 1] - rb);
    if (jj < rc) { it.type = 2; it.v = v; it.idx = rb + (uint32_t)jj; break; }
    jj -= rc;
  }
  return it;
}
__device__ __forceinline__ int river_items(const int t, const uint32_t *ro, const int nrl, const int q0, const int itR) {
  int n = 0;
#pragma unroll 1
  for (int v = 0; v < itR; v++) {
    const int q = q0 + t - v;
    if (q >= 0 && q < nrl) n += (int)(ro[q + 1] - ro[q]);
  }
  return n;
}

// One group's pipelined sweep.  The item space of a tick is laid out TYPE-major over the steps in flight,
//     [columns of every step | subsurface of every step | overland of every step | river of every step],
// every (type, step) segment starting on a warp boundary, and thread k takes items k, k + nthreads, ...:
// every thread gets its share of each task type (a step-major layout whose period is close to the thread
// count hands some threads nothing but columns, four times the work of a river cell, and the tick waits
// for them), and at any moment most warps of an SM run the same task's code.
template <int ML, int MODE>
__device__ __forceinline__ void pipe_group(const DevModel &m, const PipeArgs pa, const GroupDesc gd, const uint32_t *lo,
                                           const uint32_t *ro, const int tid, const int nthreads, WaveSync &ws) {
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
    // steps in flight: 0 <= tau = t - s*lag < nt1
    int s_hi = t / lag;
    if (s_hi > pa.nsteps - 1) s_hi = pa.nsteps - 1;
    int s_lo = t - nt1 + 1;
    s_lo = (s_lo <= 0) ? 0 : (s_lo + lag - 1) / lag;
    int base = 0;  // items laid out so far in this tick, modulo the thread count
    // first item of this thread in a segment that starts at `base`; advances base past the segment
    auto first_item = [&](const int cnt) -> int {
      int j = tid - base;
      if (j < 0) j += nthreads;
      base += (cnt + 31) & ~31;
      if (base >= nthreads) base %= nthreads;
      return j;
    };
    // columns: level tau
    for (int s = s_lo; s <= s_hi; s++) {
      const int lev = t - s * lag;
      if (lev >= nl) continue;
      const uint32_t b0 = lo[lev];
      const int cnt = (int)(lo[lev + 1] - b0);
      for (int j = first_item(cnt); j < cnt; j += nthreads)
        column_cell<ML, 1>(m, (int64_t)b0 + j, nullptr, nullptr, 0, (int64_t)s * pa.fstride);
    }
    // subsurface flow: level tau - 1
    for (int s = s_lo; s <= s_hi; s++) {
      const int lev = t - s * lag - 1;
      if (lev < 0 || lev >= nl) continue;
      const uint32_t b0 = lo[lev];
      const int cnt = (int)(lo[lev + 1] - b0);
      for (int j = first_item(cnt); j < cnt; j += nthreads) {
        load_sub<ML>(m, (int64_t)b0 + j, sub);
        task_subsurface<ML, MODE>(m, (int64_t)b0 + j, sub WF_PNONE);
      }
    }
    // overland flow: level tau - 2
    for (int s = s_lo; s <= s_hi; s++) {
      const int lev = t - s * lag - 2;
      if (lev < 0 || lev >= nl) continue;
      const uint32_t b0 = lo[lev];
      const int cnt = (int)(lo[lev + 1] - b0);
      for (int j = first_item(cnt); j < cnt; j += nthreads) {
        load_ovl(m, (int64_t)b0 + j, ovl);
        task_overland<MODE>(m, (int64_t)b0 + j, ovl WF_PNONE);
      }
    }
    // river: sub-step v on river level q0 + (tau - 1) - v
    if (nrl > 0) {
      for (int s = s_lo; s <= s_hi; s++) {
        const int tl = t - s * lag - 1;
        if (tl < 0) continue;
        const int cnt = river_items(tl, ro, nrl, q0, itR);
        if (cnt == 0) continue;
        for (int j = first_item(cnt); j < cnt; j += nthreads) {
          const WaveItem it = decode_river(j, tl, ro, nrl, q0, itR);
          if (it.type == 2) {
            load_riv(m, (int32_t)it.idx, riv);
            task_river<MODE>(m, (int32_t)it.idx, it.v, riv, s WF_PNONE);
          }
        }
      }
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
    const uint32_t *lo = m.lev_off + gd.lev_idx, *ro = m.rlev_off + gd.rlev_idx;
    if (nlo + nro <= m.lev_cache) {  // level offsets of the group into shared memory
      __syncthreads();
      for (int k = threadIdx.x; k < nlo; k += blockDim.x) s_lev[k] = lo[k];
      for (int k = threadIdx.x; k < nro; k += blockDim.x) s_lev[nlo + k] = ro[k];
      __syncthreads();
      lo = s_lev;
      ro = s_lev + nlo;
    }
    pipe_group<ML, MODE>(m, pa, gd, lo, ro, tid, nthreads, ws);
    if (MODE == WF_GRID) break;
  }
}

}  // namespace wf
