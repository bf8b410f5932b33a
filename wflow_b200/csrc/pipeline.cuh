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

// block size and launch bounds of the pipelined sweep per barrier kind.  The sweep is throughput-bound, so the
// grid sweep runs ONE block of 768 threads per SM (<= 85 registers): 24 warps that share the block's chunk
// counters (see pipe_group).
#ifndef WF_PIPE_GRID_THREADS
#define WF_PIPE_GRID_THREADS 768
#endif
__host__ __device__ constexpr int wf_pipe_threads(int mode) {
  return mode == WF_GRID ? WF_PIPE_GRID_THREADS : (mode == WF_CTA ? WF_GROUP_THREADS : WF_GROUP_THREADS_WIDE);
}
#ifndef WF_PIPE_GRID_MINBLOCKS
#define WF_PIPE_GRID_MINBLOCKS 1
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
// as segments (type, step) cut into chunks of 32 consecutive items = one warp's worth.  Chunks are dealt round
// robin to the blocks of the sweep (so every block gets its share of every task type), and INSIDE a block the
// warps take the block's chunks of a segment from a shared-memory counter, segment after segment:
//   * a tick ends when its slowest warp ends, and items differ widely (a column is ~4 river cells; sub-step
//     loops are data dependent): with a fixed item -> thread map 24 % of all stall cycles of the bench shard
//     were spent at the tick barrier; dealing chunks on demand lets the block's warps finish together;
//   * the warps of a block move through the segments together, so at any moment an SM runs mostly ONE task's
//     code (all four together, ~100 KB, do not fit the instruction cache: 17 % "no instruction" stalls with
//     the fixed map).
// Counters are double-buffered by tick parity: a tick zeroes the other parity's counters, which every warp
// has left behind when it passed the previous tick's barrier.
#define WF_PIPE_MAX_INFLIGHT 64
template <int ML, int MODE>
__device__ __forceinline__ void pipe_group(const DevModel &m, const PipeArgs pa, const GroupDesc gd, const uint32_t *lo,
                                           const uint32_t *ro, const int brank, const int nblocks, int (*ctr)[4 * WF_PIPE_MAX_INFLIGHT],
                                           WaveSync &ws) {
  const int itR = m.itR;
  const int nl = m.nlev - gd.lev0;                                      // levels of the group
  const int nrl = (gd.rlev0 < m.nrlev) ? m.nrlev - gd.rlev0 : 0;        // river levels of the group
  const int nt1 = nl + 2 + ((gd.rlev0 < m.nrlev) ? itR : 0);            // local ticks of one step: the column's + the lateral tasks' (own schedule, below)
  const int q0 = (gd.lev0 - 2) - m.rlev_shift - gd.rlev0;
  const int lag = pa.lag;
  const int total = nt1 + (pa.nsteps - 1) * lag;
  const int lane = (int)threadIdx.x & 31;
  SubOps<ML> sub;
  OvlOps ovl;
  RivOps riv;
  for (int t = 0; t < total; t++) {
    int *const cnt_now = ctr[t & 1];
    {  // the next tick's counters
      int *const cnt_next = ctr[(t + 1) & 1];
      for (int k = (int)threadIdx.x; k < 4 * WF_PIPE_MAX_INFLIGHT; k += (int)blockDim.x) cnt_next[k] = 0;
    }
    // steps in flight: 0 <= tau = t - s*lag < nt1  (at most WF_PIPE_MAX_INFLIGHT, checked by the host)
    int s_hi = t / lag;
    if (s_hi > pa.nsteps - 1) s_hi = pa.nsteps - 1;
    int s_lo = t - nt1 + 1;
    s_lo = (s_lo <= 0) ? 0 : (s_lo + lag - 1) / lag;
    int gchunk = 0;  // chunks laid out so far in this tick, modulo the number of blocks
    // The block's chunks of a segment of `cnt` items: c0, c0 + nblocks, ... ; a warp takes the next one.
    // Returns the first item of the chunk taken, or -1 when the block's share of the segment is used up.
    auto seg_begin = [&](const int cnt, int &c0) {
      c0 = brank - gchunk;
      if (c0 < 0) c0 += nblocks;
      gchunk += (cnt + 31) >> 5;
      if (gchunk >= nblocks) gchunk %= nblocks;
    };
    auto take = [&](int *counter, const int c0, const int cnt) -> int {
      int k = 0;
      if (lane == 0) k = atomicAdd(counter, 1);
      k = __shfl_sync(0xffffffffu, k, 0);
      const int j = ((c0 + k * nblocks) << 5) + lane;
      return ((c0 + k * nblocks) << 5) < cnt ? j : -1;
    };
    // columns: level tau
    for (int s = s_lo; s <= s_hi; s++) {
      const int lev = t - s * lag;
      if (lev >= nl) continue;
      const uint32_t b0 = lo[lev];
      const int cnt = (int)(lo[lev + 1] - b0);
      int c0;
      seg_begin(cnt, c0);
      int *counter = cnt_now + (s - s_lo);
      for (int j = take(counter, c0, cnt); j >= 0; j = take(counter, c0, cnt)) {
        if (j < cnt) column_cell<ML, 1, false, true, true>(m, (int64_t)b0 + j, nullptr, nullptr, 0, (int64_t)s * pa.fstride, nullptr);
        __syncwarp();
      }
    }
    // subsurface flow: level tau - 1
    for (int s = s_lo; s <= s_hi; s++) {
      const int lev = t - s * lag - 1;
      if (lev < 0 || lev >= nl) continue;
      const uint32_t b0 = lo[lev];
      const int cnt = (int)(lo[lev + 1] - b0);
      int c0;
      seg_begin(cnt, c0);
      int *counter = cnt_now + WF_PIPE_MAX_INFLIGHT + (s - s_lo);
      for (int j = take(counter, c0, cnt); j >= 0; j = take(counter, c0, cnt)) {
        if (j < cnt) {
          load_sub<ML>(m, (int64_t)b0 + j, sub);
          task_subsurface<ML, MODE>(m, (int64_t)b0 + j, sub, false WF_PNONE);
        }
        __syncwarp();
      }
    }
    // overland flow: level tau - 2
    for (int s = s_lo; s <= s_hi; s++) {
      const int lev = t - s * lag - 2;
      if (lev < 0 || lev >= nl) continue;
      const uint32_t b0 = lo[lev];
      const int cnt = (int)(lo[lev + 1] - b0);
      int c0;
      seg_begin(cnt, c0);
      int *counter = cnt_now + 2 * WF_PIPE_MAX_INFLIGHT + (s - s_lo);
      for (int j = take(counter, c0, cnt); j >= 0; j = take(counter, c0, cnt)) {
        if (j < cnt) {
          load_ovl(m, (int64_t)b0 + j, ovl);
          task_overland<MODE>(m, (int64_t)b0 + j, ovl, false WF_PNONE);
        }
        __syncwarp();
      }
    }
    // river: sub-step v on river level q0 + (tau - 1) - v
    if (nrl > 0) {
      for (int s = s_lo; s <= s_hi; s++) {
        const int tl = t - s * lag - 1;
        if (tl < 0) continue;
        const int cnt = river_items(tl, ro, nrl, q0, itR);
        if (cnt == 0) continue;
        int c0;
        seg_begin(cnt, c0);
        int *counter = cnt_now + 3 * WF_PIPE_MAX_INFLIGHT + (s - s_lo);
        for (int j = take(counter, c0, cnt); j >= 0; j = take(counter, c0, cnt)) {
          if (j < cnt) {
            const WaveItem it = decode_river(j, tl, ro, nrl, q0, itR);
            if (it.type == 2) {
              load_riv(m, (int32_t)it.idx, riv);
              task_river<MODE>(m, (int32_t)it.idx, it.v, riv, s WF_PNONE);
            }
          }
          __syncwarp();
        }
      }
    }
    if (t + 1 < total) {
      wave_arrive<MODE>(ws);
      wave_wait<MODE>(ws);
    } else if (MODE == WF_CTA || wf_is_cluster(MODE)) {
      __syncthreads();  // the next group of this block starts with fresh counters
    }
  }
}

template <int ML, int MODE>
__global__ void __launch_bounds__(wf_pipe_threads(MODE), wf_pipe_minblocks(MODE)) k_pipeline(const DevModel m, const PipeArgs pa) {
  extern __shared__ uint32_t s_lev[];
  __shared__ int s_ctr[2][4 * WF_PIPE_MAX_INFLIGHT];
  int crank = 0, csize = 1;
  if (wf_is_cluster(MODE)) {
    cg::cluster_group cl = cg::this_cluster();
    crank = (int)cl.block_rank();
    csize = (int)cl.num_blocks();
  }
  const int nclusters = (MODE == WF_GRID) ? 1 : (int)gridDim.x / csize;
  const int cid = (MODE == WF_GRID) ? 0 : (int)blockIdx.x / csize;
  // blocks that sweep a group together, and this block's rank among them
  const int brank = (MODE == WF_GRID) ? (int)blockIdx.x : crank;
  const int nblocks = (MODE == WF_GRID) ? (int)gridDim.x : csize;
  WaveSync ws;
  ws.counter = m.grid_counter;
  ws.nblocks = gridDim.x;
  ws.target = 0;
  for (int k = (int)threadIdx.x; k < 2 * 4 * WF_PIPE_MAX_INFLIGHT; k += (int)blockDim.x) (&s_ctr[0][0])[k] = 0;
  __syncthreads();
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
    pipe_group<ML, MODE>(m, pa, gd, lo, ro, brank, nblocks, s_ctr, ws);
    if (MODE != WF_GRID) {
      // both parities clean for the next group (its tick 0 zeroes parity 1 itself; parity 0 here)
      for (int k = (int)threadIdx.x; k < 4 * WF_PIPE_MAX_INFLIGHT; k += (int)blockDim.x) s_ctr[0][k] = 0;
      __syncthreads();
    } else {
      g = m.ngroups;  // the grid sweeps the one group
    }
  }
}

}  // namespace wf
