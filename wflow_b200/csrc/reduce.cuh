// Gauge / subcatchment accumulation kernels: segmented reductions with warp shuffles.
// Behavioural reference: wf_OutputTimeSeriesArea.writestep, wflow/wf_DynamicFramework.py:446-499
// (pcr.areaaverage / areatotal / areamaximum / areaminimum over an id map, then one sample
// per id).
#pragma once
#include <cstdint>

namespace wf {

// Two passes, both deterministic.  The cells of an id are the run perm[seg[b] .. seg[b+1]) of storage positions (sorted by
// id at wf_area_create); the runs are cut into chunks of WF_AREA_CHUNK cells, listed in chunk_id / chunk_beg / chunk_end.
//   k_area_partial  one block per chunk: float64 sum, maximum and minimum of the chunk (thread-strided partials, shuffle
//                   tree, warp partials in order) -> partial[chunk]
//   k_area_combine  one thread per id: the chunks of the id in order -> areaaverage / areatotal / areamaximum / areaminimum
// so a subcatchment of millions of cells is reduced by thousands of blocks instead of one (r1: one block per id).
// slot != nullptr: `field` is a river-order array; cells without a river slot contribute 0.
#define WF_AREA_CHUNK 16384
struct AreaPartial {
  double sum;
  float mx, mn;
};
__global__ void __launch_bounds__(256) k_area_partial(const float *__restrict__ field, const int32_t *__restrict__ slot,
                                                      const int32_t *__restrict__ perm, const int32_t *__restrict__ chunk_beg,
                                                      const int32_t *__restrict__ chunk_end, AreaPartial *__restrict__ partial) {
  const int b = blockIdx.x;
  const int beg = chunk_beg[b], end = chunk_end[b];
  double s = 0.0;
  float mx = -3.402823466e38f, mn = 3.402823466e38f;
  for (int p = beg + threadIdx.x; p < end; p += blockDim.x) {
    const int k = perm[p];
    float v;
    if (slot) {
      const int r = slot[k];
      v = (r >= 0) ? field[r] : 0.0f;
    } else {
      v = field[k];
    }
    s += (double)v;
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o));
  }
  __shared__ double ss[8];
  __shared__ float smx[8], smn[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { ss[w] = s; smx[w] = mx; smn[w] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    float a = smx[0], c = smn[0];
    for (int i = 0; i < 8; i++) { t += ss[i]; a = fmaxf(a, smx[i]); c = fminf(c, smn[i]); }
    partial[b].sum = t;
    partial[b].mx = a;
    partial[b].mn = c;
  }
}
__global__ void k_area_combine(const AreaPartial *__restrict__ partial, const int32_t *__restrict__ id_chunk0,
                               const int32_t *__restrict__ seg, int nids, int op, float *__restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nids) return;
  double t = 0.0;
  float a = -3.402823466e38f, c = 3.402823466e38f;
  for (int k = id_chunk0[b]; k < id_chunk0[b + 1]; k++) {
    t += partial[k].sum;
    a = fmaxf(a, partial[k].mx);
    c = fminf(c, partial[k].mn);
  }
  float r;
  if (op == 0) r = (float)(t / (double)(seg[b + 1] - seg[b]));  // areaaverage
  else if (op == 1) r = (float)t;                               // areatotal
  else if (op == 2) r = a;                                      // areamaximum
  else r = c;                                                   // areaminimum
  out[b] = r;
}

__global__ void k_sample(const float *__restrict__ field, const int32_t *__restrict__ pos, float *__restrict__ out,
                         int64_t n, float mv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (pos[i] >= 0) ? field[pos[i]] : mv;
}

// ---- estimate_iterations_kin_wave (wflow/wflow_funcs.py:103-116) -------------------------------------
// Courant number of every cell with Q > 0, as REAL4 map algebra evaluates it:
//   celerity = 1 / (alpha * Beta * Q ** (Beta - 1));  courant = (timestepsecs / dx) * celerity
// Keys are the float bits (positive floats order like unsigned integers); cells without flow get the
// all-ones key and are not counted.  The 95th percentile is then two order statistics found by a
// most-significant-digit-first radix select (k_select_hist: one 8-bit digit per pass).
__global__ void k_courant_keys(const float *__restrict__ Q, const float *__restrict__ alpha, const float *__restrict__ dx,
                               float beta, float timestepsecs, uint32_t *__restrict__ keys, unsigned long long *count,
                               int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t key = 0xffffffffu;
  if (i < n) {
    const float q = Q[i];
    if (q > 0.0f) {
      const float p = powf(q, beta - 1.0f);
      const float cel = 1.0f / ((alpha[i] * beta) * p);
      const float courant = (timestepsecs / dx[i]) * cel;
      if (courant == courant && courant >= 0.0f) key = __float_as_uint(courant);
    }
    keys[i] = key;
  }
  const unsigned valid = __ballot_sync(0xffffffffu, key != 0xffffffffu);
  if ((threadIdx.x & 31) == 0 && valid) atomicAdd(count, (unsigned long long)__popc(valid));
}
// histogram of the digit at `shift` over the keys whose higher digits equal `prefix` (mask selects them)
__global__ void k_select_hist(const uint32_t *__restrict__ keys, int64_t n, uint32_t mask, uint32_t prefix, int shift,
                              unsigned long long *__restrict__ hist) {
  __shared__ unsigned int sh[256];
  for (int k = threadIdx.x; k < 256; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = keys[i];
    if (key != 0xffffffffu && (key & mask) == prefix) atomicAdd(&sh[(key >> shift) & 255u], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 256; k += blockDim.x)
    if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// gauge sets: pos = storage position (-1 outside the network); rpos (river-only fields) = river-order
// position or -1 for a network cell that is not a river cell (the reference holds 0 there)
__global__ void k_sample_gauges(const float *__restrict__ field, const int32_t *__restrict__ pos,
                                const int32_t *__restrict__ rpos, float *__restrict__ out, int64_t n, float mv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pos[i] < 0) out[i] = mv;
  else if (rpos) out[i] = (rpos[i] >= 0) ? field[rpos[i]] : 0.0f;
  else out[i] = field[pos[i]];
}

}  // namespace wf
