// Gauge / subcatchment accumulation kernels: segmented reductions with warp shuffles.
// Behavioural reference: wf_OutputTimeSeriesArea.writestep, wflow/wf_DynamicFramework.py:446-499
// (pcr.areaaverage / areatotal / areamaximum / areaminimum over an id map, then one sample
// per id).
#pragma once
#include <cstdint>

namespace wf {

// One block per id.  Cells of an id are the run perm[seg[b] .. seg[b+1]) of network-order
// indices (sorted by id at wf_area_create).  Sums accumulate in float64 in a fixed order
// (thread-strided partials, shuffle tree, then warp partials in order): deterministic.
// slot != nullptr: `field` is a river-order array; cells without a river slot contribute 0.
__global__ void __launch_bounds__(256) k_area_reduce(const float *__restrict__ field, const int32_t *__restrict__ slot,
                                                     const int32_t *__restrict__ perm, const int32_t *__restrict__ seg,
                                                     int op, float *__restrict__ out) {
  const int b = blockIdx.x;
  const int beg = seg[b], end = seg[b + 1];
  double s = 0.0;
  float mx = -3.402823466e38f, mn = 3.402823466e38f;
  for (int p = beg + threadIdx.x; p < end; p += blockDim.x) {
    int k = perm[p];
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
    float r;
    if (op == 0) r = (float)(t / (double)(end - beg));   // areaaverage
    else if (op == 1) r = (float)t;                      // areatotal
    else if (op == 2) r = a;                             // areamaximum
    else r = c;                                          // areaminimum
    out[b] = r;
  }
}

__global__ void k_sample(const float *__restrict__ field, const int32_t *__restrict__ pos, float *__restrict__ out,
                         int64_t n, float mv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (pos[i] >= 0) ? field[pos[i]] : mv;
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
