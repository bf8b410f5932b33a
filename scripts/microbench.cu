// Latency microbenchmark of the two Newton solvers (development aid): one warp, dependent calls.
#include <cstdio>
#include <cuda_runtime.h>
#include "../wflow_b200/csrc/lateral.cuh"
using namespace wf;

__global__ void bench_kw(double Qin, double Qold, double q, double alpha, double beta, double dT, double dX, int reps,
                         long long *cyc, double *out) {
  double acc = 0.0, qb;
  double qo = Qold + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    double v = kinematic_wave(Qin, qo, q, alpha, beta, dT, dX, qb);
    acc += v; qo = Qold + 1e-9 * v + threadIdx.x * 1e-3;  // dependent chain
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { *cyc = (t1 - t0) / reps; *out = acc; }
}
__global__ void bench_ssf(double ssf_in, double ssf_old, double zi_old, double r, double Kh, double Ks, double slope,
                          double neff, double f, double D, int reps, long long *cyc, double *out) {
  double acc = 0.0;
  double s, z, e, so = ssf_old * (1.0 + threadIdx.x * 1e-3);
  long long t0 = clock64();
  for (int k = 0; k < reps; k++) {
    kinematic_wave_ssf(ssf_in, so, zi_old, r, Kh, Ks, slope, neff, f, D, exp(-f * D), 1.0, 1.414e6, 7.07e5, 1e9, s, z, e);
    acc += s + z; so = ssf_old * (1.0 + 1e-12 * z + threadIdx.x * 1e-3);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { *cyc = (t1 - t0) / reps; *out = acc; }
}
__global__ void bench_ops(int reps, long long *cyc, double *out) {
  double x = 1.234 + threadIdx.x * 1e-3;
  long long t[8];
  t[0] = clock64(); for (int k = 0; k < reps; k++) x = log(x) + 2.0;
  t[1] = clock64(); for (int k = 0; k < reps; k++) x = exp(x * 1e-3) + 1.0;
  t[2] = clock64(); for (int k = 0; k < reps; k++) x = 3.0 / x + 1.0;
  t[3] = clock64(); for (int k = 0; k < reps; k++) x = pow(x, 0.6) + 1.0;
  t[4] = clock64(); for (int k = 0; k < reps; k++) x = fma(x, 0.999, 0.5);
  t[5] = clock64();
  float y = (float)x;
  for (int k = 0; k < reps; k++) y = powf(y, 0.6f) + 1.0f;
  t[6] = clock64(); for (int k = 0; k < reps; k++) y = 3.0f / y + 1.0f;
  t[7] = clock64();
  if (threadIdx.x == 0) { for (int i = 0; i < 7; i++) cyc[i] = (t[i + 1] - t[i]) / reps; *out = x + y; }
}
int main() {
  long long *cyc; double *out;
  cudaMallocManaged(&cyc, 64); cudaMallocManaged(&out, 8);
  const double beta = (double)0.6f;
  bench_ops<<<1, 32>>>(1000, cyc, out); cudaDeviceSynchronize();
  printf("cycles/op (dependent, 1 warp): log %lld exp %lld div %lld pow %lld dfma %lld powf %lld divf %lld\n", cyc[0], cyc[1], cyc[2], cyc[3], cyc[4], cyc[5], cyc[6]);
  bench_kw<<<1, 32>>>(0.02, 0.01, 1e-5, 5.0, beta, 3600, 1414, 200, cyc, out); cudaDeviceSynchronize();
  printf("kinematic_wave overland-like: %lld cycles/call\n", cyc[0]);
  bench_kw<<<1, 32>>>(50.0, 40.0, 1e-3, 8.0, beta, 900, 1414, 200, cyc, out); cudaDeviceSynchronize();
  printf("kinematic_wave river-like: %lld cycles/call\n", cyc[0]);
  bench_kw<<<1, 32>>>(0.0, 0.0, 1e-6, 5.0, beta, 3600, 1414, 200, cyc, out); cudaDeviceSynchronize();
  printf("kinematic_wave dry start: %lld cycles/call\n", cyc[0]);
  bench_ssf<<<1, 32>>>(2e10, 1.5e10, 800.0, 2e6, 10.0, 100.0, 0.05, 0.45, 1e-3, 2000.0, 200, cyc, out); cudaDeviceSynchronize();
  printf("kinematic_wave_ssf: %lld cycles/call\n", cyc[0]);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
