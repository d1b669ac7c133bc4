// Latency micro-benchmarks on one warp / one block (B200): dependent DFMA chain, reciprocal,
// shared-memory round trip with barrier.  nvcc -arch=sm_100a -O3 -o ubench tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double* out, double a, double b, int n, long long* cyc) {
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dadd_chain(double* out, double a, double b, int n, long long* cyc) {
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = x + b;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ffma_chain(float* out, float a, float b, int n, long long* cyc) {
  float x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fmaf(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void rcp_chain(double* out, double a, int n, long long* cyc) {
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    double r;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    x = r + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void div_chain(double* out, double a, int n, long long* cyc) {
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = 1.0 / x + 1.0;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void smem_bar_loop(double* out, int n, long long* cyc) {
  __shared__ double s[1024];
  s[threadIdx.x] = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    double v = s[(threadIdx.x + 1) % blockDim.x];
    s[threadIdx.x] = v + 1.0;
    __syncthreads();
  }
  long long t1 = clock64();
  out[threadIdx.x] = s[threadIdx.x];
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void bar_only(int n, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dfma_tput(double* out, double a, double b, int n, long long* cyc) {
  double x[8];
  for (int k = 0; k < 8; ++k) x[k] = a + k;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fma(x[k], b, a);
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; float* outf; long long* cyc; long long h;
  cudaMalloc(&out, 8192); cudaMalloc(&outf, 8192); cudaMalloc(&cyc, 8);
  const int n = 4096;
#define RUN(name, call, threads)                                                   \
  call; cudaDeviceSynchronize(); call; cudaDeviceSynchronize();                   \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                 \
  printf("%-28s threads=%4d  %8.2f cycles/iter\n", name, threads, (double)h / n);
  RUN("DFMA dependent chain", (dfma_chain<<<1, 32>>>(out, 1.0, 0.999, n, cyc)), 32)
  RUN("DADD dependent chain", (dadd_chain<<<1, 32>>>(out, 1.0, 0.999, n, cyc)), 32)
  RUN("FFMA dependent chain", (ffma_chain<<<1, 32>>>(outf, 1.0f, 0.999f, n, cyc)), 32)
  RUN("rcp.approx.f64 + DADD", (rcp_chain<<<1, 32>>>(out, 1.5, n, cyc)), 32)
  RUN("1.0/x + DADD", (div_chain<<<1, 32>>>(out, 1.5, n, cyc)), 32)
  RUN("DFMA chain, 256 threads", (dfma_chain<<<1, 256>>>(out, 1.0, 0.999, n, cyc)), 256)
  RUN("8 indep DFMA, 32 thr (/8)", (dfma_tput<<<1, 32>>>(out, 1.0, 0.999, n, cyc)), 32)
  RUN("8 indep DFMA, 256 thr (/8)", (dfma_tput<<<1, 256>>>(out, 1.0, 0.999, n, cyc)), 256)
  RUN("LDS+DADD+STS+BAR, 256 thr", (smem_bar_loop<<<1, 256>>>(out, n, cyc)), 256)
  RUN("LDS+DADD+STS+BAR, 64 thr", (smem_bar_loop<<<1, 64>>>(out, n, cyc)), 64)
  RUN("BAR only, 256 thr", (bar_only<<<1, 256>>>(n, cyc)), 256)
  RUN("BAR only, 1024 thr", (bar_only<<<1, 1024>>>(n, cyc)), 1024)
  return 0;
}
