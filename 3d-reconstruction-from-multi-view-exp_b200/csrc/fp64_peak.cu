// FP64 peak micro-benchmarks: register-resident DMMA.8x8x4 (mma.sync.m8n8k4.f64) and DFMA
// loops.  MEASURED_PEAKS.json has no FP64 figure, so the roofline denominator of the Schur SYRK
// (K3) and the Cholesky trailing update (K4) is measured on the box with these.
#include "ba_common.cuh"

namespace ba {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  double c[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) c[k] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[2 * k]), "+d"(c[2 * k + 1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += c[k];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(1024) dfma_peak_kernel(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  double c[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) c[k] = k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) c[k] = fma(a, c[k], b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += c[k];
  if (s == 123.456) out[0] = s;
}

// Mode 2: the two interleaved with equal FMA counts (8 DMMA = 2048 FMA per warp against 64 DFMA per
// lane): a sum near the single-kind peak says both run on the same FP64 datapath, a sum near twice
// that says the tensor path is separate.
__global__ void __launch_bounds__(256) mixed_peak_kernel(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  double c[16], d[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { c[k] = 0.0; d[k] = k; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[2 * k]), "+d"(c[2 * k + 1])
                   : "d"(a), "d"(b));
#pragma unroll
      for (int q = 0; q < 8; ++q) d[(8 * k + q) & 15] = fma(a, d[(8 * k + q) & 15], b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += c[k] + d[k];
  if (s == 123.456) out[0] = s;
}

int fp64_peak(int device, int use_dmma, double* tflops) {
  if (!tflops) { set_error("null argument"); return BA_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("no such CUDA device");
    return BA_ERR_NO_DEVICE;
  }
  BA_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BA_CUDA(cudaGetDeviceProperties(&prop, device));
  double* out = nullptr;
  BA_CUDA(cudaMalloc(&out, sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  // modes 100 + w: the DFMA loop with w warps per SM (one block per SM): FP64 rate at low occupancy
  const bool low_occ = use_dmma >= 100;
  const int blocks = prop.multiProcessorCount * (low_occ ? 1 : 8), threads = low_occ ? 32 * (use_dmma - 100) : 256,
            iters = 20000;
  if (low_occ) use_dmma = 0;
  if (threads < 32 || threads > 1024) { set_error("bad mode"); cudaFree(out); return BA_ERR_INVALID; }
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (use_dmma == 2) mixed_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + rep);
    else if (use_dmma) dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + rep);
    else dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + rep);
    g_launch_count++;
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    // DMMA.8x8x4: 8*8*4 FMA per warp instruction; DFMA: 1 FMA per thread instruction
    const double f_dmma = 2.0 * 256.0 * 8.0 * iters * (double)blocks * (threads / 32);
    const double f_dfma = 2.0 * 16.0 * iters * (double)blocks * threads;
    const double flops = use_dmma == 2 ? f_dmma + 4.0 * f_dfma : use_dmma ? f_dmma : f_dfma;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  BA_CUDA(cudaGetLastError());
  *tflops = best;
  return BA_OK;
}

}  // namespace ba
