// Gauge normalisation / de-normalisation on the device.
//
// Replaces the O(N + M) NumPy passes of reference lib/bundle_adjustment.py:
//   :23-33    c0c1_len = |R0[:, k] . (t1 - t0)|            (saved for the way back)
//   :208-240  _transform_to_normalize_coodinates:  X <- (X - t0) R0 / s,  R <- R0^T R,
//             t <- (t - t0) R0 / s,  s = sign((t1 - t0)[k]) * (R0^T (t1 - t0))[k]
//             -- the divisor takes its sign in the WORLD frame and its magnitude in camera 0's
//             frame, so it can be negative (reflected scene); kept verbatim.
//   :242-258  _inverse_transform_to_global_coordinates:  X <- (len X) R0^T + t0, R <- R0 R,
//             t <- (len t) R0^T + t0
//   :283-289  _get_K
// so that a caller with device-resident (or pinned) arrays never touches the scene on the host.
#include "ba_common.cuh"

namespace ba {

// gauge block: R0 (9, row-major), t0 (3), s, len, 2 pad
__global__ void gauge_prepare_kernel(const double* __restrict__ R, const double* __restrict__ t, int axis,
                                     double* __restrict__ g) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double d[3];
  for (int a = 0; a < 3; ++a) d[a] = t[3 + a] - t[a];
  for (int k = 0; k < 9; ++k) g[k] = R[k];
  for (int a = 0; a < 3; ++a) g[9 + a] = t[a];
  const double comp = R[0 * 3 + axis] * d[0] + R[1 * 3 + axis] * d[1] + R[2 * 3 + axis] * d[2];
  const double sg = d[axis] > 0.0 ? 1.0 : (d[axis] < 0.0 ? -1.0 : 0.0);  // numpy.sign
  g[12] = sg * comp;
  g[13] = fabs(comp);
}

__global__ void gauge_points_kernel(int64_t N, double* __restrict__ X, const double* __restrict__ g) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const double x0 = X[3 * j] - g[9], x1 = X[3 * j + 1] - g[10], x2 = X[3 * j + 2] - g[11];
  const double s = g[12];
#pragma unroll
  for (int b = 0; b < 3; ++b) X[3 * j + b] = (x0 * g[b] + x1 * g[3 + b] + x2 * g[6 + b]) / s;
}

__global__ void gauge_cams_kernel(int M, double* __restrict__ R, double* __restrict__ t,
                                  const double* __restrict__ g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double Ri[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Ri[k] = R[9 * (size_t)i + k];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
      R[9 * (size_t)i + 3 * a + b] = g[a] * Ri[b] + g[3 + a] * Ri[3 + b] + g[6 + a] * Ri[6 + b];  // R0^T R
  const double x0 = t[3 * i] - g[9], x1 = t[3 * i + 1] - g[10], x2 = t[3 * i + 2] - g[11];
  const double s = g[12];
#pragma unroll
  for (int b = 0; b < 3; ++b) t[3 * i + b] = (x0 * g[b] + x1 * g[3 + b] + x2 * g[6 + b]) / s;
}

__global__ void ungauge_points_kernel(int64_t N, const double* __restrict__ X, double* __restrict__ out,
                                      const double* __restrict__ g) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const double len = g[13];
  const double x0 = len * X[3 * j], x1 = len * X[3 * j + 1], x2 = len * X[3 * j + 2];
#pragma unroll
  for (int b = 0; b < 3; ++b) out[3 * j + b] = (x0 * g[3 * b] + x1 * g[3 * b + 1] + x2 * g[3 * b + 2]) + g[9 + b];
}

__global__ void ungauge_cams_kernel(int M, double f0, const double* __restrict__ f, const double* __restrict__ u,
                                    const double* __restrict__ R, const double* __restrict__ t,
                                    double* __restrict__ K_out, double* __restrict__ R_out,
                                    double* __restrict__ t_out, const double* __restrict__ g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const double* Ri = R + 9 * (size_t)i;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
      R_out[9 * (size_t)i + 3 * a + b] = g[3 * a] * Ri[b] + g[3 * a + 1] * Ri[3 + b] + g[3 * a + 2] * Ri[6 + b];  // R0 R
  const double len = g[13];
  const double x0 = len * t[3 * i], x1 = len * t[3 * i + 1], x2 = len * t[3 * i + 2];
#pragma unroll
  for (int b = 0; b < 3; ++b) t_out[3 * i + b] = (x0 * g[3 * b] + x1 * g[3 * b + 1] + x2 * g[3 * b + 2]) + g[9 + b];
  double* K = K_out + 9 * (size_t)i;
  K[0] = f[i]; K[1] = 0.0; K[2] = u[2 * i];
  K[3] = 0.0; K[4] = f[i]; K[5] = u[2 * i + 1];
  K[6] = 0.0; K[7] = 0.0; K[8] = f0;
}

}  // namespace ba

using namespace ba;

extern "C" {

int ba_set_state_global(ba_engine* e, const double* X, const double* R, const double* t, const double* f,
                        const double* u, int mem, void* stream) {
  if (!e || !X || !R || !t || !f || !u) { set_error("null argument"); return BA_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  BA_CUDA(cudaSetDevice(e->device));
  const cudaMemcpyKind kind = mem == BA_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const size_t d = sizeof(double);
  BA_CUDA(cudaMemcpyAsync(e->X[0], X, (size_t)3 * e->N * d, kind, s));
  BA_CUDA(cudaMemcpyAsync(e->cam[0].R, R, (size_t)9 * e->M * d, kind, s));
  BA_CUDA(cudaMemcpyAsync(e->cam[0].t, t, (size_t)3 * e->M * d, kind, s));
  BA_CUDA(cudaMemcpyAsync(e->cam[0].f, f, (size_t)e->M * d, kind, s));
  BA_CUDA(cudaMemcpyAsync(e->cam[0].u, u, (size_t)2 * e->M * d, kind, s));
  gauge_prepare_kernel<<<1, 32, 0, s>>>(e->cam[0].R, e->cam[0].t, e->axis, e->gauge);
  BA_LAUNCH_CHECK();
  gauge_points_kernel<<<(unsigned)((e->N + 255) / 256), 256, 0, s>>>(e->N, e->X[0], e->gauge);
  BA_LAUNCH_CHECK();
  gauge_cams_kernel<<<(e->M + 127) / 128, 128, 0, s>>>(e->M, e->cam[0].R, e->cam[0].t, e->gauge);
  BA_LAUNCH_CHECK();
  if (mem == BA_MEM_HOST) BA_CUDA(cudaStreamSynchronize(s));
  e->have_state = true;
  e->have_gauge = true;
  return BA_OK;
}

int ba_get_state_global(ba_engine* e, int which, double* X, double* K, double* R, double* t, int mem,
                        void* stream) {
  if (!e || (which != 0 && which != 1) || !X || !K || !R || !t) { set_error("bad argument"); return BA_ERR_INVALID; }
  if (!e->have_gauge) { set_error("ba_set_state_global must come first"); return BA_ERR_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  BA_CUDA(cudaSetDevice(e->device));
  const size_t d = sizeof(double);
  double* scratch = nullptr;  // host destinations need a device-side staging copy
  double *dX = X, *dK = K, *dR = R, *dt = t;
  if (mem == BA_MEM_HOST) {
    BA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), ((size_t)3 * e->N + (size_t)21 * e->M) * d, s));
    dX = scratch;
    dK = scratch + 3 * (size_t)e->N;
    dR = dK + 9 * (size_t)e->M;
    dt = dR + 9 * (size_t)e->M;
  }
  ungauge_points_kernel<<<(unsigned)((e->N + 255) / 256), 256, 0, s>>>(e->N, e->X[which], dX, e->gauge);
  BA_LAUNCH_CHECK();
  const CamState& c = e->cam[which];
  ungauge_cams_kernel<<<(e->M + 127) / 128, 128, 0, s>>>(e->M, e->f0, c.f, c.u, c.R, c.t, dK, dR, dt, e->gauge);
  BA_LAUNCH_CHECK();
  if (mem == BA_MEM_HOST) {
    BA_CUDA(cudaMemcpyAsync(X, dX, (size_t)3 * e->N * d, cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaMemcpyAsync(K, dK, (size_t)9 * e->M * d, cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaMemcpyAsync(R, dR, (size_t)9 * e->M * d, cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaMemcpyAsync(t, dt, (size_t)3 * e->M * d, cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaFreeAsync(scratch, s));
    BA_CUDA(cudaStreamSynchronize(s));
  }
  return BA_OK;
}

}  // extern "C"
