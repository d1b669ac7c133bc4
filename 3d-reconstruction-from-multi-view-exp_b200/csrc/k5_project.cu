// Batched re-projection of all points into all cameras: the step right after bundle adjustment in
// both reference scripts (euclidiean_reconstruction.py:63, affine_reconstruction.py:64).
//
// Replaces reference lib/camera.py:74-81 (`calc_projected_points`: a Python loop over the cameras,
// each building P = K [R^T | -R^T t] (:13) and projecting X_ext @ P^T followed by the perspective
// division (:28-32)).  K is a general 3x3 here, exactly as in the reference.
//
// One block row per camera (blockIdx.y); the 3x4 camera matrix is formed once per block in shared
// memory in the reference's order of operations (R^T t first, then the product with K); each
// thread projects one point and writes one double2 (coalesced).  HBM: 16 B written per
// (camera, point); X (24 B / point) is re-read per camera from L2.
#include "ba_common.cuh"

namespace ba {

__global__ void __launch_bounds__(256)
project_points_kernel(int64_t N, int M, const double* __restrict__ X, const double* __restrict__ K,
                      const double* __restrict__ R, const double* __restrict__ t,
                      double2* __restrict__ out) {
  __shared__ double P[12];
  const int i = blockIdx.y;
  if (threadIdx.x < 12) {
    const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
    const double* Ki = K + 9 * (size_t)i;
    const double* Ri = R + 9 * (size_t)i;
    const double* ti = t + 3 * (size_t)i;
    // E = [R^T | -R^T t]: E[k][c] = R[c][k], E[k][3] = -(R[0][k] t0 + R[1][k] t1 + R[2][k] t2)
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double e = c < 3 ? Ri[3 * c + k] : -(Ri[k] * ti[0] + Ri[3 + k] * ti[1] + Ri[6 + k] * ti[2]);
      acc += Ki[3 * r + k] * e;
    }
    P[threadIdx.x] = acc;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    const double x = X[3 * j], y = X[3 * j + 1], z = X[3 * j + 2];
    const double p = x * P[0] + y * P[1] + z * P[2] + P[3];
    const double q = x * P[4] + y * P[5] + z * P[6] + P[7];
    const double w = x * P[8] + y * P[9] + z * P[10] + P[11];
    out[(size_t)i * N + j] = make_double2(p / w, q / w);
  }
}

}  // namespace ba

using namespace ba;

extern "C" int ba_project_points(int device, int64_t n_points, int32_t n_cams, const double* X,
                                 const double* K, const double* R, const double* t, double* out,
                                 int mem, void* stream) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this engine has no CPU path");
    return BA_ERR_NO_DEVICE;
  }
  if (!X || !K || !R || !t || !out || n_points < 0 || n_cams < 0 || device < 0 || device >= ndev) {
    set_error("bad argument");
    return BA_ERR_INVALID;
  }
  if (n_points == 0 || n_cams == 0) return BA_OK;
  if (n_cams > 65535) { set_error("more than 65535 cameras per call are not supported"); return BA_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  BA_CUDA(cudaSetDevice(device));
  const size_t d = sizeof(double);
  const double *dX = X, *dK = K, *dR = R, *dt = t;
  double* dout = out;
  double* scratch = nullptr;
  if (mem == BA_MEM_HOST) {
    const size_t n_in = (size_t)3 * n_points + (size_t)21 * n_cams;
    const size_t n_out = (size_t)2 * n_points * n_cams;
    BA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), (n_in + n_out + 2) * d, s));
    double* p = scratch;
    BA_CUDA(cudaMemcpyAsync(p, X, (size_t)3 * n_points * d, cudaMemcpyHostToDevice, s)); dX = p; p += 3 * n_points;
    BA_CUDA(cudaMemcpyAsync(p, K, (size_t)9 * n_cams * d, cudaMemcpyHostToDevice, s)); dK = p; p += 9 * (size_t)n_cams;
    BA_CUDA(cudaMemcpyAsync(p, R, (size_t)9 * n_cams * d, cudaMemcpyHostToDevice, s)); dR = p; p += 9 * (size_t)n_cams;
    BA_CUDA(cudaMemcpyAsync(p, t, (size_t)3 * n_cams * d, cudaMemcpyHostToDevice, s)); dt = p; p += 3 * (size_t)n_cams;
    // 16-byte alignment of the double2 output
    p += (reinterpret_cast<uintptr_t>(p) & 8) ? 1 : 0;
    dout = p;
  }
  int64_t bx = (n_points + 255) / 256;
  if (bx > 1024) bx = 1024;
  dim3 grid((unsigned)bx, (unsigned)n_cams);
  project_points_kernel<<<grid, 256, 0, s>>>(n_points, n_cams, dX, dK, dR, dt, reinterpret_cast<double2*>(dout));
  BA_LAUNCH_CHECK();
  if (mem == BA_MEM_HOST) {
    BA_CUDA(cudaMemcpyAsync(out, dout, (size_t)2 * n_points * n_cams * d, cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaStreamSynchronize(s));
    BA_CUDA(cudaFreeAsync(scratch, s));
  }
  return BA_OK;
}
