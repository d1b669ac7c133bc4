// Round-2 diagnostic: the single-thread 12 x 12 Jacobi of k7_projective_depth.cu on the device vs the host.
#include <cmath>
#include <cstdio>
#include <cstdlib>
template <int K>
__host__ __device__ int jacobi_small(double (&a)[K][K], double (&e)[K][K], double* trace) {
  for (int r = 0; r < K; ++r)
    for (int c = 0; c < K; ++c) e[r][c] = r == c ? 1.0 : 0.0;
  int sweep = 0;
  for (; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int r = 0; r < K; ++r)
      for (int c = 0; c < K; ++c) {
        const double v = a[r][c] * a[r][c];
        if (r == c) dg += v; else off += v;
      }
    trace[sweep] = off;
    if (!(off > 1e-30 * dg)) break;
    for (int p = 0; p < K - 1; ++p)
      for (int q = p + 1; q < K; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double tau = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        for (int r = 0; r < K; ++r) {
          const double gp = a[r][p], gq = a[r][q];
          a[r][p] = c * gp - s * gq;
          a[r][q] = s * gp + c * gq;
          const double vp = e[r][p], vq = e[r][q];
          e[r][p] = c * vp - s * vq;
          e[r][q] = s * vp + c * vq;
        }
        for (int r = 0; r < K; ++r) {
          const double gp = a[p][r], gq = a[q][r];
          a[p][r] = c * gp - s * gq;
          a[q][r] = s * gp + c * gq;
        }
      }
  }
  return sweep;
}
// One warp per matrix: A and E in shared memory, lanes own rows.
template <int K>
__device__ int jacobi_warp(double* a, double* e, double* trace) {
  const int lane = threadIdx.x & 31;
  for (int k = lane; k < K * K; k += 32) e[k] = (k / K == k % K) ? 1.0 : 0.0;
  __syncwarp();
  int sweep = 0;
  for (; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int k = lane; k < K * K; k += 32) {
      const double v = a[k] * a[k];
      if (k / K == k % K) dg += v; else off += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      dg += __shfl_xor_sync(0xffffffffu, dg, o);
    }
    if (lane == 0 && trace) trace[sweep] = off;
    if (!(off > 1e-30 * dg)) break;
    for (int p = 0; p < K - 1; ++p)
      for (int q = p + 1; q < K; ++q) {
        const double apq = a[p * K + q];
        double c = 1.0, s = 0.0;
        if (apq != 0.0) {
          const double tau = (a[q * K + q] - a[p * K + p]) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          s = t * c;
        }
        __syncwarp();
        if (lane < K) {  // columns p, q of A and E, row `lane`
          const double gp = a[lane * K + p], gq = a[lane * K + q];
          a[lane * K + p] = c * gp - s * gq;
          a[lane * K + q] = s * gp + c * gq;
          const double vp = e[lane * K + p], vq = e[lane * K + q];
          e[lane * K + p] = c * vp - s * vq;
          e[lane * K + q] = s * vp + c * vq;
        }
        __syncwarp();
        if (lane < K) {  // rows p, q of A, column `lane`
          const double gp = a[p * K + lane], gq = a[q * K + lane];
          a[p * K + lane] = c * gp - s * gq;
          a[q * K + lane] = s * gp + c * gq;
        }
        __syncwarp();
      }
  }
  return sweep;
}
__global__ void kw(const double* A, double* out, double* trace, int* sweeps) {
  __shared__ double a[144], e[144];
  for (int i = threadIdx.x; i < 144; i += 32) a[i] = A[i];
  __syncwarp();
  const int sw = jacobi_warp<12>(a, e, trace);
  if (threadIdx.x == 0) {
    *sweeps = sw;
    int best = 0;
    for (int i = 1; i < 12; ++i) if (a[i * 12 + i] > a[best * 12 + best]) best = i;
    for (int i = 0; i < 12; ++i) out[i] = e[i * 12 + best];
    out[12] = a[best * 12 + best];
  }
}
__global__ void k(const double* A, double* out, double* trace, int* sweeps) {
  if (threadIdx.x != 3) return;
  double a[12][12], e[12][12];
  for (int i = 0; i < 144; ++i) a[i / 12][i % 12] = A[i];
  *sweeps = jacobi_small<12>(a, e, trace);
  int best = 0;
  for (int i = 1; i < 12; ++i) if (a[i][i] > a[best][best]) best = i;
  for (int i = 0; i < 12; ++i) out[i] = e[i][best];
  out[12] = a[best][best];
}
int main() {
  double B[50][12], A[144] = {}, a[12][12], e[12][12], tr[32] = {};
  srand(1);
  for (auto& r : B) for (int j = 0; j < 12; ++j) r[j] = (rand() / (double)RAND_MAX - 0.5) * (j % 3 == 2 ? 1.0 : 0.1);
  for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) { for (int kk = 0; kk < 50; ++kk) A[12 * i + j] += B[kk][i] * B[kk][j]; a[i][j] = A[12 * i + j]; }
  int hs = jacobi_small<12>(a, e, tr);
  int best = 0; for (int i = 1; i < 12; ++i) if (a[i][i] > a[best][best]) best = i;
  printf("host: sweeps %d lambda %.15g off trace %.3g %.3g %.3g %.3g\n", hs, a[best][best], tr[0], tr[1], tr[2], tr[3]);
  double *dA, *dout, *dtr; int* dsw;
  cudaMalloc(&dA, sizeof(A)); cudaMalloc(&dout, 13 * 8); cudaMalloc(&dtr, 32 * 8); cudaMalloc(&dsw, 4);
  cudaMemset(dtr, 0, 32 * 8);
  cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice);
  k<<<1, 32>>>(dA, dout, dtr, dsw);
  double out[13], dt[32]; int sw = -1;
  cudaError_t err = cudaDeviceSynchronize();
  cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost); cudaMemcpy(dt, dtr, sizeof(dt), cudaMemcpyDeviceToHost); cudaMemcpy(&sw, dsw, 4, cudaMemcpyDeviceToHost);
  printf("device: err %s sweeps %d lambda %.15g off trace %.3g %.3g %.3g %.3g\n", cudaGetErrorString(err), sw, out[12], dt[0], dt[1], dt[2], dt[3]);
  double d = 0; for (int i = 0; i < 12; ++i) d = fmax(d, fmin(fabs(out[i] - e[i][best]), fabs(out[i] + e[i][best])));
  printf("max eigenvector difference %.3g\n", d);
  cudaMemset(dtr, 0, 32 * 8);
  kw<<<1, 32>>>(dA, dout, dtr, dsw);
  err = cudaDeviceSynchronize();
  cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost); cudaMemcpy(dt, dtr, sizeof(dt), cudaMemcpyDeviceToHost); cudaMemcpy(&sw, dsw, 4, cudaMemcpyDeviceToHost);
  printf("device warp: err %s sweeps %d lambda %.15g off trace %.3g %.3g %.3g %.3g\n", cudaGetErrorString(err), sw, out[12], dt[0], dt[1], dt[2], dt[3]);
  d = 0; for (int i = 0; i < 12; ++i) d = fmax(d, fmin(fabs(out[i] - e[i][best]), fabs(out[i] + e[i][best])));
  printf("max eigenvector difference %.3g\n", d);
}
