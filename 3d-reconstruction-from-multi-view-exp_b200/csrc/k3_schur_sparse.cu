// K3 for sparse visibility: output-stationary Schur products, no FP64 atomics.
//
//   P[9i+r][9k+s] = sum over points j seen by both camera i and camera k (k <= i) of
//                   sum_d Y_ij[r][d] Y_kj[s][d]                (reference :132-135)
//   P[rhs][9i+a]  = sum_j sum_d Y_ij[a][d] z_j[d]               (reference :138-143)
//
// One CTA owns the blocks (i, k) of one camera i and 128 consecutive cameras k <= i, accumulated
// in shared memory (128 x 96 doubles, lane-major so the read-modify-writes are conflict-free);
// each of its 8 warps owns 16 of those cameras (16 resident warps per SM hide the L2 latency of
// the scattered Y reads better than 8 wider ones), so no two
// warps ever touch the same accumulator and the summation order is fixed (points in camera-major
// order): the result is bit-reproducible.  A warp walks the points of camera i 32 at a time (one
// point's metadata per lane, then a ballot), finds the point's cameras inside its own 16-camera
// half-group from a per-point bitmap + prefix count (built once by build_group_index_kernel), and
// for every such camera accumulates the 9x9 block with 27 lanes x 3 outputs.
//
// Work: sum_j 3 (9 m_j)(9 m_j + 1) flops as in SURVEY 8d; every Y block is read m_j / 2 times
// (from L2), every accumulator lives in shared memory until the single store at the end.
#include "ba_common.cuh"

namespace ba {

constexpr int kSR = 128;      // cameras k per CTA (8 warps x 16)
constexpr int kSThreads = 256;
constexpr int kSAcc = 96;     // doubles per camera block in shared memory: [rr][lane], 27 lanes used

// Per point j and 32-camera group g: bitmap of visible cameras and the number of the point's
// observations in lower groups (observations are sorted by camera within a point).
__global__ void build_group_index_kernel(int64_t N, int G, const int64_t* __restrict__ obs_ptr,
                                         const int32_t* __restrict__ obs_cam,
                                         uint32_t* __restrict__ grp_bits,
                                         uint16_t* __restrict__ grp_pre) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  uint32_t* bits = grp_bits + (size_t)j * G;
  for (int g = 0; g < G; ++g) bits[g] = 0u;
  for (int64_t o = obs_ptr[j]; o < obs_ptr[j + 1]; ++o) {
    const int c = obs_cam[o];
    bits[c >> 5] |= 1u << (c & 31);
  }
  uint16_t* pre = grp_pre + (size_t)j * G;
  int run = 0;
  for (int g = 0; g < G; ++g) {
    pre[g] = (uint16_t)run;
    run += __popc(bits[g]);
  }
}

int build_group_index(ba_engine* e, cudaStream_t s) {
  const int G = (e->M + 31) / 32;
  build_group_index_kernel<<<(int)((e->N + 127) / 128), 128, 0, s>>>(e->N, G, e->obs_ptr, e->obs_cam,
                                                                     e->grp_bits, e->grp_pre);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

__global__ void __launch_bounds__(kSThreads)
schur_sparse_rowblock_kernel(int M, int G, const int64_t* __restrict__ cam_ptr,
                             const int32_t* __restrict__ cm_perm,
                             const int32_t* __restrict__ obs_pt,
                             const int64_t* __restrict__ obs_ptr,
                             const uint32_t* __restrict__ grp_bits,
                             const uint16_t* __restrict__ grp_pre, const double* __restrict__ Ysp,
                             const double* __restrict__ Z, double* __restrict__ P, int ld,
                             int rhs_row, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int i = blockIdx.x, rg = blockIdx.y;
  if (rg * kSR > i) return;
  extern __shared__ double acc[];  // [kSR][kSAcc]: output (r = 3 (lane / 9) + rr, s = lane % 9) at [rr][lane]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int q = tid; q < kSR * kSAcc; q += kSThreads) acc[q] = 0.0;
  __syncthreads();

  const int k0 = rg * kSR + 16 * warp;   // this warp's 16 cameras
  const int g = k0 >> 5;                 // their 32-camera bitmap word
  const int half = (k0 >> 4) & 1;
  double* wacc = acc + warp * 16 * kSAcc;
  const bool calc = lane < 27;           // 27 lanes x 3 outputs = one 9x9 block
  const int s = lane % 9, r3 = (lane / 9) * 3;
  // cameras of the half-group that are <= i (lower triangle incl. the diagonal block)
  const uint32_t kmask = k0 > i ? 0u : (k0 + 15 <= i ? 0xffffu : ((2u << (i - k0)) - 1u));
  const bool rhs_warp = rg == 0 && warp == 0;  // also accumulates the rhs entries of camera i
  double brhs = 0.0;

  const int64_t q_lo = cam_ptr[i], q_hi = cam_ptr[i + 1];
  if (kmask != 0u || rhs_warp) {
    for (int64_t q0 = q_lo; q0 < q_hi; q0 += 32) {
      // one point of camera i per lane: observation id, its cameras in this group, where they start
      const int64_t q = q0 + lane;
      int o = 0, j = 0;
      uint32_t bits = 0u;
      int64_t base = 0;
      if (q < q_hi) {
        o = cm_perm[q];
        j = obs_pt[o];
        const uint32_t word = g < G ? grp_bits[(size_t)j * G + g] : 0u;
        bits = (half ? word >> 16 : word & 0xffffu) & kmask;
        base = obs_ptr[j] + (g < G ? grp_pre[(size_t)j * G + g] : 0) + (half ? __popc(word & 0xffffu) : 0);
      }
      unsigned todo = __ballot_sync(0xffffffffu, bits != 0u || (rhs_warp && q < q_hi));
      // The rows of Y_ij this lane needs are fetched one point ahead (the loop is bound by L2
      // latency, not by arithmetic).
      double yn[3][3];
      double zn = 0.0;
      auto fetch = [&](int src) {
        const int oo = __shfl_sync(0xffffffffu, o, src);
        const int jj = __shfl_sync(0xffffffffu, j, src);
        const double* Yi = Ysp + (size_t)oo * 27;
        if (calc) {
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int d = 0; d < 3; ++d) yn[rr][d] = Yi[d * 9 + r3 + rr];
        }
        if (rhs_warp && lane < 9) {
          const double* z = Z + 3 * (size_t)jj;
          zn = Yi[lane] * z[0] + Yi[9 + lane] * z[1] + Yi[18 + lane] * z[2];
        }
      };
      if (todo) fetch(__ffs(todo) - 1);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        uint32_t bb = __shfl_sync(0xffffffffu, bits, src);
        const int64_t bs = __shfl_sync(0xffffffffu, base, src);
        double yi[3][3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int d = 0; d < 3; ++d) yi[rr][d] = yn[rr][d];
        brhs += zn;
        if (todo) fetch(__ffs(todo) - 1);  // next point's rows, in flight during this point
        int idx = 0;
        while (bb) {
          // up to four cameras of the group at a time: all their loads are issued before the
          // shared-memory read-modify-writes
          double y[4][3];
          int bp[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            bp[u] = -1;
            if (bb) {
              bp[u] = __ffs(bb) - 1;
              bb &= bb - 1;
              if (calc) {
                const double* yk = Ysp + (size_t)(bs + idx) * 27 + s;
                y[u][0] = yk[0];
                y[u][1] = yk[9];
                y[u][2] = yk[18];
              }
              ++idx;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (bp[u] >= 0 && calc) {
              double* a = wacc + bp[u] * kSAcc + lane;
              a[0] += yi[0][0] * y[u][0] + yi[0][1] * y[u][1] + yi[0][2] * y[u][2];
              a[32] += yi[1][0] * y[u][0] + yi[1][1] * y[u][1] + yi[1][2] * y[u][2];
              a[64] += yi[2][0] * y[u][0] + yi[2][1] * y[u][1] + yi[2][2] * y[u][2];
            }
          }
        }
      }
    }
  }
  __syncthreads();
  for (int q = tid; q < kSR * kSAcc; q += kSThreads) {
    const int kl = q / kSAcc, rem = q - kl * kSAcc;
    const int rr = rem >> 5, ln = rem & 31;
    const int k = rg * kSR + kl;
    if (ln < 27 && k <= i && k < M)
      P[(size_t)(9 * i + 3 * (ln / 9) + rr) * ld + 9 * k + ln % 9] = acc[q];
  }
  if (rhs_warp && lane < 9) P[(size_t)rhs_row * ld + 9 * i + lane] = brhs;
}

int launch_schur_sparse(ba_engine* e, const ba_lm_state* ctl, cudaStream_t s) {
  const int G = (e->M + 31) / 32;
  const size_t smem = (size_t)kSR * kSAcc * sizeof(double);
  BA_CUDA(cudaFuncSetAttribute(schur_sparse_rowblock_kernel,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(e->M, (e->M + kSR - 1) / kSR);
  ProfScope ps(e, PG_SYRK, s);
  schur_sparse_rowblock_kernel<<<grid, kSThreads, smem, s>>>(e->M, G, e->cam_ptr, e->cm_perm, e->obs_pt,
                                                      e->obs_ptr, e->grp_bits, e->grp_pre, e->Ysp,
                                                      e->Z, e->P(), e->n_pad, e->rhs_row, ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

}  // namespace ba
