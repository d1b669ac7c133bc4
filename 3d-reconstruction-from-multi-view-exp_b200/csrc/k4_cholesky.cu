// K4 (factor/solve part): blocked right-looking Cholesky of the reduced camera system and the
// back substitution, replacing reference lib/bundle_adjustment.py:146 (`np.linalg.solve`, LAPACK
// dgesv).  The system is symmetric positive definite, so LU with pivoting is not needed; the
// survey measured <= 3e-14 relative effect on every per-iteration cost.
//
// Layout: S is n_pad x n_pad row-major (lower triangle used); row `rhs_row` = n carries the
// right-hand side and is swept along as one more row, so after the factorisation it holds
// y = L^-1 b.  Only L^T x = y is left for chol_backsolve_kernel.
//
// Per 64-column panel two launches:
//   chol_panel_kernel   every block eliminates the tall panel [diagonal block; its 64 rows] in
//                       shared memory (re-factoring the diagonal block per block is cheaper than
//                       a third launch, and the rows below need no separate triangular solve).
//                       Block 0 carries the rows of the identity instead, which yields
//                       W = L_D^-T for the back substitution (no serial solve there either).
//                       Writes L rows in place and a k-major copy Lt for the update.
//   chol_update_kernel  trailing update S -= L_panel L_panel^T on 64x64 tiles.
#include <cstdlib>

#include "ba_common.cuh"

namespace ba {

constexpr int NB = kCholNB;

// Reciprocal to full double precision (<= ~1 ulp) with a short dependency chain: the hardware
// seed (MUFU.RCP64H, ~2^-20) and two Newton steps.  FP64 latency dominates the pivot chain below.
__device__ __forceinline__ double rcp_newton(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  return r;
}

constexpr int kPanelThreads = 8 * NB;  // 4 threads per row of the tall panel (16 columns each)
constexpr int kPanelCols = 16;          // columns per thread

__global__ void __launch_bounds__(kPanelThreads)
chol_panel_kernel(double* __restrict__ S, int ld, int n_rows, int k0, int nb,
                  double* __restrict__ Lt, double* __restrict__ W, ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  extern __shared__ double psm[];
  // tall panel T[128][65]: rows 0..63 the diagonal block, rows 64..127 this block's rows
  double(*T)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(psm);
  double* col = psm + 2 * NB * (NB + 1);  // [2][128] published pivot column (double buffered)
  double* piv = col + 4 * NB;             // [64] pivots d_k
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  // block 0 carries the rows of the identity (-> W = L_D^-T), block b >= 1 the 64 rows from rbase
  const int rbase = k0 + nb + ((int)blockIdx.x - 1) * NB;
  if (tid == 0) s_fail = 0;
  {
    // coalesced loads, all issued before the first use (L2 latency paid once)
    double vd[8], va[8];
    const int c = tid & 63;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = (tid >> 6) + 8 * u;
      vd[u] = (r == c) ? 1.0 : 0.0;
      if (r < nb && c <= r) vd[u] = S[(size_t)(k0 + r) * ld + k0 + c];
      va[u] = (blockIdx.x == 0 && r == c) ? 1.0 : 0.0;
      if (blockIdx.x != 0 && rbase + r < n_rows && c < nb) va[u] = S[(size_t)(rbase + r) * ld + k0 + c];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = (tid >> 6) + 8 * u;
      T[r][c] = vd[u];
      T[NB + r][c] = va[u];
    }
  }
  __syncthreads();
  // Right-looking elimination with unscaled columns (A = Lu D^-1 Lu^T):
  //   a[r][c] -= a[r][k] a[c][k] / d_k   for c > k  (diagonal block: c <= r matters).
  // The rows below the diagonal block ride along, so there is no separate triangular solve.
  // Thread (r, cg) keeps columns 16 cg .. 16 cg + 15 of row r in registers; warps are uniform in
  // cg.  Per column: the owners publish column k of all 128 rows, one barrier, then every thread
  // still holding live columns does <= 16 FMAs with broadcast loads of the pivot row.  The whole
  // column step is one dependent chain (publish -> barrier -> reciprocal -> scale -> FMA ->
  // publish), so the 4-way split of a row is what shortens the panel: 16 instead of 64 FMAs behind
  // every barrier, 16 warps instead of 4 to hide the FP64 and shared-memory latencies.
  const int r = tid & (2 * NB - 1);  // row of the tall panel
  const int cg = tid >> 7;           // column group
  double a[kPanelCols];
#pragma unroll
  for (int j = 0; j < kPanelCols; ++j) a[j] = T[r][kPanelCols * cg + j];
  for (int g = 0; g < NB / kPanelCols; ++g) {
#pragma unroll
    for (int kk = 0; kk < kPanelCols; ++kk) {
      const int k = kPanelCols * g + kk;
      double* ck = col + (kk & 1) * (2 * NB);
      if (cg == g) ck[r] = a[kk];
      __syncthreads();
      if (cg >= g) {  // warp-uniform; groups < g hold finished columns only
        const double d = ck[k];
        if (r == k && cg == g) piv[k] = d;
        const double la = ck[r] * rcp_newton(d);
        const double2* pr = reinterpret_cast<const double2*>(ck + kPanelCols * cg);
        if (cg > g) {
#pragma unroll
          for (int jp = 0; jp < kPanelCols / 2; ++jp) {
            const double2 cv = pr[jp];
            a[2 * jp] = fma(-la, cv.x, a[2 * jp]);
            a[2 * jp + 1] = fma(-la, cv.y, a[2 * jp + 1]);
          }
        } else {
          // No per-row predicates: rows r <= k and the upper triangle of the diagonal block
          // accumulate values nobody reads.
#pragma unroll
          for (int jp = 0; jp < kPanelCols / 2; ++jp) {
            if (2 * jp + 1 > kk) {
              const double2 cv = pr[jp];
              if (2 * jp > kk) a[2 * jp] = fma(-la, cv.x, a[2 * jp]);
              a[2 * jp + 1] = fma(-la, cv.y, a[2 * jp + 1]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kPanelCols; ++j) T[r][kPanelCols * cg + j] = a[j];
  __syncthreads();
  if (tid < NB) {
    const double d = piv[tid];
    if (tid < nb && !(d > 0.0)) s_fail = 1;
    piv[tid] = rsqrt(d);  // = 1 / L[tid][tid]
  }
  __syncthreads();
  for (int q = tid; q < 2 * NB * NB; q += kPanelThreads)  // L[r][c] = Lu[r][c] / sqrt(d_c)
    T[q >> 6][q & 63] *= piv[q & 63];
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid == 0 && s_fail) ctl->chol_fail = 1;
    // L_D (k-major) for the trailing update: Lt[m][k0 + r] = L_D[r][m]
    for (int q = tid; q < NB * NB; q += kPanelThreads) {
      const int m = q >> 6, rr = q & 63;
      if (m < nb && rr < nb) Lt[(size_t)m * ld + k0 + rr] = rr >= m ? T[rr][m] : 0.0;
    }
    // W = L_D^-T from the identity rows
    for (int q = tid; q < NB * NB; q += kPanelThreads) W[q] = T[NB + (q >> 6)][q & 63];
  } else {
    for (int q = tid; q < NB * NB; q += kPanelThreads) {
      const int rr = q >> 6, c = q & 63;
      if (rbase + rr < n_rows && c < nb) S[(size_t)(rbase + rr) * ld + k0 + c] = T[NB + rr][c];
    }
    for (int q = tid; q < NB * NB; q += kPanelThreads) {
      const int c = q >> 6, rr = q & 63;
      if (rbase + rr < n_rows && c < nb) Lt[(size_t)c * ld + rbase + rr] = T[NB + rr][c];
    }
  }
}

// Trailing update S[r][c] -= sum_m L[r][k0+m] L[c][k0+m] on 64x64 tiles of the lower triangle
// (rows/cols >= k0+nb, rows < n_rows).  The diagonal blocks of S keep their pre-factor content:
// nothing reads L_D from S afterwards (the back substitution uses W = L_D^-T).
__global__ void __launch_bounds__(256)
chol_update_kernel(double* __restrict__ S, int ld, int n_rows, int k0, int nb,
                   const double* __restrict__ Lt, const ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  const int tid = threadIdx.x;
  const int t0 = k0 + nb;
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  const int r0 = t0 + ti * 64, c0 = t0 + tj * 64;
  if (r0 >= n_rows) return;
  constexpr int MH = NB;  // the whole panel depth in one pass: all loads in flight at once
  extern __shared__ double usm[];
  double(*sA)[64 + 1] = reinterpret_cast<double(*)[64 + 1]>(usm);
  double(*sB)[64 + 1] = reinterpret_cast<double(*)[64 + 1]>(usm + MH * 65);
  const int ty = tid / 16, tx = tid % 16;  // thread owns rows ty+16*i, cols tx+16*j
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int m0 = 0; m0 < nb; m0 += MH) {
    __syncthreads();
    {
      double va[16], vb[16];
      const int x = tid & 63;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int m = m0 + (tid >> 6) + 4 * u;
        va[u] = (m < nb && r0 + x < n_rows) ? Lt[(size_t)m * ld + r0 + x] : 0.0;
        vb[u] = (m < nb && c0 + x < n_rows) ? Lt[(size_t)m * ld + c0 + x] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        sA[(tid >> 6) + 4 * u][x] = va[u];
        sB[(tid >> 6) + 4 * u][x] = vb[u];
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int m = 0; m < MH; ++m) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[m][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[m][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + ty + 16 * i, c = c0 + tx + 16 * j;
      if (r < n_rows && c <= r) S[(size_t)r * ld + c] -= acc[i][j];
    }
}

// ---- trailing update on the FP64 tensor cores (large systems) ---------------------------------
// Same update, 128x128 tiles, the whole 64-deep panel resident in shared memory (k-major rows of
// Lt, 16-byte cp.async, rows padded to 132 doubles: conflict-free 8-byte fragment loads), 16 warps
// each accumulating a 32x32 sub-tile with DMMA.8x8x4, then S -= acc in place.  Per tile 128 KB of
// L2 reads feed 2.1 MFLOP (the FMA kernel above moves 64 KB per 0.5 MFLOP and is bound by its
// shared-memory fragment loads: 8 LDS per 16 DFMA).
__device__ __forceinline__ void upd_cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void upd_dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int kUT = 128;        // tile edge
constexpr int kULDS = kUT + 4;  // padded row of the staged panel

__global__ void __launch_bounds__(512)
chol_update_dmma_kernel(double* __restrict__ S, int ld, int n_rows, int k0,
                        const double* __restrict__ Lt, const ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  extern __shared__ __align__(16) double dsm[];
  double* sA = dsm;               // [64][132] rows r0.. of the panel, k-major
  double* sB = dsm + NB * kULDS;  // [64][132] rows c0.. (unused on diagonal tiles)
  const int t0 = k0 + NB;
  int ti = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= (int)blockIdx.x) ++ti;
  while (ti * (ti + 1) / 2 > (int)blockIdx.x) --ti;
  const int tj = blockIdx.x - ti * (ti + 1) / 2;
  const bool diag = ti == tj;
  const int r0 = t0 + ti * kUT, c0 = t0 + tj * kUT;
  // stage the panel: 64 k-rows x 128 columns per operand = 4096 16-byte pieces, 8 per thread
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int q = threadIdx.x + u * 512;
    const int m = q >> 6, pc = q & 63;
    const int ca = r0 + 2 * pc, cb = c0 + 2 * pc;
    upd_cp_async16_zfill(sA + m * kULDS + 2 * pc, Lt + (size_t)m * ld + (ca < n_rows ? ca : 0), ca < n_rows ? 16 : 0);
    if (!diag)
      upd_cp_async16_zfill(sB + m * kULDS + 2 * pc, Lt + (size_t)m * ld + (cb < n_rows ? cb : 0), cb < n_rows ? 16 : 0);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = warp >> 2, wc = warp & 3;
  const int row0 = wr * 32 + (lane >> 2), col0 = wc * 32 + (lane >> 2);
  const int kq = lane & 3;
  // 8-row fragments of this warp that reach below n_rows; on diagonal tiles the sub-tiles
  // strictly above the diagonal (wc > wr) are skipped altogether
  int vm = 0, vn = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) vm += (r0 + wr * 32 + 8 * i) < n_rows;
#pragma unroll
  for (int j = 0; j < 4; ++j) vn += (c0 + wc * 32 + 8 * j) < n_rows;
  if (vm == 0 || vn == 0 || (diag && wc > wr)) return;
  const double* a = sA;
  const double* b = diag ? sA : sB;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  if (vm == 4 && vn == 4) {
#pragma unroll 4
    for (int kk = 0; kk < NB; kk += 4) {
      double fa[4], fb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) fa[i] = a[(kk + kq) * kULDS + row0 + 8 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) fb[j] = b[(kk + kq) * kULDS + col0 + 8 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) upd_dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
  } else {
#pragma unroll 2
    for (int kk = 0; kk < NB; kk += 4) {
      double fa[4], fb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) fa[i] = a[(kk + kq) * kULDS + row0 + 8 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) fb[j] = b[(kk + kq) * kULDS + col0 + 8 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (i < vm && j < vn) upd_dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
  }
  // S -= acc on the lower triangle; an accumulator pair sits at (row, col), (row, col + 1)
  const int orow = r0 + wr * 32 + (lane >> 2);
  const int ocol = c0 + wc * 32 + 2 * (lane & 3);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = orow + 8 * i, c = ocol + 8 * j;
      if (r >= n_rows || c > r) continue;
      double* p = S + (size_t)r * ld + c;
      if (c + 1 <= r) {
        double2 v = *reinterpret_cast<double2*>(p);
        v.x -= acc[i][j][0];
        v.y -= acc[i][j][1];
        *reinterpret_cast<double2*>(p) = v;
      } else {
        *p -= acc[i][j][0];
      }
    }
}

// Back substitution L^T x = y (y = row rhs_row of the factor) by 64-column blocks from the
// last one: rhs_B = y_B - L[below, B]^T x_below (GEMV), x_B = W_B rhs_B with W_B = L_BB^-T from
// the panel kernel.  One block; x lives in shared memory.
__global__ void __launch_bounds__(1024)
chol_backsolve_kernel(const double* __restrict__ S, int ld, int n, int rhs_row,
                      const double* __restrict__ W, double* __restrict__ dxi,
                      const ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  extern __shared__ double sm[];
  double* x = sm;             // [n]
  double* red = sm + n;       // [16][64]
  double* rhs = red + 16 * 64;  // [64]
  const int tid = threadIdx.x;
  const int tx = tid & 63, ty = tid >> 6;  // 64 columns x 16 row groups
  const int nblk = (n + 63) / 64;
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int b0 = blk * 64;
    const int b1 = b0 + 64 < n ? b0 + 64 : n;
    const int w = b1 - b0;
    // W row for the second product and y_B are fetched now so their L2 latency overlaps the GEMV
    const int wj = tid >> 4, wpart = tid & 15;
    const double* Wj = W + (size_t)blk * NB * NB + (size_t)wj * NB;
    double wv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wv[i] = Wj[wpart + 16 * i];
    const double yb = (ty == 0 && tx < w) ? S[(size_t)rhs_row * ld + b0 + tx] : 0.0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // independent chains: FP64 latency is long
    if (tx < w) {
      int k = b1 + ty;
      for (; k + 48 < n; k += 64) {
        const double s0 = S[(size_t)k * ld + b0 + tx], s1 = S[(size_t)(k + 16) * ld + b0 + tx];
        const double s2 = S[(size_t)(k + 32) * ld + b0 + tx], s3 = S[(size_t)(k + 48) * ld + b0 + tx];
        a0 = fma(s0, x[k], a0);
        a1 = fma(s1, x[k + 16], a1);
        a2 = fma(s2, x[k + 32], a2);
        a3 = fma(s3, x[k + 48], a3);
      }
      for (; k < n; k += 16) a0 = fma(S[(size_t)k * ld + b0 + tx], x[k], a0);
    }
    red[ty * 64 + tx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ty == 0) {
      double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int g = 0; g < 16; ++g) p[g & 3] += red[g * 64 + tx];
      rhs[tx] = yb - ((p[0] + p[1]) + (p[2] + p[3]));
    }
    __syncthreads();
    {
      // x_B[j] = sum_c W[j][c] rhs[c]; thread (j = tid / 16, part = tid % 16)
      const int j = wj, part = wpart;
      double s = (wv[0] * rhs[part] + wv[1] * rhs[part + 16]) +
                 (wv[2] * rhs[part + 32] + wv[3] * rhs[part + 48]);
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (part == 0 && j < w) x[b0 + j] = s;
    }
    __syncthreads();
  }
  for (int k = tid; k < n; k += 1024) dxi[k] = x[k];
}

int launch_cholesky_solve(ba_engine* e, bool conditional, cudaStream_t s) {
  const int use_ctl = conditional ? 1 : 0;
  constexpr size_t kPanelSmem = (2 * NB * (NB + 1) + 5 * NB) * sizeof(double);
  BA_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)kPanelSmem));
  constexpr size_t kUpdateSmem = 2 * NB * 65 * sizeof(double);
  BA_CUDA(cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)kUpdateSmem));
  constexpr size_t kDmmaSmem = 2 * NB * kULDS * sizeof(double);
  BA_CUDA(cudaFuncSetAttribute(chol_update_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)kDmmaSmem));
  // 128-tiles need enough of them to fill the GPU; small trailing matrices keep the 64-tile kernel
  constexpr int kDmmaUpdateMinRows = 2048;
  static const bool no_dmma_update = std::getenv("BA_CHOL_NO_DMMA") != nullptr;  // A/B timing only
  const int n = e->n_full, n_rows = e->rhs_row + 1, ld = e->n_pad;
  int panel = 0;
  for (int k0 = 0; k0 < n; k0 += NB, ++panel) {
    const int nb = n - k0 < NB ? n - k0 : NB;
    const int below = n_rows - (k0 + nb);
    const int pblocks = 1 + (below + NB - 1) / NB;
    chol_panel_kernel<<<pblocks, kPanelThreads, kPanelSmem, s>>>(e->P(), ld, n_rows, k0, nb, e->Lt,
                                              e->Winv + (size_t)panel * NB * NB, e->ctl, use_ctl);
    BA_LAUNCH_CHECK();
    if (below <= 0) continue;  // nothing below the last panel
    if (nb == NB && below >= kDmmaUpdateMinRows && !no_dmma_update) {
      const int nt = (below + kUT - 1) / kUT;
      chol_update_dmma_kernel<<<nt * (nt + 1) / 2, 512, kDmmaSmem, s>>>(e->P(), ld, n_rows, k0, e->Lt,
                                                                       e->ctl, use_ctl);
    } else {
      const int nt = (below + 63) / 64;
      dim3 grid(nt, nt);
      chol_update_kernel<<<grid, 256, kUpdateSmem, s>>>(e->P(), ld, n_rows, k0, nb, e->Lt, e->ctl, use_ctl);
    }
    BA_LAUNCH_CHECK();
  }
  const size_t smem = ((size_t)n + 16 * 64 + 64) * sizeof(double);
  if (smem > 220 * 1024) {
    set_error("reduced system too large for the single-block back substitution (n=%d)", n);
    return BA_ERR_INVALID;
  }
  BA_CUDA(cudaFuncSetAttribute(chol_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  chol_backsolve_kernel<<<1, 1024, smem, s>>>(e->P(), ld, n, e->rhs_row, e->Winv, e->dxi, e->ctl,
                                              use_ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

}  // namespace ba
