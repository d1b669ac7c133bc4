// K3: Schur-complement products  P = sum_j Y_j Y_j^T  (and the rhs row sum_j z_j^T Y_j^T).
//
// Replaces reference lib/bundle_adjustment.py:132-143 -- `(FtEinv @ matF).sum(axis=0)`, which
// materialises an (N, n, n) temporary, and `(FtEinv @ delta_X_E).sum(axis=0)`.
//
// Dense visibility: Y^T is stored k-major, Yt[3N_pad][ld] (ld = n_pad), and P = Yt^T Yt is a
// symmetric rank-k update computed on the FP64 tensor cores (DMMA.8x8x4 via
// mma.sync.m8n8k4.f64 -- tcgen05 has no f64 kind).  Only lower-triangle tiles are computed;
// K (= 3 x points) is split across CTAs to fill the 148 SMs and the per-split partial tiles are
// summed in a fixed order (deterministic, no FP64 atomics).  z_j occupies column `rhs_row` of
// Yt, so row `rhs_row` of P is the rhs term for free.
//
// Compute roofline (FP64 tensor): flops = 2 * (#lower tiles * TILE^2) * 3N.
//
// Sparse visibility: per-point outer products of the visible 9x3 blocks, scattered into P
// with FP64 reductions (first correct version; see DESIGN.md for the planned tile-gather
// formulation).
#include "ba_common.cuh"

namespace ba {

constexpr int kKC = 16;      // k rows per pipeline stage
constexpr int kStages = 3;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// TILE x TILE output tile per CTA; WR x WC warps, each owning a (TILE/WR) x (TILE/WC) sub-tile.
template <int TILE, int WR, int WC>
__global__ void __launch_bounds__(WR* WC * 32)
syrk_dmma_kernel(const double* __restrict__ Yt, int ld, int64_t n_chunks, int chunks_per_split,
                 int n_tiles, double* __restrict__ part, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  constexpr int NT = WR * WC * 32;
  constexpr int LDS = TILE + 4;  // stride = 4 (mod 16) doubles: conflict-free fragment loads
  constexpr int WM = TILE / WR, WN = TILE / WC;
  constexpr int FM = WM / 8, FN = WN / 8;
  extern __shared__ __align__(16) double smem[];

  // lower-triangle tile (ti >= tj) from the linear index
  const int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const bool diag = ti == tj;

  const int split = blockIdx.y;
  const int64_t c_lo = (int64_t)split * chunks_per_split;
  int64_t c_hi = c_lo + chunks_per_split;
  if (c_hi > n_chunks) c_hi = n_chunks;
  const int nk = (int)(c_hi > c_lo ? c_hi - c_lo : 0);

  double* sA = smem;                                  // [stage][kKC][LDS]
  double* sB = smem + (size_t)kStages * kKC * LDS;    // unused for diagonal tiles
  const double* gA = Yt + (size_t)c_lo * kKC * ld + (size_t)ti * TILE;
  const double* gB = Yt + (size_t)c_lo * kKC * ld + (size_t)tj * TILE;

  auto load_stage = [&](int stage, int chunk) {
    constexpr int CPR = TILE / 2;  // 16-byte pieces per row
    const double* a = gA + (size_t)chunk * kKC * ld;
    const double* b = gB + (size_t)chunk * kKC * ld;
    double* da = sA + (size_t)stage * kKC * LDS;
    double* db = sB + (size_t)stage * kKC * LDS;
#pragma unroll
    for (int q = threadIdx.x; q < kKC * CPR; q += NT) {
      const int row = q / CPR, pc = q % CPR;
      cp_async16(da + row * LDS + 2 * pc, a + (size_t)row * ld + 2 * pc);
      if (!diag) cp_async16(db + row * LDS + 2 * pc, b + (size_t)row * ld + 2 * pc);
    }
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = warp / WC, wc = warp % WC;
  const int row0 = wr * WM + (lane >> 2);
  const int col0 = wc * WN + (lane >> 2);
  const int kq = lane & 3;

  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    const int nxt = kc + kStages - 1;
    if (nxt < nk) load_stage(nxt % kStages, nxt);
    cp_async_commit();
    const double* a = sA + (size_t)(kc % kStages) * kKC * LDS;
    const double* b = diag ? a : sB + (size_t)(kc % kStages) * kKC * LDS;
#pragma unroll
    for (int kk = 0; kk < kKC; kk += 4) {
      double fa[FM], fb[FN];
#pragma unroll
      for (int i = 0; i < FM; ++i) fa[i] = a[(kk + kq) * LDS + row0 + 8 * i];
#pragma unroll
      for (int j = 0; j < FN; ++j) fb[j] = b[(kk + kq) * LDS + col0 + 8 * j];
#pragma unroll
      for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
  }
  cp_async_wait<0>();

  // partial tile [split][tile][TILE][TILE]
  double* out = part + ((size_t)split * n_tiles + t) * TILE * TILE;
  const int orow = wr * WM + (lane >> 2);
  const int ocol = wc * WN + 2 * (lane & 3);
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j)
      *reinterpret_cast<double2*>(out + (size_t)(orow + 8 * i) * TILE + ocol + 8 * j) =
          make_double2(acc[i][j][0], acc[i][j][1]);
}

// P[tile] = sum over splits (fixed order); also stages the local U / dF blocks into the reduce
// buffer so a sharded run all-reduces everything in one buffer.
template <int TILE>
__global__ void __launch_bounds__(256)
syrk_reduce_kernel(const double* __restrict__ part, int n_tiles, int splits, double* __restrict__ P,
                   int ld, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  for (int q = threadIdx.x; q < TILE * TILE / 2; q += blockDim.x) {
    const int r = q / (TILE / 2), c = 2 * (q % (TILE / 2));
    double2 s = make_double2(0.0, 0.0);
    for (int sp = 0; sp < splits; ++sp) {
      const double2 v = *reinterpret_cast<const double2*>(
          part + ((size_t)sp * n_tiles + t) * TILE * TILE + (size_t)r * TILE + c);
      s.x += v.x;
      s.y += v.y;
    }
    *reinterpret_cast<double2*>(P + (size_t)(ti * TILE + r) * ld + tj * TILE + c) = s;
  }
}

__global__ void stage_camera_blocks_kernel(int n, const double* __restrict__ src,
                                           double* __restrict__ dst, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[k];
}

// ---- sparse visibility ---------------------------------------------------------------------
__global__ void zero_kernel(double* __restrict__ p, int64_t n, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) p[k] = 0.0;
}

__global__ void __launch_bounds__(128)
schur_sparse_atomic_kernel(int64_t N, const int64_t* __restrict__ obs_ptr,
                           const int32_t* __restrict__ obs_cam, const double* __restrict__ Ysp,
                           const double* __restrict__ Z, double* __restrict__ P, int ld,
                           int rhs_row, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < N; j += nwarps) {
    const int64_t lo = obs_ptr[j], hi = obs_ptr[j + 1];
    const int m = (int)(hi - lo);
    const double z0 = Z[3 * j], z1 = Z[3 * j + 1], z2 = Z[3 * j + 2];
    for (int a = 0; a < m; ++a) {
      const int ia = obs_cam[lo + a];
      double ya[27];
      const double* Ya = Ysp + (size_t)(lo + a) * 27;
#pragma unroll
      for (int k = 0; k < 27; ++k) ya[k] = Ya[k];
      if (lane < 9) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k)
          if (k == lane) v = z0 * ya[k] + z1 * ya[9 + k] + z2 * ya[18 + k];
        atomicAdd(P + (size_t)rhs_row * ld + 9 * ia + lane, v);
      }
      for (int b = lane; b <= a; b += 32) {
        const int ib = obs_cam[lo + b];
        double yb[27];
        const double* Yb = Ysp + (size_t)(lo + b) * 27;
#pragma unroll
        for (int k = 0; k < 27; ++k) yb[k] = Yb[k];
        double* dst = P + (size_t)(9 * ia) * ld + 9 * ib;  // ia >= ib: lower triangle
#pragma unroll
        for (int r = 0; r < 9; ++r)
#pragma unroll
          for (int s = 0; s < 9; ++s)
            atomicAdd(dst + (size_t)r * ld + s,
                      ya[r] * yb[s] + ya[9 + r] * yb[9 + s] + ya[18 + r] * yb[18 + s]);
      }
    }
  }
}

template <int TILE, int WR, int WC>
static int launch_syrk(ba_engine* e, const ba_lm_state* ctl, cudaStream_t s) {
  const int nt1 = e->n_pad / TILE;
  const int n_tiles = nt1 * (nt1 + 1) / 2;
  const int64_t n_chunks = e->k_pad / kKC;
  const int splits = e->syrk_splits;
  const int cps = (int)((n_chunks + splits - 1) / splits);
  const size_t smem = (size_t)2 * kStages * kKC * (TILE + 4) * sizeof(double);
  BA_CUDA(cudaFuncSetAttribute(syrk_dmma_kernel<TILE, WR, WC>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_tiles, splits);
  {
    ProfScope ps(e, PG_SYRK, s);
    syrk_dmma_kernel<TILE, WR, WC><<<grid, WR * WC * 32, smem, s>>>(e->Yt, e->n_pad, n_chunks, cps,
                                                                   n_tiles, e->Spart, ctl);
    BA_LAUNCH_CHECK();
  }
  syrk_reduce_kernel<TILE><<<n_tiles, 256, 0, s>>>(e->Spart, n_tiles, splits, e->P(), e->n_pad, ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// Number of K splits so that tiles x splits fills the machine (~2 CTAs per SM, whole waves).
int syrk_choose_splits(int n_pad, int tile, int64_t k_pad, int num_sms) {
  const int nt1 = n_pad / tile;
  const int n_tiles = nt1 * (nt1 + 1) / 2;
  const int64_t n_chunks = k_pad / kKC;
  const int slots = 2 * num_sms;
  int splits = (slots + n_tiles - 1) / n_tiles;
  if (n_tiles >= slots) splits = 1;
  // each split should still stream a few hundred chunks
  const int64_t max_splits = n_chunks / 64 > 0 ? n_chunks / 64 : 1;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

int launch_k3(ba_engine* e, bool conditional, cudaStream_t s) {
  const ba_lm_state* ctl = conditional ? e->ctl : nullptr;
  // local U / dF into the reduce buffer
  {
    const int n = e->M * 90;
    stage_camera_blocks_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, e->Uloc, e->U(), ctl);
    BA_LAUNCH_CHECK();
  }
  if (e->dense) {
    if (e->syrk_tile == 128) return launch_syrk<128, 2, 4>(e, ctl, s);
    return launch_syrk<64, 2, 2>(e, ctl, s);
  }
  const int64_t np = (int64_t)e->n_pad * e->n_pad;
  zero_kernel<<<e->num_sms * 8, 256, 0, s>>>(e->P(), np, ctl);
  BA_LAUNCH_CHECK();
  int64_t blocks = (e->N + 3) / 4;
  const int64_t cap = (int64_t)e->num_sms * 16;
  schur_sparse_atomic_kernel<<<(int)(blocks < cap ? blocks : cap), 128, 0, s>>>(
      e->N, e->obs_ptr, e->obs_cam, e->Ysp, e->Z, e->P(), e->n_pad, e->rhs_row, ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

}  // namespace ba
