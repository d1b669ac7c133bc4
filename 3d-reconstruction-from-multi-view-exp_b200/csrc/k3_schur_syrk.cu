// K3: Schur-complement products  P = sum_j Y_j Y_j^T  (and the rhs row sum_j z_j^T Y_j^T).
//
// Replaces reference lib/bundle_adjustment.py:132-143 -- `(FtEinv @ matF).sum(axis=0)`, which
// materialises an (N, n, n) temporary, and `(FtEinv @ delta_X_E).sum(axis=0)`.
//
// Dense visibility: Y^T is stored k-major, Yt[3N_pad][ld] (ld = n_pad, a multiple of 8), and
// P = Yt^T Yt is a symmetric rank-k update on the FP64 tensor cores (DMMA.8x8x4 through
// mma.sync.m8n8k4.f64; tcgen05 has no f64 kind).  Only lower-triangle tiles are computed; edge
// tiles skip the 8x8 fragments beyond n_pad.  K (= 3 x points) is split over many more CTAs than
// SM slots so the hardware block scheduler evens out the tail, and the per-split partial tiles are
// summed in a fixed order (deterministic; no FP64 atomics).  z_j occupies column `rhs_row` of
// Yt, so row `rhs_row` of P is the rhs term for free.
//
// Pipeline: cp.async (LDGSTS) 16-byte copies, 3 stages of KC k-rows, shared-memory rows padded
// to TILE+4 doubles (stride = 4 mod 16: the 8-byte fragment loads of a half-warp hit 16
// distinct bank pairs -- ncu: 0 bank conflicts).
//
// Compute roofline (FP64 tensor): algorithmic flops = 3N * n (n + 1), n = 9M - 7.
//
// Sparse visibility: see k3_schur_sparse.cu (output-stationary, no atomics).
#include "ba_common.cuh"

namespace ba {

constexpr int kStages = 3;

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
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

__device__ __forceinline__ void tile_from_linear(int t, int& ti, int& tj) {
  ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  tj = t - ti * (ti + 1) / 2;
}

// TILE x TILE output tile per CTA; WR x WC warps, each owning a (TILE/WR) x (TILE/WC) sub-tile.
// SUB = false: the tile of this split is written to part[split][tile] (Schur product).
// SUB = true:  one split; the tile is subtracted in place from the lower triangle of the n_store x
//              n_store matrix `part` (leading dimension ld) -- the rank-k trailing update of the
//              blocked Cholesky (k4_cholesky.cu), where Yt is the k-major copy of a block column.
template <int TILE, int WR, int WC, int KC, bool SUB>
__global__ void __launch_bounds__(WR* WC * 32)
syrk_dmma_kernel(const double* __restrict__ Yt, int ld, int n_valid, int64_t n_chunks,
                 int chunks_per_split, int n_tiles, double* __restrict__ part,
                 const ba_lm_state* ctl, int n_store) {
  if (ctl && ctl->done) return;
  constexpr int NT = WR * WC * 32;
  constexpr int LDS = TILE + 4;
  constexpr int WM = TILE / WR, WN = TILE / WC;
  constexpr int FM = WM / 8, FN = WN / 8;
  constexpr int CPR = TILE / 2;               // 16-byte pieces per row
  constexpr int PER_THREAD = KC * CPR / NT;   // pieces per thread per operand per stage
  static_assert(KC * CPR % NT == 0, "stage copy must divide evenly");
  extern __shared__ __align__(16) double smem[];

  const int t = blockIdx.x;
  int ti, tj;
  tile_from_linear(t, ti, tj);
  const bool diag = ti == tj;

  const int split = blockIdx.y;
  const int64_t c_lo = (int64_t)split * chunks_per_split;
  int64_t c_hi = c_lo + chunks_per_split;
  if (c_hi > n_chunks) c_hi = n_chunks;
  const int nk = (int)(c_hi > c_lo ? c_hi - c_lo : 0);

  double* sA = smem;                                // [stage][KC][LDS]
  double* sB = smem + (size_t)kStages * KC * LDS;   // unused for diagonal tiles

  // per-thread copy slots: fixed (row, piece) positions; columns beyond n_valid are zero-filled
  int soff[PER_THREAD];
  int goffA[PER_THREAD], goffB[PER_THREAD];
  int bytesA[PER_THREAD], bytesB[PER_THREAD];
#pragma unroll
  for (int u = 0; u < PER_THREAD; ++u) {
    const int q = threadIdx.x + u * NT;
    const int row = q / CPR, pc = q % CPR;
    soff[u] = row * LDS + 2 * pc;
    const int ca = ti * TILE + 2 * pc, cb = tj * TILE + 2 * pc;
    bytesA[u] = ca < n_valid ? 16 : 0;
    bytesB[u] = cb < n_valid ? 16 : 0;
    goffA[u] = row * ld + (ca < n_valid ? ca : 0);
    goffB[u] = row * ld + (cb < n_valid ? cb : 0);
  }
  const double* gbase = Yt + (size_t)c_lo * KC * ld;

  auto load_stage = [&](int stage, int chunk) {
    const double* g = gbase + (size_t)chunk * KC * ld;
    double* da = sA + (size_t)stage * KC * LDS;
    double* db = sB + (size_t)stage * KC * LDS;
#pragma unroll
    for (int u = 0; u < PER_THREAD; ++u) {
      cp_async16_zfill(da + soff[u], g + goffA[u], bytesA[u]);
      if (!diag) cp_async16_zfill(db + soff[u], g + goffB[u], bytesB[u]);
    }
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = warp / WC, wc = warp % WC;
  const int row0 = wr * WM + (lane >> 2);
  const int col0 = wc * WN + (lane >> 2);
  const int kq = lane & 3;

  // fragments of this warp inside the valid part of P (n_valid is a multiple of 8)
  int vm = 0, vn = 0;
#pragma unroll
  for (int i = 0; i < FM; ++i) vm += (ti * TILE + wr * WM + 8 * i) < n_valid;
#pragma unroll
  for (int j = 0; j < FN; ++j) vn += (tj * TILE + wc * WN + 8 * j) < n_valid;
  // on a diagonal tile the warps whose sub-tile lies strictly above the diagonal have nothing to
  // compute (nothing above the diagonal is read anywhere): they only help with the copies
  if (diag && WM == WN && wc > wr) vm = 0;
  const bool full = vm == FM && vn == FN;

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
    const double* a = sA + (size_t)(kc % kStages) * KC * LDS;
    const double* b = diag ? a : sB + (size_t)(kc % kStages) * KC * LDS;
    if (full) {
#pragma unroll
      for (int kk = 0; kk < KC; kk += 4) {
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
    } else if (vm > 0 && vn > 0) {
#pragma unroll
      for (int kk = 0; kk < KC; kk += 4) {
        double fa[FM], fb[FN];
#pragma unroll
        for (int i = 0; i < FM; ++i) fa[i] = a[(kk + kq) * LDS + row0 + 8 * i];
#pragma unroll
        for (int j = 0; j < FN; ++j) fb[j] = b[(kk + kq) * LDS + col0 + 8 * j];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j)
            if (i < vm && j < vn) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
      }
    }
  }
  cp_async_wait<0>();

  const int orow = wr * WM + (lane >> 2);
  const int ocol = wc * WN + 2 * (lane & 3);
  if (SUB) {
    // S -= acc on the lower triangle; an accumulator pair sits at (r, c), (r, c + 1), c even
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
      for (int j = 0; j < FN; ++j) {
        const int r = ti * TILE + orow + 8 * i, c = tj * TILE + ocol + 8 * j;
        if (r >= n_store || c > r) continue;
        double* p = part + (size_t)r * ld + c;
        if (c + 1 <= r) {
          double2 v = *reinterpret_cast<double2*>(p);
          v.x -= acc[i][j][0];
          v.y -= acc[i][j][1];
          *reinterpret_cast<double2*>(p) = v;
        } else {
          *p -= acc[i][j][0];
        }
      }
    return;
  }
  // partial tile [split][tile][TILE][TILE]
  double* out = part + ((size_t)split * n_tiles + t) * TILE * TILE;
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j)
      *reinterpret_cast<double2*>(out + (size_t)(orow + 8 * i) * TILE + ocol + 8 * j) =
          make_double2(acc[i][j][0], acc[i][j][1]);
}

// P[tile] = sum over splits in a fixed order, restricted to the valid part of P.
template <int TILE>
__global__ void __launch_bounds__(256)
syrk_reduce_kernel(const double* __restrict__ part, int n_tiles, int splits, double* __restrict__ P,
                   int ld, int n_valid, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int t = blockIdx.x;
  int ti, tj;
  tile_from_linear(t, ti, tj);
  // blockIdx.y picks a slab of rows of the tile; 4 independent loads in flight per thread
  const int per = TILE * TILE / 2 / gridDim.y;
  for (int q = blockIdx.y * per + threadIdx.x; q < (blockIdx.y + 1) * per; q += blockDim.x) {
    const int r = q / (TILE / 2), c = 2 * (q % (TILE / 2));
    if (ti * TILE + r >= n_valid || tj * TILE + c >= n_valid) continue;
    const double* src = part + (size_t)t * TILE * TILE + (size_t)r * TILE + c;
    const size_t stride = (size_t)n_tiles * TILE * TILE;
    double2 s = make_double2(0.0, 0.0);
    int sp = 0;
    for (; sp + 4 <= splits; sp += 4) {
      const double2 v0 = *reinterpret_cast<const double2*>(src + (size_t)sp * stride);
      const double2 v1 = *reinterpret_cast<const double2*>(src + (size_t)(sp + 1) * stride);
      const double2 v2 = *reinterpret_cast<const double2*>(src + (size_t)(sp + 2) * stride);
      const double2 v3 = *reinterpret_cast<const double2*>(src + (size_t)(sp + 3) * stride);
      s.x = (((s.x + v0.x) + v1.x) + v2.x) + v3.x;
      s.y = (((s.y + v0.y) + v1.y) + v2.y) + v3.y;
    }
    for (; sp < splits; ++sp) {
      const double2 v = *reinterpret_cast<const double2*>(src + (size_t)sp * stride);
      s.x += v.x;
      s.y += v.y;
    }
    *reinterpret_cast<double2*>(P + (size_t)(ti * TILE + r) * ld + tj * TILE + c) = s;
  }
}

// Local U / dF blocks into the reduce buffer: a sharded run all-reduces one flat buffer.
__global__ void stage_camera_blocks_kernel(int n, const double* __restrict__ src,
                                           double* __restrict__ dst, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[k];
}

// ---- host side --------------------------------------------------------------------------------
static inline int syrk_kc(int tile) { return tile == 128 ? 32 : 16; }
static inline int syrk_occupancy(int tile) { return tile == 128 ? 1 : 4; }

// Number of K splits: enough CTAs (tiles x splits) to keep every SM slot busy through the tail,
// each still streaming >= 768 k-rows, preferring counts that fill whole waves.
int syrk_choose_splits(int n_pad, int tile, int64_t k_pad, int num_sms) {
  const int nt1 = (n_pad + tile - 1) / tile;
  const int n_tiles = nt1 * (nt1 + 1) / 2;
  const int64_t n_chunks = k_pad / syrk_kc(tile);
  const int slots = num_sms * syrk_occupancy(tile);
  const int64_t min_chunks = 768 / syrk_kc(tile);
  int64_t s_max = n_chunks / min_chunks;
  if (s_max < 1) s_max = 1;
  int64_t s_cap = (int64_t)24 * slots / n_tiles;
  if (s_cap < 1) s_cap = 1;
  // All tiles of one split stream the same k-slab of Yt (k_pad / splits rows x n_pad columns).
  // Tiles drift apart (diagonal and edge tiles do less work), so the slab is only served from L2
  // if it fits there as a whole: ncu showed 18.6 GB of DRAM reads for the 4.3 GB operand of C3
  // with 271 MB slabs.  Ask for slabs of at most 48 MB (of the 126 MB L2) when K allows it.
  const int64_t slab_cap = (int64_t)48 << 20;
  int64_t s_l2 = ((int64_t)k_pad * n_pad * 8 + slab_cap - 1) / slab_cap;
  if (s_l2 > s_max) s_l2 = s_max;
  if (s_cap < s_l2) s_cap = s_l2;
  const int s_hi = (int)(s_max < s_cap ? s_max : s_cap);
  int best = s_hi;
  double best_eff = 0.0;
  for (int s = s_hi; s >= (s_hi + 1) / 2 && s >= 1; --s) {
    const int64_t items = (int64_t)n_tiles * s;
    const int64_t waves = (items + slots - 1) / slots;
    const double eff = (double)items / (double)(waves * slots);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = s;
    }
  }
  return best;
}

template <int TILE, int WR, int WC, int KC>
static int launch_syrk(ba_engine* e, const ba_lm_state* ctl, cudaStream_t s) {
  const int nt1 = (e->n_pad + TILE - 1) / TILE;
  const int n_tiles = nt1 * (nt1 + 1) / 2;
  const int64_t n_chunks = e->k_pad / KC;
  const int splits = e->syrk_splits;
  const int cps = (int)((n_chunks + splits - 1) / splits);
  const size_t smem = (size_t)2 * kStages * KC * (TILE + 4) * sizeof(double);
  BA_CUDA(cudaFuncSetAttribute(syrk_dmma_kernel<TILE, WR, WC, KC, false>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_tiles, splits);
  {
    ProfScope ps(e, PG_SYRK, s);
    syrk_dmma_kernel<TILE, WR, WC, KC, false><<<grid, WR * WC * 32, smem, s>>>(
        e->Yt, e->n_pad, e->n_pad, n_chunks, cps, n_tiles, e->Spart, ctl, 0);
    BA_LAUNCH_CHECK();
  }
  const int slabs = n_tiles >= 4 * e->num_sms ? 1 : (n_tiles >= e->num_sms ? 4 : 8);
  syrk_reduce_kernel<TILE><<<dim3(n_tiles, slabs), 256, 0, s>>>(e->Spart, n_tiles, splits, e->P(),
                                                               e->n_pad, e->n_pad, ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// S[t0:, t0:] -= Lt[:, t0:]^T Lt[:, t0:]  (lower triangle, rows < n_rows) with Lt k-major
// [depth][ld]: the wide trailing update of the two-level blocked Cholesky.
template <int TILE, int WR, int WC, int KC>
static int launch_syrk_sub(double* S, int ld, int n_rows, int t0, const double* Lt, int depth,
                           const ba_lm_state* ctl, cudaStream_t s) {
  const int n_store = n_rows - t0;
  const int n_valid = (n_store + 7) / 8 * 8;  // <= ld - t0: ld is a multiple of 8, t0 of 64
  const int nt1 = (n_valid + TILE - 1) / TILE;
  const int n_tiles = nt1 * (nt1 + 1) / 2;
  const int n_chunks = depth / KC;
  const size_t smem = (size_t)2 * kStages * KC * (TILE + 4) * sizeof(double);
  BA_CUDA(cudaFuncSetAttribute(syrk_dmma_kernel<TILE, WR, WC, KC, true>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  syrk_dmma_kernel<TILE, WR, WC, KC, true><<<dim3(n_tiles, 1), WR * WC * 32, smem, s>>>(
      Lt + t0, ld, n_valid, n_chunks, n_chunks, n_tiles, S + (size_t)t0 * ld + t0, ctl, n_store);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

int launch_chol_wide_update(double* S, int ld, int n_rows, int t0, const double* Lt, int depth,
                            const ba_lm_state* ctl, cudaStream_t s) {
  // 128-tiles (one CTA per SM) once there are enough of them, 64-tiles (four per SM) below
  if (n_rows - t0 >= 2560) return launch_syrk_sub<128, 4, 4, 32>(S, ld, n_rows, t0, Lt, depth, ctl, s);
  return launch_syrk_sub<64, 2, 2, 16>(S, ld, n_rows, t0, Lt, depth, ctl, s);
}

int launch_k3(ba_engine* e, bool conditional, cudaStream_t s) {
  const ba_lm_state* ctl = conditional ? e->ctl : nullptr;
  {
    const int n = e->M * 90;
    stage_camera_blocks_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, e->Uloc, e->U(), ctl);
    BA_LAUNCH_CHECK();
  }
  if (e->dense) {
    if (e->syrk_tile == 128) return launch_syrk<128, 4, 4, 32>(e, ctl, s);
    return launch_syrk<64, 2, 2, 16>(e, ctl, s);
  }
  return launch_schur_sparse(e, ctl, s);
}

}  // namespace ba
