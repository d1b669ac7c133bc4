// K3 for sparse visibility: pair-stationary Schur products, accumulators in registers.
//
//   P[9i+a][9k+b] = sum over points j seen by both camera i and camera k (k <= i) of
//                   sum_d Y_ij[a][d] Y_kj[b][d]                (reference :132-135)
//   P[rhs][9i+a]  = sum_j sum_d Y_ij[a][d] z_j[d]               (reference :138-143)
//
// With Y_ij = Jc_ij^T T_ij (Jc the 2x9 camera Jacobian, T = 2 Jx L_j^-T the 2x3 point factor, both
// written by K2b into the camera-major array Ycm[q][24] = [Jc row 0 | Jc row 1 | T row 0 | T row 1])
// the pair product factors through a 2x2 matrix:
//   Y_ij Y_kj^T = Jc_ij^T (T_ij T_kj^T) Jc_kj      -- 210 instead of 243 FMAs, 192 instead of 216 B.
//
// schur_pairs_kernel: ONE WARP PER CAMERA PAIR (i, k), k < i.  The warp intersects the two cameras'
// point bitmaps 32 words at a time; a hit's position in either camera's slice of Ycm is the prefix
// count stored next to the bitmap word (+ a popc).  Hits are queued; every 32 hits make a round:
// the 64 blocks of the round are gathered into shared memory with 16-byte cp.async (12 lanes per
// block: whole sectors), double buffered so the gather of round r+1 is in flight while round r is
// computed; then every lane takes ONE common point and accumulates its 9x9 contribution into 81
// registers.  There is no shared-memory or global read-modify-write at all; the accumulators meet
// once per pair in a fixed-order butterfly.  Order of summation depends on the data only: runs
// are bit-reproducible.
//
// Bound: every pair-point moves 2 x 192 B from L2 to the SM for 210 FMAs, i.e. the kernel is bound
// by L2 -> SM bandwidth (the Y blocks of C4 are 19 GB, each read m_j = 100 times), not by FP64.
//
// schur_diag_kernel: the diagonal blocks P[9i..][9i..] and the rhs row, a segmented reduction over
// the camera's contiguous slice of Ycm (fixed chunking and order).
#include "ba_common.cuh"

namespace ba {

constexpr int kYB = kYcm;          // doubles per block in Ycm (24 = 192 B)
constexpr int kYS = 26;            // doubles per block in shared memory (208 B = 13 x 16: odd -> conflict-free LDS.128)
constexpr int kPairTile = 32;      // pairs are scheduled in kPairTile x kPairTile tiles of (i, k)
constexpr int kDiagPart = 54;      // 45 unique entries of the diagonal block + 9 rhs entries

// ---- index: per camera bitmap over points + running prefix count --------------------------------
__global__ void bitmap_fill_kernel(int64_t nobs, int64_t Wp, const int32_t* __restrict__ obs_cam,
                                   const int32_t* __restrict__ obs_pt, uint2* __restrict__ bitpre) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nobs; o += stride) {
    const int j = obs_pt[o];
    atomicOr(&bitpre[(size_t)obs_cam[o] * Wp + (j >> 5)].x, 1u << (j & 31));
  }
}

// One block per camera: y = number of set bits in the words before this one.
__global__ void __launch_bounds__(1024)
bitmap_prefix_kernel(int64_t Wp, uint2* __restrict__ bitpre) {
  __shared__ uint32_t part[1024];
  uint2* row = bitpre + (size_t)blockIdx.x * Wp;
  const int64_t per = (Wp + 1023) / 1024;
  const int64_t lo = per * threadIdx.x, hi = lo + per < Wp ? lo + per : Wp;
  uint32_t s = 0;
  for (int64_t w = lo; w < hi; ++w) s += __popc(row[w].x);
  part[threadIdx.x] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over the 1024 partial counts
  for (int off = 1; off < 1024; off <<= 1) {
    const uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;
  for (int64_t w = lo; w < hi; ++w) {
    row[w].y = run;
    run += __popc(row[w].x);
  }
}

__global__ void invert_perm_kernel(int64_t n, const int32_t* __restrict__ perm, int32_t* __restrict__ inv) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) inv[perm[q]] = (int32_t)q;
}

int build_pair_index(ba_engine* e, cudaStream_t s) {
  BA_CUDA(cudaMemsetAsync(e->bitpre, 0, (size_t)e->M * e->Wp * sizeof(uint2), s));
  bitmap_fill_kernel<<<e->num_sms * 8, 256, 0, s>>>(e->nobs, e->Wp, e->obs_cam, e->obs_pt, e->bitpre);
  BA_LAUNCH_CHECK();
  bitmap_prefix_kernel<<<e->M, 1024, 0, s>>>(e->Wp, e->bitpre);
  BA_LAUNCH_CHECK();
  invert_perm_kernel<<<e->num_sms * 8, 256, 0, s>>>(e->nobs, e->cm_perm, e->cm_pos);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// ---- off-diagonal pairs -------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Pair number t -> (i, k), k < i.  Pairs are ordered tile by tile (kPairTile x kPairTile cameras),
// so that the warps resident at any time share few camera slices and walk them in step (L2 reuse).
__device__ __forceinline__ bool pair_from_linear(int64_t t, int M, int& i, int& k) {
  const int64_t per_tile = (int64_t)kPairTile * kPairTile;
  const int64_t tile = t / per_tile;
  const int r = (int)(t - tile * per_tile);
  int ti = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
  while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tile) ++ti;
  while ((int64_t)ti * (ti + 1) / 2 > tile) --ti;
  const int tk = (int)(tile - (int64_t)ti * (ti + 1) / 2);
  i = ti * kPairTile + r / kPairTile;
  k = tk * kPairTile + r % kPairTile;
  return i < M && k < i;
}

struct PairSmem {
  double stage[2][64 * kYS];  // [stage][block: 0..31 camera i side, 32..63 camera k side][kYS]
  uint2 queue[64];            // (position in slice i, position in slice k); [0, 32) = the next round
};

__global__ void __launch_bounds__(32, 8)
schur_pairs_kernel(int M, int64_t Wp, const uint2* __restrict__ bitpre,
                   const int64_t* __restrict__ cam_ptr, const double* __restrict__ Ycm,
                   double* __restrict__ P, int ld, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  int i, k;
  if (!pair_from_linear(blockIdx.x, M, i, k)) return;
  __shared__ __align__(16) PairSmem sm;
  const int lane = threadIdx.x;

  // two bitmap words (64 points) per lane and batch
  const uint4* bi = reinterpret_cast<const uint4*>(bitpre + (size_t)i * Wp) + lane;
  const uint4* bk = reinterpret_cast<const uint4*>(bitpre + (size_t)k * Wp) + lane;
  // gather mapping: lanes 0..11 copy the 12 pieces of one block, lanes 12..23 of the next one;
  // the piece offset is folded into the base pointers
  const int g_sub = lane >= 12 ? 1 : 0;
  const int g_piece = lane - 12 * g_sub;
  const bool g_on = lane < 24;
  const double* Yi = Ycm + (size_t)cam_ptr[i] * kYB + 2 * g_piece;
  const double* Yk = Ycm + (size_t)cam_ptr[k] * kYB + 2 * g_piece;
  const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(sm.queue) + 8u * g_sub;
  const uint32_t s_addr = (uint32_t)__cvta_generic_to_shared(sm.stage[0]) + (g_sub * kYS + 2 * g_piece) * 8u;

  double acc[9][9];
#pragma unroll
  for (int a = 0; a < 9; ++a)
#pragma unroll
    for (int b = 0; b < 9; ++b) acc[a][b] = 0.0;

  int qn = 0;           // entries queued
  int rounds = 0;       // rounds whose gather has been issued
  int cnt_prev = 0;     // valid lanes of the round waiting in stage (rounds - 1) & 1

  auto compute = [&](int st, int cnt) {
    if (lane < cnt) {
      const double* pi = sm.stage[st] + lane * kYS;
      const double* pk = sm.stage[st] + (32 + lane) * kYS;
      const double2 ti0 = *reinterpret_cast<const double2*>(pi + 18);  // Ti row 0: [0], [1]
      const double2 ti1 = *reinterpret_cast<const double2*>(pi + 20);  // Ti[0][2], Ti[1][0]
      const double2 ti2 = *reinterpret_cast<const double2*>(pi + 22);  // Ti[1][1], Ti[1][2]
      const double2 tk0 = *reinterpret_cast<const double2*>(pk + 18);
      const double2 tk1 = *reinterpret_cast<const double2*>(pk + 20);
      const double2 tk2 = *reinterpret_cast<const double2*>(pk + 22);
      // G = Ti Tk^T (2x2)
      const double g00 = ti0.x * tk0.x + ti0.y * tk0.y + ti1.x * tk1.x;
      const double g01 = ti0.x * tk1.y + ti0.y * tk2.x + ti1.x * tk2.y;
      const double g10 = ti1.y * tk0.x + ti2.x * tk0.y + ti2.y * tk1.x;
      const double g11 = ti1.y * tk1.y + ti2.x * tk2.x + ti2.y * tk2.y;
      // rows of G Jc_k (2x9)
      double t0[9], t1[9];
      {
        double jk[18];
#pragma unroll
        for (int u = 0; u < 9; ++u) {
          const double2 v = *reinterpret_cast<const double2*>(pk + 2 * u);
          jk[2 * u] = v.x;
          jk[2 * u + 1] = v.y;
        }
#pragma unroll
        for (int b = 0; b < 9; ++b) {
          t0[b] = g00 * jk[b] + g01 * jk[9 + b];
          t1[b] = g10 * jk[b] + g11 * jk[9 + b];
        }
      }
      double ji[18];
#pragma unroll
      for (int u = 0; u < 9; ++u) {
        const double2 v = *reinterpret_cast<const double2*>(pi + 2 * u);
        ji[2 * u] = v.x;
        ji[2 * u + 1] = v.y;
      }
#pragma unroll
      for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int b = 0; b < 9; ++b) acc[a][b] = fma(ji[a], t0[b], fma(ji[9 + a], t1[b], acc[a][b]));
    }
  };

  // Issue the gather of the next round (queue[0, cnt)), move the rest of the queue down, then
  // compute the previous round.
  auto round = [&](int cnt) {
    const uint32_t sa = s_addr + (uint32_t)(rounds & 1) * (64 * kYS * 8);
    if (cnt == 32) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        uint2 en;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(en.x), "=r"(en.y) : "r"(q_addr + 16u * u));
        if (g_on) {
          cp_async16(sa + 2 * u * kYS * 8, Yi + (size_t)en.x * kYB);
          cp_async16(sa + (32 + 2 * u) * kYS * 8, Yk + (size_t)en.y * kYB);
        }
      }
    } else {
      for (int u = 0; 2 * u < cnt; ++u) {
        const uint2 en = sm.queue[2 * u + g_sub];
        if (g_on && 2 * u + g_sub < cnt) {
          cp_async16(sa + 2 * u * kYS * 8, Yi + (size_t)en.x * kYB);
          cp_async16(sa + (32 + 2 * u) * kYS * 8, Yk + (size_t)en.y * kYB);
        }
      }
    }
    cp_commit();
    qn -= cnt;
    {
      const uint2 up = sm.queue[32 + lane];
      __syncwarp();
      sm.queue[lane] = up;
    }
    if (rounds > 0) {
      cp_wait<1>();
      __syncwarp();
      compute((rounds & 1) ^ 1, cnt_prev);
    }
    __syncwarp();
    cnt_prev = cnt;
    ++rounds;
  };

  // bitmap words are fetched one batch ahead
  uint4 wi = bi[0], wk = bk[0];
  for (int64_t w0 = 0; w0 < Wp; w0 += 64) {
    const uint4 ci = wi, ck = wk;
    if (w0 + 64 < Wp) {
      wi = bi[(w0 + 64) >> 1];
      wk = bk[(w0 + 64) >> 1];
    }
    const uint64_t mi = ((uint64_t)ci.z << 32) | ci.x, mk = ((uint64_t)ck.z << 32) | ck.x;
    uint64_t c = mi & mk;
    while (__any_sync(0xffffffffu, c != 0ull)) {
      // exclusive prefix of the hit counts -> queue slots; hits that do not fit wait for the next pass
      const int n = __popcll(c);
      int incl = n;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int pos = qn + incl - n;
      while (c != 0ull && pos < 64) {
        const int b = __ffsll((long long)c) - 1;
        c &= c - 1;
        const uint64_t below = (1ull << b) - 1ull;
        sm.queue[pos++] = make_uint2(ci.y + __popcll(mi & below), ck.y + __popcll(mk & below));
      }
      qn = qn + total < 64 ? qn + total : 64;
      __syncwarp();
      while (qn >= 32) round(32);
    }
  }
  if (qn > 0) round(qn);
  if (rounds > 0) {
    cp_wait<0>();
    __syncwarp();
    compute((rounds - 1) & 1, cnt_prev);
  }

  // fixed-order butterfly over the 32 lanes, then lane e % 32 stores entry e
#pragma unroll
  for (int a = 0; a < 9; ++a)
#pragma unroll
    for (int b = 0; b < 9; ++b) {
      double v = acc[a][b];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == ((a * 9 + b) & 31)) P[(size_t)(9 * i + a) * ld + 9 * k + b] = v;
    }
}

// ---- diagonal blocks and rhs --------------------------------------------------------------------
// grid (chunks, M): block (c, i) reduces chunk c of camera i's slice into 54 numbers.
__global__ void __launch_bounds__(128)
schur_diag_kernel(const int64_t* __restrict__ cam_ptr, const int32_t* __restrict__ cm_perm,
                  const int32_t* __restrict__ obs_pt, const double* __restrict__ Ycm,
                  const double* __restrict__ Z, double* __restrict__ Dpart, const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int i = blockIdx.y;
  const int chunk = blockIdx.x, nchunks = gridDim.x;
  const int64_t seg_lo = cam_ptr[i], seg_n = cam_ptr[i + 1] - seg_lo;
  const int64_t per = (seg_n + nchunks - 1) / nchunks;
  const int64_t lo = per * chunk;
  const int64_t hi = lo + per < seg_n ? lo + per : seg_n;
  double acc[kDiagPart];
#pragma unroll
  for (int q = 0; q < kDiagPart; ++q) acc[q] = 0.0;
  for (int64_t q = lo + threadIdx.x; q < hi; q += blockDim.x) {
    const double2* row = reinterpret_cast<const double2*>(Ycm + (size_t)(seg_lo + q) * kYB);
    double v[kYB];
#pragma unroll
    for (int u = 0; u < kYB / 2; ++u) {
      const double2 t2 = row[u];
      v[2 * u] = t2.x;
      v[2 * u + 1] = t2.y;
    }
    const double* z = Z + 3 * (size_t)obs_pt[cm_perm[seg_lo + q]];
    const double z0 = z[0], z1 = z[1], z2 = z[2];
    const double* ja = v;
    const double* jb = v + 9;
    const double* ta = v + 18;
    const double* tb = v + 21;
    const double gaa = ta[0] * ta[0] + ta[1] * ta[1] + ta[2] * ta[2];
    const double gab = ta[0] * tb[0] + ta[1] * tb[1] + ta[2] * tb[2];
    const double gbb = tb[0] * tb[0] + tb[1] * tb[1] + tb[2] * tb[2];
    const double za = ta[0] * z0 + ta[1] * z1 + ta[2] * z2;
    const double zb = tb[0] * z0 + tb[1] * z1 + tb[2] * z2;
    int idx = 0;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const double t0 = gaa * ja[r] + gab * jb[r];
      const double t1 = gab * ja[r] + gbb * jb[r];
#pragma unroll
      for (int c = r; c < 9; ++c) acc[idx++] += t0 * ja[c] + t1 * jb[c];
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) acc[45 + r] += ja[r] * za + jb[r] * zb;
  }
  __shared__ double sred[4][kDiagPart];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < kDiagPart; ++q) {
    const double s = warp_sum(acc[q]);
    if (lane == 0) sred[warp][q] = s;
  }
  __syncthreads();
  if (threadIdx.x < kDiagPart) {
    const double s = ((sred[0][threadIdx.x] + sred[1][threadIdx.x]) + sred[2][threadIdx.x]) +
                     sred[3][threadIdx.x];
    Dpart[((size_t)i * nchunks + chunk) * kDiagPart + threadIdx.x] = s;
  }
}

__global__ void schur_diag_finish_kernel(int nchunks, const double* __restrict__ Dpart,
                                         double* __restrict__ P, int ld, int rhs_row,
                                         const ba_lm_state* ctl) {
  if (ctl && ctl->done) return;
  const int i = blockIdx.x;
  const int q = threadIdx.x;  // 0..89: 81 entries of the block, 9 of the rhs
  if (q >= 90) return;
  int src, r = 0, c = 0;
  if (q < 81) {
    r = q / 9;
    c = q % 9;
    const int lo = r < c ? r : c, hi = r < c ? c : r;
    src = lo * 9 - lo * (lo - 1) / 2 + (hi - lo);
  } else {
    src = 45 + (q - 81);
  }
  double s = 0.0;
  for (int ch = 0; ch < nchunks; ++ch) s += Dpart[((size_t)i * nchunks + ch) * kDiagPart + src];
  if (q < 81) {
    if (c <= r) P[(size_t)(9 * i + r) * ld + 9 * i + c] = s;
  } else {
    P[(size_t)rhs_row * ld + 9 * i + (q - 81)] = s;
  }
}

int launch_schur_sparse(ba_engine* e, const ba_lm_state* ctl, cudaStream_t s) {
  ProfScope ps(e, PG_SYRK, s);
  dim3 dgrid(e->cam_chunks, e->M);
  schur_diag_kernel<<<dgrid, 128, 0, s>>>(e->cam_ptr, e->cm_perm, e->obs_pt, e->Ycm, e->Z, e->Upart, ctl);
  BA_LAUNCH_CHECK();
  schur_diag_finish_kernel<<<e->M, 96, 0, s>>>(e->cam_chunks, e->Upart, e->P(), e->n_pad, e->rhs_row, ctl);
  BA_LAUNCH_CHECK();
  const int nt = (e->M + kPairTile - 1) / kPairTile;
  const int64_t n_items = (int64_t)nt * (nt + 1) / 2 * kPairTile * kPairTile;
  if (n_items >= ((int64_t)1 << 31)) {
    set_error("too many camera pairs for one launch (M=%d)", e->M);
    return BA_ERR_INVALID;
  }
  BA_CUDA(cudaFuncSetAttribute(schur_pairs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                               (int)cudaSharedmemCarveoutMaxShared));
  schur_pairs_kernel<<<(unsigned)n_items, 32, 0, s>>>(e->M, e->Wp, e->bitpre, e->cam_ptr, e->Ycm, e->P(),
                                                     e->n_pad, ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

}  // namespace ba
