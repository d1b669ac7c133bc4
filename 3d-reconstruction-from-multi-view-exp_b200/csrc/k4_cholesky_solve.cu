// K4: reduced camera system -- assembly, blocked Cholesky factor/solve, camera update, point
// back-substitution with the trial cost, and the Levenberg-Marquardt decision, all on device.
//
// Replaces reference lib/bundle_adjustment.py:
//   :123-125, :135, :143  A = G_c - sum F^T E^-1 F,  b = sum F^T E^-1 d_P - d_F   (assemble)
//   :146                  np.linalg.solve(A, b) (LAPACK dgesv/LU) -> Cholesky; A is SPD
//   :152                  dX_j = -E_j^-1 (F_j dxi + d_P_j)                          (back-subst.)
//   :260-281 + lib/utils.py:10-29  parameter update incl. Rodrigues                (camera update)
//   :159-162              trial cost
//   :164-195              accept / reject, damping schedule, termination           (decide)
//
// The 7 gauge unknowns (:62-72) are pinned (unit diagonal, zero row/column/rhs) rather than
// deleted, so the system keeps the regular 9-per-camera layout.  The right-hand side travels
// as one extra row of the matrix (row `rhs_row`), so the forward substitution L y = b is done
// by the factorisation itself; only L^T x = y remains.
#include "ba_common.cuh"

namespace ba {

// ---- assemble ---------------------------------------------------------------------------------
// In place on the reduce buffer: P (sum Y Y^T | rhs row) -> S (lower triangle + rhs row).
__global__ void __launch_bounds__(256)
assemble_kernel(double* __restrict__ S, int ld, int n_full, int rhs_row, int axis,
                const double* __restrict__ U, const double* __restrict__ GCAM, double c_host,
                ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  const double damp = 1.0 + (use_ctl ? ctl->c : c_host);
  const int r = blockIdx.x;
  double* row = S + (size_t)r * ld;
  if (r < n_full) {
    const int i = r / 9, a = r - 9 * i;
    const bool pin_r = (gauge_mask(i, axis) >> a) & 1u;
    for (int cc = threadIdx.x; cc <= r; cc += blockDim.x) {
      const int k = cc / 9, b = cc - 9 * k;
      double v = -row[cc];
      if (k == i) {
        const double u = U[(size_t)i * 81 + a * 9 + b];
        v = (a == b ? u * damp : u) + v;  // G_c - sum(...)  (:123-125, :135)
      }
      if (pin_r && cc == r) v = 1.0;
      row[cc] = v;
    }
  } else if (r == rhs_row) {
    for (int cc = threadIdx.x; cc <= r; cc += blockDim.x)
      row[cc] = cc < n_full ? row[cc] - GCAM[cc] : 1.0;  // b (:143)
  } else {
    for (int cc = threadIdx.x; cc <= r; cc += blockDim.x) row[cc] = cc == r ? 1.0 : 0.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->chol_fail = 0;
}

int launch_assemble(ba_engine* e, bool conditional, double c_host, cudaStream_t s) {
  assemble_kernel<<<e->n_pad, 256, 0, s>>>(e->P(), e->n_pad, e->n_full, e->rhs_row, e->axis, e->U(),
                                           e->GCAM(), c_host, e->ctl, conditional ? 1 : 0);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// ---- camera update (:263-281, lib/utils.py:10-29) -----------------------------------------------
__global__ void camera_update_kernel(int M, double f0, const double* __restrict__ dxi,
                                     const double* __restrict__ f, const double* __restrict__ u,
                                     const double* __restrict__ R, const double* __restrict__ t,
                                     double* __restrict__ f2, double* __restrict__ u2,
                                     double* __restrict__ R2, double* __restrict__ t2,
                                     double* __restrict__ tab2, const ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const double* d = dxi + 9 * (size_t)i;
  const double fi = f[i] + d[0];
  const double u0 = u[2 * i] + d[1], v0 = u[2 * i + 1] + d[2];
  const double t0 = t[3 * i] + d[3], t1 = t[3 * i + 1] + d[4], t2v = t[3 * i + 2] + d[5];
  const double w0 = d[6], w1 = d[7], w2 = d[8];
  double dR[9];
  if (w0 == 0.0 && w1 == 0.0 && w2 == 0.0) {  // exact identity (lib/utils.py:14-15)
    dR[0] = 1; dR[1] = 0; dR[2] = 0; dR[3] = 0; dR[4] = 1; dR[5] = 0; dR[6] = 0; dR[7] = 0; dR[8] = 1;
  } else {
    const double th = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double l0 = w0 / th, l1 = w1 / th, l2 = w2 / th;
    const double c = cos(th), sn = sin(th), oc = 1.0 - c;
    dR[0] = oc * (l0 * l0) + c;       dR[1] = oc * (l0 * l1) - sn * l2;  dR[2] = oc * (l0 * l2) + sn * l1;
    dR[3] = oc * (l1 * l0) + sn * l2; dR[4] = oc * (l1 * l1) + c;        dR[5] = oc * (l1 * l2) - sn * l0;
    dR[6] = oc * (l2 * l0) - sn * l1; dR[7] = oc * (l2 * l1) + sn * l0;  dR[8] = oc * (l2 * l2) + c;
  }
  const double* Ri = R + 9 * (size_t)i;
  double Rn[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
      Rn[3 * a + b] = dR[3 * a] * Ri[b] + dR[3 * a + 1] * Ri[3 + b] + dR[3 * a + 2] * Ri[6 + b];
  f2[i] = fi;
  u2[2 * i] = u0; u2[2 * i + 1] = v0;
  t2[3 * i] = t0; t2[3 * i + 1] = t1; t2[3 * i + 2] = t2v;
#pragma unroll
  for (int k = 0; k < 9; ++k) R2[9 * (size_t)i + k] = Rn[k];
  double* T = tab2 + (size_t)i * kCamTab;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T[k] = fi * Rn[3 * k] + u0 * Rn[3 * k + 2];
    T[3 + k] = fi * Rn[3 * k + 1] + v0 * Rn[3 * k + 2];
    T[6 + k] = f0 * Rn[3 * k + 2];
  }
  T[9] = t0; T[10] = t1; T[11] = t2v;
  T[12] = 1.0 / fi; T[13] = u0 / f0; T[14] = v0 / f0; T[15] = 1.0 / f0;  // as cam_prep_kernel
}

// ---- point back-substitution + trial cost (:152, :155, :159-162) -------------------------------
// MF (dense scenes, see dense_matrix_free): Y is not read back (216 B per observation); its product
// with the camera step is re-derived from the camera table of the linearisation state, X_j and L_j^-1
// (obs_jacobian, scaled_point_rows: the rows K2b formed Y from);
// both camera tables and the camera step sit in shared memory.
template <bool DENSE, bool MF>
__global__ void __launch_bounds__(256)
point_update_cost_kernel(int64_t N, int M, const int64_t* __restrict__ obs_ptr,
                         const int32_t* __restrict__ obs_cam, const double2* __restrict__ xy,
                         const double* __restrict__ Yt, int ld, const double* __restrict__ Ysp,
                         const double* __restrict__ Z, const double* __restrict__ LINV,
                         const double* __restrict__ dxi, const double* __restrict__ X,
                         double* __restrict__ X2, const double* __restrict__ tab2, double f0,
                         double* __restrict__ cost_part, const ba_lm_state* ctl, int use_ctl,
                         const double* __restrict__ tab0, int axis) {
  static_assert(DENSE || !MF, "the matrix-free variant is the dense one");
  if (use_ctl && ctl->done) return;
  __shared__ double scratch[32];
  extern __shared__ double pu_smem[];  // MF: [table of the linearisation state | trial table | dxi]
  double* s_tab0 = pu_smem;
  double* s_tab2 = pu_smem + (MF ? tab_smem_doubles(M) : 0);
  double* s_dxi = s_tab2 + (MF ? tab_smem_doubles(M) : 0);
  if (MF) {
    for (int k = threadIdx.x; k < M * kCamTab; k += blockDim.x) {
      s_tab0[(k >> 4) * kTabStride + (k & 15)] = tab0[k];
      s_tab2[(k >> 4) * kTabStride + (k & 15)] = tab2[k];
    }
    for (int k = threadIdx.x; k < 9 * M; k += blockDim.x) s_dxi[k] = dxi[k];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double cost = 0.0;
  for (int64_t j = warp; j < N; j += nwarps) {
    const int64_t lo = DENSE ? j * M : obs_ptr[j];
    const int64_t hi = DENSE ? lo + M : obs_ptr[j + 1];
    const double* m = LINV + 6 * (size_t)j;
    double s0 = 0, s1 = 0, s2 = 0;
    if (MF) {
      const double xj0 = X[3 * j], xj1 = X[3 * j + 1], xj2 = X[3 * j + 2];
      const double m00 = m[0], m10 = m[1], m11 = m[2], m20 = m[3], m21 = m[4], m22 = m[5];
      for (int64_t o = lo + lane; o < hi; o += 32) {
        const int i = (int)(o - lo);
        const uint32_t mask = gauge_mask(i, axis);
        ObsJacobian J;
        obs_jacobian(s_tab0 + (size_t)i * kTabStride, xj0, xj1, xj2, J);
        double ta[3], tb[3];
        scaled_point_rows(J.ax, J.bx, m00, m10, m11, m20, m21, m22, ta, tb);
        // sum_a Y[a][d] dxi[a] = ta[d] (Jc row a . dxi) + tb[d] (Jc row b . dxi): 24 FMAs instead of
        // forming the 27 entries of Y (the step agrees with the stored-Y path to rounding, not bitwise)
        const double* d = s_dxi + 9 * i;
        double ua = 0.0, ub = 0.0;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const bool pin = (mask >> a) & 1u;
          const double da = pin ? 0.0 : d[a];
          ua = fma(J.ja[a], da, ua);
          ub = fma(J.jb[a], da, ub);
        }
        s0 = fma(ta[0], ua, fma(tb[0], ub, s0));
        s1 = fma(ta[1], ua, fma(tb[1], ub, s1));
        s2 = fma(ta[2], ua, fma(tb[2], ub, s2));
      }
    } else {
    for (int64_t o = lo + lane; o < hi; o += 32) {
      const int i = DENSE ? (int)(o - lo) : obs_cam[o];
      const double* d = dxi + 9 * (size_t)i;
      if (DENSE) {
        const double* y0 = Yt + (size_t)(3 * j) * ld + 9 * (size_t)i;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const double da = d[a];
          s0 = fma(y0[a], da, s0);
          s1 = fma(y0[ld + a], da, s1);
          s2 = fma(y0[2 * (size_t)ld + a], da, s2);
        }
      } else {
        const double* y = Ysp + (size_t)o * 27;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const double da = d[a];
          s0 = fma(y[a], da, s0);
          s1 = fma(y[9 + a], da, s1);
          s2 = fma(y[18 + a], da, s2);
        }
      }
    }
    }
    s0 = warp_sum(s0) + Z[3 * j];
    s1 = warp_sum(s1) + Z[3 * j + 1];
    s2 = warp_sum(s2) + Z[3 * j + 2];
    // dX = -L^-T v with m = L^-1 (lower): (L^-T v)[b] = sum_{d >= b} m[d][b] v[d]
    const double x0 = X[3 * j] - (m[0] * s0 + m[1] * s1 + m[3] * s2);
    const double x1 = X[3 * j + 1] - (m[2] * s1 + m[4] * s2);
    const double x2 = X[3 * j + 2] - (m[5] * s2);
    if (lane == 0) {
      X2[3 * j] = x0; X2[3 * j + 1] = x1; X2[3 * j + 2] = x2;
    }
    for (int64_t o = lo + lane; o < hi; o += 32) {
      const int i = DENSE ? (int)(o - lo) : obs_cam[o];
      const double* T = MF ? s_tab2 + (size_t)i * kTabStride : tab2 + (size_t)i * kCamTab;
      const double2 mm = xy[o];
      cost += obs_cost(T, x0, x1, x2, mm.x, mm.y);
    }
  }
  const double tot = block_sum(cost, scratch);
  if (threadIdx.x == 0) cost_part[blockIdx.x] = tot;
}

__global__ void cost_finish2_kernel(const double* __restrict__ part, int n, double* out,
                                    const ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) acc += part[k];
  const double tot = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    out[0] = tot;
    // travels with the trial cost through whatever sums it over the ranks (see ba_cost_buffer)
    out[1] = (use_ctl && ctl->status == BA_ERR_SINGULAR) ? 1.0 : 0.0;
  }
}

int launch_update_trial(ba_engine* e, bool conditional, cudaStream_t s) {
  const int use_ctl = conditional ? 1 : 0;
  const CamState &c0 = e->cam[0], &c1 = e->cam[1];
  camera_update_kernel<<<(e->M + 127) / 128, 128, 0, s>>>(e->M, e->f0, e->dxi, c0.f, c0.u, c0.R, c0.t,
                                                           c1.f, c1.u, c1.R, c1.t, e->camtab[1],
                                                           e->ctl, use_ctl);
  BA_LAUNCH_CHECK();
  const int grid = balanced_blocks((e->N + 7) / 8, (int64_t)e->num_sms * 8);  // 8 warps = 8 points per block
  const double2* xy = reinterpret_cast<const double2*>(e->obs_xy);
  const bool mf = dense_matrix_free(e);
  const size_t smem = mf ? (2 * tab_smem_doubles(e->M) + 9 * (size_t)e->M) * sizeof(double) : 0;
#define BA_PU_LAUNCH(D, F)                                                                                       \
  do {                                                                                                           \
    if (smem > 48 * 1024)                                                                                        \
      BA_CUDA(cudaFuncSetAttribute(point_update_cost_kernel<D, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                   (int)smem));                                                                  \
    point_update_cost_kernel<D, F><<<grid, 256, smem, s>>>(                                                      \
        e->N, e->M, e->obs_ptr, e->obs_cam, xy, e->Yt, e->n_pad, e->Ysp, e->Z, e->LINV, e->dxi, e->X[0], e->X[1], \
        e->camtab[1], e->f0, e->cost_part, e->ctl, use_ctl, e->camtab[0], e->axis);                              \
  } while (0)
  if (mf) BA_PU_LAUNCH(true, true);
  else if (e->dense) BA_PU_LAUNCH(true, false);
  else BA_PU_LAUNCH(false, false);
#undef BA_PU_LAUNCH
  BA_LAUNCH_CHECK();
  cost_finish2_kernel<<<1, 256, 0, s>>>(e->cost_part, grid, e->cost_buf + 1, e->ctl, use_ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// ---- LM control (:100-101, :164-195) -----------------------------------------------------------
__global__ void lm_begin_kernel(ba_lm_state* ctl, double scale, double tol, int max_iter,
                                int max_retries) {
  ctl->E = 0.0;
  ctl->E_trial = 0.0;
  ctl->c = 0.0001;  // :100
  ctl->delta = 0.0;
  ctl->scale_factor = scale;
  ctl->delta_tol = tol;
  ctl->count = 0;
  ctl->max_iter = max_iter;
  ctl->solves = 0;
  ctl->iter_solves = 0;
  ctl->need_linearize = 1;
  ctl->accepted = 0;
  ctl->done = 0;
  ctl->status = BA_OK;
  ctl->chol_fail = 0;
  ctl->max_retries = max_retries;
}

int launch_lm_begin(ba_engine* e, double scale, double tol, int max_iter, int max_retries,
                    cudaStream_t s) {
  lm_begin_kernel<<<1, 1, 0, s>>>(e->ctl, scale, tol, max_iter, max_retries);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// Adopt the (all-reduced) initial cost as E before the first solve (:85-87).
__global__ void lm_adopt_kernel(ba_lm_state* ctl, const double* cost_buf) {
  if (ctl->done) return;
  if (ctl->solves == 0) ctl->E = cost_buf[0];
}

int launch_lm_adopt(ba_engine* e, cudaStream_t s) {
  lm_adopt_kernel<<<1, 1, 0, s>>>(e->ctl, e->cost_buf);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

__global__ void lm_decide_kernel(ba_lm_state* ctl, const double* cost_buf, ba_iter_record* rec) {
  if (ctl->done) return;
  const double E_ = cost_buf[1];
  ctl->E_trial = E_;
  ctl->solves += 1;
  ctl->iter_solves += 1;
  // a singular point block on ANY rank (the flag is summed over the ranks with the cost)
  if (cost_buf[2] > 0.0 && ctl->status == BA_OK) ctl->status = BA_ERR_SINGULAR;
  if (ctl->status != BA_OK) {  // singular block / failed barrier: stop, the host raises
    ctl->done = 1;
    ctl->accepted = 0;
    return;
  }
  if (ctl->chol_fail || E_ > ctl->E) {  // reject (:164-165); a non-SPD system counts as a reject
    ctl->c *= ctl->scale_factor;
    ctl->accepted = 0;
    ctl->need_linearize = 0;
    if (ctl->iter_solves >= ctl->max_retries) {
      ctl->status = BA_ERR_STALL;
      ctl->done = 1;
    }
    return;
  }
  // accept (:166-173): the trial becomes the state, also on the terminating iteration
  ctl->accepted = 1;
  ctl->need_linearize = 1;
  ctl->count += 1;
  const double delta = fabs(E_ - ctl->E);  // :186
  ctl->delta = delta;
  if (ctl->count <= kMaxRecords) {
    ba_iter_record& r = rec[ctl->count - 1];
    r.E_prev = ctl->E;
    r.E = E_;
    r.delta = delta;
    r.c = ctl->c;
    r.solves = ctl->iter_solves;
    r.count = ctl->count;
  }
  ctl->iter_solves = 0;
  if (delta <= ctl->delta_tol || ctl->count >= ctl->max_iter) {  // :191
    ctl->done = 1;
    ctl->E = E_;  // the reference leaves E stale here; reported cost is E_trial either way
  } else {
    ctl->E = E_;                   // :194
    ctl->c /= ctl->scale_factor;   // :195
  }
}

// Commit the accepted trial state (:169-173).  Runs right after lm_decide_kernel; `done` may
// already be set by an accepting, terminating decide, so it keys on `accepted` only.  After
// termination every producer kernel is a no-op, so a repeated copy is idempotent.
__global__ void lm_commit_kernel(int64_t n_x, int n_cam_doubles, const double* __restrict__ X2,
                                 double* __restrict__ X, const double* __restrict__ cam2,
                                 double* __restrict__ cam, const double* __restrict__ tab2,
                                 double* __restrict__ tab, int n_tab, ba_lm_state* ctl) {
  if (!ctl->accepted) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t k = g; k < n_x; k += stride) X[k] = X2[k];
  for (int64_t k = g; k < n_cam_doubles; k += stride) cam[k] = cam2[k];
  for (int64_t k = g; k < n_tab; k += stride) tab[k] = tab2[k];
}

int launch_decide(ba_engine* e, cudaStream_t s) {
  lm_decide_kernel<<<1, 1, 0, s>>>(e->ctl, e->cost_buf, e->rec);
  BA_LAUNCH_CHECK();
  const int64_t nx = 3 * e->N;
  int64_t blocks = (nx + 255) / 256;
  const int64_t cap = (int64_t)e->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // camera state of one slot is a single allocation [f | u | R | t] = 15 M doubles
  lm_commit_kernel<<<(int)blocks, 256, 0, s>>>(nx, 15 * e->M, e->X[1], e->X[0], e->cam[1].f,
                                               e->cam[0].f, e->camtab[1], e->camtab[0],
                                               e->M * kCamTab, e->ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

}  // namespace ba
