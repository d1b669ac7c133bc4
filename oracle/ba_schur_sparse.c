/* Sparsity-aware CPU restatement of the Schur reduction of the reference's bundle adjustment.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/ba_oracle.py): built by oracle/build_c.py into
 * oracle/_ref/libba_oracle.so, called through ctypes by oracle/ba_oracle.py::reduced_system_sparse.
 * Nothing of the product path links or loads it.
 *
 * What it restates: reference lib/bundle_adjustment.py:132-143
 *     FtEinv = matF^T @ matEinv                 (:132)
 *     A      = matGc - (FtEinv @ matF).sum(0)   (:135)
 *     b      = (FtEinv @ d_P).sum(0) - d_F      (:138-143)
 * The reference forms these on dense (N, 3, n) / (N, n, n) arrays, i.e. it multiplies the zero
 * blocks of every camera that does not see point j.  Here point j only touches the (9 m_j) x
 * (9 m_j) sub-matrix of the cameras that see it (m_j views): per pair of its observations (a, b)
 *     A[9 cam_a : +9, 9 cam_b : +9] -= W_a^T Vinv_j W_b,          W_o = block of matF (3 x 9)
 * which is the count SURVEY.md section 8d asks a CPU baseline to be judged on:
 *     sum_j 3 (9 m_j)(9 m_j + 1) flops  (lower triangle)  instead of  2 * 3 * n^2 * N.
 * Only pairs with cam_a >= cam_b are computed (observations of a point are sorted by camera);
 * the upper triangle is mirrored at the end.
 *
 * Threads (OpenMP): camera rows are dealt to threads in contiguous ranges of equal triangular
 * work; every thread scans all points and updates only block rows it owns -- no atomics, no
 * private copies of A, and a summation order per entry that does not depend on the thread
 * count (points ascending), so the result is bit-reproducible.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int ba_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* A: [n][n] row-major, n = 9 * n_cams; on entry the block-diagonal damped U (matGc), on exit
 *    the full reduced matrix.  b: [n]; on entry -d_F, on exit the reduced right-hand side.
 * W: [nobs][3][9], Vinv: [n_points][3][3], g_pt: [n_points][3] (d_P),
 * ptr: [n_points + 1] CSR offsets, cam: [nobs] camera of each observation (ascending per point).
 * Returns the number of 9x9 pair blocks accumulated (for the flop count). */
int64_t ba_oracle_schur_sparse(int64_t n_points, int32_t n_cams, const int64_t* ptr,
                               const int64_t* cam, const double* W, const double* Vinv,
                               const double* g_pt, double* A, double* b) {
  const int64_t n = 9 * (int64_t)n_cams;
  int64_t pairs_total = 0;
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  /* row ranges of equal triangular work: rows [lo, hi) cost ~ hi^2 - lo^2 */
  int32_t* bounds = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nthreads + 1));
  for (int t = 0; t <= nthreads; ++t) {
    double frac = (double)t / (double)nthreads;
    double r = (double)n_cams * __builtin_sqrt(frac);
    bounds[t] = (int32_t)(r + 0.5);
  }
  bounds[0] = 0;
  bounds[nthreads] = n_cams;

#pragma omp parallel num_threads(nthreads) reduction(+ : pairs_total)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    const int32_t row_lo = bounds[tid], row_hi = bounds[tid + 1];
    double T[27]; /* W_a^T Vinv_j : 9 x 3 */
    for (int64_t j = 0; j < n_points; ++j) {
      const double* Vi = Vinv + 9 * j;
      const double* gj = g_pt + 3 * j;
      for (int64_t a = ptr[j]; a < ptr[j + 1]; ++a) {
        const int64_t ca = cam[a];
        if (ca < row_lo || ca >= row_hi) continue;
        const double* Wa = W + 27 * a; /* [3][9] */
        for (int p = 0; p < 9; ++p)
          for (int d = 0; d < 3; ++d)
            T[3 * p + d] = Wa[p] * Vi[d] + Wa[9 + p] * Vi[3 + d] + Wa[18 + p] * Vi[6 + d];
        /* rhs: b[9 ca + p] += sum_d T[p][d] g_j[d]   (:138-143) */
        double* bp = b + 9 * ca;
        for (int p = 0; p < 9; ++p) bp[p] += T[3 * p] * gj[0] + T[3 * p + 1] * gj[1] + T[3 * p + 2] * gj[2];
        for (int64_t o = ptr[j]; o <= a; ++o) { /* cam[o] <= ca */
          const double* Wb = W + 27 * o;
          double* blk = A + (9 * ca) * n + 9 * cam[o];
          for (int p = 0; p < 9; ++p) {
            const double t0 = T[3 * p], t1 = T[3 * p + 1], t2 = T[3 * p + 2];
            double* row = blk + p * n;
            for (int q = 0; q < 9; ++q) row[q] -= t0 * Wb[q] + t1 * Wb[9 + q] + t2 * Wb[18 + q];
          }
          ++pairs_total;
        }
      }
    }
  }
  free(bounds);
  /* mirror the strictly-lower camera blocks into the upper triangle; a diagonal block was
   * accumulated in full (o == a gives the whole symmetric 9x9) */
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t i = 0; i < n_cams; ++i)
    for (int64_t k = 0; k < i; ++k)
      for (int p = 0; p < 9; ++p)
        for (int q = 0; q < 9; ++q) A[(9 * k + q) * n + 9 * i + p] = A[(9 * i + p) * n + 9 * k + q];
  return pairs_total;
}
