// chol.cuh -- diagonal-block kernel of the blocked FP64 Cholesky / triangular inversion.
// Replaces MATLAB chol(.,'lower') at solvers/lasso.m:168,172, huberfit.m:166, lad.m:134 and the
// per-iteration refactorisation of unwrappedadmm.m:139 (W\d) with one cached factor.
//
// Blocked right-looking algorithm (driver in engine.cu):
//   for each NB-wide panel:  [this kernel] L11 = chol(A11), X11 = inv(L11)
//                            [DMMA GEMM]   L21 = A21 * X11'
//                            [DMMA GEMM]   A22 -= L21 * L21'   (lower tiles only)
// and the inverse factor W = inv(L) is assembled from the X11 blocks by recursive doubling
// (W21 = -W22 * L21 * W11, two DMMA GEMMs per level).
#pragma once
#include "common.cuh"

namespace admmb200 {

constexpr int CHOL_NB = 128;
constexpr int CHOL_LDS = CHOL_NB + 1;
constexpr int CHOL_DIAG_THREADS = 512;
constexpr int CHOL_DIAG_SMEM = (CHOL_NB * CHOL_LDS + CHOL_NB) * 8;  // block (factored, then inverted in place) + one column

// One CTA: in-place lower Cholesky of the nb x nb block A (nb <= 128), strict upper part of the
// block zeroed, X = inv(L) written as a full nb x nb block (upper part zero).
// *fail is set to (1 + global index of the offending pivot) when a pivot is not positive.
__global__ void __launch_bounds__(CHOL_DIAG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int nb, double* __restrict__ X, int64_t ldx,
                  int* fail, int pivot_base) {
  extern __shared__ __align__(16) double sm[];
  double* L = sm;                          // L[r + c*CHOL_LDS]
  double* lk = sm + CHOL_NB * CHOL_LDS;    // stash of one column during the in-place inversion
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int idx = tid; idx < nb * nb; idx += nt) {
    int r = idx % nb, c = idx / nb;
    L[r + c * CHOL_LDS] = (r >= c) ? A[r + c * lda] : 0.0;
  }
  __syncthreads();

  // right-looking Cholesky: three barriers per column
  for (int j = 0; j < nb; ++j) {
    double d = L[j + j * CHOL_LDS];
    if (!(d > 0.0)) {
      if (tid == 0) atomicCAS(fail, 0, pivot_base + j + 1);
    }
    double inv = 1.0 / sqrt(d);
    __syncthreads();  // everyone has read d before the column is scaled
    for (int r = j + tid; r < nb; r += nt) L[r + j * CHOL_LDS] *= inv;  // L[j][j] becomes sqrt(d)
    __syncthreads();
    // trailing update of the lower triangle: A[r][c] -= L[r][j] * L[c][j],  j < c <= r < nb
    int rem = nb - j - 1;
    for (int idx = tid; idx < rem * rem; idx += nt) {
      int rr = idx % rem, cc = idx / rem;
      if (rr >= cc) {
        int r = j + 1 + rr, c = j + 1 + cc;
        L[r + c * CHOL_LDS] -= L[r + j * CHOL_LDS] * L[c + j * CHOL_LDS];
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nb * nb; idx += nt) {
    int r = idx % nb, c = idx / nb;
    A[r + c * lda] = L[r + c * CHOL_LDS];
  }
  if (!X) return;
  __syncthreads();

  // in-place inverse, all columns at once (forward substitution on L * V = I):
  //   step k:  V[k][c] /= L[k][k] (c <= k);   V[r][c] -= L[r][k] * V[k][c]  (r > k, c <= k)
  // V[r][c] (r > k >= c) shares storage with L[r][c], which is dead once step c has run; column
  // k of L is stashed in lk before it is overwritten by V[r][k] = -L[r][k] * V[k][k].
  for (int k = 0; k < nb; ++k) {
    const double dk = 1.0 / L[k + k * CHOL_LDS];
    __syncthreads();  // dk read by everyone before row k is rewritten
    for (int c = tid; c < k; c += nt) L[k + c * CHOL_LDS] *= dk;
    if (tid == 0) L[k + k * CHOL_LDS] = dk;
    for (int r = k + 1 + tid; r < nb; r += nt) lk[r] = L[r + k * CHOL_LDS];
    __syncthreads();
    const int rem = nb - k - 1, cols = k + 1;
    for (int idx = tid; idx < rem * cols; idx += nt) {
      int rr = idx % rem, c = idx / rem;
      int r = k + 1 + rr;
      double cur = (c == k) ? 0.0 : L[r + c * CHOL_LDS];
      L[r + c * CHOL_LDS] = cur - lk[r] * L[k + c * CHOL_LDS];
    }
    // next step reads L[k+1][k+1]: untouched by this step (only columns <= k are written)
  }
  __syncthreads();
  for (int idx = tid; idx < nb * nb; idx += nt) {
    int r = idx % nb, c = idx / nb;
    X[r + c * ldx] = L[r + c * CHOL_LDS];
  }
}

}  // namespace admmb200
