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
constexpr int CHOL_NBO = 512;                // outer panel width (K of the big trailing update)
constexpr int CHOL_LDS = CHOL_NB + 4;        // row stride 132: (q + 4r) mod 16 distinct -> conflict-free LDS.64
constexpr int CHOL_DIAG_THREADS = 512;
constexpr int CHOL_DIAG_SMEM = (CHOL_NB * CHOL_LDS + 2 * CHOL_NB) * 8;

// One CTA: in-place lower Cholesky of the nb x nb block A (nb <= 128), strict upper part of the
// block zeroed, X = inv(L) written as a full nb x nb block (upper part zero).
// *fail is set to (1 + global index of the offending pivot) when a pivot is not positive.
//
// Left-looking, 4 threads per row (thread t: row t>>2, terms k = (t&3) mod 4): every step is a
// register-accumulated dot product of two shared-memory rows (no read-modify-write of shared
// memory, so the loads pipeline), a 2-step shuffle reduction and two barriers.  The inverse is
// built row by row with the same "row i . row c" kernel: V' is kept in the upper triangle of the
// same array (S(c,k) = V[k][c], k > c), its diagonal in dinv[].
#ifdef CHOL_PROFILE
__device__ long long chol_prof[8];
#define CHOL_T(i) if (threadIdx.x == 0) chol_prof[i] = clock64();
#else
#define CHOL_T(i)
#endif

__global__ void __launch_bounds__(CHOL_DIAG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int nb, double* __restrict__ X, int64_t ldx,
                  int* fail, int pivot_base) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                          // S(row, col) = S[col + row*CHOL_LDS]
  double* tcol = sm + CHOL_NB * CHOL_LDS;  // unscaled column j
  double* dinv = tcol + CHOL_NB;           // 1 / L[j][j]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int r = tid >> 2, q = tid & 3;
  CHOL_T(0)

  for (int idx = tid; idx < CHOL_NB * CHOL_NB; idx += nt) {
    int rr = idx & (CHOL_NB - 1), c = idx >> 7;
    S[c + rr * CHOL_LDS] = (rr >= c && rr < nb) ? A[rr + c * lda] : 0.0;
  }
  __syncthreads();
  CHOL_T(1)

  for (int j = 0; j < nb; ++j) {
    // t[r] = A[r][j] - sum_{k<j} L[r][k] * L[j][k]   for r >= j
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (r >= j && r < nb) {
      const double* pr = S + r * CHOL_LDS + q;
      const double* pj = S + j * CHOL_LDS + q;
      int k = 0;
      for (; k + 12 + q < j; k += 16) {
        a0 = fma(pr[k], pj[k], a0);
        a1 = fma(pr[k + 4], pj[k + 4], a1);
        a2 = fma(pr[k + 8], pj[k + 8], a2);
        a3 = fma(pr[k + 12], pj[k + 12], a3);
      }
      for (; k + q < j; k += 4) a0 = fma(pr[k], pj[k], a0);
    }
    double acc = (a0 + a1) + (a2 + a3);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0 && r >= j && r < nb) tcol[r] = S[j + r * CHOL_LDS] - acc;
    __syncthreads();
    const double d = tcol[j];
    const double inv = rsqrt(d);
    if (q == 0 && r >= j && r < nb) {
      S[j + r * CHOL_LDS] = tcol[r] * inv;          // r == j: d * rsqrt(d) = sqrt(d)
      if (r == j) {
        dinv[j] = inv;
        if (!(d > 0.0)) atomicCAS(fail, 0, pivot_base + j + 1);
      }
    }
    __syncthreads();
  }
  CHOL_T(2)
  for (int idx = tid; idx < nb * nb; idx += nt) {
    int rr = idx % nb, c = idx / nb;
    A[rr + c * lda] = (rr >= c) ? S[c + rr * CHOL_LDS] : 0.0;
  }
  if (!X) return;
  CHOL_T(3)

  // inverse: V[i][c] = -dinv[i] * ( L[i][c]*V[c][c] + sum_{c<k<i} L[i][k] * V[k][c] ),  c < i
  const int c = r;  // this thread group owns column c of V
  for (int i = 1; i < nb; ++i) {
    double a0 = 0.0, a1 = 0.0;
    if (c < i) {
      const double* pi = S + i * CHOL_LDS;
      const double* pc = S + c * CHOL_LDS;
      int k = c + 1 + q;
      for (; k + 4 < i; k += 8) {
        a0 = fma(pi[k], pc[k], a0);
        a1 = fma(pi[k + 4], pc[k + 4], a1);
      }
      for (; k < i; k += 4) a0 = fma(pi[k], pc[k], a0);
    }
    double acc = a0 + a1;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0 && c < i) S[i + c * CHOL_LDS] = -dinv[i] * fma(S[c + i * CHOL_LDS], dinv[c], acc);
    __syncthreads();
  }
  CHOL_T(4)
  for (int idx = tid; idx < nb * nb; idx += nt) {
    int rr = idx % nb, cc = idx / nb;
    X[rr + cc * ldx] = (rr > cc) ? S[rr + cc * CHOL_LDS] : (rr == cc ? dinv[cc] : 0.0);
  }
  CHOL_T(5)
}

}  // namespace admmb200
