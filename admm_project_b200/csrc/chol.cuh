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
constexpr int CHOL_LDS = 130;                // row stride: private-row accesses 2-way conflicted at worst, rows 16-byte aligned
constexpr int CHOL_DIAG_THREADS = 256;       // 8 warps, up to 255 registers each (32x32 tiles live in registers)
constexpr int CHOL_DIAG_SMEM = (CHOL_NB * CHOL_LDS + CHOL_NB) * 8;

#ifdef CHOL_PROFILE
__device__ long long chol_prof[8];
#define CHOL_T(i) if (threadIdx.x == 0) chol_prof[i] = clock64();
#else
#define CHOL_T(i)
#endif

// One CTA: in-place lower Cholesky of the nb x nb block A (nb <= 128), strict upper part of the
// block zeroed, X = inv(L) written as a full nb x nb block (upper part zero).
// *fail is set to (1 + global index of the offending pivot) when a pivot is not positive.
//
// A dependent "store -> __syncthreads -> load" step through shared memory costs ~650 cycles on B200
// (tools/cu/lat_probe.cu), so a column-at-a-time kernel spends 128 x 2 x 650 cycles waiting.  Here
// the block is a 4 x 4 grid of 32 x 32 tiles that live in REGISTERS of one warp (lane = row or
// column): the sequential part (tile Cholesky, tile triangular solve / inverse) talks through warp
// shuffles and broadcast shared-memory reads only, and CTA-wide barriers separate tile phases
// (13 for the factorisation, 4 for the inverse).  S(r,c) holds L in the lower triangle; the
// inverse is kept transposed in the strict upper triangle (S(c,r) = V[r][c], r > c), its diagonal
// in dinv[].
#define CHOL_S(r, c) S[(r) * CHOL_LDS + (c)]
__global__ void __launch_bounds__(CHOL_DIAG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int nb, double* __restrict__ X, int64_t ldx,
                  int* fail, int pivot_base) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;
  double* dinv = sm + CHOL_NB * CHOL_LDS;   // 1 / L[j][j]
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  CHOL_T(0)
  // rows / columns >= nb are padded with the identity, so every tile is full
  for (int idx = tid; idx < CHOL_NB * CHOL_NB; idx += nt) {
    const int r = idx & (CHOL_NB - 1), c = idx >> 7;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nb && c < nb) v = (r >= c) ? A[r + c * lda] : 0.0;
    CHOL_S(r, c) = v;
  }
  __syncthreads();
  CHOL_T(1)

  for (int kb = 0; kb < 4; ++kb) {
    const int o = 32 * kb;
    if (warp == 0) {   // ---- Cholesky of the diagonal tile, lane = row
      // column j is published through shared memory (one STS + __syncwarp per column) and read back
      // as broadcasts; 31-j shuffles per column were 3x slower
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = CHOL_S(o + lane, o + c);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        if (!(d > 0.0) && lane == 0 && o + j < nb) atomicCAS(fail, 0, pivot_base + o + j + 1);
        const double inv = rsqrt(d);
        a[j] = (lane == j) ? d * inv : ((lane > j) ? a[j] * inv : 0.0);    // L[r][j]
        if (lane == 0) dinv[o + j] = inv;
        CHOL_S(o + lane, o + j) = a[j];
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < 32; ++c) a[c] = fma(-a[j], CHOL_S(o + c, o + j), a[c]);   // L[c][j] broadcast
      }
    }
    __syncthreads();
    if (warp < 3 - kb) {   // ---- tiles below: X = A * inv(Lkk)', lane = row
      const int ro = 32 * (kb + 1 + warp);
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = CHOL_S(ro + lane, o + c);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double x = a[j] * dinv[o + j];
        a[j] = x;
#pragma unroll
        for (int c = j + 1; c < 32; ++c) a[c] = fma(-x, CHOL_S(o + c, o + j), a[c]);
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) CHOL_S(ro + lane, o + c) = a[c];
    }
    __syncthreads();
    {   // ---- trailing tiles (i, j), kb < j <= i: C -= L_i * L_j', two 16-column halves per tile
      const int nrem = 3 - kb;                     // remaining tile rows
      const int nunits = nrem * (nrem + 1);        // (tiles) x 2 halves
      for (int u = warp; u < nunits; u += CHOL_DIAG_THREADS / 32) {
        const int tile = u >> 1, half = u & 1;
        int ti = 0, rem = tile;                     // tile -> (ti >= tj) in the lower triangle of nrem x nrem
        while (rem > ti) { rem -= ti + 1; ++ti; }
        const int tj = rem;
        const int ro = 32 * (kb + 1 + ti), co = 32 * (kb + 1 + tj) + 16 * half;
        double li[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) li[k] = CHOL_S(ro + lane, o + k);
#pragma unroll
        for (int cc = 0; cc < 16; cc += 4) {       // 4 independent accumulation chains
          double acc[4];
          const double2* lj[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[q] = CHOL_S(ro + lane, co + cc + q);
            lj[q] = reinterpret_cast<const double2*>(&CHOL_S(co + cc + q, o));
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const double2 v = lj[q][k];            // L_j[cc+q][2k], [2k+1] (broadcast)
              acc[q] = fma(-li[2 * k], v.x, acc[q]);
              acc[q] = fma(-li[2 * k + 1], v.y, acc[q]);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) CHOL_S(ro + lane, co + cc + q) = acc[q];
        }
      }
    }
    __syncthreads();
  }
  CHOL_T(2)
  for (int idx = tid; idx < nb * nb; idx += nt) {
    const int r = idx % nb, c = idx / nb;
    A[r + c * lda] = (r >= c) ? CHOL_S(r, c) : 0.0;
  }
  if (!X) return;
  CHOL_T(3)

  // ---- inverse of the diagonal tiles, lane = column of V; stored transposed in the strict upper part
  if (warp < 4) {
    const int o = 32 * warp;
    double v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      v[k] *= dinv[o + k];
#pragma unroll
      for (int r = k + 1; r < 32; ++r) v[r] = fma(-CHOL_S(o + r, o + k), v[k], v[r]);
    }
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (r > lane) CHOL_S(o + lane, o + r) = v[r];
  }
  __syncthreads();
  // ---- off-diagonal tiles V_ij = -V_ii * sum_{k=j}^{i-1} L_ik * V_kj, by tile diagonals, lane = column
  for (int d = 1; d < 4; ++d) {
    if (warp < 4 - d) {
      const int tj = warp, ti = warp + d;
      const int jo = 32 * tj, io = 32 * ti;
      double T[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) T[r] = 0.0;
      for (int tk = tj; tk < ti; ++tk) {
        const int ko = 32 * tk;
        double vk[32];                               // V_kj[t][lane]
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          double x = CHOL_S(jo + lane, ko + t);      // transposed storage (strict upper)
          if (tk == tj) x = (t > lane) ? x : (t == lane ? dinv[jo + lane] : 0.0);
          vk[t] = x;
        }
#pragma unroll
        for (int r = 0; r < 32; r += 4) {            // 4 independent accumulation chains
          const double2* li[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) li[q] = reinterpret_cast<const double2*>(&CHOL_S(io + r + q, ko));
#pragma unroll
          for (int t = 0; t < 16; ++t) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const double2 l2 = li[q][t];           // L_ik[r+q][2t], [2t+1] (broadcast)
              T[r + q] = fma(l2.x, vk[2 * t], T[r + q]);
              T[r + q] = fma(l2.y, vk[2 * t + 1], T[r + q]);
            }
          }
        }
      }
      // out[r] = -(V_ii[r][r] T[r] + sum_{t<r} V_ii[r][t] T[t]); column sweeps keep 31-t chains independent
      double outv[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) outv[r] = dinv[io + r] * T[r];
#pragma unroll
      for (int t = 0; t < 31; ++t) {
#pragma unroll
        for (int r = t + 1; r < 32; ++r) outv[r] = fma(CHOL_S(io + t, io + r), T[t], outv[r]);   // V_ii[r][t], transposed storage
      }
#pragma unroll
      for (int r = 0; r < 32; ++r) CHOL_S(jo + lane, io + r) = -outv[r];
    }
    __syncthreads();
  }
  CHOL_T(4)
  for (int idx = tid; idx < nb * nb; idx += nt) {
    const int r = idx % nb, c = idx / nb;
    X[r + c * ldx] = (r > c) ? CHOL_S(c, r) : (r == c ? dinv[c] : 0.0);
  }
  CHOL_T(5)
}
#undef CHOL_S

}  // namespace admmb200
