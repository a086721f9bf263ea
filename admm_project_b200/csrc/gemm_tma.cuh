// gemm_tma.cuh -- the Gram contraction C = alpha * A'B (+ beta*C, + rho on the diagonal) with BOTH operands read
// K-major straight out of column-major matrices (D'*D of solvers/lasso.m:168, huberfit.m:166, lad.m:134,
// unwrappedadmm.m:115): FP64 DMMA tiles (mma.sync m8n8k4 -- tcgen05 / wgmma have no FP64 kind) fed by TMA.
//
//   * operand tiles are 128 (rows of op) x 16 (k) doubles = 128 B per row, fetched by cp.async.bulk.tensor.2d
//     (SASS UTMALDG) with the 128-byte hardware swizzle into a 6-stage ring; a dedicated producer warp issues two
//     tensor loads per stage and arms the stage's mbarrier with the byte count, consumer warps wait on it -- no
//     thread computes a load address, predicate or zero-fill (the TMA unit clips at the matrix edge);
//   * 8 consumer warps as 2 (M) x 4 (N), warp tile 64 x 32, 64 FP64 accumulators per lane (same as gemm.cuh);
//   * the swizzled layout has no padding, so the fragment loads avoid bank conflicts by PERMUTING K inside a
//     16-wide tile: DMMA step j uses k = {2j, 2j+1, 8+2j, 9+2j} (lane t -> k = 8(t>>1) + 2j + (t&1)).  A and B use the
//     same permutation, so the sum over k is the same set of products; within a half-warp the 16 lanes then touch
//     16 distinct 8-byte words of 8 distinct 16-byte chunks -- all 32 banks exactly once;
//   * lower-tile skipping, grouped tile order and deterministic split-K as in gemm.cuh.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm.cuh"
#include "symtri.cuh"   // mbarrier helpers

namespace admmb200 {

constexpr int GT_BM = 128, GT_BN = 128, GT_BK = 16, GT_STAGES = 6;
constexpr int GT_CONSUMERS = 256, GT_THREADS = GT_CONSUMERS + 32;
constexpr int GT_TILE_BYTES = 128 * GT_BK * 8;                  // 16384
constexpr int GT_SMEM_BYTES = GT_STAGES * 2 * GT_TILE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

struct GemmTmaArgs {
  int64_t M, N, K;
  double* C; int64_t ldc;
  double alpha, beta, diag_add;
  int lower_only;
  int splits; int64_t k_per_split; double* ws;   // ws: [splits][M*N]
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmTmaArgs p) {
  extern __shared__ unsigned char gt_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gt_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = base;                                        // [STAGES][128][128 B]
  unsigned char* sB = base + GT_STAGES * GT_TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + 2 * GT_STAGES * GT_TILE_BYTES);
  uint64_t* empty = full + GT_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int64_t bm = blockIdx.x, bn = blockIdx.y;
  {
    constexpr int GROUP = 12;
    const int64_t tm = gridDim.x, tn = gridDim.y;
    const int64_t pid = (int64_t)blockIdx.y * tm + blockIdx.x;
    const int64_t per_group = GROUP * tn;
    const int64_t g = pid / per_group, first_m = g * GROUP;
    const int64_t gsz = min((int64_t)GROUP, tm - first_m);
    bm = first_m + (pid % per_group) % gsz;
    bn = (pid % per_group) / gsz;
  }
  const int64_t m0 = bm * GT_BM, n0 = bn * GT_BN;
  if (p.lower_only && n0 > m0) return;
  const int split = blockIdx.z;
  const int64_t kbeg = (int64_t)split * p.k_per_split;
  const int64_t kend = min(p.K, kbeg + p.k_per_split);
  const int nk = (int)((kend - kbeg + GT_BK - 1) / GT_BK);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, GT_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == GT_CONSUMERS / 32) {
    // ---- producer warp: one lane drives the TMA unit
    if (lane == 0) {
      for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % GT_STAGES;
        const unsigned ph = (unsigned)((kt / GT_STAGES) & 1);
        mbar_wait(empty + s, ph ^ 1u);                            // fresh barrier: passes at once on the first lap
        mbar_expect_tx(full + s, 2 * GT_TILE_BYTES);
        const int k0 = (int)(kbeg + (int64_t)kt * GT_BK);
        tma_load_2d(sA + s * GT_TILE_BYTES, &tmA, k0, (int)m0, full + s);
        tma_load_2d(sB + s * GT_TILE_BYTES, &tmB, k0, (int)n0, full + s);
      }
    }
    return;
  }

  // ---- consumer warps
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  const int wrow = wm * 64, wcol = wn * 32;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // byte offset inside a 128-byte row of element k = 8(t>>1) + 2j + (t&1) under the 128B swizzle (row & 7 == g)
  int koff[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) koff[j] = (((4 * (t >> 1) + j) ^ g) << 4) + (t & 1) * 8;
  const int arow = (wrow + g) * 128, brow = (wcol + g) * 128;

  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt % GT_STAGES;
    const unsigned ph = (unsigned)((kt / GT_STAGES) & 1);
    mbar_wait(full + s, ph);
    const unsigned char* a_s = sA + s * GT_TILE_BYTES + arow;
    const unsigned char* b_s = sB + s * GT_TILE_BYTES + brow;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double a[8], b[4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) a[mi] = *reinterpret_cast<const double*>(a_s + mi * 1024 + koff[j]);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = *reinterpret_cast<const double*>(b_s + ni * 1024 + koff[j]);
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }

  if (p.splits > 1) {
    double* W = p.ws + (int64_t)split * p.M * p.N;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t r = m0 + wrow + mi * 8 + g, c = n0 + wcol + ni * 8 + 2 * t + e;
          if (r < p.M && c < p.N) W[r + c * p.M] = acc[mi][ni][e];
        }
  } else {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t r = m0 + wrow + mi * 8 + g, c = n0 + wcol + ni * 8 + 2 * t + e;
          if (r < p.M && c < p.N) {
            double v = p.alpha * acc[mi][ni][e];
            if (p.beta != 0.0) v += p.beta * p.C[r + c * p.ldc];
            if (r == c) v += p.diag_add;
            p.C[r + c * p.ldc] = v;
          }
        }
  }
}

}  // namespace admmb200
