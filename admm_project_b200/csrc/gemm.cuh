// gemm.cuh -- FP64 GEMM / SYRK on the DMMA tensor pipe (mma.sync m8n8k4 f64; tcgen05 has no FP64
// kind).  This is the only dense contraction on the reference's path: D'*D and D*D'
// (solvers/lasso.m:168,172; huberfit.m:166; lad.m:134; unwrappedadmm.m:115), and it also carries
// the Cholesky trailing update, the panel solves and the inverse-factor products.
//
// C(MxN, column-major) = alpha * op(A)(MxK) * op(B)(KxN) + beta * C [+ diag_add on the diagonal]
//
// CTA tile 128x128x16, 8 warps (2 along M x 4 along N, warp tile 64x32 = 8x4 DMMA tiles,
// 64 FP64 accumulators per lane), 4-stage cp.async (LDGSTS) pipeline, padded shared memory so
// every fragment LDS.64 is bank-conflict free.  Split-K (deterministic two-pass reduction) keeps
// all 148 SMs busy when the output has few tiles (n = 784 / 1024 Gram matrices of tall D).
#pragma once
#include "common.cuh"

namespace admmb200 {

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 16, GEMM_STAGES = 4, GEMM_THREADS = 256;
constexpr int GEMM_LDK = GEMM_BK + 4;    // K-major tile: [128][20]  (row stride 160 B = 32 mod 128)
constexpr int GEMM_LDM = GEMM_BM + 4;    // M-major tile: [16][132]  (row stride 1056 B = 32 mod 128)
constexpr int GEMM_TILE_DOUBLES = 128 * GEMM_LDK;  // 2560 >= 16*132 = 2112
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * 2 * GEMM_TILE_DOUBLES * 8;  // 163840

struct GemmArgs {
  int64_t M, N, K;
  const double* A; int64_t lda;
  const double* B; int64_t ldb;
  double* C; int64_t ldc;
  double alpha, beta, diag_add;
  int lower_only;
  int a_lower, b_lower;   // op(A) / op(B) is lower triangular: restrict the K range per tile
  int a_upper;            // op(A) is upper triangular
  int b_upper;            // op(B) is upper triangular: op(B)[k][j] = 0 for k > j
  int batch; int64_t strideA, strideB, strideC;
  int splits; int64_t k_per_split; double* ws;   // ws: [batch][splits][M*N]
  int bm;                 // CTA tile height (128 or 256)
  int tri_skip;           // split-K with a triangular op(A): work units whose K range is empty neither run nor get summed
  int persist;            // > 0: the grid is `persist` CTAs that walk the tm x tn x nz tiles round-robin (background GEMMs of
  int64_t tm, tn; int nz; //      the look-ahead Cholesky keep off some SMs so the latency-bound chain always finds a free one)
};

// Loads one ROWS x 16 operand tile (ROWS = 128 or 64).  KMAJOR: element (r,k) at g[k + r*ld]; else at g[r + k*ld].
template <bool KMAJOR, int VEC, int ROWS>
__device__ __forceinline__ void gemm_load_tile(double* s, const double* __restrict__ g, int64_t ld,
                                               int64_t r0, int64_t R, int64_t k0, int64_t Kend, int tid) {
  if (KMAJOR) {
    if (VEC == 2) {
#pragma unroll
      for (int it = 0; it < (ROWS + 31) / 32; ++it) {
        int r = (tid >> 3) + 32 * it, c = (tid & 7) * 2;
        if (ROWS < 32 && r >= ROWS) continue;      // a 16-row tile (few right-hand sides): half the threads load
        int64_t gr = r0 + r, gk = k0 + c;
        int64_t left = (gr < R) ? (Kend - gk) : 0;
        int nb = left >= 2 ? 16 : (left == 1 ? 8 : 0);
        const double* src = nb ? (g + gk + gr * ld) : g;
        cp_async16(&s[r * GEMM_LDK + c], src, nb);
      }
    } else {
#pragma unroll
      for (int it = 0; it < ROWS / 16; ++it) {
        int r = (tid >> 4) + 16 * it, c = tid & 15;
        int64_t gr = r0 + r, gk = k0 + c;
        int nb = (gr < R && gk < Kend) ? 8 : 0;
        const double* src = nb ? (g + gk + gr * ld) : g;
        cp_async8(&s[r * GEMM_LDK + c], src, nb);
      }
    }
  } else {
    constexpr int GEMM_LDM = (ROWS > 128) ? ROWS + 4 : admmb200::GEMM_LDM;   // 260 doubles: row stride 2080 B = 32 mod 128
    constexpr int CH2 = ROWS / 2;            // 16-byte chunks per k-row
    constexpr int KR2 = GEMM_THREADS / CH2;  // k-rows per pass
    if (VEC == 2) {
#pragma unroll
      for (int it = 0; it < GEMM_BK / KR2; ++it) {
        int kk = tid / CH2 + KR2 * it, c = (tid % CH2) * 2;
        int64_t gr = r0 + c, gk = k0 + kk;
        int64_t left = (gk < Kend) ? (R - gr) : 0;
        int nb = left >= 2 ? 16 : (left == 1 ? 8 : 0);
        const double* src = nb ? (g + gr + gk * ld) : g;
        cp_async16(&s[kk * GEMM_LDM + c], src, nb);
      }
    } else {
      constexpr int KR1 = GEMM_THREADS / ROWS;
#pragma unroll
      for (int it = 0; it < GEMM_BK / KR1; ++it) {
        int kk = tid / ROWS + KR1 * it, c = tid % ROWS;
        int64_t gr = r0 + c, gk = k0 + kk;
        int nb = (gr < R && gk < Kend) ? 8 : 0;
        const double* src = nb ? (g + gr + gk * ld) : g;
        cp_async8(&s[kk * GEMM_LDM + c], src, nb);
      }
    }
  }
}

// BN = 128: 8 warps as 2 (M) x 4 (N), warp tile 64 x 32.  BN = 64 (skinny right-hand sides, e.g. the
// 64-lambda batch): 8 warps as 4 (M) x 2 (N), warp tile 32 x 32, so no DMMA is spent on padding columns.
// BM = 256, BN = 64 (tall triangular op(A) times a few right-hand sides: the lambda batch on an 8192-wide factor):
// 8 warps as 4 (M) x 2 (N) with the full 64 x 32 warp tile, 3 stages; K is cut into uniform chunks so the
// (row tile, K chunk) units of a triangular operand are equal pieces of work for the 148 SMs.  BN = 32 / 16 with
// BM = 256 (K-major right-hand sides only): the lambda columns of ONE rank when the batch is split over 2 / 4 / 8
// GPUs -- warp tile 64 x 16 / 64 x 8, so a rank with 8 columns does a quarter of the DMMA work of the 64-wide tile.
template <int BM> struct GemmCfg {
  static constexpr int STAGES = (BM == 256) ? 3 : GEMM_STAGES;
  static constexpr int LDM = (BM > 128) ? BM + 4 : GEMM_LDM;
  static constexpr int A_TILE = (BM * GEMM_LDK > GEMM_BK * LDM) ? BM * GEMM_LDK : GEMM_BK * LDM;
  static constexpr int SMEM_BYTES = STAGES * (A_TILE + GEMM_TILE_DOUBLES) * 8;
};
template <bool AK, bool BK, int VEC, int BN, int BM = 128>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_f64_dmma_kernel(GemmArgs p) {
  constexpr int WM = (BM == 256) ? 4 : ((BM == 64) ? 2 : ((BN == 128) ? 2 : 4));    // warps along M
  constexpr int WN = 8 / WM;
  constexpr int MI = BM / WM / 8;
  constexpr int NI = BN / WN / 8;            // 4, or 2 for the 64 x 64 tile (warp tile 32 x 16)
  constexpr int GEMM_BM = BM;
  constexpr int GEMM_STAGES = GemmCfg<BM>::STAGES;
  constexpr int A_TILE = GemmCfg<BM>::A_TILE;
  constexpr int LDM_A = GemmCfg<BM>::LDM;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  // grouped tile order (GROUP x GROUP super-tiles): the ~148 CTAs resident at any time then share
  // ~12 A panels and ~12 B panels instead of 64 + 3, which cuts the DRAM re-reads of the Gram ~3x
  // (ncu: 69.6 GB for a 4.3 GB matrix with the natural order).
  const int64_t tm = p.persist ? p.tm : (int64_t)gridDim.x, tn = p.persist ? p.tn : (int64_t)gridDim.y;
  const int64_t per_z = tm * tn, total = p.persist ? per_z * p.nz : 1;
  bool first_tile = true;
  for (int64_t lin = p.persist ? (int64_t)blockIdx.x : 0; lin < total; lin += p.persist ? (int64_t)gridDim.x : 1) {
  if (!first_tile) __syncthreads();     // every warp is done with the shared-memory ring of the tile before
  first_tile = false;
  const int bz = p.persist ? (int)(lin / per_z) : (int)blockIdx.z;
  int64_t bm, bn;
  {
    constexpr int GROUP = 12;
    const int64_t pid = p.persist ? lin - (int64_t)bz * per_z : (int64_t)blockIdx.y * tm + blockIdx.x;
    const int64_t per_group = GROUP * tn;
    const int64_t g = pid / per_group, first_m = g * GROUP;
    const int64_t gsz = min((int64_t)GROUP, tm - first_m);
    bm = first_m + (pid % per_group) % gsz;
    bn = (pid % per_group) / gsz;
  }
  const int64_t m0 = bm * GEMM_BM, n0 = bn * BN;
  if (p.lower_only && n0 > m0) continue;  // tile strictly above the diagonal
  const int batch = bz / p.splits, split = bz - batch * p.splits;
  const double* A = p.A + (int64_t)batch * p.strideA;
  const double* B = p.B + (int64_t)batch * p.strideB;
  int64_t kbeg = (int64_t)split * p.k_per_split;
  int64_t kend = min(p.K, kbeg + p.k_per_split);
  if (p.b_lower) kbeg = max(kbeg, n0);                      // op(B)[k][j] = 0 for k < j
  if (p.a_lower) kend = min(kend, m0 + (int64_t)GEMM_BM);   // op(A)[i][k] = 0 for k > i
  if (p.a_upper) kbeg = max(kbeg, m0);                      // op(A)[i][k] = 0 for k < i
  if (p.b_upper) kend = min(kend, n0 + (int64_t)BN);        // op(B)[k][j] = 0 for k > j
  const int nk = (int)((kend - kbeg + GEMM_BK - 1) / GEMM_BK);
  if (p.tri_skip && nk <= 0) continue;    // empty unit of a triangular operand: the reduce pass skips it too

  double* sA = smem;
  double* sB = smem + GEMM_STAGES * A_TILE;

  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % WM, wn = warp / WM;
  const int wrow = wm * (MI * 8);

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < nk) {
      gemm_load_tile<AK, VEC, GEMM_BM>(sA + s * A_TILE, A, p.lda, m0, p.M, kbeg + (int64_t)s * GEMM_BK, kend, tid);
      gemm_load_tile<BK, VEC, BN>(sB + s * GEMM_TILE_DOUBLES, B, p.ldb, n0, p.N, kbeg + (int64_t)s * GEMM_BK, kend, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    {
      int kn = kt + GEMM_STAGES - 1;
      if (kn < nk) {
        int slot = kn % GEMM_STAGES;
        gemm_load_tile<AK, VEC, GEMM_BM>(sA + slot * A_TILE, A, p.lda, m0, p.M, kbeg + (int64_t)kn * GEMM_BK, kend, tid);
        gemm_load_tile<BK, VEC, BN>(sB + slot * GEMM_TILE_DOUBLES, B, p.ldb, n0, p.N, kbeg + (int64_t)kn * GEMM_BK, kend, tid);
      }
      cp_async_commit();
    }
    const double* a_s = sA + (kt % GEMM_STAGES) * A_TILE;
    const double* b_s = sB + (kt % GEMM_STAGES) * GEMM_TILE_DOUBLES;
#pragma unroll
    for (int ks = 0; ks < GEMM_BK / 4; ++ks) {
      double a[MI], b[NI];
      const int k = ks * 4 + t;
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        int r = wrow + mi * 8 + g;
        a[mi] = AK ? a_s[r * GEMM_LDK + k] : a_s[k * LDM_A + r];
      }
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        int c = wn * (NI * 8) + ni * 8 + g;
        b[ni] = BK ? b_s[c * GEMM_LDK + k] : b_s[k * GEMM_LDM + c];
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();

  // epilogue
  if (p.splits > 1) {
    double* W = p.ws + ((int64_t)batch * p.splits + split) * p.M * p.N;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          int64_t r = m0 + wrow + mi * 8 + g, c = n0 + wn * (NI * 8) + ni * 8 + 2 * t + e;
          if (r < p.M && c < p.N) W[r + c * p.M] = acc[mi][ni][e];
        }
  } else {
    double* C = p.C + (int64_t)batch * p.strideC;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          int64_t r = m0 + wrow + mi * 8 + g, c = n0 + wn * (NI * 8) + ni * 8 + 2 * t + e;
          if (r < p.M && c < p.N) {
            double v = p.alpha * acc[mi][ni][e];
            if (p.beta != 0.0) v += p.beta * C[r + c * p.ldc];
            if (r == c) v += p.diag_add;
            C[r + c * p.ldc] = v;
          }
        }
  }
  }   // tile loop
}

// second pass of split-K: fixed summation order over the splits (deterministic)
__global__ void gemm_splitk_reduce_kernel(GemmArgs p) {
  const int64_t total = p.M * p.N;
  const int batch = blockIdx.y;
  const double* W = p.ws + (int64_t)batch * p.splits * total;
  double* C = p.C + (int64_t)batch * p.strideC;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx % p.M, c = idx / p.M;
    if (p.lower_only && (c / p.bm) > (r / p.bm)) continue;         // lower_only is used with square tiles only
    double s = 0.0;
    int k0 = 0, k1 = p.splits;
    if (p.tri_skip) {            // the same K clamps the GEMM kernel applied to row tile r / bm
      const int64_t m0 = (r / p.bm) * p.bm;
      if (p.a_lower) k1 = (int)min((int64_t)k1, (m0 + p.bm + p.k_per_split - 1) / p.k_per_split);
      if (p.a_upper) k0 = (int)(m0 / p.k_per_split);
    }
    for (int k = k0; k < k1; ++k) s += W[(int64_t)k * total + idx];
    double v = p.alpha * s;
    if (p.beta != 0.0) v += p.beta * C[r + c * p.ldc];
    if (r == c) v += p.diag_add;
    C[r + c * p.ldc] = v;
  }
}

// copy the lower triangle onto the upper one (full symmetric result of a lower-only SYRK)
__global__ void symmetrize_lower_kernel(double* C, int64_t n, int64_t ldc) {
  __shared__ double tile[32][33];
  int bx = blockIdx.x, by = blockIdx.y;  // tile (row-block by, col-block bx) with by >= bx read
  if (by < bx) return;
  int64_t r = (int64_t)by * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = (int64_t)bx * 32 + j;
    tile[j][threadIdx.x] = (r < n && c < n) ? C[r + c * ldc] : 0.0;
  }
  __syncthreads();
  // write transposed: element (c, r) <- (r, c) for r > c
  int64_t cc = (int64_t)bx * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t rr = (int64_t)by * 32 + j;
    if (rr < n && cc < n && rr > cc) C[cc + rr * ldc] = tile[threadIdx.x][j];
  }
}

// out (cols x rows, ldo) = in (rows x cols, ldi) transposed
__global__ void transpose_kernel(const double* __restrict__ in, int64_t rows, int64_t cols, int64_t ldi,
                                 double* __restrict__ out, int64_t ldo) {
  __shared__ double tile[32][33];
  int64_t r = (int64_t)blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = (int64_t)blockIdx.y * 32 + j;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[r + c * ldi] : 0.0;
  }
  __syncthreads();
  int64_t oc = (int64_t)blockIdx.y * 32 + threadIdx.x;  // row index in `out`
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t orow = (int64_t)blockIdx.x * 32 + j;          // column index in `out`
    if (oc < cols && orow < rows) out[oc + orow * ldo] = tile[threadIdx.x][j];
  }
}

// zero the strict upper triangle
__global__ void zero_upper_kernel(double* A, int64_t n, int64_t lda) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t c = blockIdx.y;
  if (r < n && r < c) A[r + c * lda] = 0.0;
}

}  // namespace admmb200
