// onepass.cuh -- the A = D iteration (linear SVM / Huber / LAD, unwrapped.cuh) with D read ONCE.
//
// The two-pass iteration streams D twice: D_g*x (fused with the prox) and D_g'*[rhs, dz, u].  Both
// products of a row block need only that block of D, so a CTA keeps a tile of R rows x ALL n columns
// in shared memory and does both from one read:
//     w = T x  ->  z-prox / u-update / norms of the R rows  ->  d += T'[rhs, dz, u]
// R*8 contiguous bytes per column is what decides the HBM efficiency of such a tile; measured on B200
// (tools/cu/tile_read_probe.cu): R = 32 streams at 7.2 TB/s, R = 16 at 4.7, R = 8 at 2.4 -- so the tile
// height is 32 (n <= ~800), 24 (n <= ~1030) or 16 rows, bounded by 227 KB of shared memory, and the
// two-pass kernels stay for wider matrices.
//
// Pipeline inside the single CTA per SM (512 threads): the tile is loaded with cp.async in 4 column
// chunks.  The D*x phase starts on chunk 0 while chunks 1..3 are still in flight; the D' phase releases
// chunk c as soon as it has been used, and the NEXT tile's chunk c is issued into it immediately, so
// loads are in flight during both compute phases.  Column j of the tile sits at T[j*(R+2) + i].
#pragma once
#include "unwrapped.cuh"

namespace admmb200 {

constexpr int OP_THREADS = 512;
constexpr int OP_NCH = 4;
constexpr int OP_MAXCOLS = 3;          // columns per thread in the D' phase: n <= 1536

struct OnepassArgs {
  UwArgs uw;                           // alg == 0 only
  int nv;                              // 1 (nodualerror) or 3
  int64_t ntiles, npad;
  double* dpart;                       // [gridDim.x][nv][npad] per-CTA partial D'[.]
  double* partials;                    // [gridDim.x][UW_NRED]
};

template <int R> struct OnepassCfg {
  static constexpr int RS = R + 2, G = OP_THREADS / R;
  static size_t smem_bytes(int64_t n, int64_t npad) {
    return (size_t)(n * RS + npad + G * R + 3 * R + (OP_THREADS / 32) * UW_NRED) * 8;
  }
};

__device__ __forceinline__ void cp_async16_zfill(double* smem_dst, const double* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(src_bytes)
               : "memory");
}

template <int R>
__global__ void __launch_bounds__(OP_THREADS, 1) uw_onepass_kernel(OnepassArgs a) {
  const UwArgs& u = a.uw;
  if (u.ctl->done) return;
  constexpr int RS = OnepassCfg<R>::RS, RP = R / 2, G = OnepassCfg<R>::G;
  extern __shared__ __align__(16) double sm[];
  const int64_t n = u.n, m = u.m;
  double* T = sm;
  double* xs = T + n * RS;
  double* wpart = xs + a.npad;
  double* rs = wpart + G * R;
  double* redsh = rs + 3 * R;
  const int tid = threadIdx.x, it = u.ctl->it;
  const int64_t CW = (n + OP_NCH - 1) / OP_NCH;
  for (int64_t j = tid; j < n; j += OP_THREADS) xs[j] = u.x[j];

  auto issue = [&](int64_t tile, int c) {               // column chunk c of `tile` -> shared memory
    const int64_t cbeg = c * CW, cend = min(n, cbeg + CW), row0 = tile * R;
    const int64_t units = (cend - cbeg) * RP;
    for (int64_t idx = tid; idx < units; idx += OP_THREADS) {
      const int64_t j = cbeg + idx / RP;
      const int q = (int)(idx % RP);
      const int64_t row = row0 + 2 * q;
      const int bytes = (int)min((int64_t)16, max((int64_t)0, (m - row) * 8));   // rows past m read as zero
      cp_async16_zfill(T + j * RS + 2 * q, u.D + (bytes > 0 ? row : 0) + j * u.ld, bytes);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int64_t tile = blockIdx.x;
  if (tile < a.ntiles)
    for (int c = 0; c < OP_NCH; ++c) issue(tile, c);
  double acc[3][OP_MAXCOLS];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < OP_MAXCOLS; ++c) acc[k][c] = 0.0;
  double racc[UW_NRED];
#pragma unroll
  for (int k = 0; k < UW_NRED; ++k) racc[k] = 0.0;
  const int ri = tid % R, rg = tid / R;

  for (; tile < a.ntiles; tile += gridDim.x) {
    // the R row threads fetch their z, u, aux now; the values are needed after the D*x phase
    const int64_t row = tile * R + tid;
    double zp = 0.0, uold = 0.0, aux = 0.0;
    if (tid < R && row < m) { zp = u.z[row]; uold = u.u[row]; aux = u.aux[row]; }
    // ---- w = T x, chunk by chunk as the loads land
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < OP_NCH; ++c) {
      if (c == 0) cp_async_wait<OP_NCH - 1>();
      else if (c == 1) cp_async_wait<OP_NCH - 2>();
      else if (c == 2) cp_async_wait<OP_NCH - 3>();
      else cp_async_wait<0>();
      __syncthreads();
      const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
      if (rg < G)
        for (int64_t j = cbeg + rg; j < cend; j += G) s = fma(T[j * RS + ri], xs[j], s);
    }
    if (rg < G) wpart[rg * R + ri] = s;
    __syncthreads();
    if (tid < R) {
      double w = 0.0;
#pragma unroll 4
      for (int g = 0; g < G; ++g) w += wpart[g * R + tid];          // fixed order
      double rv = 0.0, dzv = 0.0, uv = 0.0;
      if (row < m) {
        const UwRowOut o = uw_row_core(u, zp, uold, uold, aux, 0.0, w, racc);
        u.z[row] = o.z;
        u.u[row] = o.u;
        if (u.zvals) {
          u.zvals[(int64_t)it * m + row] = o.z;
          u.uvals[(int64_t)it * m + row] = o.u;
        }
        rv = (u.kind >= UW_HUBER) ? (aux + o.z - o.u) : (o.z - o.u);
        dzv = o.dz;
        uv = o.u;
      }
      rs[tid] = rv; rs[R + tid] = dzv; rs[2 * R + tid] = uv;
    }
    __syncthreads();
    // ---- d += T'[rhs, dz, u]; chunk c is released -- and refilled with the next tile -- as soon as it is used
    const int64_t next = tile + gridDim.x;
#pragma unroll
    for (int c = 0; c < OP_NCH; ++c) {
      const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
#pragma unroll
      for (int k = 0; k < OP_MAXCOLS; ++k) {
        const int64_t j = tid + (int64_t)k * OP_THREADS;
        if (j >= cbeg && j < cend) {
          const double* col = T + j * RS;
          if (a.nv == 3) {
            double a0 = acc[0][k], a1 = acc[1][k], a2 = acc[2][k];
#pragma unroll 8
            for (int i = 0; i < R; ++i) {
              const double t = col[i];
              a0 = fma(t, rs[i], a0);
              a1 = fma(t, rs[R + i], a1);
              a2 = fma(t, rs[2 * R + i], a2);
            }
            acc[0][k] = a0; acc[1][k] = a1; acc[2][k] = a2;
          } else {
            double a0 = acc[0][k];
#pragma unroll 8
            for (int i = 0; i < R; ++i) a0 = fma(col[i], rs[i], a0);
            acc[0][k] = a0;
          }
        }
      }
      __syncthreads();
      if (next < a.ntiles) issue(next, c);
    }
  }
  // per-CTA partials; uw_onepass_finish_kernel sums them in CTA order
  double* dp = a.dpart + (int64_t)blockIdx.x * a.nv * a.npad;
#pragma unroll
  for (int k = 0; k < OP_MAXCOLS; ++k) {
    const int64_t j = tid + (int64_t)k * OP_THREADS;
    if (j < n) {
      dp[j] = acc[0][k];
      if (a.nv == 3) { dp[a.npad + j] = acc[1][k]; dp[2 * a.npad + j] = acc[2][k]; }
    }
  }
  block_reduce_store<UW_NRED>(racc, a.partials + (int64_t)blockIdx.x * UW_NRED, redsh);
}

// d_k = sum over CTAs of their partials (fixed order), scalars likewise
__global__ void uw_onepass_finish_kernel(const double* dpart, int nparts, int nv, int64_t n, int64_t npad, double* d,
                                         const double* partials, double* scalars, const LoopCtl* ctl) {
  if (ctl->done) return;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (int64_t)nv * n) {
    const int64_t k = t / n, j = t % n;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += dpart[((int64_t)p * nv + k) * npad + j];
    d[k * npad + j] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < UW_NRED) {
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partials[(int64_t)p * UW_NRED + threadIdx.x];
    scalars[threadIdx.x] = s;
  }
}

}  // namespace admmb200
