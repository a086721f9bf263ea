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
// Pipeline inside the single CTA per SM (512 threads): the tile is loaded with cp.async in NCH column
// chunks.  The D*x phase starts on chunk 0 while the later chunks are still in flight; the D' phase
// releases chunk c as soon as it has been used, and the NEXT tile's chunk c is issued into it
// immediately, so loads are in flight during both compute phases.  Column j of the tile sits at
// T[j*(R+2) + i]: 16-byte aligned columns, and both access patterns (row pairs of one column in the D*x
// phase, one column per lane in the D' phase) are free of bank conflicts with LDS.128.
#pragma once
#include "unwrapped.cuh"
#include "p2p.cuh"

namespace admmb200 {

constexpr int OP_THREADS = 512;
constexpr int OP_SLOTS = 256;          // D' phase: thread = (column slot, row half)
constexpr int OP_MAXCOLS = 6;          // columns per slot: n <= 1536

struct OnepassArgs {
  UwArgs uw;                           // alg == 0 only
  int nv;                              // 1 (nodualerror) or 3
  int64_t ntiles, npad;
  double* dpart;                       // [gridDim.x][2 row halves][nv][npad] per-CTA partial D'[.]
  double* partials;                    // [gridDim.x][UW_NRED]
};

template <int R> struct OnepassCfg {
  static constexpr int RS = R + 2;                      // column stride of the tile: 16-byte aligned, conflict-free LDS.128
  static constexpr int RP = R / 2;                      // row pairs
  static constexpr int JSTEP = OP_THREADS / RP;         // column groups of the load / D*x mapping
  static constexpr int ACTIVE = JSTEP * RP;             // threads that take part in it (504 of 512 for R = 24)
  static size_t smem_bytes(int64_t n, int64_t npad) {
    return (size_t)(n * RS + npad + JSTEP * R + 3 * R + (OP_THREADS / 32) * UW_NRED) * 8;
  }
};

__device__ __forceinline__ void cp_async16_zfill(double* smem_dst, const double* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(src_bytes)
               : "memory");
}

template <int R, int OP_NCH>
__global__ void __launch_bounds__(OP_THREADS, 1) uw_onepass_kernel(OnepassArgs a) {
  const UwArgs& u = a.uw;
  if (u.ctl->done) return;
  using Cfg = OnepassCfg<R>;
  constexpr int RS = Cfg::RS, RP = Cfg::RP, JSTEP = Cfg::JSTEP, ACTIVE = Cfg::ACTIVE, RH = R / 2;
  extern __shared__ __align__(16) double sm[];
  const int64_t n = u.n, m = u.m;
  double* T = sm;
  double* xs = T + n * RS;
  double* wpart = xs + a.npad;          // [JSTEP][R]
  double* rs = wpart + JSTEP * R;       // [3][R]: rhs, dz, u of the tile's rows
  double* redsh = rs + 3 * R;
  const int tid = threadIdx.x, it = u.ctl->it;
  const int64_t CW = (n + OP_NCH - 1) / OP_NCH;
  for (int64_t j = tid; j < n; j += OP_THREADS) xs[j] = u.x[j];
  // load / D*x mapping: thread = (row pair q, column group j0); D' mapping: thread = (column slot sc, row half hh)
  const int q = tid % RP, j0 = tid / RP;
  const int sc = tid % OP_SLOTS, hh = tid / OP_SLOTS;

  auto issue = [&](int64_t tile, int c) {               // column chunk c of `tile` -> shared memory
    if (tid < ACTIVE) {
      const int64_t cbeg = c * CW, cend = min(n, cbeg + CW), row = tile * R + 2 * q;
      const int bytes = (int)min((int64_t)16, max((int64_t)0, (m - row) * 8));   // rows past m read as zero
      const double* src = u.D + (bytes > 0 ? row : 0) + (cbeg + j0) * u.ld;      // a valid address even when bytes == 0
      double* dst = T + (cbeg + j0) * RS + 2 * q;
      for (int64_t j = cbeg + j0; j < cend; j += JSTEP) {
        cp_async16_zfill(dst, src, bytes);
        src += (int64_t)JSTEP * u.ld;
        dst += JSTEP * RS;
      }
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
  const bool rowthread = (tid < 8 * R) && ((tid & 7) == 0);   // one of 8 lanes that sum a row's partial dots
  const int myrow = tid >> 3;

  for (; tile < a.ntiles; tile += gridDim.x) {
    // the row threads fetch their z, u, aux now; the values are needed after the D*x phase
    const int64_t row = tile * R + myrow;
    double zp = 0.0, uold = 0.0, aux = 0.0;
    if (rowthread && row < m) { zp = u.z[row]; uold = u.u[row]; aux = u.aux[row]; }
    // ---- w = T x, chunk by chunk as the loads land: a row PAIR per thread (one LDS.128 per column)
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int c = 0; c < OP_NCH; ++c) {
      if (OP_NCH - 1 - c == 3) cp_async_wait<3>();
      else if (OP_NCH - 1 - c == 2) cp_async_wait<2>();
      else if (OP_NCH - 1 - c == 1) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncthreads();
      if (tid < ACTIVE) {
        const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
        const double* tp = T + (cbeg + j0) * RS + 2 * q;
#pragma unroll 4
        for (int64_t j = cbeg + j0; j < cend; j += JSTEP) {
          const double2 t = *reinterpret_cast<const double2*>(tp);
          const double xj = xs[j];
          s0 = fma(t.x, xj, s0);
          s1 = fma(t.y, xj, s1);
          tp += JSTEP * RS;
        }
      }
    }
    if (tid < ACTIVE) *reinterpret_cast<double2*>(wpart + j0 * R + 2 * q) = make_double2(s0, s1);
    __syncthreads();
    if (tid < 8 * R) {                                   // whole warps: 8 lanes per row, fixed summation order
      double w = 0.0;
      for (int g = tid & 7; g < JSTEP; g += 8) w += wpart[g * R + myrow];
      w += __shfl_xor_sync(0xffffffffu, w, 4);
      w += __shfl_xor_sync(0xffffffffu, w, 2);
      w += __shfl_xor_sync(0xffffffffu, w, 1);
      if ((tid & 7) == 0) {
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
        rs[myrow] = rv; rs[R + myrow] = dzv; rs[2 * R + myrow] = uv;
      }
    }
    __syncthreads();
    // ---- d += T'[rhs, dz, u] on this thread's half of the rows; chunk c is released -- and refilled with the
    // next tile -- as soon as it has been used
    const int64_t next = tile + gridDim.x;
#pragma unroll
    for (int c = 0; c < OP_NCH; ++c) {
      const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
#pragma unroll
      for (int k = 0; k < OP_MAXCOLS; ++k) {
        const int64_t j = sc + (int64_t)k * OP_SLOTS;
        if (j >= cbeg && j < cend) {
          const double* col = T + j * RS + hh * RH;
          const double* rr = rs + hh * RH;
          if (a.nv == 3) {
            double a0 = acc[0][k], a1 = acc[1][k], a2 = acc[2][k];
#pragma unroll
            for (int i = 0; i < RH; i += 2) {
              const double2 t = *reinterpret_cast<const double2*>(col + i);
              const double2 r0 = *reinterpret_cast<const double2*>(rr + i);
              const double2 r1 = *reinterpret_cast<const double2*>(rr + R + i);
              const double2 r2 = *reinterpret_cast<const double2*>(rr + 2 * R + i);
              a0 = fma(t.x, r0.x, a0); a0 = fma(t.y, r0.y, a0);
              a1 = fma(t.x, r1.x, a1); a1 = fma(t.y, r1.y, a1);
              a2 = fma(t.x, r2.x, a2); a2 = fma(t.y, r2.y, a2);
            }
            acc[0][k] = a0; acc[1][k] = a1; acc[2][k] = a2;
          } else {
            double a0 = acc[0][k];
#pragma unroll
            for (int i = 0; i < RH; i += 2) {
              const double2 t = *reinterpret_cast<const double2*>(col + i);
              const double2 r0 = *reinterpret_cast<const double2*>(rr + i);
              a0 = fma(t.x, r0.x, a0); a0 = fma(t.y, r0.y, a0);
            }
            acc[0][k] = a0;
          }
        }
      }
      __syncthreads();
      if (next < a.ntiles) issue(next, c);
    }
  }
  // per-CTA, per-row-half partials; uw_onepass_finish_kernel sums them in a fixed order
  double* dp = a.dpart + ((int64_t)blockIdx.x * 2 + hh) * a.nv * a.npad;
#pragma unroll
  for (int k = 0; k < OP_MAXCOLS; ++k) {
    const int64_t j = sc + (int64_t)k * OP_SLOTS;
    if (j < n) {
      dp[j] = acc[0][k];
      if (a.nv == 3) { dp[a.npad + j] = acc[1][k]; dp[2 * a.npad + j] = acc[2][k]; }
    }
  }
  block_reduce_store<UW_NRED>(racc, a.partials + (int64_t)blockIdx.x * UW_NRED, redsh);
}

// d_k = sum over the CTAs' partials, scalars likewise.  One WARP per output: lane l adds parts l, l+32, ...
// in order, then a fixed xor tree -- a thread per output would walk ~300 dependent L2 loads (20+ us).
__global__ void __launch_bounds__(256) uw_onepass_finish_kernel(const double* dpart, int ndparts, int nparts, int nv, int64_t n,
                                                                int64_t npad, double* d, const double* partials,
                                                                double* scalars, const LoopCtl* ctl, P2PDev mail, int use_mail) {
  if (ctl->done) return;
  if (use_mail && *mail.err) return;
  // row-sharded runs: the sums of THIS rank go straight into every rank's mailbox (p2p.cuh) -- message layout
  // [d (nv x npad) ; scalars] -- and the last CTA raises the flags; uw_epilogue_kernel waits and adds the ranks up
  const unsigned long long seq = use_mail ? *mail.seq : 0ull;
  const int par = (int)(seq & 1);
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // output index
  const int64_t nout = (int64_t)nv * n;
  if (w < nout + UW_NRED) {
    double s = 0.0;
    if (w < nout) {
      const int64_t k = w / n, j = w % n;
      for (int p = lane; p < ndparts; p += 32) s += dpart[((int64_t)p * nv + k) * npad + j];
    } else {
      for (int p = lane; p < nparts; p += 32) s += partials[(int64_t)p * UW_NRED + (w - nout)];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const int64_t idx = (w < nout) ? (w / n) * npad + (w % n) : (int64_t)nv * npad + (w - nout);
      if (use_mail) p2p_store(mail, par, idx, s);
      else if (w < nout) d[idx] = s;
      else scalars[w - nout] = s;
    }
  }
  if (use_mail) p2p_signal(mail, par, seq, gridDim.x);
}

}  // namespace admmb200
