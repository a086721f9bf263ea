// tv.cuh -- 1-D total variation denoising (solvers/totalvariation.m:122-164, getProxOps.m:172-199,
// xminTotalVariation :1044-1048).  Constraint D*x - z = 0 with the sparse difference operator
// (D x)_i = x_i - x_{i+1} (i < n), (D x)_n = x_n.
//
// x-update: (I + rho*D'D) x = s + rho*D'(z - u).  The reference rebuilds and factorises the sparse
// matrix every iteration (:1047); the matrix is the constant tridiagonal  diag = [1+rho, 1+2rho, ...],
// off-diagonal -rho, so the engine runs the Thomas recurrences as two block-parallel affine scans
//   forward :  y_i = r_i + (rho/delta_{i-1}) y_{i-1}        delta_i = t_i - rho^2/delta_{i-1}
//   backward:  x_i = (y_i + rho x_{i+1}) / delta_i
// in ONE kernel: each CTA owns a segment and starts both recurrences K elements outside it; the
// recurrences are contractions (rho/delta < 1), so the influence of the unknown carry decays as
// (rho/delta*)^K and K is chosen for < 4e-18.  1/delta_i comes from a host-built table (the pivot
// sequence reaches its fixed point after a few dozen to a few thousand entries).
// z/u pass: tv_prox_kernel fuses D*x, the relaxed soft threshold, the u-update and every norm
// (including the D' stencils of the dual residual) into one sweep.
#pragma once
#include "common.cuh"
#include "prox.cuh"

namespace admmb200 {

constexpr int TV_THREADS = 512;
// E consecutive elements per thread: E = 8 (4096-element segments, 70 KB of shared memory, two CTAs
// per SM so one CTA's loads overlap another's scans) for small halos, E = 16 (8192) for large rho.
template <int E> struct TvCfg {
  static constexpr int SEG = TV_THREADS * E;                 // elements per CTA (segment + both halos)
  static constexpr int SMEM_DOUBLES = SEG + SEG / E + 32;    // padded: one spare slot per thread chunk
  static constexpr int SMEM_BYTES = 2 * SMEM_DOUBLES * 8 + 64 * 8 * 2;
};
template <int E> __host__ __device__ constexpr int TV_PAD(int i) { return i + i / E; }

struct TvSolveArgs {
  int64_t n;
  const double *s, *z, *u;
  double* x;
  double rho;
  const double* invdelta;   // table of 1/delta_i, i < ntab; beyond: inv_star (the fixed point)
  double inv_star;
  int ntab;
  int halo;                 // K
  const int* done;
};

struct Affine { double A, B; };   // v -> A*v + B
__device__ __forceinline__ Affine compose(const Affine& second, const Affine& first) {
  return Affine{second.A * first.A, fma(second.A, first.B, second.B)};
}

// inclusive scan of per-thread affine maps over the CTA, returns the map of all threads BEFORE this
// one (exclusive prefix) applied to a zero carry, i.e. the carry-in value of this thread.
__device__ __forceinline__ double block_affine_carry(Affine mine, double* shA, double* shB) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Affine inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double pa = __shfl_up_sync(0xffffffffu, inc.A, d);
    const double pb = __shfl_up_sync(0xffffffffu, inc.B, d);
    if (lane >= d) inc = compose(inc, Affine{pa, pb});
  }
  if (lane == 31) { shA[warp] = inc.A; shB[warp] = inc.B; }
  __syncthreads();
  if (warp == 0) {
    Affine w = (lane < TV_THREADS / 32) ? Affine{shA[lane], shB[lane]} : Affine{1.0, 0.0};
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double pa = __shfl_up_sync(0xffffffffu, w.A, d);
      const double pb = __shfl_up_sync(0xffffffffu, w.B, d);
      if (lane >= d) w = compose(w, Affine{pa, pb});
    }
    if (lane < TV_THREADS / 32) { shA[32 + lane] = w.A; shB[32 + lane] = w.B; }
  }
  __syncthreads();
  // exclusive prefix of this thread: (inclusive of previous lane) after (inclusive of previous warps)
  const double pa = __shfl_up_sync(0xffffffffu, inc.A, 1);
  const double pb = __shfl_up_sync(0xffffffffu, inc.B, 1);
  Affine excl = (lane > 0) ? Affine{pa, pb} : Affine{1.0, 0.0};
  if (warp > 0) excl = compose(excl, Affine{shA[32 + warp - 1], shB[32 + warp - 1]});
  __syncthreads();
  return excl.B;   // carry-in of the whole CTA is 0
}

template <int TV_E>
__global__ void __launch_bounds__(TV_THREADS, (TV_E == 8 ? 2 : 1)) tv_solve_kernel(TvSolveArgs a) {
  if (a.done && *a.done) return;
  constexpr int TV_SEG = TvCfg<TV_E>::SEG;
  constexpr int TV_SMEM_DOUBLES = TvCfg<TV_E>::SMEM_DOUBLES;
#define TV_PAD(i) TV_PAD<TV_E>(i)
  extern __shared__ __align__(16) double sm[];
  double* bufA = sm;                         // w = z - u, later y, later x
  double* bufB = sm + TV_SMEM_DOUBLES;       // s
  double* shA = sm + 2 * TV_SMEM_DOUBLES;    // 64 doubles
  double* shB = shA + 64;
  const int tid = threadIdx.x;
  const int64_t S = TV_SEG - 2 * a.halo;                  // outputs per CTA
  const int64_t out0 = (int64_t)blockIdx.x * S;
  const int64_t g0 = out0 - a.halo;                       // global index of local element 0 (may be < 0)
  const double rho = a.rho;

  // coalesced load of w = z - u (with one extra element on the left) and s
  for (int j = tid; j < TV_SEG + 1; j += TV_THREADS) {
    const int64_t i = g0 + j - 1;                         // local slot j holds global i = g0 + j - 1
    double w = 0.0;
    if (i >= 0 && i < a.n) w = a.z[i] - a.u[i];
    bufA[TV_PAD(j)] = w;
  }
  for (int j = tid; j < TV_SEG; j += TV_THREADS) {
    const int64_t i = g0 + j;
    bufB[TV_PAD(j)] = (i >= 0 && i < a.n) ? a.s[i] : 0.0;
  }
  __syncthreads();

  auto invd = [&](int64_t i) -> double {                  // 1/delta_i
    return i < a.ntab ? a.invdelta[i] : a.inv_star;
  };

  // ---- forward: y_i = r_i + fa_i * y_{i-1},  fa_i = rho/delta_{i-1} (0 at i = 0 and outside [0,n))
  double r[TV_E], fa[TV_E];
  const int j0 = tid * TV_E;
#pragma unroll
  for (int e = 0; e < TV_E; ++e) {
    const int j = j0 + e;
    const int64_t i = g0 + j;
    const bool in = (i >= 0 && i < a.n);
    const double w = bufA[TV_PAD(j + 1)], wl = bufA[TV_PAD(j)];   // w_i, w_{i-1} (w_{-1} = 0)
    r[e] = in ? fma(rho, w - wl, bufB[TV_PAD(j)]) : 0.0;           // s + rho*D'(z-u), totalvariation.m / :1047
    fa[e] = (in && i > 0) ? rho * invd(i - 1) : 0.0;
  }
  Affine m{1.0, 0.0};
#pragma unroll
  for (int e = 0; e < TV_E; ++e) m = Affine{fa[e] * m.A, fma(fa[e], m.B, r[e])};
  __syncthreads();   // everyone has read bufA/bufB
  double carry = block_affine_carry(m, shA, shB);
#pragma unroll
  for (int e = 0; e < TV_E; ++e) {
    carry = fma(fa[e], carry, r[e]);
    bufA[TV_PAD(j0 + e)] = carry;                        // y_i at local slot j
  }
  __syncthreads();

  // ---- backward: x_i = invd_i*y_i + (rho*invd_i) * x_{i+1}, scanned over reversed local order
  double bb[TV_E], ba[TV_E];
#pragma unroll
  for (int e = 0; e < TV_E; ++e) {
    const int j = TV_SEG - 1 - (j0 + e);
    const int64_t i = g0 + j;
    const bool in = (i >= 0 && i < a.n);
    const double id = in ? invd(i) : 0.0;
    bb[e] = id * bufA[TV_PAD(j)];
    ba[e] = (in && i < a.n - 1) ? rho * id : 0.0;
  }
  Affine mb{1.0, 0.0};
#pragma unroll
  for (int e = 0; e < TV_E; ++e) mb = Affine{ba[e] * mb.A, fma(ba[e], mb.B, bb[e])};
  __syncthreads();
  carry = block_affine_carry(mb, shA, shB);
#pragma unroll
  for (int e = 0; e < TV_E; ++e) {
    carry = fma(ba[e], carry, bb[e]);
    bufB[TV_PAD(TV_SEG - 1 - (j0 + e))] = carry;         // x_i
  }
  __syncthreads();
  for (int j = a.halo + tid; j < a.halo + S; j += TV_THREADS) {
    const int64_t i = g0 + j;
    if (i < a.n) a.x[i] = bufB[TV_PAD(j)];
  }
#undef TV_PAD
}

// ---------------------------------------------------------------------------------------------
// Exact x-update for ANY rho (the windowed kernels above need the recurrences to forget their carry within a
// halo of ~40 sqrt(rho) elements; the reference accepts any rho, getProxOps.m:1047).  The two Thomas recurrences
// are scans of affine maps, and affine maps compose exactly, so the segments are chained through their
// aggregates instead of through a halo:
//   pass 1   every CTA composes the forward maps of its 8192-element segment            -> (A_b, B_b)
//   chain    carry into segment b:  c_b = B_{b-1} + A_{b-1} c_{b-1}                       (one thread, nseg steps)
//   pass 2   the forward recurrence again, now from the true carry: y (kept) -- and the aggregate of the
//            segment's backward maps
//   chain    from the right
//   pass 3   the backward recurrence from the true carry: x
// 9 vector passes instead of 4; used only when the halo does not fit (rho >~ 2500).
// ---------------------------------------------------------------------------------------------
struct TvExactArgs {
  int64_t n;
  const double *s, *z, *u;
  double *y, *x;
  double rho;
  const double* invdelta; double inv_star; int ntab;
  double *segA, *segB;        // [nseg] aggregates of the pass being run
  const double* cin;          // [nseg] carry into each segment (passes 2 and 3)
  const int* done;
};

// exclusive prefix (as a map) of this thread's affine map over the CTA in thread order, and the CTA total
__device__ __forceinline__ Affine block_affine_prefix(Affine mine, double* shA, double* shB, Affine& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Affine inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double pa = __shfl_up_sync(0xffffffffu, inc.A, d);
    const double pb = __shfl_up_sync(0xffffffffu, inc.B, d);
    if (lane >= d) inc = compose(inc, Affine{pa, pb});
  }
  if (lane == 31) { shA[warp] = inc.A; shB[warp] = inc.B; }
  __syncthreads();
  if (warp == 0) {
    Affine w = (lane < TV_THREADS / 32) ? Affine{shA[lane], shB[lane]} : Affine{1.0, 0.0};
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double pa = __shfl_up_sync(0xffffffffu, w.A, d);
      const double pb = __shfl_up_sync(0xffffffffu, w.B, d);
      if (lane >= d) w = compose(w, Affine{pa, pb});
    }
    if (lane < TV_THREADS / 32) { shA[32 + lane] = w.A; shB[32 + lane] = w.B; }
  }
  __syncthreads();
  const double pa = __shfl_up_sync(0xffffffffu, inc.A, 1);
  const double pb = __shfl_up_sync(0xffffffffu, inc.B, 1);
  Affine excl = (lane > 0) ? Affine{pa, pb} : Affine{1.0, 0.0};
  if (warp > 0) excl = compose(excl, Affine{shA[32 + warp - 1], shB[32 + warp - 1]});
  total = Affine{shA[32 + TV_THREADS / 32 - 1], shB[32 + TV_THREADS / 32 - 1]};
  __syncthreads();
  return excl;
}

// PASS 1: forward aggregates.  PASS 2: y from the true carry + backward aggregates.  PASS 3: x from the true carry.
template <int PASS>
__global__ void __launch_bounds__(TV_THREADS, 1) tv_exact_kernel(TvExactArgs a) {
  if (a.done && *a.done) return;
  constexpr int TV_E = 16;
  constexpr int TV_SEG = TvCfg<TV_E>::SEG;
  constexpr int TV_SMEM_DOUBLES = TvCfg<TV_E>::SMEM_DOUBLES;
#define TV_PAD(i) TV_PAD<TV_E>(i)
  extern __shared__ __align__(16) double sm[];
  double* bufA = sm;
  double* bufB = sm + TV_SMEM_DOUBLES;
  double* shA = sm + 2 * TV_SMEM_DOUBLES;
  double* shB = shA + 64;
  const int tid = threadIdx.x;
  const int64_t g0 = (int64_t)blockIdx.x * TV_SEG;        // global index of local element 0
  const double rho = a.rho;
  auto invd = [&](int64_t i) -> double { return i < a.ntab ? a.invdelta[i] : a.inv_star; };
  const int j0 = tid * TV_E;
  Affine total;
  if (PASS <= 2) {
    for (int j = tid; j < TV_SEG + 1; j += TV_THREADS) {
      const int64_t i = g0 + j - 1;
      double w = 0.0;
      if (i >= 0 && i < a.n) w = a.z[i] - a.u[i];
      bufA[TV_PAD(j)] = w;
    }
    for (int j = tid; j < TV_SEG; j += TV_THREADS) {
      const int64_t i = g0 + j;
      bufB[TV_PAD(j)] = (i < a.n) ? a.s[i] : 0.0;
    }
    __syncthreads();
    double r[TV_E], fa[TV_E];
#pragma unroll
    for (int e = 0; e < TV_E; ++e) {
      const int j = j0 + e;
      const int64_t i = g0 + j;
      const bool in = (i < a.n);
      const double w = bufA[TV_PAD(j + 1)], wl = bufA[TV_PAD(j)];
      r[e] = in ? fma(rho, w - wl, bufB[TV_PAD(j)]) : 0.0;
      fa[e] = (in && i > 0) ? rho * invd(i - 1) : (in ? 0.0 : 1.0);   // past the end: identity maps, so the aggregate is the segment's
      if (!in) r[e] = 0.0;
    }
    Affine m{1.0, 0.0};
#pragma unroll
    for (int e = 0; e < TV_E; ++e) m = Affine{fa[e] * m.A, fma(fa[e], m.B, r[e])};
    __syncthreads();
    const Affine excl = block_affine_prefix(m, shA, shB, total);
    if (PASS == 1) {
      if (tid == 0) { a.segA[blockIdx.x] = total.A; a.segB[blockIdx.x] = total.B; }
      return;
    }
    double carry = fma(excl.A, a.cin[blockIdx.x], excl.B);
#pragma unroll
    for (int e = 0; e < TV_E; ++e) {
      carry = fma(fa[e], carry, r[e]);
      bufA[TV_PAD(j0 + e)] = carry;                        // y_i
    }
    __syncthreads();
    for (int j = tid; j < TV_SEG; j += TV_THREADS) {
      const int64_t i = g0 + j;
      if (i < a.n) a.y[i] = bufA[TV_PAD(j)];
    }
  } else {
    for (int j = tid; j < TV_SEG; j += TV_THREADS) {
      const int64_t i = g0 + j;
      bufA[TV_PAD(j)] = (i < a.n) ? a.y[i] : 0.0;
    }
    __syncthreads();
  }
  // ---- backward maps in reversed local order: x_i = invd_i*y_i + (rho*invd_i) x_{i+1}
  double bb[TV_E], ba[TV_E];
#pragma unroll
  for (int e = 0; e < TV_E; ++e) {
    const int j = TV_SEG - 1 - (j0 + e);
    const int64_t i = g0 + j;
    const bool in = (i < a.n);
    const double id = in ? invd(i) : 0.0;
    bb[e] = id * bufA[TV_PAD(j)];
    ba[e] = (in && i < a.n - 1) ? rho * id : (in ? 0.0 : 1.0);      // past the end: identity
  }
  Affine mb{1.0, 0.0};
#pragma unroll
  for (int e = 0; e < TV_E; ++e) mb = Affine{ba[e] * mb.A, fma(ba[e], mb.B, bb[e])};
  __syncthreads();
  const Affine exclb = block_affine_prefix(mb, shA, shB, total);
  if (PASS == 2) {
    if (tid == 0) { a.segA[blockIdx.x] = total.A; a.segB[blockIdx.x] = total.B; }
    return;
  }
  double carry = fma(exclb.A, a.cin[blockIdx.x], exclb.B);
#pragma unroll
  for (int e = 0; e < TV_E; ++e) {
    carry = fma(ba[e], carry, bb[e]);
    bufB[TV_PAD(TV_SEG - 1 - (j0 + e))] = carry;           // x_i
  }
  __syncthreads();
  for (int j = tid; j < TV_SEG; j += TV_THREADS) {
    const int64_t i = g0 + j;
    if (i < a.n) a.x[i] = bufB[TV_PAD(j)];
  }
#undef TV_PAD
}

// carry into every segment from the aggregates: forward (reverse = 0: c_0 = 0, c_b = B_{b-1} + A_{b-1} c_{b-1}) or from the
// right (reverse = 1).  One warp; lane 0 walks the chain (nseg = n / 8192 steps of one FMA).
__global__ void tv_chain_kernel(const double* __restrict__ segA, const double* __restrict__ segB, int64_t nseg, double* cin,
                                int reverse, const int* done) {
  if (done && *done) return;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double c = 0.0;
  if (!reverse) {
    for (int64_t b = 0; b < nseg; ++b) { cin[b] = c; c = fma(segA[b], c, segB[b]); }
  } else {
    for (int64_t b = nseg - 1; b >= 0; --b) { cin[b] = c; c = fma(segA[b], c, segB[b]); }
  }
}

constexpr int TVP_THREADS = 256;
constexpr int TVP_E = 4;

struct TvProxArgs {
  int64_t n;
  const double *x, *s;
  const double *z, *u;     // current iterates (read by neighbours too, hence double-buffered)
  double *znew, *unew;
  double lambda;
  double* partials;     // [gridDim.x][TVP_NRED]
  LoopCtl* ctl;
  LoopParams lp;
  double *xvals, *zvals, *uvals;
  const double *v, *uhat;  // fast / accelerated ADMM (admm.m:503-529): the prox and the u-update start from uhat
};
constexpr int TVP_NRED = 10;

// FAST: the fast / accelerated variants (admm.m:267-298, 503-600).  z and u of the previous iteration are simply the
// half of the double buffer this kernel reads, so nothing is copied; the residual norms need the NEW v, so the scalar
// epilogue moves to tv_accel_kernel and the last CTA here only fixes the predictor weight / restart.
template <bool FAST>
__global__ void __launch_bounds__(TVP_THREADS, 2) tv_prox_kernel(TvProxArgs a) {
  LoopCtl* ctl = a.ctl;
  if (ctl->done) return;
  __shared__ double sh[(TVP_THREADS / 32) * TVP_NRED];
  __shared__ bool is_last;
  const int it = ctl->it;
  const double rho = a.lp.rho, relax = a.lp.relax, thr = a.lambda / rho;
  const int64_t n = a.n;
  double r[TVP_NRED];
#pragma unroll
  for (int k = 0; k < TVP_NRED; ++k) r[k] = 0.0;

  for (int64_t base = ((int64_t)blockIdx.x * TVP_THREADS + threadIdx.x) * TVP_E; base < n;
       base += (int64_t)gridDim.x * TVP_THREADS * TVP_E) {
    // registers: x[base-1 .. base+9], z[base-1 .. base+8], u[base-1 .. base+7], s[base .. base+7]
    double xr[TVP_E + 3], zr[TVP_E + 2], ur[TVP_E + 1], sr[TVP_E];
    if (base >= 2 && base + TVP_E + 2 <= n) {          // interior: 16-byte loads (base is a multiple of 8)
      xr[0] = a.x[base - 1]; zr[0] = a.z[base - 1]; ur[0] = a.u[base - 1];
#pragma unroll
      for (int k = 0; k < TVP_E / 2 + 1; ++k) {
        const double2 v = *reinterpret_cast<const double2*>(a.x + base + 2 * k);
        xr[1 + 2 * k] = v.x; xr[2 + 2 * k] = v.y;
      }
#pragma unroll
      for (int k = 0; k < TVP_E / 2; ++k) {
        const double2 vz = *reinterpret_cast<const double2*>(a.z + base + 2 * k);
        const double2 vu = *reinterpret_cast<const double2*>(a.u + base + 2 * k);
        const double2 vs = *reinterpret_cast<const double2*>(a.s + base + 2 * k);
        zr[1 + 2 * k] = vz.x; zr[2 + 2 * k] = vz.y;
        ur[1 + 2 * k] = vu.x; ur[2 + 2 * k] = vu.y;
        sr[2 * k] = vs.x; sr[2 * k + 1] = vs.y;
      }
      zr[TVP_E + 1] = a.z[base + TVP_E];
    } else {
#pragma unroll
      for (int k = 0; k < TVP_E + 3; ++k) { const int64_t i = base - 1 + k; xr[k] = (i >= 0 && i < n) ? a.x[i] : 0.0; }
#pragma unroll
      for (int k = 0; k < TVP_E + 2; ++k) { const int64_t i = base - 1 + k; zr[k] = (i >= 0 && i < n) ? a.z[i] : 0.0; }
#pragma unroll
      for (int k = 0; k < TVP_E + 1; ++k) { const int64_t i = base - 1 + k; ur[k] = (i >= 0 && i < n) ? a.u[i] : 0.0; }
#pragma unroll
      for (int k = 0; k < TVP_E; ++k) { const int64_t i = base + k; sr[k] = (i < n) ? a.s[i] : 0.0; }
    }
    double uo[TVP_E + 1], vr[TVP_E + 1];      // FAST: u of the previous iteration (ur then holds uhat), v
    if (FAST) {
#pragma unroll
      for (int k = 0; k < TVP_E + 1; ++k) {
        const int64_t i = base - 1 + k;
        const bool in = (i >= 0 && i < n);
        uo[k] = ur[k];
        ur[k] = in ? a.uhat[i] : 0.0;
        vr[k] = in ? a.v[i] : 0.0;
      }
    }
    // element i = base - 1 + k; k = 0 is the left neighbour, recomputed only for the D' stencils of
    // the dual residual.  Values of the previous element are carried in scalars (few live registers).
    double zl = 0.0, ul = 0.0, dzl = 0.0;
#pragma unroll
    for (int k = 0; k < TVP_E + 1; ++k) {
      const int64_t i = base - 1 + k;
      const double x0 = xr[k], x1 = xr[k + 1], x2 = xr[k + 2];
      const double Dx = (i < n - 1) ? (x0 - x1) : x0;
      const double zp = zr[k], up = ur[k];
      double Axh = Dx, w;
      if (relax != 1.0) {
        // admm.m:517 Axhat = relax*A(x) - (1-relax)*(B(zprev) - c); the z-prox then applies D to the
        // vector in x's slot (getProxOps.m:199 `u + D*x` with x := Axhat, admm.m:521)
        Axh = relax * Dx - (1.0 - relax) * (-zp - 0.0);
        const double Dx1 = (i + 1 < n - 1) ? (x1 - x2) : x1;
        const double Axh1 = relax * Dx1 - (1.0 - relax) * (-zr[k + 1] - 0.0);
        w = up + ((i < n - 1) ? (Axh - Axh1) : Axh);
      } else {
        w = up + Dx;
      }
      const double zn = soft_threshold(w, thr);
      const double un = up + (Axh + (-zn) - 0.0);
      const double dz = zn - zp;
      if (k >= 1 && i < n) {
        const double du = un - (FAST ? uo[k] : up);
        const double pr = Dx + (-zn) - 0.0;
        if (FAST) {
          const double e1 = un - up, e2 = zn - vr[k];
          r[8] = fma(e1, e1, r[8]);                 // ||u - uhat||^2
          r[9] = fma(e2, e2, r[9]);                 // ||B(z - v)||^2, B = -I
        }
        const double dtdz = rho * ((i > 0) ? (dz - dzl) : dz);       // rho*At(B(z-zprev)) up to sign
        const double dtu = rho * ((i > 0) ? (un - ul) : un);         // rho*At(u)
        const double xs = x0 - sr[k - 1];
        r[0] = fma(pr, pr, r[0]);
        r[1] = fma(Dx, Dx, r[1]);
        r[2] = fma(zn, zn, r[2]);
        r[3] = fma(dtdz, dtdz, r[3]);
        r[4] = fma(dtu, dtu, r[4]);
        r[5] = fma(dz, dz, r[5]);
        r[6] = fma(du, du, r[6]);
        r[7] += 0.5 * xs * xs + ((i < n - 1) ? a.lambda * fabs(Dx) : 0.0);   // totalvariation.m objective
        // neighbouring threads read the OLD values of these elements: the new iterate goes to the
        // other half of the double buffer
        a.znew[i] = zn;
        a.unew[i] = un;
        if (a.xvals) {
          a.xvals[(int64_t)it * n + i] = x0;
          a.zvals[(int64_t)it * n + i] = zn;
          a.uvals[(int64_t)it * n + i] = un;
        }
      }
      zl = zn; ul = un; dzl = dz;
    }
    (void)zl;
  }
  block_reduce_store<TVP_NRED>(r, a.partials + (int64_t)blockIdx.x * TVP_NRED, sh);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&ctl->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < TVP_NRED) {
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(a.partials + (int64_t)b * TVP_NRED + threadIdx.x);
    sh[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ctl->ticket = 0;
    if (FAST) {
      // kept for tv_accel_kernel: ||Ax - z||^2, ||Ax||^2, ||z||^2, ||rho At(u)||^2, ||dz||^2, ||du||^2, -, objective
      ctl->sums[0] = sh[0]; ctl->sums[1] = sh[1]; ctl->sums[2] = sh[2]; ctl->sums[3] = sh[4];
      ctl->sums[4] = sh[5]; ctl->sums[5] = sh[6]; ctl->sums[6] = 0.0; ctl->sums[7] = sh[7];
      accel_decide(ctl, a.lp, sh[8], sh[9]);
      return;
    }
    double red[8] = {sh[0], sh[1], sh[2], 0.0, sh[3], sh[4], sh[5], sh[6]};
    loop_epilogue(ctl, a.lp, red, (double)n, (double)n, sh[7]);
  }
}

// Acceleration pass of the fast variants for total variation (admm.m:562-600): v = z + gamma (z - zprev),
// uhat = u + gamma (u - uprev), or v = zprev, uhat = uprev on a restart; the dual residual of the fast variant,
// rho * ||At(B(z - v))|| with At = D' (admm.m:629-633: (D'e)_i = e_i - e_{i-1}); last CTA: the scalar epilogue.
struct TvAccelArgs {
  int64_t n;
  const double *z, *u, *zprev, *uprev;    // new half / old half of the double buffers
  double *v, *uhat;
  double* partials;                       // [gridDim.x]
  LoopCtl* ctl;
  LoopParams lp;
};

__global__ void __launch_bounds__(256) tv_accel_kernel(TvAccelArgs a) {
  LoopCtl* ctl = a.ctl;
  if (ctl->done) return;
  __shared__ double sh[256 / 32];
  __shared__ bool is_last;
  const double gamma = ctl->gamma, rho = a.lp.rho;
  const int restart = ctl->restart;
  double r1[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double z = a.z[i], u = a.u[i], zp = a.zprev[i], up = a.uprev[i];
    const double v = restart ? zp : z + gamma * (z - zp);
    const double uh = restart ? up : u + gamma * (u - up);
    a.v[i] = v;
    a.uhat[i] = uh;
    double e = z - v;
    if (i > 0) {                          // the left neighbour's z - v, recomputed (its thread may sit in another CTA)
      const double zl = a.z[i - 1], zpl = a.zprev[i - 1];
      const double vl = restart ? zpl : zl + gamma * (zl - zpl);
      e -= zl - vl;
    }
    r1[0] = fma(e, e, r1[0]);
  }
  block_reduce_store<1>(r1, a.partials + blockIdx.x, sh);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&ctl->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  double sdt = 0.0;
  for (unsigned b = 0; b < gridDim.x; ++b) sdt += __ldcg(a.partials + b);
  ctl->ticket = 0;
  const double* sm = ctl->sums;
  double red[8] = {sm[0], sm[1], sm[2], 0.0, rho * rho * sdt, sm[3], sm[4], sm[5]};
  loop_epilogue(ctl, a.lp, red, (double)a.n, (double)a.n, sm[7]);
}

// ---------------------------------------------------------------------------------------------
// Fused iteration (small halos): x-update + z/u pass in ONE kernel, everything register-resident.
// A thread owns TVF_E consecutive elements for every phase: 256-bit loads of z, u, s -> forward
// affine scan (y) -> reversed affine scan (x) -> neighbour x's through 3 shared-memory slots per
// thread -> relaxed soft threshold, u-update, all norms -> 256-bit stores of z, u.  x never goes to
// memory inside the loop (the next x-update reads only s, z, u): 5 vector passes per iteration
// instead of 10; the host materialises x once after the loop (xonly = 1 on the half the last
// iteration read) or the kernel writes it into the history when options.history asks for it.
// Outputs of a segment: local [hl, SEG - hr) with hl >= halo + 1, hr >= halo + 2 (the prox stencil
// reads x_{i-1} .. x_{i+2}), both multiples of TVF_E so a thread is all output or all halo.
//
// Two bodies per kernel.  FAST (every segment whose whole window lies in 1 .. n-3 and past the
// pivot table): no predicates, the recurrence multiplier is the constant c = rho/delta*, so a
// thread's affine map is (c^E, B) and the block scans carry only B with powers of c precomputed on
// the host -- 3 barriers per segment.  SLOW (the first and last one or two segments, and every
// segment of a small problem): the general predicated code in rolled loops on local arrays.
// ---------------------------------------------------------------------------------------------
constexpr int TVF_E = 8;
constexpr int TVF_TMAX = 256;                              // CTA sizes built: 128 (4 CTAs / SM) and 256 (2 CTAs / SM)

struct TvFusedArgs {
  int64_t n, S, nseg;
  const double *s, *z, *u;
  double *znew, *unew;
  double* x;                // written only when xonly
  double rho, lambda;
  const double* invdelta;
  double inv_star;
  double c;                 // rho*inv_star
  double cp[TVF_E + 1];     // c^k, k = 0 .. E
  double pw[5];             // c^(E*2^k): the map of 2^k consecutive fast threads
  double aw;                // c^(E*32): a whole warp
  int64_t seg_lo, seg_hi;   // segments [seg_lo, seg_hi) take the predicate-free body
  int ntab, hl, hr, xonly;
  double* partials;         // [gridDim.x][8]
  LoopCtl* ctl;
  LoopParams lp;
  double *xvals, *zvals, *uvals;
};

__device__ __forceinline__ void ld256(const double* p, double* v) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, const double* v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

// Carry-in of this thread for an affine scan over the CTA in thread order (REV = false) or in
// reversed thread order (REV = true); the carry into the first thread in scan order is 0.
template <int T, bool REV>
__device__ __forceinline__ double block_affine_carry_t(Affine mine, double* shA, double* shB) {
  constexpr int W = T / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Affine inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double pa = REV ? __shfl_down_sync(0xffffffffu, inc.A, d) : __shfl_up_sync(0xffffffffu, inc.A, d);
    const double pb = REV ? __shfl_down_sync(0xffffffffu, inc.B, d) : __shfl_up_sync(0xffffffffu, inc.B, d);
    if (REV ? (lane + d < 32) : (lane >= d)) inc = compose(inc, Affine{pa, pb});
  }
  if (lane == (REV ? 0 : 31)) { shA[warp] = inc.A; shB[warp] = inc.B; }
  __syncthreads();
  if (warp == 0) {
    const int src = REV ? (W - 1 - lane) : lane;          // warps in scan order
    Affine w = (lane < W) ? Affine{shA[src], shB[src]} : Affine{1.0, 0.0};
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double pa = __shfl_up_sync(0xffffffffu, w.A, d);
      const double pb = __shfl_up_sync(0xffffffffu, w.B, d);
      if (lane >= d) w = compose(w, Affine{pa, pb});
    }
    if (lane < W) { shA[32 + src] = w.A; shB[32 + src] = w.B; }
  }
  __syncthreads();
  const double pa = REV ? __shfl_down_sync(0xffffffffu, inc.A, 1) : __shfl_up_sync(0xffffffffu, inc.A, 1);
  const double pb = REV ? __shfl_down_sync(0xffffffffu, inc.B, 1) : __shfl_up_sync(0xffffffffu, inc.B, 1);
  Affine excl = (REV ? (lane < 31) : (lane > 0)) ? Affine{pa, pb} : Affine{1.0, 0.0};
  const int prevw = REV ? warp + 1 : warp - 1;            // the warp before this one in scan order
  if (prevw >= 0 && prevw < W) excl = compose(excl, Affine{shA[32 + prevw], shB[32 + prevw]});
  __syncthreads();
  return excl.B;
}

// The same carry when every thread's map is (c^E, B): only B travels.  lanepow[k] = c^(E*k).  shW
// holds the TVF_W warp totals of THIS scan (the two scans of a segment use different arrays, so no
// trailing barrier is needed: the next write comes two barriers later).
template <int T, bool REV>
__device__ __forceinline__ double tvf_carry(double B, const TvFusedArgs& a, const double* lanepow, double* shW) {
  constexpr int TVF_W = T / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double inc = B;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int d = 1 << k;
    const double p = REV ? __shfl_down_sync(0xffffffffu, inc, d) : __shfl_up_sync(0xffffffffu, inc, d);
    if (REV ? (lane + d < 32) : (lane >= d)) inc = fma(a.pw[k], p, inc);
  }
  if (lane == (REV ? 0 : 31)) shW[warp] = inc;
  __syncthreads();
  // carry into this warp: Horner over the warps before it in scan order (at most 7 broadcast loads)
  double wc = 0.0;
  if (REV) {
    for (int v = TVF_W - 1; v > warp; --v) wc = fma(a.aw, wc, shW[v]);
  } else {
    for (int v = 0; v < warp; ++v) wc = fma(a.aw, wc, shW[v]);
  }
  const double pB = REV ? __shfl_down_sync(0xffffffffu, inc, 1) : __shfl_up_sync(0xffffffffu, inc, 1);
  const double exB = (REV ? (lane < 31) : (lane > 0)) ? pB : 0.0;
  return fma(lanepow[REV ? 31 - lane : lane], wc, exB);
}

// ---- FAST body: every element the thread touches is interior and past the pivot table
template <int T, bool RELAX1>
__device__ __forceinline__ void tvf_fast_segment(const TvFusedArgs& a, int64_t i0, bool is_out, int it, double* racc,
                                                 const double* lanepow, double* shW, double (*xe)[TVF_TMAX]) {
  constexpr int TVF_W = T / 32;
  constexpr int E = TVF_E;
  const int tid = threadIdx.x;
  const double rho = a.rho, c = a.c, ids = a.inv_star;
  double zr[E + 2], ur[E + 1], sr[E];                      // z[i0-1 .. i0+E], u[i0-1 .. i0+E-1], s[i0 .. i0+E-1]
  ld256(a.z + i0, zr + 1); ld256(a.z + i0 + 4, zr + 5);
  ld256(a.u + i0, ur + 1); ld256(a.u + i0 + 4, ur + 5);
  ld256(a.s + i0, sr); ld256(a.s + i0 + 4, sr + 4);
  zr[0] = a.z[i0 - 1]; ur[0] = a.u[i0 - 1]; zr[E + 1] = a.z[i0 + E];
  // forward: y_i = r_i + c*y_{i-1}, r = s + rho*D'(z - u)
  double r[E];
  {
    double wl = zr[0] - ur[0];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const double w = zr[e + 1] - ur[e + 1];
      r[e] = fma(rho, w - wl, sr[e]);
      wl = w;
    }
  }
  // local prefixes first (y with a zero carry-in), so that after the block scan every element is ONE
  // independent fma away: y_e = c^(e+1)*carry + prefix_e -- the dependent chain is walked once, not twice
#pragma unroll
  for (int e = 1; e < E; ++e) r[e] = fma(c, r[e - 1], r[e]);
  double carry = tvf_carry<T, false>(r[E - 1], a, lanepow, shW);
  const double ym1 = carry;                                // y_{i0-1}
#pragma unroll
  for (int e = 0; e < E; ++e) r[e] = ids * fma(a.cp[e + 1], carry, r[e]);   // y_i / delta_i
  // backward: x_i = y_i/delta + c*x_{i+1}, prefixes from the right
#pragma unroll
  for (int e = E - 2; e >= 0; --e) r[e] = fma(c, r[e + 1], r[e]);
  carry = tvf_carry<T, true>(r[0], a, lanepow, shW + TVF_W);
  double xx[E + 3];                                        // x[i0-1 .. i0+E+1]
#pragma unroll
  for (int e = 0; e < E; ++e) xx[e + 1] = fma(a.cp[E - e], carry, r[e]);
  if (a.xonly) {
    if (is_out) { st256(a.x + i0, xx + 1); st256(a.x + i0 + 4, xx + 5); }
    return;
  }
  if (RELAX1) {
    // the neighbours' x the stencil needs come out of the recurrences themselves: the carry of the
    // reversed scan IS x_{i0+E}, and x_{i0-1} = y_{i0-1}/delta + c*x_{i0} -- no exchange, no barrier
    if (!is_out) return;                                   // halo threads: nothing to produce
    xx[E + 1] = carry;
    xx[0] = fma(c, xx[1], ids * ym1);
    xx[E + 2] = 0.0;
  } else {
    xe[0][tid] = xx[1]; xe[1][tid] = xx[2]; xe[2][tid] = xx[E];
    __syncthreads();
    if (!is_out) return;
    xx[0] = xe[2][tid - 1];                                // outputs never sit in the first / last thread
    xx[E + 1] = xe[0][tid + 1];
    xx[E + 2] = xe[1][tid + 1];
  }
  const double relax = a.lp.relax, thr = a.lambda / rho, lambda = a.lambda;
  double zn_o[E], un_o[E];
  double ul = 0.0, dzl = 0.0;
#pragma unroll
  for (int k = 0; k < E + 1; ++k) {                        // element i0-1+k; k = 0 only feeds the D' stencils
    const double x0 = xx[k], x1 = xx[k + 1];
    const double Dx = x0 - x1;
    const double zp = zr[k], up = ur[k];
    double Axh = Dx, w;
    if (!RELAX1) {
      // admm.m:517 / :521 with getProxOps.m:199 applying D to the vector in x's slot
      Axh = relax * Dx - (1.0 - relax) * (-zp - 0.0);
      const double Dx1 = x1 - xx[k + 2];
      const double Axh1 = relax * Dx1 - (1.0 - relax) * (-zr[k + 1] - 0.0);
      w = up + (Axh - Axh1);
    } else {
      w = up + Dx;
    }
    const double zn = soft_threshold(w, thr);
    const double un = up + (Axh + (-zn) - 0.0);
    const double dz = zn - zp;
    if (k >= 1) {
      const int e = k - 1;
      zn_o[e] = zn; un_o[e] = un;
      const double du = un - up;
      const double pr = Dx + (-zn) - 0.0;
      const double dtdz = rho * (dz - dzl);
      const double dtu = rho * (un - ul);
      const double xs = x0 - sr[e];
      racc[0] = fma(pr, pr, racc[0]);
      racc[1] = fma(Dx, Dx, racc[1]);
      racc[2] = fma(zn, zn, racc[2]);
      racc[3] = fma(dtdz, dtdz, racc[3]);
      racc[4] = fma(dtu, dtu, racc[4]);
      racc[5] = fma(dz, dz, racc[5]);
      racc[6] = fma(du, du, racc[6]);
      racc[7] += 0.5 * xs * xs + lambda * fabs(Dx);
    }
    ul = un; dzl = dz;
  }
  st256(a.znew + i0, zn_o); st256(a.znew + i0 + 4, zn_o + 4);
  st256(a.unew + i0, un_o); st256(a.unew + i0 + 4, un_o + 4);
  if (a.xvals) {
    const int64_t o = (int64_t)it * a.n + i0;              // column `it` of the history: not 32-byte aligned in general
#pragma unroll
    for (int e = 0; e < E; ++e) { a.xvals[o + e] = xx[e + 1]; a.zvals[o + e] = zn_o[e]; a.uvals[o + e] = un_o[e]; }
  }
}

// ---- SLOW body: the general predicated iteration on local arrays (edge segments, small problems)
template <int T>
__device__ __noinline__ void tvf_slow_segment(const TvFusedArgs& a, int64_t i0, int jlo, int jhi, int it, double* red,
                                              double* shA, double* shB, double (*xe)[TVF_TMAX]) {
  constexpr int E = TVF_E;
  const int tid = threadIdx.x;
  const double rho = a.rho, relax = a.lp.relax, thr = a.lambda / rho;
  const int64_t n = a.n;
  double zr[E + 2], ur[E + 1], sr[E], idv[E + 1], r[E], fa[E], xx[E + 3];
#pragma unroll 1
  for (int k = 0; k < E + 2; ++k) { const int64_t i = i0 - 1 + k; zr[k] = (i >= 0 && i < n) ? a.z[i] : 0.0; }
#pragma unroll 1
  for (int k = 0; k < E + 1; ++k) {
    const int64_t i = i0 - 1 + k;
    ur[k] = (i >= 0 && i < n) ? a.u[i] : 0.0;
    idv[k] = (i >= 0 && i < a.ntab) ? a.invdelta[i] : a.inv_star;     // 1/delta_i
  }
#pragma unroll 1
  for (int k = 0; k < E; ++k) { const int64_t i = i0 + k; sr[k] = (i >= 0 && i < n) ? a.s[i] : 0.0; }
  // forward: y_i = r_i + fa_i*y_{i-1}; r = s + rho*D'(z-u), fa_i = rho/delta_{i-1}
  Affine m{1.0, 0.0};
#pragma unroll 1
  for (int e = 0; e < E; ++e) {
    const int64_t i = i0 + e;
    const bool in = (i >= 0 && i < n);
    const double w = zr[e + 1] - ur[e + 1], wl = zr[e] - ur[e];
    r[e] = in ? fma(rho, w - wl, sr[e]) : 0.0;
    fa[e] = (in && i > 0) ? rho * idv[e] : 0.0;
    m = Affine{fa[e] * m.A, fma(fa[e], m.B, r[e])};
  }
  double carry = block_affine_carry_t<T, false>(m, shA, shB);
  Affine mb{1.0, 0.0};
#pragma unroll 1
  for (int e = 0; e < E; ++e) {
    carry = fma(fa[e], carry, r[e]);
    r[e] = carry;                                          // y_i
  }
  // backward: x_i = y_i/delta_i + (rho/delta_i)*x_{i+1}, scanned from the right
#pragma unroll 1
  for (int e = E - 1; e >= 0; --e) {
    const int64_t i = i0 + e;
    const bool in = (i >= 0 && i < n);
    const double id = in ? idv[e + 1] : 0.0;
    r[e] = id * r[e];
    fa[e] = (in && i < n - 1) ? rho * id : 0.0;
    mb = Affine{fa[e] * mb.A, fma(fa[e], mb.B, r[e])};
  }
  carry = block_affine_carry_t<T, true>(mb, shA, shB);
#pragma unroll 1
  for (int e = E - 1; e >= 0; --e) {
    carry = fma(fa[e], carry, r[e]);
    xx[e + 1] = carry;
  }
  if (a.xonly) {
#pragma unroll 1
    for (int e = 0; e < E; ++e)
      if (e >= jlo && e < jhi && i0 + e < n) a.x[i0 + e] = xx[e + 1];
    return;
  }
  xe[0][tid] = xx[1]; xe[1][tid] = xx[2]; xe[2][tid] = xx[E];
  __syncthreads();
  xx[0] = tid > 0 ? xe[2][tid - 1] : 0.0;
  xx[E + 1] = tid < T - 1 ? xe[0][tid + 1] : 0.0;
  xx[E + 2] = tid < T - 1 ? xe[1][tid + 1] : 0.0;
  // z/u pass on elements i0-1+k; k = 0 is the left neighbour, recomputed only for the D' stencils of
  // the dual residual (same arithmetic as tv_prox_kernel)
  double ul = 0.0, dzl = 0.0;
#pragma unroll 1
  for (int k = 0; k < E + 1; ++k) {
    const int64_t i = i0 - 1 + k;
    const double x0 = xx[k], x1 = xx[k + 1], x2 = xx[k + 2];
    const double Dx = (i < n - 1) ? (x0 - x1) : x0;
    const double zp = zr[k], up = ur[k];
    double Axh = Dx, w;
    if (relax != 1.0) {
      // admm.m:517 / :521 with getProxOps.m:199 applying D to the vector in x's slot
      Axh = relax * Dx - (1.0 - relax) * (-zp - 0.0);
      const double Dx1 = (i + 1 < n - 1) ? (x1 - x2) : x1;
      const double Axh1 = relax * Dx1 - (1.0 - relax) * (-zr[k + 1] - 0.0);
      w = up + ((i < n - 1) ? (Axh - Axh1) : Axh);
    } else {
      w = up + Dx;
    }
    const double zn = soft_threshold(w, thr);
    const double un = up + (Axh + (-zn) - 0.0);
    const double dz = zn - zp;
    const int e = k - 1;
    if (k >= 1 && e >= jlo && e < jhi && i < n) {
      const double du = un - up;
      const double pr = Dx + (-zn) - 0.0;
      const double dtdz = rho * ((i > 0) ? (dz - dzl) : dz);
      const double dtu = rho * ((i > 0) ? (un - ul) : un);
      const double xs = x0 - sr[e];
      red[0] = fma(pr, pr, red[0]);
      red[1] = fma(Dx, Dx, red[1]);
      red[2] = fma(zn, zn, red[2]);
      red[3] = fma(dtdz, dtdz, red[3]);
      red[4] = fma(dtu, dtu, red[4]);
      red[5] = fma(dz, dz, red[5]);
      red[6] = fma(du, du, red[6]);
      red[7] += 0.5 * xs * xs + ((i < n - 1) ? a.lambda * fabs(Dx) : 0.0);
      a.znew[i] = zn;
      a.unew[i] = un;
      if (a.xvals) {
        a.xvals[(int64_t)it * n + i] = x0;
        a.zvals[(int64_t)it * n + i] = zn;
        a.uvals[(int64_t)it * n + i] = un;
      }
    }
    ul = un; dzl = dz;
  }
}

template <int T, bool RELAX1>
__global__ void __launch_bounds__(T, 512 / T) tv_fused_kernel(TvFusedArgs a) {
  constexpr int TVF_W = T / 32;
  LoopCtl* ctl = a.ctl;
  if (!a.xonly && ctl->done) return;
  __shared__ double shA[64], shB[64];                      // slow-body scans
  __shared__ double shW[2 * TVF_W];                        // fast-body scans: warp totals, forward / reversed
  __shared__ double lanepow[32];                           // c^(E*k)
  __shared__ double xe[3][TVF_TMAX];                          // first, second and last x of every thread
  __shared__ double red_sh[(T / 32) * 8];
  __shared__ bool is_last;
  const int tid = threadIdx.x;
  const int it = ctl->it;
  const int64_t n = a.n;
  if (tid < 32) {
    double p = 1.0;
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if ((tid >> k) & 1) p *= a.pw[k];
    lanepow[tid] = p;
  }
  __syncthreads();
  double racc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) racc[k] = 0.0;
  const int j0 = tid * TVF_E;
  const int jlo = a.hl - j0, jhi = T * TVF_E - a.hr - j0;    // outputs of this thread's chunk: e in [jlo, jhi)
  const bool is_out = (jlo <= 0 && jhi >= TVF_E);          // hl, hr are multiples of E: all or nothing
  // CTA-uniform: the segment's window lies inside 1 .. n-3 and past the pivot table (host-computed range)
  for (int64_t seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const int64_t i0 = seg * a.S - a.hl + j0;              // global index of this thread's first element
    if (seg >= a.seg_lo && seg < a.seg_hi) {
      tvf_fast_segment<T, RELAX1>(a, i0, is_out, it, racc, lanepow, shW, xe);
    } else {
      // the callee's sums come back through memory; racc itself must stay in registers
      // (and the callee gets its own copy of the arguments: handing it a reference to the kernel
      // parameters makes the compiler read them from a stack copy everywhere, fast body included)
      double red[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      const TvFusedArgs la = a;
      tvf_slow_segment<T>(la, i0, jlo, jhi, it, red, shA, shB, xe);
#pragma unroll
      for (int k = 0; k < 8; ++k) racc[k] += red[k];
    }
  }
  if (a.xonly) return;
  block_reduce_store<8>(racc, a.partials + (int64_t)blockIdx.x * 8, red_sh);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&ctl->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(a.partials + (int64_t)b * 8 + threadIdx.x);
    red_sh[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double red[8] = {red_sh[0], red_sh[1], red_sh[2], 0.0, red_sh[3], red_sh[4], red_sh[5], red_sh[6]};
    ctl->ticket = 0;
    loop_epilogue(ctl, a.lp, red, (double)n, (double)n, red_sh[7]);
  }
}

}  // namespace admmb200
