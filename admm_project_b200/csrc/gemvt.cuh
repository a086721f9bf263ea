// gemvt.cuh -- out_k = scale * D' * v_k (+ addscale * addend), k < NV, for a RECTANGULAR column-major D:
// the transposed products of the path (lasso.m:160 D'*s; getProxOps.m:1204,1514; unwrappedadmm.m:116-121,
// 133; admm.m:624 At(B(z-zprev)), :654 At(u)).  HBM-bound, one pass over D for NV = 1 or 3 vectors.
//
// Work item = 256 rows x 4 columns (8 KB of D, 16 x LDG.128 in flight per lane): the vector segment is
// loaded once and reused for the 4 columns (and D for the NV vectors), so L1/L2 traffic for v is
// NV/4 of the matrix traffic instead of NV times with one column per item.  Tall matrices are cut into
// P row panels so there are enough (panel, column-group) units for 148 CTAs; a unit is owned by one
// CTA, its items go to warps in a static round-robin, every warp keeps per-lane running sums while
// it stays inside a unit and reduces (shuffles) only when it leaves it; warp results are parked in
// (unit, column, vector, warp) slots and combined in a fixed order; panel results are summed in
// panel order by panel_reduce_kernel.  Bitwise reproducible, independent of scheduling.
#pragma once
#include "common.cuh"

namespace admmb200 {

constexpr int GEMVT_THREADS = 512;
constexpr int GEMVT_WARPS = GEMVT_THREADS / 32;
constexpr int GEMVT_CG = 4;        // columns per item
constexpr int GEMVT_ROWS = 256;    // rows per item

struct GemvtArgs {
  const double* M; int64_t ld;     // ld even, M 16-byte aligned
  int64_t rows, cols;
  const double* v[3];
  double* out[3];                  // P == 1: the result (cols); P > 1: workspace [P][cols] per vector
  // NV > 3 (class batches): vector k at vbase + k*vstride, output k at obase + k*ostride
  const double* vbase; int64_t vstride;
  double* obase; int64_t ostride;
  int P; int64_t per;              // row panels of `per` rows (multiple of GEMVT_ROWS)
  int64_t ngroups;                 // ceil(cols / GEMVT_CG)
  int64_t units_per_cta;           // ceil(P * ngroups / gridDim.x)
  double scale; const double* addend; double addscale;   // addend only with P == 1
  const int* done;
};

template <int NV, int THREADS = GEMVT_THREADS>
__global__ void __launch_bounds__(THREADS, 1) gemvt_kernel(GemvtArgs a) {
  constexpr int GEMVT_WARPS = THREADS / 32;     // shadows the namespace constant: slots are sized per launch
  if (a.done && *a.done) return;
  extern __shared__ __align__(16) double slots[];   // [units_of_cta][GEMVT_CG][NV][GEMVT_WARPS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nunits = (int64_t)a.P * a.ngroups;
  const int64_t u0 = (int64_t)blockIdx.x * a.units_per_cta;
  const int64_t u1 = min(nunits, u0 + a.units_per_cta);
  if (u1 <= u0) return;
  const int nu = (int)(u1 - u0);
  for (int i = tid; i < nu * GEMVT_CG * NV * GEMVT_WARPS; i += THREADS) slots[i] = 0.0;
  __syncthreads();
  const int64_t ipu = a.per / GEMVT_ROWS;           // items per unit (the last panel may have empty ones)
  const int64_t nitems = (int64_t)nu * ipu;

  double acc[GEMVT_CG][NV];
#pragma unroll
  for (int c = 0; c < GEMVT_CG; ++c)
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[c][k] = 0.0;
  int cur = -1;
  auto flush = [&]() {
    if (cur < 0) return;
#pragma unroll
    for (int c = 0; c < GEMVT_CG; ++c)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const double s = warp_sum(acc[c][k]);
        if (lane == 0) slots[((cur * GEMVT_CG + c) * NV + k) * GEMVT_WARPS + warp] = s;
        acc[c][k] = 0.0;
      }
  };
  for (int64_t item = warp; item < nitems; item += GEMVT_WARPS) {
    const int ul = (int)(item / ipu);
    const int64_t seg = item - (int64_t)ul * ipu;
    const int64_t unit = u0 + ul;
    const int64_t p = unit / a.ngroups, g = unit - p * a.ngroups;
    const int64_t r0 = p * a.per + seg * GEMVT_ROWS;
    const int64_t rend = min(a.rows, (p + 1) * a.per);
    if (ul != cur) { flush(); cur = ul; }
    if (r0 >= rend) continue;
    const int nrows = (int)min((int64_t)GEMVT_ROWS, rend - r0);
    const int64_t c0 = g * GEMVT_CG;
    const int ncols = (int)min((int64_t)GEMVT_CG, a.cols - c0);
    const int nvec = nrows >> 1;
    const double* base = a.M + c0 * a.ld + r0;
    double2 mv[GEMVT_CG][4];
#pragma unroll
    for (int c = 0; c < GEMVT_CG; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i;
        mv[c][i] = make_double2(0.0, 0.0);
        if (c < ncols && idx < nvec) mv[c][i] = ldg_stream2(base + (int64_t)c * a.ld + 2 * idx);
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        double2 vv[NV];                       // all NV vector loads of this row pair in flight together
#pragma unroll
        for (int k = 0; k < NV; ++k)
          vv[k] = ldg_nc2((NV > 3 ? a.vbase + (int64_t)k * a.vstride : a.v[k < 3 ? k : 0]) + r0 + 2 * idx);
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
          for (int c = 0; c < GEMVT_CG; ++c) acc[c][k] = fma(mv[c][i].y, vv[k].y, fma(mv[c][i].x, vv[k].x, acc[c][k]));
      }
    }
    if ((nrows & 1) && lane == 31) {   // odd tail row (only at the very end of an odd-length matrix)
      const int64_t r = r0 + nrows - 1;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const double vv = ldg_nc1((NV > 3 ? a.vbase + (int64_t)k * a.vstride : a.v[k < 3 ? k : 0]) + r);
#pragma unroll
        for (int c = 0; c < GEMVT_CG; ++c)
          if (c < ncols) acc[c][k] = fma(ldg_stream1(a.M + (c0 + c) * a.ld + r), vv, acc[c][k]);
      }
    }
  }
  flush();
  __syncthreads();
  for (int i = tid; i < nu * GEMVT_CG * NV; i += THREADS) {
    const int k = i % NV, c = (i / NV) % GEMVT_CG, ul = i / (NV * GEMVT_CG);
    const int64_t unit = u0 + ul;
    const int64_t p = unit / a.ngroups, g = unit - p * a.ngroups;
    const int64_t col = g * GEMVT_CG + c;
    if (col >= a.cols) continue;
    const double* sl = slots + (int64_t)i * GEMVT_WARPS;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < GEMVT_WARPS; ++w) s += sl[w];
    double* o = (NV > 3) ? a.obase + (int64_t)k * a.ostride : ((k == 0) ? a.out[0] : (k == 1 ? a.out[1] : a.out[2]));
    if (a.P == 1) {
      s *= a.scale;
      if (a.addend) s += a.addscale * a.addend[col];
      o[col] = s;
    } else {
      o[p * a.cols + col] = s;
    }
  }
}

}  // namespace admmb200
