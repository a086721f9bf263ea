// tri.cuh -- the per-iteration x-update kernels: streaming products with a cached triangular
// factor (getProxOps.m:1200,1204 `U \ (L \ y)`; :1514 `Rt \ (R \ .)`; unwrappedadmm.m:139 `W \ d`).
//
// coldot_kernel computes out[j] = sum_{r in rows(j)} M[r + j*ld] * v[r] for every column j, where
// rows(j) is [0,m) (FULL: a GEMV with the transpose), [j,n) (LOWER) or [0,j] (UPPER).
// HBM-bound: every matrix byte is read exactly once with 16-byte streaming loads (16 of them in
// flight per lane), v stays in L1/L2.  The host builds a PLAN once per shape: columns are split
// between CTAs by equal AREA (not equal count) and every column into items of <= 1024 rows; a warp
// takes one item at a time, item partials are combined per column in a fixed order, so the result
// is bitwise reproducible and independent of scheduling.
#pragma once
#include "common.cuh"

namespace admmb200 {

enum { COLDOT_FULL = 0, COLDOT_LOWER = 1, COLDOT_UPPER = 2 };
constexpr int COLDOT_ITEM = 1024;      // rows per item (16 x LDG.128 per lane)
constexpr int COLDOT_THREADS = 512;

struct ColdotItem {
  int col;      // column index
  int row0;     // first row of the item
  int nrows;    // 1..COLDOT_ITEM
  int pad;
};

struct ColdotArgs {
  const double* M; int64_t ld;    // ld must be even and M 16-byte aligned
  const double* v;                // 16-byte aligned
  double* out;
  double scale;                   // out = scale * dot (+ addscale * addend[j])
  const double* addend; double addscale;
  const int* done;                // device stop flag (may be NULL): kernel exits when *done != 0
  const int* cta_col;             // [grid+1] column range of each CTA
  const int* col_item;            // [cols+1] item range of each column
  const ColdotItem* items;
};

__global__ void __launch_bounds__(COLDOT_THREADS, 1) coldot_kernel(ColdotArgs a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(16) double partial[];   // one slot per item of this CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int c0 = a.cta_col[blockIdx.x], c1 = a.cta_col[blockIdx.x + 1];
  if (c1 <= c0) return;
  const int it0 = a.col_item[c0], it1 = a.col_item[c1];

  for (int item = it0 + warp; item < it1; item += nwarps) {
    const ColdotItem d = a.items[item];
    const double* col = a.M + (int64_t)d.col * a.ld;
    int r0 = d.row0;
    const int r1 = d.row0 + d.nrows;
    double s0 = 0.0, s1 = 0.0;
    if (r0 & 1) {  // ld even and bases 16B-aligned: address parity == row parity
      if (lane == 0) s0 = __ldcs(col + r0) * __ldg(a.v + r0);
      r0 += 1;
    }
    const int nvec = (r1 - r0) >> 1;
    const double2* mp = reinterpret_cast<const double2*>(col + r0);
    const double2* vp = reinterpret_cast<const double2*>(a.v + r0);
    double2 mv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int idx = lane + 32 * i;
      mv[i] = (idx < nvec) ? __ldcs(mp + idx) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int idx = lane + 32 * i;
      if (idx < nvec) {
        double2 vv = __ldg(vp + idx);
        s0 = fma(mv[i].x, vv.x, s0);
        s1 = fma(mv[i].y, vv.y, s1);
      }
    }
    if (((r1 - r0) & 1) && lane == 31) {
      int r = r1 - 1;
      s1 = fma(__ldcs(col + r), __ldg(a.v + r), s1);
    }
    double s = warp_sum(s0 + s1);
    if (lane == 0) partial[item - it0] = s;
  }
  __syncthreads();
  for (int c = c0 + tid; c < c1; c += blockDim.x) {
    double s = 0.0;
    for (int it = a.col_item[c]; it < a.col_item[c + 1]; ++it) s += partial[it - it0];
    s *= a.scale;
    if (a.addend) s += a.addscale * a.addend[c];
    a.out[c] = s;
  }
}

}  // namespace admmb200
