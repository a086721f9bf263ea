// tri.cuh -- the per-iteration x-update kernels: streaming products with a cached triangular
// factor (getProxOps.m:1200,1204 `U \ (L \ y)`; :1514 `Rt \ (R \ .)`; unwrappedadmm.m:139 `W \ d`).
//
// coldot_kernel computes out[j] = sum_{r in rows(j)} M[r + j*ld] * v[r] for every column j, where
// rows(j) is [0,m) (FULL: a GEMV with the transpose), [j,n) (LOWER) or [0,j] (UPPER).
// HBM-bound: every matrix byte is read exactly once with 16-byte no-L1-allocate loads, v stays in
// L1/L2.  Work is split statically between CTAs by equal AREA (not equal columns), inside a CTA
// into (column, 512-row segment) items handed to warps; segment partials are combined in a
// fixed order, so the result is bitwise reproducible and independent of scheduling.
#pragma once
#include "common.cuh"

namespace admmb200 {

enum { COLDOT_FULL = 0, COLDOT_LOWER = 1, COLDOT_UPPER = 2 };
constexpr int COLDOT_SEG = 512;
constexpr int COLDOT_THREADS = 512;

struct ColdotArgs {
  const double* M; int64_t ld;    // ld must be even and M 16-byte aligned
  int64_t rows, cols;             // FULL: rows x cols; LOWER/UPPER: rows == cols == n
  int mode;
  const double* v;                // length rows, 16-byte aligned
  double* out;                    // length cols
  double scale;                   // out = scale * dot (+ addend[j] * addscale if addend)
  const double* addend; double addscale;
  const int* done;                // device stop flag (may be NULL): kernel exits when *done != 0
  int max_cols_per_cta, max_items_per_cta;
};

__device__ __forceinline__ int64_t coldot_len(const ColdotArgs& a, int64_t j) {
  return a.mode == COLDOT_FULL ? a.rows : (a.mode == COLDOT_LOWER ? a.rows - j : j + 1);
}
// elements in columns [0, j)
__device__ __forceinline__ int64_t coldot_area(const ColdotArgs& a, int64_t j) {
  if (a.mode == COLDOT_FULL) return j * a.rows;
  if (a.mode == COLDOT_LOWER) return j * a.rows - j * (j - 1) / 2;
  return j * (j + 1) / 2;
}

__global__ void __launch_bounds__(COLDOT_THREADS, 2) coldot_kernel(ColdotArgs a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(16) unsigned char smraw[];
  int* seg_prefix = reinterpret_cast<int*>(smraw);                                  // [max_cols+1]
  double* partial = reinterpret_cast<double*>(smraw + (((a.max_cols_per_cta + 1) * 4 + 15) & ~15));
  __shared__ int64_t s_c0, s_c1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  if (tid == 0) {
    // equal-area column range of this CTA: smallest c with area(c) >= b * total / G
    const int64_t total = coldot_area(a, a.cols);
    int64_t bounds[2];
    for (int e = 0; e < 2; ++e) {
      int64_t b = blockIdx.x + e;
      int64_t target = (int64_t)(((__int128)total * b) / gridDim.x);
      int64_t lo = 0, hi = a.cols;
      if (b >= gridDim.x) lo = a.cols;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (coldot_area(a, mid) >= target) hi = mid; else lo = mid + 1;
      }
      bounds[e] = lo;
    }
    s_c0 = bounds[0];
    s_c1 = bounds[1];
  }
  __syncthreads();
  const int64_t c0 = s_c0, c1 = s_c1;
  const int ncols = (int)(c1 - c0);
  if (ncols <= 0) return;

  // prefix of segment counts (serial scan by one warp is enough: ncols is a few hundred)
  if (warp == 0) {
    int run = 0;
    for (int base = 0; base < ncols; base += 32) {
      int c = base + lane;
      int cnt = 0;
      if (c < ncols) cnt = (int)((coldot_len(a, c0 + c) + COLDOT_SEG - 1) / COLDOT_SEG);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int nb = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nb;
      }
      if (c < ncols) seg_prefix[c] = run + incl - cnt;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) seg_prefix[ncols] = run;
  }
  __syncthreads();
  const int nitems = seg_prefix[ncols];

  for (int item = warp; item < nitems; item += nwarps) {
    // column of this item: last c with seg_prefix[c] <= item
    int lo = 0, hi = ncols - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (seg_prefix[mid] <= item) lo = mid; else hi = mid - 1;
    }
    const int64_t j = c0 + lo;
    const int seg = item - seg_prefix[lo];
    const int64_t rlo = (a.mode == COLDOT_LOWER) ? j : 0;
    const int64_t rhi = (a.mode == COLDOT_UPPER) ? j + 1 : a.rows;
    int64_t r0 = rlo + (int64_t)seg * COLDOT_SEG;
    const int64_t r1 = min(rhi, r0 + COLDOT_SEG);
    const double* col = a.M + j * a.ld;
    double s0 = 0.0, s1 = 0.0;
    if (r0 & 1) {  // ld even and bases 16B-aligned: address parity == row parity
      if (lane == 0) s0 = ldg_stream1(col + r0) * __ldg(a.v + r0);
      r0 += 1;
    }
    const int nvec = (int)((r1 - r0) >> 1);
    double2 mv[8], vv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int idx = lane + 32 * i;
      if (idx < nvec) {
        mv[i] = ldg_stream2(col + r0 + 2 * idx);
        vv[i] = __ldg(reinterpret_cast<const double2*>(a.v + r0) + idx);
      } else {
        mv[i] = make_double2(0.0, 0.0);
        vv[i] = make_double2(0.0, 0.0);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s0 = fma(mv[i].x, vv[i].x, s0);
      s1 = fma(mv[i].y, vv[i].y, s1);
    }
    if (((r1 - r0) & 1) && lane == 31) {
      int64_t r = r1 - 1;
      s1 = fma(ldg_stream1(col + r), __ldg(a.v + r), s1);
    }
    double s = warp_sum(s0 + s1);
    if (lane == 0) partial[item] = s;
  }
  __syncthreads();
  for (int c = tid; c < ncols; c += blockDim.x) {
    double s = 0.0;
    for (int it = seg_prefix[c]; it < seg_prefix[c + 1]; ++it) s += partial[it];
    s *= a.scale;
    if (a.addend) s += a.addscale * a.addend[c0 + c];
    a.out[c0 + c] = s;
  }
}

}  // namespace admmb200
