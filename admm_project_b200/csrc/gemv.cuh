// gemv.cuh -- out = alpha * D * v (+ beta * w), D m x n column-major: the `D*x`, `A(x)` products of
// getProxOps.m:1088,1128,911,810, admm.m:535 and `D*y` of xminLASSO's fat branch (:1204).
// HBM-bound: each thread owns two consecutive rows (16-byte loads, a warp covers 512 contiguous
// bytes of every column), columns are split into chunks across blockIdx.y so small-m problems
// still fill 148 SMs; chunk partials are summed in a fixed order by the last CTA of a row block.
#pragma once
#include "common.cuh"

namespace admmb200 {

constexpr int GEMVN_THREADS = 256;
constexpr int GEMVN_ROWS = 2 * GEMVN_THREADS;

struct GemvNArgs {
  const double* D; int64_t ld, m, n;
  const double* v;            // length n
  double* out;                // length m
  double alpha, beta; const double* w;
  int64_t cols_per_chunk;
  double* ws;                 // [gridDim.y][m] chunk partials (unused when gridDim.y == 1)
  unsigned* tickets;          // [gridDim.x], zero on entry, left zero on exit
  const int* done;
};

template <int VEC>
__global__ void __launch_bounds__(GEMVN_THREADS) gemvn_kernel(GemvNArgs a) {
  if (a.done && *a.done) return;
  const int tid = threadIdx.x;
  const int64_t r = (int64_t)blockIdx.x * GEMVN_ROWS + 2 * tid;
  const int64_t cbeg = (int64_t)blockIdx.y * a.cols_per_chunk;
  const int64_t cend = min(a.n, cbeg + a.cols_per_chunk);
  double s0 = 0.0, s1 = 0.0;
  if (r < a.m) {
    const bool two = (r + 1 < a.m);
    const double* p = a.D + r + cbeg * a.ld;
    int64_t c = cbeg;
    if (VEC == 2 && two) {
      for (; c + 8 <= cend; c += 8) {
        double2 d[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = ldg_stream2(p + (int64_t)k * a.ld);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          double vv = ldg_nc1(a.v + c + k);
          s0 = fma(d[k].x, vv, s0);
          s1 = fma(d[k].y, vv, s1);
        }
        p += 8 * a.ld;
      }
      for (; c < cend; ++c) {
        double2 d = ldg_stream2(p);
        double vv = ldg_nc1(a.v + c);
        s0 = fma(d.x, vv, s0);
        s1 = fma(d.y, vv, s1);
        p += a.ld;
      }
    } else {
      for (; c < cend; ++c) {
        double vv = ldg_nc1(a.v + c);
        s0 = fma(ldg_stream1(p), vv, s0);
        if (two) s1 = fma(ldg_stream1(p + 1), vv, s1);
        p += a.ld;
      }
    }
  }
  if (gridDim.y == 1) {
    if (r < a.m) {
      a.out[r] = a.alpha * s0 + (a.w ? a.beta * a.w[r] : 0.0);
      if (r + 1 < a.m) a.out[r + 1] = a.alpha * s1 + (a.w ? a.beta * a.w[r + 1] : 0.0);
    }
    return;
  }
  if (r < a.m) {
    a.ws[(int64_t)blockIdx.y * a.m + r] = s0;
    if (r + 1 < a.m) a.ws[(int64_t)blockIdx.y * a.m + r + 1] = s1;
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned t = atomicAdd(&a.tickets[blockIdx.x], 1u);
    is_last = (t == gridDim.y - 1);
    if (is_last) a.tickets[blockIdx.x] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    int64_t rr = r + e;
    if (rr < a.m) {
      double s = 0.0;
      for (unsigned k = 0; k < gridDim.y; ++k) s += __ldcg(a.ws + (int64_t)k * a.m + rr);
      a.out[rr] = a.alpha * s + (a.w ? a.beta * a.w[rr] : 0.0);
    }
  }
}

}  // namespace admmb200
