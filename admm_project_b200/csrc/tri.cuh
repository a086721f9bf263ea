// tri.cuh -- streaming "column dot" products: the per-iteration x-update with a cached triangular
// factor (getProxOps.m:1200,1204 `U \ (L \ y)`; :1514 `Rt \ (R \ .)`; unwrappedadmm.m:139 `W \ d`)
// and every transposed product D'*v on the path (lasso.m:160, getProxOps.m:1204,1514,
// unwrappedadmm.m:116-121,133, admm.m:624,654).
//
// coldot_kernel<NV> computes out_k[j] = sum_{r in rows(j)} M[r + j*ld] * v_k[r], k < NV, for every
// column j, where rows(j) is [0,m) (FULL), [j,n) (LOWER) or [0,j] (UPPER).  NV = 3 shares one pass
// over D between the three transposed products an iteration of the A = D problems needs.
// HBM-bound: every matrix byte is read exactly once with 16-byte streaming loads (16 in flight per
// lane), the vectors stay in L1/L2.  The host builds a PLAN once per shape:
//   - tall FULL matrices are cut into P row panels (a "virtual column" = one column of one panel,
//     panel results are summed in a fixed order by panel_reduce_kernel);
//   - virtual columns are visited in a folded order (longest, shortest, 2nd longest, ...) so every
//     CTA gets the same mix of long and short columns, and split between CTAs by equal cost;
//   - every virtual column is cut into items of <= 1024 rows; warp w of a CTA takes items
//     w, w+16, ... (a static map), keeps a running sum per column and parks it in a
//     (column, warp) slot; slots are combined in a fixed order.
// The result is therefore bitwise reproducible and independent of scheduling.
#pragma once
#include "common.cuh"

namespace admmb200 {

enum { COLDOT_FULL = 0, COLDOT_LOWER = 1, COLDOT_UPPER = 2 };
constexpr int COLDOT_ITEM = 1024;      // rows per item (16 x LDG.128 per lane)
constexpr int COLDOT_THREADS = 640;
constexpr int COLDOT_WARPS = COLDOT_THREADS / 32;

struct ColdotItem {
  int col;      // matrix column
  int row0;     // first row of the item
  int nrows;    // 1..COLDOT_ITEM
  int pos;      // position of its virtual column in the visiting order
};

struct ColdotArgs {
  const double* M; int64_t ld;    // ld must be even and M 16-byte aligned
  const double* v[3];             // 16-byte aligned
  double* out[3];                 // length = number of virtual columns (panel-major when P > 1)
  double scale;                   // out = scale * dot (+ addscale * addend[j]); only with P == 1
  const double* addend; double addscale;
  const int* done;                // device stop flag (may be NULL): kernel exits when *done != 0
  const int* cta_pos;             // [grid+1] range of positions of each CTA
  const int* pos_item;            // [nvcols+1] item range of each position
  const int* order;               // [nvcols] position -> output index (virtual column)
  const ColdotItem* items;
  int dbg_rows, dbg_cols, dbg_nitems, dbg_npos;   // bounds for the self-check
  int* dbg;                                        // [8] first violation
};

template <int NV>
__global__ void __launch_bounds__(COLDOT_THREADS, 1) coldot_kernel(ColdotArgs a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(16) double slots[];   // [npos_of_cta][NV][COLDOT_WARPS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = a.cta_pos[blockIdx.x], p1 = a.cta_pos[blockIdx.x + 1];
  if (p1 <= p0) return;
  const int it0 = a.pos_item[p0], it1 = a.pos_item[p1];
  if (a.dbg && (p0 < 0 || p1 > a.dbg_npos || it0 < 0 || it1 > a.dbg_nitems || it1 < it0)) {
    if (tid == 0 && atomicCAS(a.dbg, 0, 2) == 0) { a.dbg[1] = blockIdx.x; a.dbg[2] = p0; a.dbg[3] = p1; a.dbg[4] = it0; a.dbg[5] = it1; }
    return;
  }
  for (int i = tid; i < (p1 - p0) * NV * COLDOT_WARPS; i += COLDOT_THREADS) slots[i] = 0.0;
  __syncthreads();

  int item = it0 + warp;
  ColdotItem dnext = (item < it1) ? a.items[item] : ColdotItem{0, 0, 0, 0};
  int cur_pos = -1;
  double run[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) run[k] = 0.0;
  for (; item < it1; item += COLDOT_WARPS) {
    const ColdotItem d = dnext;
    if (item + COLDOT_WARPS < it1) dnext = a.items[item + COLDOT_WARPS];   // next descriptor, off the critical path
    if (d.pos != cur_pos) {
      if (cur_pos >= 0 && lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) slots[((cur_pos - p0) * NV + k) * COLDOT_WARPS + warp] = run[k];
      }
      cur_pos = d.pos;
#pragma unroll
      for (int k = 0; k < NV; ++k) run[k] = 0.0;
    }
    if (a.dbg && (d.col < 0 || d.col >= a.dbg_cols || d.row0 < 0 || d.nrows < 1 || d.nrows > COLDOT_ITEM ||
                  d.row0 + d.nrows > a.dbg_rows || d.pos < p0 || d.pos >= p1 || item >= a.dbg_nitems)) {
      if (lane == 0 && atomicCAS(a.dbg, 0, 1) == 0) {
        a.dbg[1] = item; a.dbg[2] = d.col; a.dbg[3] = d.row0; a.dbg[4] = d.nrows; a.dbg[5] = d.pos; a.dbg[6] = p0; a.dbg[7] = p1;
      }
      continue;
    }
    const double* col = a.M + (int64_t)d.col * a.ld;
    int r0 = d.row0;
    const int r1 = d.row0 + d.nrows;
    double s0[NV], s1[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) s0[k] = s1[k] = 0.0;
    if (r0 & 1) {  // ld even and bases 16B-aligned: address parity == row parity
      if (lane == 0) {
        const double mval = ldg_stream1(col + r0);
#pragma unroll
        for (int k = 0; k < NV; ++k) s0[k] = mval * ldg_nc1(a.v[k] + r0);
      }
      r0 += 1;
    }
    const int nvec = (r1 - r0) >> 1;
    if (a.dbg) {
      const long long lastm = (long long)d.col * a.ld + r0 + 2LL * nvec;   // one past the last vector element
      if (lastm > (long long)a.ld * a.dbg_cols || r0 + 2 * nvec > a.dbg_rows || ((uintptr_t)(col + r0) & 15) ||
          ((uintptr_t)(a.v[0] + r0) & 15)) {
        if (lane == 0 && atomicCAS(a.dbg, 0, 3) == 0) {
          a.dbg[1] = item; a.dbg[2] = d.col; a.dbg[3] = r0; a.dbg[4] = nvec; a.dbg[5] = (int)(lastm >> 20); a.dbg[6] = (int)((uintptr_t)(col + r0) & 15); a.dbg[7] = (int)((uintptr_t)(a.v[0] + r0) & 15);
        }
        continue;
      }
    }
    const double2* mp = reinterpret_cast<const double2*>(col + r0);
    // NB: the loads must be `asm volatile` (ldg_stream2): __ldcs is a NON-volatile asm in the CUDA
    // headers, the compiler may hoist it out of the (idx < nvec) guard, and the speculated load then
    // runs up to 8 KB past the end of a short column -- past the end of the allocation for the last
    // columns of W (seen on B200 as a layout-dependent cudaErrorIllegalAddress).
    double2 mv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int idx = lane + 32 * i;
      mv[i] = make_double2(0.0, 0.0);
      if (idx < nvec) mv[i] = ldg_stream2(reinterpret_cast<const double*>(mp + idx));
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double2* vp = reinterpret_cast<const double2*>(a.v[k] + r0);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        int idx = lane + 32 * i;
        if (idx < nvec) {
          double2 vv = ldg_nc2(reinterpret_cast<const double*>(vp + idx));
          s0[k] = fma(mv[i].x, vv.x, s0[k]);
          s1[k] = fma(mv[i].y, vv.y, s1[k]);
        }
      }
    }
    if (((r1 - r0) & 1) && lane == 31) {
      int r = r1 - 1;
      const double mval = ldg_stream1(col + r);
#pragma unroll
      for (int k = 0; k < NV; ++k) s1[k] = fma(mval, ldg_nc1(a.v[k] + r), s1[k]);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) run[k] += warp_sum(s0[k] + s1[k]);
  }
  if (cur_pos >= 0 && lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) slots[((cur_pos - p0) * NV + k) * COLDOT_WARPS + warp] = run[k];
  }
  __syncthreads();
  for (int i = tid; i < (p1 - p0) * NV; i += COLDOT_THREADS) {
    const int pos = p0 + i / NV, k = i % NV;
    const double* sl = slots + (int64_t)i * COLDOT_WARPS;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < COLDOT_WARPS; ++w) s += sl[w];
    s *= a.scale;
    const int c = a.order[pos];
    if (a.addend) s += a.addscale * a.addend[c];
    double* o = (k == 0) ? a.out[0] : (k == 1 ? a.out[1] : a.out[2]);
    o[c] = s;
  }
}

// y = Wkk * y (trans = 0) or Wkk' * y (trans = 1) in place, Wkk an nb x nb (<= 128) lower-triangular
// block (the inverted diagonal block of the factor): the diagonal step of the blocked substitution
// (ADMM_B200_XSOLVE_SUBST).  One CTA of 128 threads.
__global__ void __launch_bounds__(128) tri_block_mv_kernel(const double* __restrict__ Wkk, int64_t ld, int nb, double* y,
                                                           int trans, const int* done) {
  if (done && *done) return;
  __shared__ double v[128];
  const int i = threadIdx.x;
  if (i < nb) v[i] = y[i];
  __syncthreads();
  if (i >= nb) return;
  double s = 0.0;
  if (!trans) {
    for (int j = 0; j <= i; ++j) s = fma(Wkk[i + (int64_t)j * ld], v[j], s);        // row i of Wkk
  } else {
    for (int j = i; j < nb; ++j) s = fma(Wkk[j + (int64_t)i * ld], v[j], s);        // column i of Wkk
  }
  y[i] = s;
}

// dst = src unless the loop has ended (a plain memcpy would keep running after `done` and clobber x)
__global__ void vec_copy_kernel(double* dst, const double* src, int64_t n, const int* done) {
  if (done && *done) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// out_k[j] = scale * sum_p ws_k[p*cols + j]  (fixed order), k < nv
struct PanelReduceArgs {
  const double* ws[3];
  double* out[3];
  int nv, panels;
  int64_t cols;
  double scale;
  const int* done;
};
__global__ void panel_reduce_kernel(PanelReduceArgs a) {
  if (a.done && *a.done) return;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.cols) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < a.nv) {
      double s = 0.0;
      for (int p = 0; p < a.panels; ++p) s += a.ws[k][(int64_t)p * a.cols + j];
      a.out[k][j] = a.scale * s;
    }
  }
}

// strided variant for class batches: out[k*ostride + j] = scale * sum_p ws[(k*panels + p)*cols + j]
__global__ void panel_reduce_strided_kernel(const double* ws, double* out, int64_t ostride, int nv, int panels, int64_t cols,
                                            double scale) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (j >= cols || k >= nv) return;
  double s = 0.0;
  for (int p = 0; p < panels; ++p) s += ws[((int64_t)k * panels + p) * cols + j];
  out[(int64_t)k * ostride + j] = scale * s;
}

}  // namespace admmb200
