// uwbatch.cuh -- NB independent A = D problems that share D (the 10 one-vs-all linear SVMs of
// examples/mnistsvm.m:121-156 / BASELINE.json configs[2]) advanced together: D is read twice per
// iteration for ALL classes instead of twice per class.
//   pass 1  uwb_gemm_prox_kernel<NB> : AX = D_g * X (m x NB) with register accumulators per class, then the
//           per-(row, class) z-prox / u-update / next rhs / norm sums (uw_row of unwrapped.cuh);
//   pass 2  gemvt_strided_kernel<NB> : D_g' * R for the NB right-hand sides in one sweep (gemvt.cuh);
//   one allreduce of NB x [d ; scalars], the x-updates of all classes as two triangular DMMA GEMMs,
//   uwb_epilogue_kernel : one CTA per class, its own LoopCtl / stop test; a class that has stopped is frozen.
#pragma once
#include "common.cuh"
#include "prox.cuh"
#include "unwrapped.cuh"

namespace admmb200 {

struct UwbArgs {
  const double* D; int64_t ld, m, n;
  const double* X; int64_t ldx;          // n x nb
  double *Z, *U, *R; const double* AUX;  // m x nb, column stride ldm
  int64_t ldm;
  int nb;
  double rho, C;
  int kind;
  int64_t cols_per_chunk;
  double* ws;                            // [chunks][nb][m]
  unsigned* tickets;                     // one per row block
  double* partials;                      // [row blocks][nb][UW_NRED]
  unsigned* grid_ticket;
  double* cb; int64_t cb_stride, scal_off;   // class k: cb + k*cb_stride, scalars at + scal_off
  LoopCtl* ctl;                          // [nb]
};

template <int NB>
__global__ void __launch_bounds__(UW_THREADS) uwb_gemm_prox_kernel(UwbArgs a) {
  extern __shared__ __align__(16) double Xs[];     // [chunk cols][NB]
  __shared__ double sh[(UW_THREADS / 32) * UW_NRED];
  __shared__ bool is_last, is_last_grid;
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * UW_ROWS + 2 * tid;
  const int64_t cbeg = (int64_t)blockIdx.y * a.cols_per_chunk;
  const int64_t cend = min(a.n, cbeg + a.cols_per_chunk);
  constexpr int CH = 128;                         // columns of X staged in shared memory at a time
  constexpr int NB2 = (NB + 1) / 2;               // classes in pairs: LDS.128 broadcasts
  double acc0[NB], acc1[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) acc0[k] = acc1[k] = 0.0;
  const bool two = (r0 + 1 < a.m);
  const bool vec = ((a.ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.D) & 15) == 0);
  for (int64_t cc0 = cbeg; cc0 < cend; cc0 += CH) {
    const int ncol = (int)min((int64_t)CH, cend - cc0);
    __syncthreads();                               // previous chunk fully consumed
    for (int i = tid; i < ncol * 2 * NB2; i += UW_THREADS) {
      const int c = i / (2 * NB2), k = i - c * (2 * NB2);
      Xs[i] = (k < a.nb) ? a.X[cc0 + c + (int64_t)k * a.ldx] : 0.0;
    }
    __syncthreads();
    if (r0 < a.m) {
      const double* p = a.D + r0 + cc0 * a.ld;
      for (int c = 0; c < ncol; c += 8) {
        double2 d[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          d[q] = make_double2(0.0, 0.0);
          if (c + q < ncol) {
            if (vec && two) d[q] = ldg_stream2(p + (int64_t)q * a.ld);
            else { d[q].x = ldg_stream1(p + (int64_t)q * a.ld); if (two) d[q].y = ldg_stream1(p + (int64_t)q * a.ld + 1); }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (c + q < ncol) {
            const double2* xs = reinterpret_cast<const double2*>(Xs + (c + q) * 2 * NB2);
#pragma unroll
            for (int k2 = 0; k2 < NB2; ++k2) {
              const double2 xv = xs[k2];           // two classes per broadcast load
              acc0[2 * k2] = fma(d[q].x, xv.x, acc0[2 * k2]);
              acc1[2 * k2] = fma(d[q].y, xv.x, acc1[2 * k2]);
              if (2 * k2 + 1 < NB) {
                acc0[2 * k2 + 1] = fma(d[q].x, xv.y, acc0[2 * k2 + 1]);
                acc1[2 * k2 + 1] = fma(d[q].y, xv.y, acc1[2 * k2 + 1]);
              }
            }
          }
        }
        p += 8 * a.ld;
      }
    }
  }
  if (gridDim.y > 1) {
    if (r0 < a.m) {
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        if (k < a.nb) {
          double* w = a.ws + ((int64_t)blockIdx.y * a.nb + k) * a.m;
          w[r0] = acc0[k];
          if (r0 + 1 < a.m) w[r0 + 1] = acc1[k];
        }
      }
    }
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
    for (int k = 0; k < NB; ++k) {
      acc0[k] = acc1[k] = 0.0;
      if (k < a.nb) {
        for (unsigned ch = 0; ch < gridDim.y; ++ch) {
          const double* w = a.ws + ((int64_t)ch * a.nb + k) * a.m;
          if (r0 < a.m) acc0[k] += __ldcg(w + r0);
          if (r0 + 1 < a.m) acc1[k] += __ldcg(w + r0 + 1);
        }
      }
    }
  }
  // per class: prox / u-update / rhs / norm sums on this CTA's rows
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (k >= a.nb) break;
    const LoopCtl* ck = a.ctl + k;
    if (ck->done) continue;                      // frozen class (uniform over the CTA)
    UwArgs v;
    v.m = a.m; v.z = a.Z + (int64_t)k * a.ldm; v.u = a.U + (int64_t)k * a.ldm; v.aux = a.AUX + (int64_t)k * a.ldm;
    v.rvec = a.R + (int64_t)k * a.ldm; v.dzvec = nullptr; v.rho = a.rho; v.relax = 1.0; v.C = a.C; v.kind = a.kind;
    v.zvals = v.uvals = nullptr; v.alg = 0; v.v = v.uhat = nullptr; v.zprev = v.uprev = nullptr;
    double r[UW_NRED];
#pragma unroll
    for (int q = 0; q < UW_NRED; ++q) r[q] = 0.0;
    if (r0 < a.m) uw_row(v, r0, acc0[k], 0, r);
    if (r0 + 1 < a.m) uw_row(v, r0 + 1, acc1[k], 0, r);
    block_reduce_store<UW_NRED>(r, a.partials + ((int64_t)blockIdx.x * a.nb + k) * UW_NRED, sh);
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned t = atomicAdd(a.grid_ticket, 1u);
    is_last_grid = (t == gridDim.x - 1);
    if (is_last_grid) *a.grid_ticket = 0;
  }
  __syncthreads();
  if (!is_last_grid) return;
  __threadfence();
  for (int i = tid; i < a.nb * UW_NRED; i += UW_THREADS) {   // fixed-order sum over row blocks
    const int k = i / UW_NRED, q = i - k * UW_NRED;
    if (a.ctl[k].done) continue;
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(a.partials + ((int64_t)b * a.nb + k) * UW_NRED + q);
    a.cb[(int64_t)k * a.cb_stride + a.scal_off + q] = s;
  }
}

// R = Z - U (svm) or AUX + Z - U (huber / lad) for every class: rhs of the first x-update
__global__ void uwb_first_rhs_kernel(int64_t m, int nb, int64_t ldm, const double* Z, const double* U, const double* AUX,
                                     int kind, double* R) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= m || k >= nb) return;
  const int64_t o = i + (int64_t)k * ldm;
  R[o] = (kind >= UW_HUBER) ? (AUX[o] + Z[o] - U[o]) : (Z[o] - U[o]);
}

struct UwbEpiArgs {
  int64_t n;
  const double* X; int64_t ldx;
  double* XK;                      // x of the last iteration each class ran
  const double* cb; int64_t cb_stride, scal_off;
  double m_total;
  int kind;
  double C;
  LoopCtl* ctl;
  LoopParams lp;
  int64_t hist_stride;
  int* done_count;
};

__global__ void __launch_bounds__(256) uwb_epilogue_kernel(UwbEpiArgs a) {
  const int k = blockIdx.x;
  LoopCtl* ctl = a.ctl + k;
  if (ctl->done) return;
  __shared__ double sh[8];
  const double* x = a.X + (int64_t)k * a.ldx;
  double* xk = a.XK + (int64_t)k * a.ldx;
  double r[1] = {0.0};
  for (int64_t i = threadIdx.x; i < a.n; i += blockDim.x) {
    const double v = x[i];
    xk[i] = v;
    r[0] = fma(v, v, r[0]);
  }
  double tot[1];
  block_reduce_store<1>(r, tot, sh);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double xx = tot[0];
    const double* sc = a.cb + (int64_t)k * a.cb_stride + a.scal_off;
    LoopParams lp = a.lp;
    const int64_t ho = (int64_t)k * a.hist_stride;
    lp.pnorm += ho; lp.dnorm += ho; lp.perr += ho; lp.derr += ho; lp.hn += ho; lp.obj += ho;
    double red[8] = {sc[0], sc[1], sc[2], sc[3], 0.0, 0.0, sc[4], sc[5]};
    double obj;
    if (a.kind <= UW_SVM_01) obj = 0.5 * xx + a.C * sc[6];
    else if (a.kind == UW_HUBER) obj = 0.5 * sc[6];
    else obj = sc[6];
    loop_epilogue(ctl, lp, red, a.m_total, a.m_total, obj);
    if (ctl->done) atomicAdd(a.done_count, 1);
  }
}

}  // namespace admmb200
