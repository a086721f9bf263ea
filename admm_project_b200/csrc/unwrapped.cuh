// unwrapped.cuh -- the per-iteration kernels of the A = D problems (constraint D*x - z = c):
// linear SVM through unwrapped ADMM / transpose reduction (solvers/unwrappedadmm.m, linearsvm.m,
// getProxOps.m:1084-1143,1158-1180), Huber fitting and least absolute deviations
// (huberfit.m, lad.m, getProxOps.m:1511-1515,1529-1539,808-810).
//
// The reference touches D five times per iteration (D*x inside the z-prox, A(x) at admm.m:535,
// D'*(.) in the x-update, At(B(z-zprev)) :624, At(u) :654).  Here an iteration is TWO passes:
//   pass 1  uw_gemv_prox_kernel : Ax = D_g*x fused with the relaxed z-prox, the u-update, the rhs of
//           the next x-update and the partial sums of every norm (rows of this rank only);
//   pass 2  coldot_kernel<NV>   : D_g'*[rhs, z-zprev, u] in one sweep (tri.cuh);
// then one allreduce of [d ; D'dz ; D'u ; 8 scalars] when rows are sharded over GPUs
// (unwrappedadmm.m:96-141 does the same sum over parfor slices), and
//   uw_epilogue_kernel : residual norms, tolerances, H-norm, objective, stop tests (admm.m:618-722).
#pragma once
#include "common.cuh"
#include "prox.cuh"
#include "p2p.cuh"

namespace admmb200 {

enum { UW_SVM_HINGE = 0, UW_SVM_01 = 1, UW_HUBER = 2, UW_LAD = 3 };
constexpr int UW_NRED = 10;
constexpr int UW_THREADS = 256;
constexpr int UW_ROWS = 2 * UW_THREADS;

struct UwArgs {
  const double* D; int64_t ld, m, n;   // this rank's rows
  const double* x;                     // n
  double *z, *u;                       // m, in/out
  const double* aux;                   // ell (svm) or s (huber / lad)
  double* rvec;                        // rhs of the next x-update: z-u (svm) or s+z-u (huber / lad)
  double* dzvec;                       // z - zprev (NULL when nodualerror)
  double rho, relax, C;
  int kind;
  int64_t cols_per_chunk;
  double* ws; unsigned* tickets;       // column-chunk partials of the GEMV, one ticket per row block
  double* partials;                    // [gridDim.x][UW_NRED]
  unsigned* grid_ticket;
  double* scalars;                     // [UW_NRED] sums over this rank's rows
  const LoopCtl* ctl;
  double *zvals, *uvals;               // optional history (m x maxiters)
  // fast / accelerated ADMM (admm.m:506-529): prox and u-update use uhat; z/u of the previous
  // iteration are kept for the acceleration pass, which also writes rvec / dzvec
  int alg;
  const double *v, *uhat;
  double *zprev, *uprev;
};

__device__ __forceinline__ double huber1(double v) {  // CVX huber(v, 1)
  const double a = fabs(v);
  return a <= 1.0 ? a * a : 2.0 * a - 1.0;
}

// z-prox, u-update and norm terms of one row (admm.m:515-548 with A = D, B = -1) -- arithmetic only:
// zp / uold are the row's current z / u, up the u the prox uses (uhat for the fast variants), vprev the
// predictor v (fast variants).  Returns z, u; the sums go to r[].
struct UwRowOut { double z, u, dz, c; };
__device__ __forceinline__ UwRowOut uw_row_core(const UwArgs& a, double zp, double uold, double up, double aux,
                                                double vprev, double Ax, double (&r)[UW_NRED]) {
  const double c = (a.kind >= UW_HUBER) ? aux : 0.0;
  double xh = Ax;
  if (a.relax != 1.0) xh = a.relax * Ax - (1.0 - a.relax) * (-zp - c);   // admm.m:517
  double z, obj;
  if (a.kind == UW_SVM_HINGE) {            // getProxOps.m:1088-1096
    const double w = xh + up, v = aux * w;
    z = w + aux * fmax(fmin(1.0 - v, a.C / a.rho), 0.0);
    obj = fmax(1.0 - aux * Ax, 0.0);       // linearsvm.m:233
  } else if (a.kind == UW_SVM_01) {        // getProxOps.m:1100, minz01 :1158-1180 with t = rho/C
    const double w = xh + up, v = aux * w;
    const double y = (v >= 1.0 || v < 1.0 - sqrt(2.0 / (a.rho / a.C))) ? v : 1.0;
    z = aux * y;
    obj = (1.0 - aux * Ax) > 0.0 ? 1.0 : 0.0;   // pos(sign(.)), linearsvm.m:236
  } else if (a.kind == UW_HUBER) {         // getProxOps.m:1529-1539
    const double v = xh + up - aux;
    z = 1.0 / (1.0 + a.rho) * (a.rho * v + soft_threshold(v, 1.0 + 1.0 / a.rho));
    obj = huber1(z);                       // huberfit.m:180
  } else {                                 // getProxOps.m:808-810
    z = soft_threshold(xh + up - aux, 1.0 / a.rho);
    obj = fabs(z);                         // lad.m:148
  }
  const double u = up + (xh + (-z) - c);   // admm.m:542/548
  const double dz = z - zp, du = u - uold, pr = Ax + (-z) - c;
  if (a.alg != 0) {
    const double e1 = u - up, e2 = z - vprev;
    r[7] = fma(e1, e1, r[7]);      // ||u - uhat||^2
    r[8] = fma(e2, e2, r[8]);      // ||B(z - v)||^2
  }
  r[0] = fma(pr, pr, r[0]);
  r[1] = fma(Ax, Ax, r[1]);
  r[2] = fma(z, z, r[2]);
  r[3] = fma(c, c, r[3]);
  r[4] = fma(dz, dz, r[4]);
  r[5] = fma(du, du, r[5]);
  r[6] += obj;
  return UwRowOut{z, u, dz, c};
}

__device__ __forceinline__ void uw_row(const UwArgs& a, int64_t i, double Ax, int it, double (&r)[UW_NRED]) {
  const double zp = a.z[i], uold = a.u[i], aux = a.aux[i];
  const double up = a.alg ? a.uhat[i] : uold;
  const UwRowOut o = uw_row_core(a, zp, uold, up, aux, a.alg ? a.v[i] : 0.0, Ax, r);
  a.z[i] = o.z;
  a.u[i] = o.u;
  if (a.alg == 0) {
    a.rvec[i] = (a.kind >= UW_HUBER) ? (aux + o.z - o.u) : (o.z - o.u);
    if (a.dzvec) a.dzvec[i] = o.dz;
  } else {
    a.zprev[i] = zp;
    a.uprev[i] = uold;
  }
  if (a.zvals) {
    a.zvals[(int64_t)it * a.m + i] = o.z;
    a.uvals[(int64_t)it * a.m + i] = o.u;
  }
}

template <int VEC>
__global__ void __launch_bounds__(UW_THREADS) uw_gemv_prox_kernel(UwArgs a) {
  if (a.ctl->done) return;
  __shared__ double sh[(UW_THREADS / 32) * UW_NRED];
  __shared__ bool is_last, is_last_grid;
  const int tid = threadIdx.x;
  const int it = a.ctl->it;
  const int64_t r0 = (int64_t)blockIdx.x * UW_ROWS + 2 * tid;
  const int64_t cbeg = (int64_t)blockIdx.y * a.cols_per_chunk;
  const int64_t cend = min(a.n, cbeg + a.cols_per_chunk);
  double s0 = 0.0, s1 = 0.0;
  if (r0 < a.m) {
    const bool two = (r0 + 1 < a.m);
    const double* p = a.D + r0 + cbeg * a.ld;
    int64_t c = cbeg;
    if (VEC == 2 && two) {
      for (; c + 8 <= cend; c += 8) {
        double2 d[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = ldg_stream2(p + (int64_t)k * a.ld);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double vv = ldg_nc1(a.x + c + k);
          s0 = fma(d[k].x, vv, s0);
          s1 = fma(d[k].y, vv, s1);
        }
        p += 8 * a.ld;
      }
      for (; c < cend; ++c) {
        const double2 d = ldg_stream2(p);
        const double vv = ldg_nc1(a.x + c);
        s0 = fma(d.x, vv, s0);
        s1 = fma(d.y, vv, s1);
        p += a.ld;
      }
    } else {
      for (; c < cend; ++c) {
        const double vv = ldg_nc1(a.x + c);
        s0 = fma(ldg_stream1(p), vv, s0);
        if (two) s1 = fma(ldg_stream1(p + 1), vv, s1);
        p += a.ld;
      }
    }
  }
  if (gridDim.y > 1) {
    // column chunks: park the partial, the last CTA of this row block sums them in chunk order
    if (r0 < a.m) {
      a.ws[(int64_t)blockIdx.y * a.m + r0] = s0;
      if (r0 + 1 < a.m) a.ws[(int64_t)blockIdx.y * a.m + r0 + 1] = s1;
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
    s0 = s1 = 0.0;
    for (unsigned k = 0; k < gridDim.y; ++k) {
      if (r0 < a.m) s0 += __ldcg(a.ws + (int64_t)k * a.m + r0);
      if (r0 + 1 < a.m) s1 += __ldcg(a.ws + (int64_t)k * a.m + r0 + 1);
    }
  }
  double r[UW_NRED];
#pragma unroll
  for (int k = 0; k < UW_NRED; ++k) r[k] = 0.0;
  if (r0 < a.m) uw_row(a, r0, s0, it, r);
  if (r0 + 1 < a.m) uw_row(a, r0 + 1, s1, it, r);
  block_reduce_store<UW_NRED>(r, a.partials + (int64_t)blockIdx.x * UW_NRED, sh);
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
  if (tid < UW_NRED) {   // fixed-order sum over row blocks
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(a.partials + (int64_t)b * UW_NRED + tid);
    a.scalars[tid] = s;
  }
}

// rhs of the first x-update from the initial iterates: z0 - u0 (unwrappedadmm.m:133) or
// s + z0 - u0 (getProxOps.m:1514)
__global__ void uw_first_rhs_kernel(int64_t m, const double* z, const double* u, const double* aux, int kind,
                                    double* rvec) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) rvec[i] = (kind >= UW_HUBER) ? (aux[i] + z[i] - u[i]) : (z[i] - u[i]);
}

// one thread: predictor weight / restart of this iteration from the (rank-summed) scalars
__global__ void uw_accel_decide_kernel(LoopCtl* ctl, LoopParams lp, const double* scalars) {
  if (ctl->done) return;
  accel_decide(ctl, lp, scalars[7], scalars[8]);
}

// acceleration pass over this rank's rows (admm.m:562-600) + rhs of the next x-update from (v, uhat)
__global__ void uw_accel_kernel(int64_t m, const double* z, const double* u, const double* zprev, const double* uprev,
                                const double* aux, int kind, double* v, double* uhat, double* rvec, double* dzvec,
                                const LoopCtl* ctl) {
  if (ctl->done) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double gamma = ctl->gamma;
  const double zi = z[i], ui = u[i], zp = zprev[i], up = uprev[i];
  const double vn = ctl->restart ? zp : zi + gamma * (zi - zp);
  const double uh = ctl->restart ? up : ui + gamma * (ui - up);
  v[i] = vn;
  uhat[i] = uh;
  rvec[i] = (kind >= UW_HUBER) ? (aux[i] + vn - uh) : (vn - uh);
  if (dzvec) dzvec[i] = zi - vn;       // B(z - v) up to sign, admm.m:631
}

struct UwEpiArgs {
  int64_t n;
  const double* x;                 // n (replicated)
  const double *dzv, *duv;         // D'(z - zprev), D'u summed over ranks (NULL when nodualerror)
  const double* scalars;           // UW_NRED sums over all rows of all ranks
  double m_total;                  // numel(Ax) = numel(Bz) = global row count
  int kind;
  double C;
  LoopCtl* ctl;
  LoopParams lp;
  double* xvals;                   // optional history (n x maxiters)
  // row-sharded runs whose message went straight into the mailboxes (uw_onepass_finish_kernel): wait for every
  // rank, add the ranks up in rank order into msg[0 .. msg_count) (= [d ; D'dz ; D'u ; scalars]) first
  P2PDev mail;
  int use_mail;
  double* msg;
  int64_t msg_count;
};

__global__ void __launch_bounds__(256) uw_epilogue_kernel(UwEpiArgs a) {
  LoopCtl* ctl = a.ctl;
  if (ctl->done) return;
  __shared__ double sh[8 * 3];
  if (a.use_mail) {
    if (*a.mail.err) return;
    const unsigned long long seq = *a.mail.seq;
    const int par = (int)(seq & 1);
    const bool ok = p2p_wait(a.mail, par, seq);
    if (ok) {
      for (int64_t i = threadIdx.x; i < a.msg_count; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < a.mail.nranks; ++r) acc += __ldcg(a.mail.slot(a.mail.rank, par, r) + i);
        a.msg[i] = acc;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) *a.mail.seq = seq + 1;
    if (!ok) {                       // a peer stopped taking part: end the loop, the host raises (p2p_check)
      if (threadIdx.x == 0) { ctl->status = 3; __threadfence(); ctl->done = 1; }
      return;
    }
  }
  const int it = ctl->it;
  double r[3] = {0.0, 0.0, 0.0};
  for (int64_t i = threadIdx.x; i < a.n; i += blockDim.x) {
    const double x = a.x[i];
    r[0] = fma(x, x, r[0]);
    if (a.dzv) {
      const double p = a.lp.rho * a.dzv[i], q = a.lp.rho * a.duv[i];   // rho*At(B(z-zprev)), rho*At(u)
      r[1] = fma(p, p, r[1]);
      r[2] = fma(q, q, r[2]);
    }
    if (a.xvals) a.xvals[(int64_t)it * a.n + i] = x;
  }
  double tot[3];
  block_reduce_store<3>(r, tot, sh);   // tot valid in threads 0..2 only -> go through shared
  __syncthreads();
  if (threadIdx.x < 3) sh[threadIdx.x] = tot[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    double red[8];
    red[0] = a.scalars[0];
    red[1] = a.scalars[1];
    red[2] = a.scalars[2];
    red[3] = a.scalars[3];
    red[4] = sh[1];
    red[5] = sh[2];
    red[6] = a.scalars[4];
    red[7] = a.scalars[5];
    double obj;
    if (a.kind <= UW_SVM_01) obj = 0.5 * sh[0] + a.C * a.scalars[6];   // linearsvm.m:233,236
    else if (a.kind == UW_HUBER) obj = 0.5 * a.scalars[6];             // huberfit.m:180
    else obj = a.scalars[6];                                           // lad.m:148
    loop_epilogue(ctl, a.lp, red, a.m_total, a.m_total, obj);
  }
}

}  // namespace admmb200
