// prox.cuh -- the fused per-iteration pass of admm.m:515-560 + :603-722 for the problems whose
// constraint is x - z = 0 (A = 1, B = -1, c = 0: lasso.m:231-239, basispursuit.m:130-137, and the
// nonneg / box z-prox of getProxOps.m:1378-1382,1470-1474):
//
//   xhat = relax*x + (1-relax)*zprev            (admm.m:517, only when relax ~= 1)
//   z    = prox(xhat + u)                       (getProxOps.m:933-938 soft threshold, ...)
//   u    = u + (xhat - z)                       (admm.m:542/548)
//   y    = rhs of the NEXT x-update             (getProxOps.m:1196: rho*(z-u) + Dts)
//   partial sums of every norm admm.m:618-658,682 needs, reduced block -> grid in fixed order,
//   and, in the last CTA to finish, the scalar epilogue: pnorm/dnorm/perr/derr/Hnormsq/objective
//   history, H-norm divergence test (admm.m:686-701) and both stop tests (admm.m:706-722).
//
// One launch replaces ~8 interpreter temporaries and 5-7 separate norm passes of the reference.
#pragma once
#include "common.cuh"

namespace admmb200 {

// device-resident loop control (the host only polls `done` every check_every iterations)
struct LoopCtl {
  int it;            // iterations completed (MATLAB's i after the body ran)
  int done;          // non-zero: every kernel of later iterations exits immediately
  int status;        // ADMM_B200_CONVERGED_* ...
  unsigned ticket;   // last-block election
  double objpart;    // objective term produced by an earlier kernel of the iteration
  double pad;
  // fast / accelerated ADMM (admm.m:267-298, 562-600)
  double acurr, d, dprev, gamma;   // alpha_k, d_k, d_{k-1}, (alpha_{k-1} - 1) / alpha_k
  int restart, pad2;
  double sums[8];                  // norm sums of the z/u pass, consumed by the acceleration pass
};

struct LoopParams {
  double rho, relax, abstol, reltol, convtol, hnormtol, eps;
  long long maxiters;
  int domaxiters, stopcond, nodualerror, convtest, objevals, use_hnorm, raw;  // raw: no stop test
  int alg;                                         // 0 ADMM, 1 fast ADMM, 2 accelerated ADMM with restart
  double nrestart, dvaltol;
  double *pnorm, *dnorm, *perr, *derr, *hn, *obj;  // device history, length maxiters
  double *dvals, *avals, *rst;
};

enum { PROX_SOFT = 0, PROX_NONNEG = 1, PROX_BOX = 2, PROX_GIVEN = 3 };   // GIVEN: z was computed by a solve (model problem)
enum { NEXT_LASSO = 0, NEXT_DIFF = 1 };
constexpr int PROX_NRED = 10;
constexpr int PROX_THREADS = 256;

struct ProxIdentArgs {
  int64_t n;
  const double* x;        // x-update result
  double *z, *u;          // in/out
  const double* dts;      // Dts (lasso) or NULL
  double* y;              // rhs of the next x-update
  const double *lb, *ub;  // box bounds (length n) or NULL
  const double* zin;      // PROX_GIVEN: the z-update result (getProxOps.m:1011)
  double thresh;          // lambda/rho (lasso), 1/rho (bp)
  double objscale;        // objective = objpart + objscale * sum|z| (lasso: lambda) / sum|x| (bp)
  int kind, next, obj_l1_of_x;
  double* partials;       // [gridDim.x][PROX_NRED]
  LoopCtl* ctl;
  LoopParams lp;
  double *xvals, *zvals, *uvals;  // optional history (n x maxiters)
  // batch of independent columns (regularisation path): blockIdx.y = column
  int64_t ld;                     // column stride of x, z, u, y (0 for a single problem)
  const double* thresh_v;         // per-column threshold (NULL: thresh)
  const double* objscale_v;       // per-column objective scale (NULL: objscale)
  int64_t hist_stride;            // per-column stride of the scalar histories
  int* done_count;                // number of columns whose loop has ended (NULL for a single problem)
  double* xkeep;                  // batch: x of the last iteration a column ran (the GEMMs keep overwriting x)
  // fast / accelerated ADMM: the prox uses uhat, z/u of the previous iteration are kept for the acceleration pass
  const double *v, *uhat;
  double *zprev, *uprev;
};

__device__ __forceinline__ double soft_threshold(double v, double t) {
  // sign(v).*subplus(abs(v) - t), getProxOps.m:937
  double a = fabs(v) - t;
  a = a > 0.0 ? a : 0.0;
  return v > 0.0 ? a : (v < 0.0 ? -a : 0.0 * a);
}

template <int NRED>
__device__ __forceinline__ void block_reduce_store(double (&r)[NRED], double* dst, double* sh /*[nwarps][NRED]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NRED; ++k) r[k] = warp_sum(r[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NRED; ++k) sh[warp * NRED + k] = r[k];
  __syncthreads();
  if (threadIdx.x < NRED) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += sh[w * NRED + threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

// Scalar epilogue shared by every problem family.  red[] holds:
//  0 ||Ax+Bz-c||^2  1 ||Ax||^2  2 ||Bz||^2  3 ||c||^2  4 ||rho*At(B(z-zprev))||^2  5 ||rho*At(u)||^2
//  6 ||B(z-zprev)||^2  7 ||u-uprev||^2 ;  M1 = numel(Ax), M2 = numel(Bz); obj = objective value.
__device__ inline void loop_epilogue(LoopCtl* ctl, const LoopParams& lp, const double* red, double M1,
                                     double M2, double obj) {
  const int i = ctl->it + 1;  // MATLAB's i
  const int slot = lp.raw ? 0 : i - 1;
  const double pn = sqrt(red[0]);
  const double dn = lp.nodualerror ? __longlong_as_double(0x7ff8000000000000LL) : sqrt(red[4]);
  const double pe = sqrt(M1) * lp.abstol + lp.reltol * fmax(fmax(sqrt(red[1]), sqrt(red[2])), sqrt(red[3]));
  const double de = lp.nodualerror ? __longlong_as_double(0x7ff8000000000000LL)
                                   : sqrt(M2) * lp.abstol + lp.reltol * sqrt(red[5]);
  const double hn = lp.rho * red[6] + lp.rho * (lp.rho * lp.rho * red[7]);  // admm.m:305-306 on w = [x;z;rho*u]
  if (lp.alg != 2) {          // admm.m:618-658: the accelerated variant records no residual norms
    lp.pnorm[slot] = pn;
    lp.dnorm[slot] = dn;
    lp.perr[slot] = pe;
    lp.derr[slot] = de;
  }
  if (lp.use_hnorm) lp.hn[slot] = hn;
  if (lp.objevals) lp.obj[slot] = obj;
  int done = 0, status = 0;
  if (!lp.raw) {
    if (lp.convtest && i >= 2 && lp.alg == 0) {  // admm.m:689-700
      const double H2 = hn, H1 = lp.hn[i - 2];
      if (H1 > lp.eps && H2 > H1 && !((H2 - H1) <= H1 * lp.convtol)) { done = 1; status = 4; }
    }
    if (lp.alg == 2) {                                          // admm.m:706-707
      if (!done && i >= 2 && fabs(ctl->d - ctl->dprev) <= lp.dvaltol * ctl->dprev) { done = 1; status = 5; }
    } else if (!done && (lp.stopcond == 0 || lp.stopcond == 2) && !lp.domaxiters && pn < pe &&
               (lp.nodualerror || dn < de)) { done = 1; status = 1; }  // admm.m:710-713
    if (!done && (lp.stopcond == 1 || lp.stopcond == 2) && !lp.domaxiters && i > 2 && hn <= lp.hnormtol) {
      done = 1; status = 2;                                    // admm.m:719-722
    }
    if (!done && i >= lp.maxiters) { done = 1; status = 3; }
  }
  ctl->it = lp.raw ? ctl->it : i;
  ctl->status = status;
  ctl->objpart = 0.0;
  __threadfence();
  ctl->done = done;
}

// admm.m:562-600: predictor/corrector weight of this iteration (one thread).  s_uuh = ||u - uhat||^2,
// s_zv = ||B(z - v)||^2 with the v / uhat the iteration started from.
__device__ inline void accel_decide(LoopCtl* ctl, const LoopParams& lp, double s_uuh, double s_zv) {
  const int slot = lp.raw ? 0 : ctl->it;
  const double aprev = ctl->acurr;
  double acurr = 0.5 * (1.0 + sqrt(1.0 + 4.0 * aprev * aprev));
  int restart = 0;
  if (lp.alg == 2) {
    const double dprev = ctl->d;
    double d = 1.0 / lp.rho * s_uuh + lp.rho * s_zv;
    if (!(d < lp.nrestart * dprev)) {       // restart (admm.m:591-596)
      acurr = 1.0;
      restart = 1;
      d = dprev / lp.nrestart;
    }
    ctl->dprev = dprev;
    ctl->d = d;
    lp.dvals[slot] = d;
    lp.rst[slot] = (double)restart;
  }
  ctl->gamma = (aprev - 1.0) / acurr;
  ctl->acurr = acurr;
  ctl->restart = restart;
  lp.avals[slot] = acurr;
}

__global__ void ctl_init_kernel(LoopCtl* ctl, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LoopCtl c;
  memset(&c, 0, sizeof(c));
  c.acurr = 1.0;
  c.d = __longlong_as_double(0x7ff0000000000000LL);      // Inf (admm.m:283-284)
  c.dprev = c.d;
  ctl[i] = c;
}

__global__ void __launch_bounds__(PROX_THREADS) prox_ident_kernel(ProxIdentArgs a) {
  const int col = blockIdx.y;
  LoopCtl* ctl = a.ctl + col;
  if (ctl->done) return;   // a column that has stopped is frozen
  if (col > 0 || a.ld) {
    const int64_t off = (int64_t)col * a.ld;
    a.x += off; a.z += off; a.u += off; a.y += off;
    if (a.xkeep) a.xkeep += off;
    a.partials += (int64_t)col * gridDim.x * PROX_NRED;
    const int64_t ho = (int64_t)col * a.hist_stride;
    a.lp.pnorm += ho; a.lp.dnorm += ho; a.lp.perr += ho; a.lp.derr += ho; a.lp.hn += ho; a.lp.obj += ho;
    if (a.thresh_v) a.thresh = a.thresh_v[col];
    if (a.objscale_v) a.objscale = a.objscale_v[col];
  }
  __shared__ double sh[(PROX_THREADS / 32) * PROX_NRED];
  __shared__ bool is_last;
  const double rho = a.lp.rho, relax = a.lp.relax;
  const int it = ctl->it;
  double r[PROX_NRED];
#pragma unroll
  for (int k = 0; k < PROX_NRED; ++k) r[k] = 0.0;

  const bool fastmode = a.lp.alg != 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = a.x[i], zp = a.z[i], uold = a.u[i];
    const double up = fastmode ? a.uhat[i] : uold;      // admm.m:506-529: the fast variants use uhat
    double xh = x;
    if (relax != 1.0) xh = relax * x - (1.0 - relax) * (-zp - 0.0);
    const double v = xh + up;
    double z;
    if (a.kind == PROX_SOFT) z = soft_threshold(v, a.thresh);
    else if (a.kind == PROX_NONNEG) z = fmax(v, 0.0);
    else if (a.kind == PROX_GIVEN) z = a.zin[i];
    else z = fmin(a.ub[i], fmax(a.lb[i], v));
    const double u = up + (xh + (-z) - 0.0);
    a.z[i] = z;
    a.u[i] = u;
    if (a.xkeep) a.xkeep[i] = x;
    if (!fastmode) a.y[i] = (a.next == NEXT_LASSO) ? (rho * (z - u) + a.dts[i]) : (z - u);
    if (a.xvals) {
      a.xvals[(int64_t)it * a.n + i] = x;
      a.zvals[(int64_t)it * a.n + i] = z;
      a.uvals[(int64_t)it * a.n + i] = u;
    }
    const double pr = x + (-z) - 0.0, dz = z - zp, du = u - uold;
    r[0] = fma(pr, pr, r[0]);
    r[1] = fma(x, x, r[1]);
    r[2] = fma(z, z, r[2]);
    r[3] = fma(dz, dz, r[3]);
    r[4] = fma(u, u, r[4]);
    r[5] = fma(du, du, r[5]);
    r[6] += a.obj_l1_of_x ? fabs(x) : fabs(z);
    if (fastmode) {
      a.zprev[i] = zp;
      a.uprev[i] = uold;
      const double e1 = u - up, e2 = z - a.v[i];
      r[7] = fma(e1, e1, r[7]);                 // ||u - uhat||^2
      r[8] = fma(e2, e2, r[8]);                 // ||B(z - v)||^2
    }
  }
  block_reduce_store<PROX_NRED>(r, a.partials + (int64_t)blockIdx.x * PROX_NRED, sh);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&ctl->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fixed-order grid reduction by the last CTA
  if (threadIdx.x < PROX_NRED) {
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(a.partials + (int64_t)b * PROX_NRED + threadIdx.x);
    sh[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ctl->ticket = 0;
    if (fastmode) {
      // the residual norms need the NEW v (admm.m:629-633), so the epilogue runs in the acceleration
      // pass; here: keep the sums and fix this iteration's predictor weight / restart
#pragma unroll
      for (int k = 0; k < 7; ++k) ctl->sums[k] = sh[k];
      ctl->sums[7] = ctl->objpart + a.objscale * sh[6];
      accel_decide(ctl, a.lp, sh[7], sh[8]);
      return;
    }
    double red[8];
    red[0] = sh[0];                  // ||x - z||^2
    red[1] = sh[1];                  // ||A x||^2 = ||x||^2
    red[2] = sh[2];                  // ||B z||^2 = ||z||^2
    red[3] = 0.0;                    // c = 0
    red[4] = rho * rho * sh[3];      // ||rho * At(B(dz))||^2, At = 1
    red[5] = rho * rho * sh[4];      // ||rho * At(u)||^2
    red[6] = sh[3];
    red[7] = sh[5];
    const double obj = ctl->objpart + a.objscale * sh[6];
    loop_epilogue(ctl, a.lp, red, (double)a.n, (double)a.n, obj);
    if (a.done_count && ctl->done) atomicAdd(a.done_count, 1);
  }
}

// Acceleration pass of the fast variants for the x - z = 0 problems (admm.m:562-600):
//   v = z + gamma*(z - zprev), uhat = u + gamma*(u - uprev)   or, on a restart, v = zprev, uhat = uprev;
//   y = rhs of the next x-update from (v, uhat) (admm.m:506); sum ||z - v||^2 for the dual residual of
//   the fast variant (admm.m:629-633); last CTA: the scalar epilogue.
struct AccelIdentArgs {
  int64_t n;
  const double *z, *u, *zprev, *uprev, *dts;
  double *v, *uhat, *y;
  int next;
  double* partials;       // [gridDim.x]
  LoopCtl* ctl;
  LoopParams lp;
};

__global__ void __launch_bounds__(PROX_THREADS) accel_ident_kernel(AccelIdentArgs a) {
  LoopCtl* ctl = a.ctl;
  if (ctl->done) return;
  __shared__ double sh[PROX_THREADS / 32];
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
    a.y[i] = (a.next == NEXT_LASSO) ? (rho * (v - uh) + a.dts[i]) : (v - uh);
    const double e = z - v;
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
  double szv = 0.0;
  for (unsigned b = 0; b < gridDim.x; ++b) szv += __ldcg(a.partials + b);
  ctl->ticket = 0;
  const double* sm = ctl->sums;
  double red[8];
  red[0] = sm[0];
  red[1] = sm[1];
  red[2] = sm[2];
  red[3] = 0.0;
  red[4] = rho * rho * szv;        // rho*norm(At(B(z - v))), At = 1 (admm.m:631)
  red[5] = rho * rho * sm[4];
  red[6] = sm[3];
  red[7] = sm[5];
  loop_epilogue(ctl, a.lp, red, (double)a.n, (double)a.n, sm[7]);
}

// objective term 0.5*||D x - s||^2 from r = D*x computed by the GEMV kernel (lasso.m:227)
__global__ void half_sqdist_kernel(const double* __restrict__ r, const double* __restrict__ s, int64_t m,
                                   LoopCtl* ctl, int accumulate = 0) {
  if (ctl->done) return;
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
    double d = r[i] - s[i];
    acc = fma(d, d, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    ctl->objpart = accumulate ? ctl->objpart + 0.5 * t : 0.5 * t;
  }
}

// Model problem (getProxOps.m:1011): right-hand side of the z-update, Qts + rho*(xhat + u), with the
// relaxed xhat of admm.m:517 (A = 1, B = -1, c = 0) and uhat in place of u for the fast variants.
__global__ void model_rhs_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ z,
                                 const double* __restrict__ u, const double* __restrict__ qts, double rho, double relax,
                                 double* __restrict__ t, const LoopCtl* ctl) {
  if (ctl->done) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xh = x[i];
  if (relax != 1.0) xh = relax * xh - (1.0 - relax) * (-z[i] - 0.0);
  t[i] = qts[i] + rho * (xh + u[i]);
}

// objective 1/2*x'*P*x + q'*x + r from t = P*x (quadraticprogram.m: options.obj)
__global__ void quad_obj_kernel(const double* __restrict__ x, const double* __restrict__ t, const double* __restrict__ negq,
                                double r, int64_t n, LoopCtl* ctl) {
  if (ctl->done) return;
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc = fma(x[i], 0.5 * t[i] - negq[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sh[w];
    ctl->objpart = tot + r;
  }
}

// A[i][i] += shift;  v = -v
__global__ void add_diag_kernel(double* A, int64_t lda, int64_t n, double shift) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[i + i * lda] += shift;
}

// W_jj = 1 where the Gram matrix has an exactly zero diagonal entry (an all-zero column of D); counts them
__global__ void fix_zero_diag_kernel(double* A, int64_t lda, int64_t n, int* count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && A[i + i * lda] == 0.0) {
    A[i + i * lda] = 1.0;
    atomicAdd(count, 1);
  }
}

__global__ void add_diag_negate_kernel(double* A, int64_t lda, int64_t n, double shift, double* v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    A[i + i * lda] += shift;
    v[i] = -v[i];
  }
}

// y0 = rho*(z0 - u0) + Dts  /  z0 - u0  before the first iteration
__global__ void first_rhs_kernel(int64_t n, const double* z, const double* u, const double* dts, double rho,
                                 int next, double* y) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (next == NEXT_LASSO) ? (rho * (z[i] - u[i]) + dts[i]) : (z[i] - u[i]);
}

}  // namespace admmb200
