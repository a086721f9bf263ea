// persist.cuh -- the A = D iteration (linear SVM by unwrapped ADMM / transpose reduction, unwrappedadmm.m:96-141;
// Huber / LAD with nodualerror) as ONE persistent cooperative kernel per burst of iterations, with the cross-GPU
// exchange done by the kernel itself over NVLink peer memory.
//
// Algebra.  With R = chol(W), W = sum_g D_g'D_g (unwrappedadmm.m:114-122) and Q_g = D_g * inv(R)' (cached at
// setup, one DMMA GEMM per rank; Q = D inv(R)' has orthonormal columns), the x-update and the product A(x) of an
// iteration collapse:
//     x  = W \ d,  d = sum_g D_g' r_g        (unwrappedadmm.m:127-139)
//     Ax = D_g x = Q_g t,   t = inv(R) d = sum_g Q_g' r_g
// so an iteration needs only the n-vector t:   [t] -> w = Q_g t -> z-prox / u-update / norms -> t' = sum_g Q_g' r_g,
// no triangular solve and no x inside the loop (x = inv(R)' t is formed once after the loop, or per iteration by the
// non-persistent path when the objective / history is asked for).  Same iterates as the reference up to rounding
// (tests/test_gpu_persist.py: steps equal, 1e-9).
//
// One iteration inside the kernel (grid = one CTA per SM, cooperative launch):
//   D  every CTA walks its row tiles of Q_g (R rows x all n columns in shared memory, cp.async, as onepass.cuh):
//      w = T t -> prox of the R rows -> acc += T' r; writes its partial t' and norm sums        [grid barrier]
//   R  the n + 10 outputs are split over the CTAs: fixed-order sum over the CTAs' partials, stored straight into
//      EVERY rank's mailbox (p2p.cuh; remote stores over NVLink), last CTA raises this rank's flags
//   E  every CTA waits for all ranks' flags on its own mailbox, adds the ranks up in rank order into its shared
//      memory copy of t, and evaluates the stop tests (admm.m:618-722) redundantly -- identical inputs, identical
//      decision -- so no second grid barrier and no broadcast; CTA 0 records the histories.
// A single rank runs the same code on a local mailbox.  The first tile of the next iteration is prefetched
// across the barrier; a CTA that owns a single tile keeps it in shared memory for the whole burst.
#pragma once
#include <cooperative_groups.h>

#include "onepass.cuh"
#include "p2p.cuh"

namespace admmb200 {

namespace cg = cooperative_groups;

struct PersistArgs {
  UwArgs uw;                 // uw.D = Q_g (ld), uw.z / u / aux, rho, relax, C, kind; uw.x unused; alg == 0
  int64_t ntiles, npad;
  double* dpart;             // [gridDim.x][2][npad]
  double* partials;          // [gridDim.x][UW_NRED]
  double* tcur;              // [npad] t the next iteration starts from (in: from the host-side first rhs / last launch)
  double* tlast;             // [npad] t the LAST executed iteration used (x = inv(R)' tlast after the loop)
  P2PDev mail;
  LoopCtl* ctl;
  LoopParams lp;
  int burst;
  double m_total;
  long long* prof;           // optional [gridDim.x][8] accumulated SM cycles per phase (ADMM_B200_PERSIST_PROF)
};

template <int R, int NCH>
__global__ void __launch_bounds__(OP_THREADS, 1) uw_persist_kernel(PersistArgs a) {
  const UwArgs& u = a.uw;
  cg::grid_group grid = cg::this_grid();
  using Cfg = OnepassCfg<R>;
  constexpr int RS = Cfg::RS, RP = Cfg::RP, JSTEP = Cfg::JSTEP, ACTIVE = Cfg::ACTIVE, RH = R / 2;
  extern __shared__ __align__(16) double sm[];
  const int64_t n = u.n, m = u.m;
  double* T = sm;
  double* ts = T + n * RS;              // t of the current iteration
  double* wpart = ts + a.npad;          // [JSTEP][R]
  double* rs = wpart + JSTEP * R;       // [3][R] (only the rhs row is used)
  double* redsh = rs + 3 * R;           // [16][UW_NRED]
  __shared__ double scal[UW_NRED];
  __shared__ int s_stop;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.ctl->done || *a.mail.err) return;          // uniform over the grid: both only change at the end of a launch
  int it = a.ctl->it;
  unsigned long long seq = *a.mail.seq;
  double hn_prev = (a.lp.use_hnorm && it >= 1) ? a.lp.hn[it - 1] : 0.0;
  const int64_t CW = (n + NCH - 1) / NCH;
  for (int64_t j = tid; j < a.npad; j += OP_THREADS) ts[j] = (j < n) ? a.tcur[j] : 0.0;
  const int q = tid % RP, j0 = tid / RP;
  const int sc = tid % OP_SLOTS, hh = tid / OP_SLOTS;
  const int64_t first = blockIdx.x;
  const bool single_tile = (first + gridDim.x >= a.ntiles);      // this CTA owns one tile: it stays in shared memory

  auto issue = [&](int64_t tile, int c) {               // column chunk c of `tile` -> shared memory
    if (tid < ACTIVE) {
      const int64_t cbeg = c * CW, cend = min(n, cbeg + CW), row = tile * R + 2 * q;
      const int bytes = (int)min((int64_t)16, max((int64_t)0, (m - row) * 8));
      const double* src = u.D + (bytes > 0 ? row : 0) + (cbeg + j0) * u.ld;
      double* dst = T + (cbeg + j0) * RS + 2 * q;
      for (int64_t j = cbeg + j0; j < cend; j += JSTEP) {
        cp_async16_zfill(dst, src, bytes);
        src += (int64_t)JSTEP * u.ld;
        dst += JSTEP * RS;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (first < a.ntiles)
    for (int c = 0; c < NCH; ++c) issue(first, c);
  bool tile_resident = false;           // single_tile: the tile is already in shared memory from the previous iteration
  const bool rowthread = (tid < 8 * R) && ((tid & 7) == 0);
  const int myrow = tid >> 3;

  long long tk = clock64();
  auto tick = [&](int k) {
    if (a.prof && tid == 0) {
      const long long now = clock64();
      a.prof[(int64_t)blockIdx.x * 8 + k] += now - tk;
      tk = now;
    }
  };
  for (int b = 0; b < a.burst; ++b) {
    double acc[OP_MAXCOLS];
#pragma unroll
    for (int c = 0; c < OP_MAXCOLS; ++c) acc[c] = 0.0;
    double racc[UW_NRED];
#pragma unroll
    for (int k = 0; k < UW_NRED; ++k) racc[k] = 0.0;
    const bool more = (b + 1 < a.burst);
    // ---------------- phase D: this CTA's row tiles -------------------------------------------------------
    for (int64_t tile = first; tile < a.ntiles; tile += gridDim.x) {
      const int64_t row = tile * R + myrow;
      double zp = 0.0, uold = 0.0, aux = 0.0;
      if (rowthread && row < m) { zp = u.z[row]; uold = u.u[row]; aux = u.aux[row]; }
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (!tile_resident) {
          if (NCH - 1 - c == 1) cp_async_wait<1>();
          else cp_async_wait<0>();
        }
        __syncthreads();
        if (tid < ACTIVE) {
          const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
          const double* tp = T + (cbeg + j0) * RS + 2 * q;
#pragma unroll 4
          for (int64_t j = cbeg + j0; j < cend; j += JSTEP) {
            const double2 t = *reinterpret_cast<const double2*>(tp);
            const double xj = ts[j];
            s0 = fma(t.x, xj, s0);
            s1 = fma(t.y, xj, s1);
            tp += JSTEP * RS;
          }
        }
      }
      if (tid < ACTIVE) *reinterpret_cast<double2*>(wpart + j0 * R + 2 * q) = make_double2(s0, s1);
      __syncthreads();
      if (tid < 8 * R) {
        double w = 0.0;
        for (int g = tid & 7; g < JSTEP; g += 8) w += wpart[g * R + myrow];
        w += __shfl_xor_sync(0xffffffffu, w, 4);
        w += __shfl_xor_sync(0xffffffffu, w, 2);
        w += __shfl_xor_sync(0xffffffffu, w, 1);
        if ((tid & 7) == 0) {
          double rv = 0.0;
          if (row < m) {
            const UwRowOut o = uw_row_core(u, zp, uold, uold, aux, 0.0, w, racc);
            u.z[row] = o.z;
            u.u[row] = o.u;
            rv = (u.kind >= UW_HUBER) ? (aux + o.z - o.u) : (o.z - o.u);
          }
          rs[myrow] = rv;
        }
      }
      __syncthreads();
      int64_t next = tile + gridDim.x;
      bool have_next = next < a.ntiles;
      if (!have_next && more && !single_tile) { next = first; have_next = true; }   // first tile of the next iteration
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
#pragma unroll
        for (int k = 0; k < OP_MAXCOLS; ++k) {
          const int64_t j = sc + (int64_t)k * OP_SLOTS;
          if (j >= cbeg && j < cend) {
            const double* col = T + j * RS + hh * RH;
            const double* rr = rs + hh * RH;
            double a0 = acc[k];
#pragma unroll
            for (int i = 0; i < RH; i += 2) {
              const double2 t = *reinterpret_cast<const double2*>(col + i);
              const double2 r0 = *reinterpret_cast<const double2*>(rr + i);
              a0 = fma(t.x, r0.x, a0); a0 = fma(t.y, r0.y, a0);
            }
            acc[k] = a0;
          }
        }
        __syncthreads();
        if (have_next) issue(next, c);
      }
      if (single_tile) tile_resident = true;
    }
    double* dp = a.dpart + ((int64_t)blockIdx.x * 2 + hh) * a.npad;
#pragma unroll
    for (int k = 0; k < OP_MAXCOLS; ++k) {
      const int64_t j = sc + (int64_t)k * OP_SLOTS;
      if (j < n) dp[j] = acc[k];
    }
    block_reduce_store<UW_NRED>(racc, a.partials + (int64_t)blockIdx.x * UW_NRED, redsh);
    __threadfence();
    tick(0);                                   // phase D
    grid.sync();
    tick(1);                                   // grid barrier (includes waiting for the slowest CTA)
    // ---------------- phase R: fixed-order sums of the CTAs' partials -> every rank's mailbox ----------------
    const int par = (int)(seq & 1);
    const unsigned fl = (unsigned)(seq + 1);
    const int nparts = (int)gridDim.x, ndparts = 2 * nparts;
    // output w belongs to (CTA w mod grid, warp w div grid): every CTA gets its ~(n + 10) / grid outputs
    for (int64_t w = (int64_t)blockIdx.x + (int64_t)gridDim.x * warp; w < n + UW_NRED; w += (int64_t)gridDim.x * (OP_THREADS / 32)) {
      // all loads first (<= 10 per lane for 2 x 148 partials), then the adds in a fixed order: a `s += load` loop
      // serialises the L2 round trips (measured 3.4 us per iteration)
      double v[10];
      const double* src = (w < n) ? a.dpart + w : a.partials + (w - n);
      const int64_t stride = (w < n) ? a.npad : UW_NRED;
      const int cnt = (w < n) ? ndparts : nparts;
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const int pidx = lane + 32 * i;
        v[i] = (pidx < cnt) ? __ldcg(src + (int64_t)pidx * stride) : 0.0;
      }
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 10; ++i) s += v[i];
      for (int pidx = lane + 320; pidx < cnt; pidx += 32) s += __ldcg(src + (int64_t)pidx * stride);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) p2p_ll_store(a.mail, par, (w < n) ? w : a.npad + (w - n), s, fl);   // value + exchange number in one word pair
    }
    tick(2);                                   // phase R sums + stores
    // ---------------- phase E: every rank's values (they validate themselves), new t, stop tests ---------------
    if (blockIdx.x == 0)                                      // the t this iteration used: x = inv(R)' tlast
      for (int64_t j = tid; j < n; j += OP_THREADS) a.tlast[j] = ts[j];
    __syncthreads();
    for (int64_t j = tid; j < n + UW_NRED; j += OP_THREADS) {
      const int64_t idx = (j < n) ? j : a.npad + (j - n);
      const double s = p2p_ll_sum(a.mail, par, idx, fl);
      if (j < n) ts[j] = s;
      else scal[j - n] = s;
    }
    __syncthreads();
    tick(4);                                   // arrival of every rank's values
    const bool ok = true;
    if (tid == 0) {
      // admm.m:618-722 with nodualerror: every CTA evaluates the same numbers, CTA 0 records them
      const LoopParams& lp = a.lp;
      const int i = it + 1;
      const double pn = sqrt(scal[0]);
      const double pe = sqrt(a.m_total) * lp.abstol + lp.reltol * fmax(fmax(sqrt(scal[1]), sqrt(scal[2])), sqrt(scal[3]));
      const double hn = lp.rho * scal[4] + lp.rho * (lp.rho * lp.rho * scal[5]);
      int done = 0, status = 0;
      if (!ok) { done = 1; status = 3; }
      if (!done && lp.convtest && i >= 2) {
        if (hn_prev > lp.eps && hn > hn_prev && !((hn - hn_prev) <= hn_prev * lp.convtol)) { done = 1; status = 4; }
      }
      if (!done && (lp.stopcond == 0 || lp.stopcond == 2) && !lp.domaxiters && pn < pe) { done = 1; status = 1; }
      if (!done && (lp.stopcond == 1 || lp.stopcond == 2) && !lp.domaxiters && i > 2 && hn <= lp.hnormtol) { done = 1; status = 2; }
      if (!done && i >= lp.maxiters) { done = 1; status = 3; }
      if (blockIdx.x == 0) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        lp.pnorm[i - 1] = pn; lp.dnorm[i - 1] = nan; lp.perr[i - 1] = pe; lp.derr[i - 1] = nan;
        if (lp.use_hnorm) lp.hn[i - 1] = hn;
        a.ctl->it = i;
        a.ctl->status = status;
        if (done) { __threadfence(); a.ctl->done = 1; }
      }
      hn_prev = hn;
      s_stop = done;
    }
    __syncthreads();
    ++seq;
    ++it;
    tick(5);                                   // read t + stop tests
    if (s_stop) break;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (blockIdx.x == 0) {
    for (int64_t j = tid; j < n; j += OP_THREADS) a.tcur[j] = ts[j];
    if (tid == 0) *a.mail.seq = seq;
  }
}

}  // namespace admmb200
