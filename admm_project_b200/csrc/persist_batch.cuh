// persist_batch.cuh -- the one-vs-all CLASS BATCH of A = D problems (examples/mnistsvm.m:121-156: ten linearsvm(D,
// ell_k, C, options) calls on the same D; BASELINE.json configs[2]) as ONE persistent cooperative kernel per burst
// of iterations: persist.cuh with NB right-hand sides per tile.
//
// Same algebra as persist.cuh (Q_g = D_g inv(R)', t_c = sum_g Q_g' r_{g,c}; x_c = inv(R)' t_c after the loop), but a
// 16-row tile of Q_g is loaded ONCE per iteration and used for all NB classes: w_c = T t_c, the NB proxes of every
// row, acc_c += T' r_c.  The tile costs 2 * 16 * n * NB FMAs (FP64-bound at NB = 10), so D is no longer streamed twice
// per iteration for the batch and the x-updates (two triangular GEMMs per iteration) are gone.
//
// Exchange: the NB * (n + 10) outputs of an iteration are OWNED by CTAs (output o -> CTA o mod grid).  The owner adds
// the CTAs' partials in a fixed order, stores the sum into every rank's mailbox with the exchange number in the same
// words (p2p.cuh, LL form), then collects the same output from every rank -- lane r spins on rank r's word -- adds
// the ranks in rank order and writes the result into this GPU's t array.  A second grid barrier makes t complete;
// every CTA then reloads it and evaluates the stop tests of every class (admm.m:618-722) redundantly.  A class that
// has stopped is frozen (its z / u / t are no longer touched); the kernel ends when all classes have.
#pragma once
#include <cooperative_groups.h>

#include "onepass.cuh"
#include "p2p.cuh"

namespace admmb200 {

constexpr int PB_R = 16, PB_RS = PB_R + 2, PB_RP = PB_R / 2, PB_JSTEP = OP_THREADS / PB_RP, PB_RH = PB_R / 2;
constexpr int PB_MAXCOLS = 2;                         // columns per thread in the T' phase (512 column slots): n <= 1024

template <int NBT> struct PersistBatchCfg {
  static size_t smem_bytes(int64_t n, int64_t npad) {
    return (size_t)(n * PB_RS + npad * NBT + (OP_THREADS / 32) * PB_R * NBT + NBT * PB_R + PB_R * NBT * UW_NRED + NBT * 16) * 8;
  }
};

struct PersistBatchArgs {
  const double* Q; int64_t ld, m, n, npad;
  double *Z, *U; const double* AUX; int64_t ldm;      // m x nb, column stride ldm
  int nb;                                             // real classes (<= NBT); the padding classes carry zeros
  double rho, C; int kind;
  int64_t ntiles;
  double* dpart;                                      // [gridDim.x][NBT][npad]
  double* partials;                                   // [gridDim.x][NBT][UW_NRED]
  double* tcur;                                       // [NBT][cbs]: t_c (cbs >= npad + 16; scalars at + npad)
  double* tlast;                                      // [NBT][cbs]: t_c of the last iteration class c ran
  int64_t cbs;
  P2PDev mail;
  LoopCtl* ctl;                                       // [NBT]
  LoopParams lp; int64_t hist_stride;
  int* done_count;
  int burst;
  double m_total;
};

template <int NBT>
__global__ void __launch_bounds__(OP_THREADS, 1) uwb_persist_kernel(PersistBatchArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  constexpr int R = PB_R, RS = PB_RS, RP = PB_RP, JSTEP = PB_JSTEP, NCH = 2, NW = OP_THREADS / 32;
  extern __shared__ __align__(16) double sm[];
  const int64_t n = a.n, m = a.m, npad = a.npad;
  double* T = sm;                                     // [n][RS]
  double* ts = T + n * RS;                            // [npad][NBT]: t_c[j] at ts[j * NBT + c]
  double* wpart = ts + npad * NBT;                    // [NW][R][NBT]
  double* rs = wpart + NW * R * NBT;                  // [NBT][R]: rhs of the tile's rows, per class
  double* redsm = rs + NBT * R;                       // [R][NBT][UW_NRED]
  double* scal = redsm + R * NBT * UW_NRED;           // [NBT][16] (10 used)
  __shared__ int cdone[16], cit[16], s_active;
  __shared__ double chn[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (*a.mail.err) return;
  if (tid < NBT) {
    cdone[tid] = (tid < a.nb) ? a.ctl[tid].done : 1;
    cit[tid] = a.ctl[tid].it;
    chn[tid] = (a.lp.use_hnorm && cit[tid] >= 1) ? a.lp.hn[(int64_t)tid * a.hist_stride + cit[tid] - 1] : 0.0;
  }
  __syncthreads();
  {
    int act = 0;
    for (int c = 0; c < NBT; ++c) act += cdone[c] ? 0 : 1;
    if (act == 0) return;                             // uniform over the grid: ctl only changes at the end of a launch
  }
  unsigned long long seq = *a.mail.seq;
  for (int64_t i = tid; i < npad * NBT; i += OP_THREADS) {
    const int64_t j = i / NBT;
    const int c = (int)(i - j * NBT);
    ts[i] = (j < n) ? a.tcur[(int64_t)c * a.cbs + j] : 0.0;
  }
  const int64_t CW = (n + NCH - 1) / NCH;
  const int q = tid % RP, j0 = tid / RP;              // phase 1: (row pair, column group)
  const int sc = tid;                                 // phase 2: column slot (columns sc, sc + 512), all R rows
  const int prow = tid % R, pcls = tid / R;           // prox: (row, class) for tid < R * NBT
  const int64_t first = blockIdx.x;
  const bool single_tile = (first + gridDim.x >= a.ntiles);

  auto issue = [&](int64_t tile, int c) {
    const int64_t cbeg = c * CW, cend = min(n, cbeg + CW), row = tile * R + 2 * q;
    const int bytes = (int)min((int64_t)16, max((int64_t)0, (m - row) * 8));
    const double* src = a.Q + (bytes > 0 ? row : 0) + (cbeg + j0) * a.ld;
    double* dst = T + (cbeg + j0) * RS + 2 * q;
    for (int64_t j = cbeg + j0; j < cend; j += JSTEP) {
      cp_async16_zfill(dst, src, bytes);
      src += (int64_t)JSTEP * a.ld;
      dst += JSTEP * RS;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (first < a.ntiles)
    for (int c = 0; c < NCH; ++c) issue(first, c);
  bool tile_resident = false;
  UwArgs u;                                            // the per-row arithmetic of unwrapped.cuh needs these fields only
  u.rho = a.rho; u.relax = 1.0; u.C = a.C; u.kind = a.kind; u.alg = 0;

  for (int b = 0; b < a.burst; ++b) {
    double acc[PB_MAXCOLS][NBT];
#pragma unroll
    for (int k = 0; k < PB_MAXCOLS; ++k)
#pragma unroll
      for (int c = 0; c < NBT; ++c) acc[k][c] = 0.0;
    if (tid < R * NBT)                                 // norm sums of this thread's (row, class), kept in shared memory
#pragma unroll
      for (int k = 0; k < UW_NRED; ++k) redsm[((int64_t)prow * NBT + pcls) * UW_NRED + k] = 0.0;
    const bool more = (b + 1 < a.burst);
    // ---------------- phase D ------------------------------------------------------------------------------------
    for (int64_t tile = first; tile < a.ntiles; tile += gridDim.x) {
      // the prox threads fetch their z, u, aux now; the values are needed after the T*t phase
      const int64_t grow = tile * R + prow;
      const bool pactive = (tid < R * NBT) && (pcls < a.nb) && !cdone[pcls] && (grow < m);
      double zp = 0.0, uold = 0.0, aux = 0.0;
      if (pactive) {
        const int64_t o = grow + (int64_t)pcls * a.ldm;
        zp = a.Z[o]; uold = a.U[o]; aux = a.AUX[o];
      }
      double s0[NBT], s1[NBT];
#pragma unroll
      for (int c = 0; c < NBT; ++c) s0[c] = s1[c] = 0.0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (!tile_resident) {
          if (NCH - 1 - c == 1) cp_async_wait<1>();
          else cp_async_wait<0>();
        }
        __syncthreads();
        const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
        const double* tp = T + (cbeg + j0) * RS + 2 * q;
        const double* tt = ts + (cbeg + j0) * NBT;
        for (int64_t j = cbeg + j0; j < cend; j += JSTEP) {
          const double2 t2 = *reinterpret_cast<const double2*>(tp);
#pragma unroll
          for (int c2 = 0; c2 < NBT / 2; ++c2) {
            const double2 x2 = *reinterpret_cast<const double2*>(tt + 2 * c2);
            s0[2 * c2] = fma(t2.x, x2.x, s0[2 * c2]);
            s1[2 * c2] = fma(t2.y, x2.x, s1[2 * c2]);
            s0[2 * c2 + 1] = fma(t2.x, x2.y, s0[2 * c2 + 1]);
            s1[2 * c2 + 1] = fma(t2.y, x2.y, s1[2 * c2 + 1]);
          }
          tp += JSTEP * RS;
          tt += JSTEP * NBT;
        }
      }
      // the 4 column groups of a warp (lane bits 3, 4) are added by shuffles; lanes 0..7 park the warp's sums
#pragma unroll
      for (int c = 0; c < NBT; ++c) {
        s0[c] += __shfl_xor_sync(0xffffffffu, s0[c], 8);
        s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 8);
        s0[c] += __shfl_xor_sync(0xffffffffu, s0[c], 16);
        s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 16);
      }
      if (lane < RP) {
        double* wp = wpart + ((int64_t)warp * R + 2 * lane) * NBT;
#pragma unroll
        for (int c = 0; c < NBT; ++c) { wp[c] = s0[c]; wp[NBT + c] = s1[c]; }
      }
      __syncthreads();
      if (tid < R * NBT) {
        double rv = 0.0;
        if (pactive) {
          double w = 0.0;
#pragma unroll
          for (int g = 0; g < NW; ++g) w += wpart[((int64_t)g * R + prow) * NBT + pcls];   // fixed order over the warps
          double rl[UW_NRED];
#pragma unroll
          for (int k = 0; k < UW_NRED; ++k) rl[k] = 0.0;
          const UwRowOut o = uw_row_core(u, zp, uold, uold, aux, 0.0, w, rl);
#pragma unroll
          for (int k = 0; k < UW_NRED; ++k) redsm[((int64_t)prow * NBT + pcls) * UW_NRED + k] += rl[k];
          const int64_t oidx = grow + (int64_t)pcls * a.ldm;
          a.Z[oidx] = o.z;
          a.U[oidx] = o.u;
          rv = (a.kind >= UW_HUBER) ? (aux + o.z - o.u) : (o.z - o.u);
        }
        rs[pcls * R + prow] = rv;
      }
      __syncthreads();
      int64_t next = tile + gridDim.x;
      bool have_next = next < a.ntiles;
      if (!have_next && more && !single_tile) { next = first; have_next = true; }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
#pragma unroll
        for (int k = 0; k < PB_MAXCOLS; ++k) {
          const int64_t j = sc + (int64_t)k * OP_THREADS;
          if (j >= cbeg && j < cend) {
            const double* col = T + j * RS;
            double tc[R];
#pragma unroll
            for (int i = 0; i < R; i += 2) {
              const double2 t2 = *reinterpret_cast<const double2*>(col + i);
              tc[i] = t2.x; tc[i + 1] = t2.y;
            }
#pragma unroll
            for (int cl = 0; cl < NBT; ++cl) {
              const double* rr = rs + cl * R;
              double s = acc[k][cl];
#pragma unroll
              for (int i = 0; i < R; i += 2) {
                const double2 r2 = *reinterpret_cast<const double2*>(rr + i);     // broadcast
                s = fma(tc[i], r2.x, s);
                s = fma(tc[i + 1], r2.y, s);
              }
              acc[k][cl] = s;
            }
          }
        }
        __syncthreads();
        if (have_next) issue(next, c);
      }
      if (single_tile) tile_resident = true;
    }
    {
      double* dp = a.dpart + (int64_t)blockIdx.x * NBT * npad;
#pragma unroll
      for (int k = 0; k < PB_MAXCOLS; ++k) {
        const int64_t j = sc + (int64_t)k * OP_THREADS;
        if (j < n)
#pragma unroll
          for (int cl = 0; cl < NBT; ++cl) dp[(int64_t)cl * npad + j] = acc[k][cl];
      }
    }
    // norm sums per class: the R prox threads of a class, fixed order
    __syncthreads();
    if (tid < NBT * UW_NRED) {
      const int cl = tid / UW_NRED, k = tid % UW_NRED;
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) s += redsm[((int64_t)r * NBT + cl) * UW_NRED + k];
      a.partials[((int64_t)blockIdx.x * NBT + cl) * UW_NRED + k] = s;
    }
    __threadfence();
    grid.sync();
    // ---------------- phase R / E1: the outputs this CTA owns ------------------------------------------------------
    const int par = (int)(seq & 1);
    const unsigned fl = (unsigned)(seq + 1);
    const int nparts = (int)gridDim.x, ndparts = nparts;
    const int64_t per = n + UW_NRED, nout = (int64_t)NBT * per;
    for (int64_t o = (int64_t)blockIdx.x + (int64_t)gridDim.x * warp; o < nout; o += (int64_t)gridDim.x * NW) {
      const int cl = (int)(o / per);
      const int64_t w = o - (int64_t)cl * per;
      double v[10];
      const double* src = (w < n) ? a.dpart + (int64_t)cl * npad + w : a.partials + (int64_t)cl * UW_NRED + (w - n);
      const int64_t stride = (w < n) ? (int64_t)NBT * npad : (int64_t)NBT * UW_NRED;
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
      for (int of = 16; of > 0; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
      if (lane == 0) p2p_ll_store(a.mail, par, (int64_t)cl * a.cbs + ((w < n) ? w : npad + (w - n)), s, fl);
    }
    if (blockIdx.x == 0) {                             // the t this iteration used, for the classes still running
      for (int64_t i = tid; i < n * NBT; i += OP_THREADS) {
        const int64_t j = i / NBT;
        const int c = (int)(i - j * NBT);
        if (!cdone[c]) a.tlast[(int64_t)c * a.cbs + j] = ts[i];
      }
    }
    for (int64_t o = (int64_t)blockIdx.x + (int64_t)gridDim.x * warp; o < nout; o += (int64_t)gridDim.x * NW) {
      const int cl = (int)(o / per);
      const int64_t w = o - (int64_t)cl * per;
      const int64_t idx = (int64_t)cl * a.cbs + ((w < n) ? w : npad + (w - n));
      double v = 0.0;
      if (lane < a.mail.nranks) {                      // lane r collects rank r's copy of this output
        const ulonglong2* wd = a.mail.ll_slot(a.mail.rank, par, lane) + idx;
        const long long t0 = clock64();
        while (!ll_load(wd, fl, v)) {
          if (clock64() - t0 > 4000000000LL) { *a.mail.err = 1; __trap(); }
        }
      }
      double s = 0.0;
      for (int r = 0; r < a.mail.nranks; ++r) s += __shfl_sync(0xffffffffu, v, r);   // rank order
      if (lane == 0) a.tcur[idx] = s;
    }
    __threadfence();
    grid.sync();
    // ---------------- phase E2: the new t of every class, stop tests ------------------------------------------------
    for (int64_t i = tid; i < n * NBT; i += OP_THREADS) {
      const int64_t j = i / NBT;
      const int c = (int)(i - j * NBT);
      if (!cdone[c]) ts[i] = __ldcg(a.tcur + (int64_t)c * a.cbs + j);
    }
    if (tid < NBT * UW_NRED) {
      const int cl = tid / UW_NRED, k = tid % UW_NRED;
      scal[cl * 16 + k] = __ldcg(a.tcur + (int64_t)cl * a.cbs + npad + k);
    }
    __syncthreads();
    if (tid < NBT && tid < a.nb && !cdone[tid]) {
      const int cl = tid;
      const LoopParams& lp = a.lp;
      const double* sc10 = scal + cl * 16;
      const int i = cit[cl] + 1;
      const double pn = sqrt(sc10[0]);
      const double pe = sqrt(a.m_total) * lp.abstol + lp.reltol * fmax(fmax(sqrt(sc10[1]), sqrt(sc10[2])), sqrt(sc10[3]));
      const double hn = lp.rho * sc10[4] + lp.rho * (lp.rho * lp.rho * sc10[5]);
      int done = 0, status = 0;
      if (lp.convtest && i >= 2) {
        const double h1 = chn[cl];
        if (h1 > lp.eps && hn > h1 && !((hn - h1) <= h1 * lp.convtol)) { done = 1; status = 4; }
      }
      if (!done && (lp.stopcond == 0 || lp.stopcond == 2) && !lp.domaxiters && pn < pe) { done = 1; status = 1; }
      if (!done && (lp.stopcond == 1 || lp.stopcond == 2) && !lp.domaxiters && i > 2 && hn <= lp.hnormtol) { done = 1; status = 2; }
      if (!done && i >= lp.maxiters) { done = 1; status = 3; }
      if (blockIdx.x == 0) {
        const int64_t ho = (int64_t)cl * a.hist_stride + (i - 1);
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        lp.pnorm[ho] = pn; lp.dnorm[ho] = nan; lp.perr[ho] = pe; lp.derr[ho] = nan;
        if (lp.use_hnorm) lp.hn[ho] = hn;
        a.ctl[cl].it = i;
        a.ctl[cl].status = status;
        if (done) { __threadfence(); a.ctl[cl].done = 1; atomicAdd(a.done_count, 1); }
      }
      chn[cl] = hn;
      cit[cl] = i;
      cdone[cl] = done;
    }
    __syncthreads();
    if (tid == 0) {
      int act = 0;
      for (int c = 0; c < NBT; ++c) act += cdone[c] ? 0 : 1;
      s_active = act;
    }
    __syncthreads();
    ++seq;
    if (s_active == 0) break;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (blockIdx.x == 0 && tid == 0) *a.mail.seq = seq;
}

}  // namespace admmb200
