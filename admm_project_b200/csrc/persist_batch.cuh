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

constexpr int PB_R = 16, PB_RS = PB_R + 2, PB_RP = PB_R / 2, PB_JSTEP = OP_THREADS / PB_RP;
constexpr int PB_TS = 12;                             // row stride of t (shared memory) and of the per-tile rhs: B fragments conflict-free
constexpr int PB_MT = 4;                              // m-tiles (8 columns) per warp per column chunk in the T' phase: n <= 1024
constexpr int PB_MAXN = 16 * PB_MT * 8 * 2;           // 1024

template <int NBT> struct PersistBatchCfg {
  static size_t smem_bytes(int64_t n, int64_t npad) {
    return (size_t)(n * PB_RS + npad * PB_TS + (OP_THREADS / 32) * PB_R * NBT + PB_R * PB_TS + 8 + PB_R * NBT * UW_NRED + NBT * 16) * 8;
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
  long long* prof;                                    // optional [gridDim.x][8] SM cycles per phase (ADMM_B200_PERSIST_PROF)
};

template <int NBT>
__global__ void __launch_bounds__(OP_THREADS, 1) uwb_persist_kernel(PersistBatchArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  constexpr int R = PB_R, RS = PB_RS, RP = PB_RP, JSTEP = PB_JSTEP, NCH = 2, NW = OP_THREADS / 32, TS = PB_TS;
  constexpr int NT = (NBT + 7) / 8;                    // n-tiles (8 classes) of the DMMA products
  extern __shared__ __align__(16) double sm[];
  const int64_t n = a.n, m = a.m, npad = a.npad;
  double* T = sm;                                     // [n][RS]
  double* ts = T + n * RS;                            // [npad][TS]: t_c[j] at ts[j * TS + c], columns NBT..TS-1 zero
  double* wpart = ts + npad * TS;                     // [NW][R][NBT]
  double* rs = wpart + NW * R * NBT;                  // [R][TS] (+8): rhs of the tile's rows, rs[row * TS + class]
  double* redsm = rs + R * TS + 8;                    // [R][NBT][UW_NRED]
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
  for (int64_t i = tid; i < npad * TS; i += OP_THREADS) {
    const int64_t j = i / TS;
    const int c = (int)(i - j * TS);
    ts[i] = (j < n && c < NBT) ? a.tcur[(int64_t)c * a.cbs + j] : 0.0;
  }
  for (int i = tid; i < R * TS + 8; i += OP_THREADS) rs[i] = 0.0;
  const int64_t CW = ((n + NCH - 1) / NCH + 7) / 8 * 8;           // column chunks end on an 8-column DMMA tile
  const int q = tid % RP, j0 = tid / RP;              // phase 1: (row pair, column group)
  const int g = lane >> 2, t = lane & 3;              // DMMA fragment coordinates: a = A[g][t], b = B[t][g], c = C[g][2t..2t+1]
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

  long long tk = clock64();
  auto tick = [&](int k) {
    if (a.prof && tid == 0) {
      const long long now = clock64();
      a.prof[(int64_t)blockIdx.x * 8 + k] += now - tk;
      tk = now;
    }
  };
  for (int b = 0; b < a.burst; ++b) {
    // accumulators of T'*r: [column chunk][m-tile of this warp][n-tile][2]
    double acc[NCH][PB_MT][NT][2];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int i = 0; i < PB_MT; ++i)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[c][i][nt][0] = acc[c][i][nt][1] = 0.0;
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
      // ---- W = T * [t_1 .. t_NB] on DMMA tiles: M = 16 rows (2 m-tiles), N = classes, K = columns of Q.  The K steps
      // (4 columns) are dealt round-robin to the 16 warps; each warp keeps a 16 x 16 partial W in its C fragments.
      double wacc[2][NT][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) wacc[mt][nt][0] = wacc[mt][nt][1] = 0.0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (!tile_resident) {
          if (NCH - 1 - c == 1) cp_async_wait<1>();
          else cp_async_wait<0>();
        }
        __syncthreads();
        const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
        for (int64_t k0 = cbeg + 4 * warp; k0 < cend; k0 += 4 * NW) {
          const double* tcol = T + (k0 + t) * RS + g;             // T[row g (+8)][column k0 + t]
          const double* trow = ts + (k0 + t) * TS + g;            // t_{class g (+8)}[column k0 + t]
          const double a0 = tcol[0], a1 = tcol[8];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double bv = trow[8 * nt];
            dmma884(wacc[0][nt][0], wacc[0][nt][1], a0, bv);
            dmma884(wacc[1][nt][0], wacc[1][nt][1], a1, bv);
          }
        }
      }
      tick(0);                                         // tile wait + T*t
      {
        double* wp = wpart + (int64_t)warp * R * NBT;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cl = nt * 8 + 2 * t + e;
              if (cl < NBT) wp[(mt * 8 + g) * NBT + cl] = wacc[mt][nt][e];
            }
      }
      __syncthreads();
      tick(1);                                         // partial W of every warp to shared memory
      if (tid < R * NBT) {
        double rv = 0.0;
        if (pactive) {
          double w = 0.0;
#pragma unroll
          for (int gw = 0; gw < NW; ++gw) w += wpart[((int64_t)gw * R + prow) * NBT + pcls];   // fixed order over the warps
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
        rs[prow * TS + pcls] = rv;
      }
      __syncthreads();
      tick(2);                                         // proxes
      int64_t next = tile + gridDim.x;
      bool have_next = next < a.ntiles;
      if (!have_next && more && !single_tile) { next = first; have_next = true; }
      // ---- acc += T' * [r_1 .. r_NB] on DMMA tiles: M = columns of Q (8 per m-tile), N = classes, K = the 16 rows.
      // The B fragments (the rhs of the tile) are loaded once; warp w takes m-tiles w, w + 16, ... of each column chunk.
      double bf[4][NT];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) bf[ks][nt] = rs[(ks * 4 + t) * TS + nt * 8 + g];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int64_t cbeg = c * CW, cend = min(n, cbeg + CW);
#pragma unroll
        for (int i = 0; i < PB_MT; ++i) {
          const int64_t jt = cbeg + 8 * (warp + NW * i);
          if (jt < cend) {
            const double* tcol = T + (jt + g) * RS + t;             // T[row t (+4 ks)][column jt + g]
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const double av = tcol[4 * ks];
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) dmma884(acc[c][i][nt][0], acc[c][i][nt][1], av, bf[ks][nt]);
            }
          }
        }
        __syncthreads();
        if (have_next) issue(next, c);
      }
      if (single_tile) tile_resident = true;
      tick(3);                                         // T'*r
    }
    {
      double* dp = a.dpart + (int64_t)blockIdx.x * NBT * npad;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < PB_MT; ++i) {
          const int64_t j = c * CW + 8 * (warp + NW * i) + g;
          if (j < min(n, (c + 1) * CW))
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int cl = nt * 8 + 2 * t + e;
                if (cl < NBT) dp[(int64_t)cl * npad + j] = acc[c][i][nt][e];
              }
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
    tick(4);                                           // partial writes + grid barrier
    // ---------------- phase R / E1: the outputs this CTA owns ------------------------------------------------------
    const int par = (int)(seq & 1);
    const unsigned fl = (unsigned)(seq + 1);
    const int nparts = (int)gridDim.x;
    const int64_t per = n + UW_NRED, nout = (int64_t)NBT * per;
    // CTA b owns the outputs [b*chunk, (b+1)*chunk).  In blocks of 64 outputs: thread (lane, warp) adds the partials of
    // CTA group (warp >> 1) for output (warp & 1)*32 + lane -- <= 20 independent L2 loads in flight per thread -- the 8
    // group sums meet in shared memory and are added in a fixed order; the owner then stores the value into every
    // rank's mailbox, and 8 lanes per output collect it back from every rank (lane r spins on rank r's word).
    const int64_t chunk = (nout + gridDim.x - 1) / gridDim.x;
    const int64_t obeg = (int64_t)blockIdx.x * chunk, oend = min(nout, obeg + chunk);
    const int pp = (nparts + 7) / 8;
    double* red = wpart;                                  // [8][64], free between the tile loop and the next iteration
    auto out_index = [&](int64_t o) {                     // position of output o in the per-rank message / t array
      const int cl = (int)(o / per);
      const int64_t w = o - (int64_t)cl * per;
      return (int64_t)cl * a.cbs + ((w < n) ? w : npad + (w - n));
    };
    if (blockIdx.x == 0) {                                // the t this iteration used, for the classes still running
      for (int64_t i = tid; i < n * NBT; i += OP_THREADS) {
        const int64_t j = i / NBT;
        const int c = (int)(i - j * NBT);
        if (!cdone[c]) a.tlast[(int64_t)c * a.cbs + j] = ts[j * TS + c];
      }
    }
    for (int64_t ob = obeg; ob < oend; ob += 64) {
      {
        const int oi = (warp & 1) * 32 + lane, g = warp >> 1;
        const int64_t o = ob + oi;
        double s = 0.0;
        if (o < oend) {
          const int cl = (int)(o / per);
          const int64_t w = o - (int64_t)cl * per;
          const double* src = (w < n) ? a.dpart + (int64_t)cl * npad + w : a.partials + (int64_t)cl * UW_NRED + (w - n);
          const int64_t stride = (w < n) ? (int64_t)NBT * npad : (int64_t)NBT * UW_NRED;
          const int p0 = g * pp, p1 = min(nparts, p0 + pp);
          double v[20];
#pragma unroll
          for (int i = 0; i < 20; ++i) v[i] = (p0 + i < p1) ? __ldcg(src + (int64_t)(p0 + i) * stride) : 0.0;
#pragma unroll
          for (int i = 0; i < 20; ++i) s += v[i];
          for (int pidx = p0 + 20; pidx < p1; ++pidx) s += __ldcg(src + (int64_t)pidx * stride);
        }
        red[g * 64 + oi] = s;
      }
      __syncthreads();
      if (tid < 64 && ob + tid < oend) {
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += red[g * 64 + tid];
        p2p_ll_store(a.mail, par, out_index(ob + tid), s, fl);
      }
      {
        const int oi = tid >> 3, r = tid & 7;
        const int64_t o = ob + oi;
        double v = 0.0;
        const bool mine = (o < oend) && (r < a.mail.nranks);
        const int64_t idx = (o < oend) ? out_index(o) : 0;
        if (mine) {
          const ulonglong2* wd = a.mail.ll_slot(a.mail.rank, par, r) + idx;
          const long long t0 = clock64();
          while (!ll_load(wd, fl, v)) {
            if (clock64() - t0 > 4000000000LL) { *a.mail.err = 1; __trap(); }
          }
        }
        double s = 0.0;
#pragma unroll
        for (int rr = 0; rr < P2P_MAXRANKS; ++rr) {
          const double vr = __shfl_sync(0xffffffffu, v, rr, 8);      // rank order inside each group of 8 lanes
          if (rr < a.mail.nranks) s += vr;
        }
        if (r == 0 && o < oend) a.tcur[idx] = s;
      }
      __syncthreads();
    }
    __threadfence();
    tick(5);                                           // owner sums, mailbox stores, collection from every rank
    grid.sync();
    // ---------------- phase E2: the new t of every class, stop tests ------------------------------------------------
    {
      const unsigned total = (unsigned)(n * NBT);                     // 32-bit index arithmetic: constant-divisor div / mod
      const unsigned cbs32 = (unsigned)a.cbs;
      for (unsigned i0 = 0; i0 < total; i0 += 16u * OP_THREADS) {      // 16 independent loads per thread, then the stores
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const unsigned i = i0 + (unsigned)tid + (unsigned)k * OP_THREADS;
          const unsigned j = i / (unsigned)NBT, c = i - j * (unsigned)NBT;
          v[k] = (i < total && !cdone[c]) ? __ldcg(a.tcur + c * cbs32 + j) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const unsigned i = i0 + (unsigned)tid + (unsigned)k * OP_THREADS;
          const unsigned j = i / (unsigned)NBT, c = i - j * (unsigned)NBT;
          if (i < total && !cdone[c]) ts[j * (unsigned)TS + c] = v[k];
        }
      }
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
    tick(6);                                           // second barrier, reload of t, stop tests
    if (s_active == 0) break;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (blockIdx.x == 0 && tid == 0) *a.mail.seq = seq;
}

}  // namespace admmb200
