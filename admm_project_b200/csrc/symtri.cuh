// symtri.cuh -- the per-iteration x-update x = U \ (L \ y) (getProxOps.m:1200; Rt \ (R \ .) :1514; W \ d
// unwrappedadmm.m:139) in ONE pass over the cached inverse factor.
//
// With W = inv(L):   x = W'(W y) = sum_i w_i (w_i . y),   w_i = row i of W = column i of WT = W' (rows 0..i,
// contiguous in memory).  A dot product and an AXPY with the SAME column: once a column sits on chip both are
// done from it, so the triangle is read from HBM once per x-update instead of twice (coldot x 2, tri.cuh) --
// n(n+1)/2 * 8 bytes = 268 MB at n = 8192.
//
// Data path: columns are fetched with the TMA engine as 1-D bulk copies (cp.async.bulk.shared.global, SASS
// UBLKCP) into a 3-stage shared-memory ring, completion on mbarriers; one thread issues, nobody executes a
// load instruction for matrix data.  A ROUND is a group of columns that fills one stage (~64 KB): columns are
// visited longest-with-shortest (their lengths sum to ~k+1), so every round moves the same number of bytes and
// the 148 CTAs finish together.  While round r is being worked on, rounds r+1 and r+2 are in flight (~130 KB
// per SM).
// Compute: 512 threads; thread t owns rows {2t, 2t+1} + 1024 q for every column; its y values and its slice of the result
// live in registers for the whole kernel.  Per round: partial dots from shared memory (conflict-free LDS.128)
// -> warp shuffle -> per-warp partials in shared memory -> barrier -> every thread sums the 16 partials in the
// same fixed order -> AXPY from shared memory into the register accumulators -> barrier -> refill the stage.
// Each CTA writes its partial x; the partials are summed in CTA order by symtri_reduce_kernel, so the result is
// bitwise reproducible.  Row-sharded runs give each rank a contiguous range of columns of equal area; the
// rank sums then go through the mailbox allreduce.
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace admmb200 {

constexpr int ST_THREADS = 512;
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_Q = 8;                               // row groups of 1024: k <= 8192 (1024 threads x 4 measured 1.5x slower)
constexpr int ST_MAXK = ST_Q * 2 * ST_THREADS;        // 8192
constexpr int ST_STAGE = 8192 + 64;                   // doubles per stage: a (longest, shortest) pair + alignment slack
constexpr int ST_NSTAGE = 3;
constexpr int ST_CPR = 16;                            // columns per round at most (small factors pack several pairs)
constexpr int ST_MAXENT = 448;                        // plan entries / rounds of a CTA kept in shared memory
constexpr int ST_MAXROUND = 127;
constexpr size_t ST_SMEM = (size_t)ST_NSTAGE * ST_STAGE * 8 + (size_t)2 * ST_CPR * ST_WARPS * 8 + 64 + (size_t)ST_MAXENT * 16 +
                           (size_t)(ST_MAXROUND + 1) * 4;

struct SymtriEnt { int col, off, len, pad; };        // column, offset in the stage (doubles, even), rows 0..len-1

struct SymtriArgs {
  const double* WT; int64_t ld;
  int k, kpad;
  int chunk;                       // doubles per bulk copy (even)
  int probe;                       // timing probe: 1 = move the data, skip the arithmetic (WRONG result)
  const double* y;
  double* xpart;                   // [gridDim.x][kpad]
  const int* done;
  const int* cta_round;            // [grid + 1]
  const int* round_ent;            // [nrounds + 1]
  const SymtriEnt* ents;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

__global__ void __launch_bounds__(ST_THREADS, 1) symtri_kernel(SymtriArgs a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(128) unsigned char st_raw[];
  double* stage = reinterpret_cast<double*>(st_raw);                                  // [NSTAGE][ST_STAGE]
  double* wsum = stage + (size_t)ST_NSTAGE * ST_STAGE;                                // [2][ST_CPR][ST_WARPS]: partial dots of two rounds
  uint64_t* full = reinterpret_cast<uint64_t*>(wsum + 2 * ST_CPR * ST_WARPS);         // [NSTAGE] (+ padding to 64 bytes)
  SymtriEnt* s_ents = reinterpret_cast<SymtriEnt*>(full + 8);                         // [ST_MAXENT]
  int* s_rent = reinterpret_cast<int*>(s_ents + ST_MAXENT);                           // [ST_MAXROUND + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = a.cta_round[blockIdx.x], r1 = a.cta_round[blockIdx.x + 1];
  // The CTA's slice of the plan goes to shared memory once: a round must not start with a dependent L2 round trip
  // for its descriptors (measured: ~40 % of the kernel when they were read from global memory every round).
  const int eb = a.round_ent[r0];
  const int ne_cta = a.round_ent[r1] - eb;
  const bool plan_in_smem = (ne_cta <= ST_MAXENT) && (r1 - r0 <= ST_MAXROUND);
  if (plan_in_smem) {
    for (int i = tid; i < ne_cta; i += ST_THREADS) s_ents[i] = a.ents[eb + i];
    for (int i = tid; i <= r1 - r0; i += ST_THREADS) s_rent[i] = a.round_ent[r0 + i] - eb;
  }
  auto rent = [&](int round) { return plan_in_smem ? s_rent[round - r0] : a.round_ent[round] - eb; };
  auto ent = [&](int e) { return plan_in_smem ? s_ents[e] : a.ents[eb + e]; };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < ST_NSTAGE; ++s) mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int round) {            // thread 0: all columns of `round` into stage (round - r0) % NSTAGE
    const int s = (round - r0) % ST_NSTAGE;
    const int e0 = rent(round), e1 = rent(round + 1);
    unsigned total = 0;
    for (int e = e0; e < e1; ++e) total += (unsigned)((ent(e).len + 1) & ~1) * 8u;
    mbar_expect_tx(full + s, total);
    for (int e = e0; e < e1; ++e) {
      const SymtriEnt en = ent(e);
      // a column can go up as several bulk copies of a.chunk doubles (ADMM_B200_SYMTRI_CHUNK); measured on B200, pieces
      // are SLOWER than one copy per column, so the default chunk is larger than any column
      const int padded = (en.len + 1) & ~1;
      double* dst = stage + (size_t)s * ST_STAGE + en.off;
      const double* src = a.WT + (int64_t)en.col * a.ld;
      for (int o = 0; o < padded; o += a.chunk) {
        const int c = min(a.chunk, padded - o);
        tma_bulk_g2s(dst + o, src + o, (unsigned)c * 8u, full + s);
      }
    }
  };
  if (tid == 0)
    for (int r = r0; r < r1 && r < r0 + ST_NSTAGE; ++r) issue(r);

  // this thread's rows: 2*tid + 1024*q (+0, +1)
  double2 yv[ST_Q], acc[ST_Q];
#pragma unroll
  for (int q = 0; q < ST_Q; ++q) {
    const int row = 2 * tid + 2 * ST_THREADS * q;
    yv[q] = make_double2(row < a.k ? a.y[row] : 0.0, row + 1 < a.k ? a.y[row + 1] : 0.0);
    acc[q] = make_double2(0.0, 0.0);
  }

  // Skewed by one round: an iteration does the DOTS of round r+1 and the AXPYs of round r, then ONE barrier -- the
  // partial dots of r+1 are complete for the next iteration, everybody is done with the stage of round r (refilled at
  // once) and with the partial-dot buffer the iteration after next overwrites.  Half the barriers of the straight
  // dot / barrier / AXPY / barrier form, and the shared-memory sweep of one round overlaps the shuffle / tree-sum
  // latency of the other.
  auto dots = [&](int round) {
    const int s = (round - r0) % ST_NSTAGE;
    const unsigned parity = (unsigned)(((round - r0) / ST_NSTAGE) & 1);
    const double* sb = stage + (size_t)s * ST_STAGE;
    double* ws = wsum + (size_t)((round - r0) & 1) * ST_CPR * ST_WARPS;
    const int e0 = rent(round), ne = rent(round + 1) - e0;
    mbar_wait(full + s, parity);
    if (a.probe) return;
    // two columns at a time (a round of a large factor is a long and a short column): two independent chains of
    // LDS -> DFMA -> shuffles per thread instead of one -- the kernel is bound by instruction latency at 16 warps per SM
    int j = 0;
    for (; j + 1 < ne; j += 2) {
      const SymtriEnt ea = ent(e0 + j), eb = ent(e0 + j + 1);
      const double* ca = sb + ea.off;
      const double* cb = sb + eb.off;
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
      for (int q = 0; q < ST_Q; ++q) {
        const int row = 2 * tid + 2 * ST_THREADS * q;
        if (row < ea.len) {
          const double2 v = *reinterpret_cast<const double2*>(ca + row);
          a0 = fma(v.x, yv[q].x, a0);
          a1 = fma(v.y, yv[q].y, a1);
        }
        if (row < eb.len) {
          const double2 v = *reinterpret_cast<const double2*>(cb + row);
          b0 = fma(v.x, yv[q].x, b0);
          b1 = fma(v.y, yv[q].y, b1);
        }
      }
      double pa = a0 + a1, pb = b0 + b1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        pa += __shfl_xor_sync(0xffffffffu, pa, o);
        pb += __shfl_xor_sync(0xffffffffu, pb, o);
      }
      if (lane == 0) { ws[j * ST_WARPS + warp] = pa; ws[(j + 1) * ST_WARPS + warp] = pb; }
    }
    for (; j < ne; ++j) {
      const SymtriEnt en = ent(e0 + j);
      const double* col = sb + en.off;
      double p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int q = 0; q < ST_Q; ++q) {
        const int row = 2 * tid + 2 * ST_THREADS * q;
        if (row < en.len) {                       // len is rounded up with a zero: row + 1 is loaded too
          const double2 v = *reinterpret_cast<const double2*>(col + row);
          p0 = fma(v.x, yv[q].x, p0);
          p1 = fma(v.y, yv[q].y, p1);
        }
      }
      const double p = warp_sum(p0 + p1);
      if (lane == 0) ws[j * ST_WARPS + warp] = p;
    }
  };
  auto tree = [&](const double* ws, int j) {        // sum of the 16 warp partials of column j: the same fixed tree in every thread
    const double2* wp = reinterpret_cast<const double2*>(ws + j * ST_WARPS);
    double2 part[ST_WARPS / 2];
#pragma unroll
    for (int w = 0; w < ST_WARPS / 2; ++w) part[w] = wp[w];              // broadcast LDS.128, all independent
#pragma unroll
    for (int w = 0; w < ST_WARPS / 2; ++w) part[w].x += part[w].y;
#pragma unroll
    for (int st = ST_WARPS / 4; st >= 1; st >>= 1)
#pragma unroll
      for (int w = 0; w < st; ++w) part[w].x += part[w + st].x;
    return part[0].x;
  };
  auto axpys = [&](int round) {
    const int s = (round - r0) % ST_NSTAGE;
    const double* sb = stage + (size_t)s * ST_STAGE;
    const double* ws = wsum + (size_t)((round - r0) & 1) * ST_CPR * ST_WARPS;
    const int e0 = rent(round), ne = rent(round + 1) - e0;
    if (a.probe) return;
    int j = 0;
    for (; j + 1 < ne; j += 2) {                      // two columns at a time, accumulated in column order
      const SymtriEnt ea = ent(e0 + j), eb = ent(e0 + j + 1);
      const double* ca = sb + ea.off;
      const double* cb = sb + eb.off;
      const double ta = tree(ws, j), tb = tree(ws, j + 1);
#pragma unroll
      for (int q = 0; q < ST_Q; ++q) {
        const int row = 2 * tid + 2 * ST_THREADS * q;
        if (row < ea.len) {
          const double2 v = *reinterpret_cast<const double2*>(ca + row);
          acc[q].x = fma(v.x, ta, acc[q].x);
          acc[q].y = fma(v.y, ta, acc[q].y);
        }
        if (row < eb.len) {
          const double2 v = *reinterpret_cast<const double2*>(cb + row);
          acc[q].x = fma(v.x, tb, acc[q].x);
          acc[q].y = fma(v.y, tb, acc[q].y);
        }
      }
    }
    for (; j < ne; ++j) {
      const SymtriEnt en = ent(e0 + j);
      const double* col = sb + en.off;
      const double t = tree(ws, j);
#pragma unroll
      for (int q = 0; q < ST_Q; ++q) {
        const int row = 2 * tid + 2 * ST_THREADS * q;
        if (row < en.len) {
          const double2 v = *reinterpret_cast<const double2*>(col + row);
          acc[q].x = fma(v.x, t, acc[q].x);
          acc[q].y = fma(v.y, t, acc[q].y);
        }
      }
    }
  };
  if (r0 < r1) {
    dots(r0);
    __syncthreads();
  }
  for (int round = r0; round < r1; ++round) {
    if (round + 1 < r1) dots(round + 1);
    axpys(round);
    __syncthreads();
    if (tid == 0 && round + ST_NSTAGE < r1) issue(round + ST_NSTAGE);
  }
  double* xp = a.xpart + (size_t)blockIdx.x * a.kpad;
#pragma unroll
  for (int q = 0; q < ST_Q; ++q) {
    const int row = 2 * tid + 2 * ST_THREADS * q;
    if (row < a.kpad) *reinterpret_cast<double2*>(xp + row) = acc[q];
  }
}

// x[r] = scale * sum over CTAs (fixed order) of xpart[cta][r] (+ addscale * addend[r]).  32 rows x 8 CTA groups per block.
// mail_on: this rank's sums go straight into every rank's mailbox (flag-in-data words) instead of x; a
// p2p_ll_gather_kernel then adds the ranks up into x (row-sharded lasso, the x-update split by columns over the ranks).
__global__ void __launch_bounds__(256) symtri_reduce_kernel(const double* __restrict__ xpart, int nparts, int kpad, int k,
                                                            double* __restrict__ x, double scale, const double* addend,
                                                            double addscale, const int* done, P2PDev mail, int mail_on) {
  if (done && *done) return;
  if (mail_on && *mail.err) return;
  const unsigned long long seq = mail_on ? *mail.seq : 0ull;
  __shared__ double sh[8][33];
  const int rl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int row = blockIdx.x * 32 + rl;
  const int per = (nparts + 7) / 8;
  double s = 0.0;
  if (row < k)
    for (int c = g * per; c < min(nparts, (g + 1) * per); ++c) s += __ldcg(xpart + (size_t)c * kpad + row);
  sh[g][rl] = s;
  __syncthreads();
  if (g == 0 && row < k) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][rl];
    t *= scale;
    if (addend) t += addscale * addend[row];
    if (mail_on) p2p_ll_store(mail, (int)(seq & 1), row, t, (unsigned)(seq + 1));
    else x[row] = t;
  }
}

}  // namespace admmb200
