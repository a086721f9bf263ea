// p2p.cuh -- one-shot allreduce of a SMALL message over NVLink peer memory (one process per GPU).
//
// The row-sharded solvers exchange ONE message per iteration: [D_g'r ; D_g'dz ; D_g'u ; scalars], 6-100 KB
// (unwrappedadmm.m:127-139 sums the same vectors over parfor slices on the client).  At that size a
// collective is pure latency, and a library allreduce between two small kernels costs more than the
// kernels (profiles/r01_scaling_unwrapped.txt: 8-GPU SVM at 19 % efficiency).  Here every rank owns a
// MAILBOX in its HBM, mapped into every peer with CUDA IPC:
//
//     mailbox = flags[2][R] (uint64)  +  slots[2][R][cap] (double)          R = ranks, 2 = parity
//
// Exchange number s (a device-resident counter, identical on every rank):
//   push : rank r stores its `count` doubles into slot[s&1][r] of EVERY rank's mailbox (remote stores over
//          NVLink; the own copy is a local store), fences at system scope, and the last CTA to finish
//          sets flag[s&1][r] = s+1 in every mailbox (st.release.sys);
//   wait : each rank spins on its OWN mailbox until all R flags read s+1 (ld.acquire.sys, local HBM/L2),
//          then sums the R slots in RANK ORDER -- every rank adds the same numbers in the same order, so
//          the result is bitwise identical everywhere and the replicated stop decision stays replicated.
// Parity double-buffering is enough: a rank can only start exchange s+2 after every rank has finished
// reading exchange s (stream order on each rank + the flags of exchange s+1).
// No host involvement, fixed kernel arguments: the exchange can sit inside a CUDA graph.
// A spin that sees no progress for ~2 s sets *err and gives up (a dead peer must not hang the GPU).
#pragma once
#include "common.cuh"

namespace admmb200 {

constexpr int P2P_MAXRANKS = 8;
constexpr int64_t P2P_CAP = 32768;          // doubles per slot (256 KB): n-vector of C2, 16-class batch of C3
constexpr int P2P_FLAG_BYTES = 256;         // flags[2][8] uint64 = 128 B, padded
constexpr int64_t P2P_LLCAP = 16384;        // values per slot of the flag-in-data (LL) region (10 classes x 800)

struct P2PDev {
  int rank, nranks;
  int64_t cap;
  unsigned char* mail[P2P_MAXRANKS];        // base of every rank's mailbox as mapped in THIS process
  unsigned long long* seq;                  // exchanges completed (device memory of this rank)
  unsigned* ticket;                         // last-CTA election of the push
  int* err;                                 // set when a wait timed out
  __device__ __forceinline__ unsigned long long* flags(int r, int par) const {
    return reinterpret_cast<unsigned long long*>(mail[r]) + par * P2P_MAXRANKS;
  }
  __device__ __forceinline__ double* slot(int r, int par, int src) const {
    return reinterpret_cast<double*>(mail[r] + P2P_FLAG_BYTES) + ((int64_t)par * nranks + src) * cap;
  }
  // LL region (behind the plain slots): every value travels as two 8-byte words {lo32 | flag<<32, hi32 | flag<<32}
  __device__ __forceinline__ ulonglong2* ll_slot(int r, int par, int src) const {
    return reinterpret_cast<ulonglong2*>(mail[r] + P2P_FLAG_BYTES + (size_t)2 * nranks * cap * 8) +
           ((int64_t)par * nranks + src) * P2P_LLCAP;
  }
};
__host__ __device__ inline size_t p2p_mailbox_bytes(int nranks) {
  return (size_t)P2P_FLAG_BYTES + (size_t)2 * nranks * P2P_CAP * 8 + (size_t)2 * nranks * P2P_LLCAP * 16;
}

struct P2PState {
  bool ready = false;
  P2PDev dev{};
  void* local = nullptr;                    // this rank's mailbox (cudaMalloc)
  void* opened[P2P_MAXRANKS] = {};          // cudaIpcOpenMemHandle results (nullptr for the own rank)
  size_t bytes = 0;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// store one value of this rank's contribution into every mailbox
__device__ __forceinline__ void p2p_store(const P2PDev& p, int par, int64_t idx, double v) {
#pragma unroll
  for (int r = 0; r < P2P_MAXRANKS; ++r)
    if (r < p.nranks) p.slot(r, par, p.rank)[idx] = v;
}

__device__ __forceinline__ unsigned atom_add_acq_rel_gpu(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

// Called by ALL threads of the CTA after their p2p_store calls; `nctas` CTAs take part.  The CTA barrier orders
// every thread's stores before thread 0, whose ONE system-scope fence then covers them all (fences are cumulative);
// measured on B200: a system fence executed by all 512 threads of 148 CTAs cost ~7 us per exchange, this form ~2.
// The ticket is an acquire-release atomic, so the last CTA to arrive has every other CTA's fenced stores in its
// past when lanes 0..R-1 of its first warp publish the flags (one st.release.sys per destination rank, in parallel).
__device__ __forceinline__ void p2p_signal(const P2PDev& p, int par, unsigned long long s, unsigned nctas) {
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned last = 0;
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned t = atom_add_acq_rel_gpu(p.ticket, 1u);
      last = (t == nctas - 1) ? 1u : 0u;
      if (last) *p.ticket = 0;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    __syncwarp();
    if (last && threadIdx.x < p.nranks) st_release_sys(p.flags(threadIdx.x, par) + p.rank, s + 1);
  }
}

// Spin (lane r of the first warp watches rank r) until every rank's flag of this exchange is up; the CTA barrier then
// carries the acquire to every thread.  Returns false on a timeout (and sets *err).  trap_on_timeout: inside a
// cooperative kernel a CTA that gave up would leave the others waiting at the grid barrier for ever, so the whole
// kernel is aborted instead (the host sees a launch failure -- loud, and the GPU stays usable).
__device__ __forceinline__ bool p2p_wait(const P2PDev& p, int par, unsigned long long s, bool trap_on_timeout = false) {
  __shared__ int p2p_bad;
  if (threadIdx.x == 0) p2p_bad = 0;
  __syncthreads();
  if (threadIdx.x < p.nranks) {
    const unsigned long long* f = p.flags(p.rank, par) + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < s + 1) {
      if (clock64() - t0 > 4000000000LL) {   // ~2 s at 1.9 GHz
        p2p_bad = 1;
        *p.err = 1;
        if (trap_on_timeout) __trap();
        break;
      }
    }
  }
  __syncthreads();
  return p2p_bad == 0;
}

// ---- flag-in-data exchange (the LL protocol of collective libraries) ------------------------------------------
// A value and the number of its exchange travel in the SAME 8-byte words: {lo32 | f<<32, hi32 | f<<32}, f = low 32
// bits of (exchange + 1).  An aligned 8-byte store is atomic, so a reader that sees f in both words has the value of
// THIS exchange -- no fence, no flag, no election: the writer just stores, the reader spins on the data itself.
// Used inside the persistent kernel, where the separate fence + ticket + flag cost more than the arithmetic.
__device__ __forceinline__ void ll_store(ulonglong2* p, double v, unsigned f) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long a = (bits & 0xffffffffull) | ((unsigned long long)f << 32);
  const unsigned long long b = (bits >> 32) | ((unsigned long long)f << 32);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ bool ll_load(const ulonglong2* p, unsigned f, double& v) {
  unsigned long long a, b;
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  v = __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
  return (unsigned)(a >> 32) == f && (unsigned)(b >> 32) == f;
}
__device__ __forceinline__ void p2p_ll_store(const P2PDev& p, int par, int64_t idx, double v, unsigned f) {
#pragma unroll
  for (int r = 0; r < P2P_MAXRANKS; ++r)
    if (r < p.nranks) ll_store(p.ll_slot(r, par, p.rank) + idx, v, f);
}
// sum over ranks (rank order) of value idx of this exchange; spins until every rank's copy has arrived
__device__ __forceinline__ double p2p_ll_sum(const P2PDev& p, int par, int64_t idx, unsigned f) {
  double s = 0.0;
  const long long t0 = clock64();
  for (int r = 0; r < p.nranks; ++r) {
    const ulonglong2* w = p.ll_slot(p.rank, par, r) + idx;
    double v;
    while (!ll_load(w, f, v)) {
      if (clock64() - t0 > 4000000000LL) {   // ~2 s: a peer is gone; abort the kernel (loud, and nobody hangs at a grid barrier)
        *p.err = 1;
        __trap();
      }
    }
    s += v;
  }
  return s;
}

// ---- generic two-kernel form: buf (count doubles, this rank's device memory) <- sum over ranks ----------
__global__ void __launch_bounds__(256) p2p_push_kernel(P2PDev p, const double* __restrict__ buf, int64_t count, const int* done) {
  if ((done && *done) || *p.err) return;
  const unsigned long long s = *p.seq;
  const int par = (int)(s & 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    p2p_store(p, par, i, buf[i]);
  p2p_signal(p, par, s, gridDim.x);
}

// one CTA per 256 outputs; CTA 0 bumps the counter after a grid-wide ticket
__global__ void __launch_bounds__(256) p2p_wait_sum_kernel(P2PDev p, double* __restrict__ buf, int64_t count, const int* done,
                                                           unsigned* ticket2) {
  if ((done && *done) || *p.err) return;
  const unsigned long long s = *p.seq;
  const int par = (int)(s & 1);
  const bool ok = p2p_wait(p, par, s);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ok && i < count) {
    double acc = 0.0;
    for (int r = 0; r < p.nranks; ++r) acc += __ldcg(p.slot(p.rank, par, r) + i);   // rank order: same sum everywhere
    buf[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(ticket2, 1u);
    if (t == gridDim.x - 1) {       // every CTA has read *p.seq and its slots
      *ticket2 = 0;
      *p.seq = s + 1;
    }
  }
}

// ---- generic form on the flag-in-data words: no fence, no flag, no election ---------------------------------------
__global__ void __launch_bounds__(256) p2p_ll_push_kernel(P2PDev p, const double* __restrict__ buf, int64_t count, const int* done) {
  if ((done && *done) || *p.err) return;
  const unsigned long long s = *p.seq;
  const int par = (int)(s & 1);
  const unsigned fl = (unsigned)(s + 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    p2p_ll_store(p, par, i, buf[i], fl);
}

// buf[i] <- sum over ranks (rank order) of value i of this exchange; the last CTA bumps the exchange counter
__global__ void __launch_bounds__(256) p2p_ll_gather_kernel(P2PDev p, double* __restrict__ buf, int64_t count, const int* done,
                                                            unsigned* ticket2) {
  if ((done && *done) || *p.err) return;
  const unsigned long long s = *p.seq;
  const int par = (int)(s & 1);
  const unsigned fl = (unsigned)(s + 1);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    double acc = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < p.nranks; ++r) {
      const ulonglong2* w = p.ll_slot(p.rank, par, r) + i;
      double v;
      bool ok;
      while (!(ok = ll_load(w, fl, v)) && clock64() - t0 <= 4000000000LL) {}
      if (!ok) { *p.err = 1; v = 0.0; }          // a peer is gone: the host raises after the loop (p2p_check)
      acc += v;
    }
    buf[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(ticket2, 1u);
    if (t == gridDim.x - 1) {
      *ticket2 = 0;
      *p.seq = s + 1;
    }
  }
}

}  // namespace admmb200
