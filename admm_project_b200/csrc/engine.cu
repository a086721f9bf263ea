// engine.cu -- libadmm_b200.so: handle, one-time setup (Gram / Cholesky / inverse factor), the
// device-resident ADMM loop and the C-ABI of include/admm_b200.h.  sm_100a only, no CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <cstring>
#include <sched.h>
#include <thread>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <vector>

#include "../../include/admm_b200.h"
#include "chol.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "gemv.cuh"
#include "prox.cuh"
#include "tri.cuh"
#include "gemvt.cuh"
#include "unwrapped.cuh"
#include "uwbatch.cuh"
#include "onepass.cuh"
#include "tv.cuh"
#include "p2p.cuh"
#include "symtri.cuh"
#include "persist.cuh"
#include "persist_batch.cuh"
#include "gemm_tma.cuh"
#include <dlfcn.h>

namespace admmb200 {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static inline int64_t tv_stride(int64_t n);
// cudaFuncSetAttribute state is per DEVICE: a process that drives several GPUs (one handle each) must
// configure every kernel once per device, so the 'already configured' marks are kept per device.
struct PerDevice {
  size_t v[64] = {};
  size_t& operator()(int dev) { return v[dev & 63]; }
};

// a device buffer the handle owns
struct DBuf {
  double* p = nullptr;
  int64_t cap = 0;  // doubles
  void ensure(int64_t n) {
    if (n <= cap) return;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    ADMM_CUDA(cudaMalloc(&p, (size_t)std::max<int64_t>(n, 1) * sizeof(double)));
    cap = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct ColdotPlan {
  int mode = 0;
  int64_t rows = 0, cols = 0;
  int grid = 0, panels = 1, max_pos = 1, nitems = 0, npos = 0;
  int *d_cta_pos = nullptr, *d_pos_item = nullptr, *d_order = nullptr;
  ColdotItem* d_items = nullptr;
};

// round plan of the one-pass x-update (symtri.cuh) for one (k, column range)
struct SymtriPlan {
  int64_t k = 0, c_lo = 0, c_hi = 0;
  int grid = 0;
  int *d_cta_round = nullptr, *d_round_ent = nullptr;
  SymtriEnt* d_ents = nullptr;
};

}  // namespace admmb200

using namespace admmb200;

struct admm_b200_handle {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, stream2 = nullptr, stream3 = nullptr, stream4 = nullptr, stream_hi = nullptr;   // stream_hi / 2 / 3: Cholesky look-ahead (high / middle / low priority); 4: the inverse factor trailing it (low)
  std::vector<cudaEvent_t> ev_pool;   // events of the look-ahead Cholesky (created on first use)
  DBuf chol_ws;                       // out-of-place panel solves of the look-ahead Cholesky
  DBuf inv_ws, inv_ws2;               // L[P, :P] * W[:P, :P] of the inverse factor's block row P; its split-K workspace
  // pinned staging ring for uploads from PAGEABLE host memory (host threads gather a row panel into it, the DMA
  // engine takes it from there): allocated on the first such upload
  static constexpr int kPinBufs = 3;
  double* pin_buf[kPinBufs] = {nullptr, nullptr, nullptr};
  cudaEvent_t pin_ev[kPinBufs] = {nullptr, nullptr, nullptr};
  size_t pin_cap = 0;                 // doubles per buffer
  int pin_next = 0;
  bool pin_failed = false;            // pinned allocation refused once: uploads go through the driver's own staging from then on
  cudaEvent_t ev_la[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp[4] = {nullptr, nullptr, nullptr, nullptr};
  double phase_ms[4] = {0, 0, 0, 0};  // gram (+Dts), cholesky, inverse factor (+transpose), total
  int64_t launches = 0;       // kernels launched (a graph replay counts the kernels it holds)
  int64_t graph_replays = 0;  // CUDA graph launches of the iteration loop

  // problem
  int kind = 0;
  int64_t m = 0, n = 0;          // D is m x n
  int64_t nA = 0, nB = 0, mc = 0;
  const double* dD = nullptr;    // device D (ldD)
  int64_t ldD = 0;
  DBuf ownD;                     // backing store when D came from the host
  DBuf s, dts;
  bool tall = true;
  double rho_setup = 1.0, lambda = 0.0;
  int xsolve = ADMM_B200_XSOLVE_INVFACTOR;
  double setup_ms = 0.0;

  // cached factor (k x k, leading dimension ldf, multiple of 16)
  int64_t k = 0, ldf = 0;
  DBuf L, W, WT;
  bool have_factor = false, have_inverse = false;
  int64_t subst_bs = 128;        // width of the inverted diagonal blocks W holds (512 after the look-ahead Cholesky)
  DBuf subst_t;                  // W_PP y_P of the substitution path

  // iterates and work vectors
  DBuf x, z, u, y, t1, t2;
  bool have_init = false;
  DBuf x0, z0, u0;

  // loop control
  LoopCtl* ctl = nullptr;
  LoopCtl* h_ctl = nullptr;      // pinned
  int* fail = nullptr;           // cholesky failure flag
  DBuf partials, hist;           // hist: 6 x cap
  int64_t hist_cap = 0;
  DBuf xvals, zvals, uvals;
  DBuf gemm_ws, gemv_ws, scratch, cd_ws;
  bool iter_ready = false;
  DBuf fv, fuhat, fzprev, fuprev;   // fast / accelerated ADMM state
  // A = D family (svm / huber / lad)
  DBuf aux, rvec, dzvec, cb, uw_partials, op_dpart;
  unsigned* grid_ticket = nullptr;
  int64_t m_total = 0;
  double svmC = 0.0;
  // quadratic objective with box / nonneg prox: P (full, for the objective), bounds, constant r
  DBuf Pfull, lb, ub;
  double qp_r = 0.0;
  // model problem: second matrix / vector, both Gram matrices (kept for rho changes), second factor
  DBuf Q2, s2, dts2, G1, G2, L2, W2, WT2, zsol;
  // total variation: double-buffered z/u, pivot table of the constant tridiagonal
  DBuf zz, uu, tvtab;
  int tv_par = 0, tv_ntab = 0, tv_halo = 0;
  bool tv_exact = false;         // rho too large for the windowed solve: segments chained through their aggregates (tv.cuh)
  DBuf tv_y, tv_seg;             // exact path: y (n) and [segA | segB | cin] (3 x nseg)
  double tv_inv_star = 0.0;
  double tv_rho = -1.0;
  // row-sharded runs: one NCCL communicator per handle (one process per GPU)
  void* comm = nullptr;
  int rank = 0, nranks = 1;
  // one-shot peer-memory allreduce over NVLink (p2p.cuh): this rank's mailbox + the peers' mapped mailboxes
  P2PState p2p;
  bool lasso_sharded = false;    // lasso set up from row shards (Gram allreduced, iterations replicated)
  int64_t generation = 0;        // bumped by every setup_*: stale (minx, minz) pairs are rejected by the host mirror
  double diag_ratio = 1.0;       // (max L_ii / min L_ii)^2 of the last factor: cheap lower bound of cond(A)
  int xsolve_eff = ADMM_B200_XSOLVE_INVFACTOR;   // x-update realisation actually used (SUBST when the guard fired)
  int64_t zero_cols = 0;         // A = D problems: all-zero columns of D handled as pinv does (x_j = 0)
  std::vector<ColdotPlan*> plans;
  DBuf Qm, tcur, tlast;          // persistent A = D iteration (persist.cuh): Q_g = D_g inv(R)', t vectors
  bool have_Q = false;
  int64_t ldq = 0;
  std::vector<SymtriPlan*> st_plans;
  DBuf st_part;                  // per-CTA partial x of the one-pass x-update
  bool xshard = false;           // lasso from row shards with a large factor: the x-update is split by columns over the ranks
  unsigned* tickets = nullptr;
  int64_t tickets_cap = 0;
};

namespace admmb200 {

// run a scope's launches on another stream of the handle; the handle's stream is restored when the scope ends
struct StreamSwap {
  admm_b200_handle* h;
  cudaStream_t saved;
  StreamSwap(admm_b200_handle* hh, cudaStream_t s) : h(hh), saved(hh->stream) { hh->stream = s; }
  ~StreamSwap() { h->stream = saved; }
};

static void allreduce_sum(admm_b200_handle* h, double* buf, int64_t count, const int* done = nullptr);

static void check_handle(admm_b200_handle* h) {
  ADMM_REQUIRE(h != nullptr, ADMM_B200_ERR_INVALID, "null handle");
  ADMM_CUDA(cudaSetDevice(h->device));
}

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ordinary (pageable) host memory: the driver would stage such a copy through its own bounce buffer with one thread
// (~12 GB/s measured for the 4.3 GB matrix of C2, against ~55 GB/s from pinned memory)
static bool is_pageable_host_ptr(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return at.type == cudaMemoryTypeUnregistered;
}

// one column segment into the pinned staging buffer.  Non-temporal stores: the buffer is only read by the DMA engine, so
// pulling its lines into the cache first (write-allocate) would cost a third of the host memory traffic for nothing.
static inline void copy_segment_nt(double* dst, const double* src, size_t count) {
#if defined(__x86_64__)
  if ((((uintptr_t)dst) & 15) == 0) {
    size_t i = 0;
    for (; i + 8 <= count; i += 8) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
      const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 2));
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 4));
      const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 6));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 2), b);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 4), c);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 6), d);
    }
    for (; i < count; ++i) dst[i] = src[i];
    return;
  }
#endif
  memcpy(dst, src, count * 8);
}

// Rows [r0, r0 + rows) of the column-major HOST matrix D (n columns, ldD) -> dst + r0 (device, leading dimension ld), on
// `stream`.  Pinned source: one strided DMA.  Pageable source: sub-panels are gathered by a few host threads into a ring
// of pinned buffers (column segments of a sub-panel are contiguous in the buffer) and the DMA engine copies from there
// while the threads fill the next buffer -- the upload runs at the PCIe rate instead of the driver's bounce-buffer rate.
static void upload_rows(admm_b200_handle* h, double* dst, int64_t ld, const double* D, int64_t ldD, int64_t r0, int64_t rows,
                        int64_t n, cudaStream_t stream, bool pageable) {
  if (rows <= 0 || n <= 0) return;
  const bool no_stage = getenv("ADMM_B200_NO_PINNED_STAGING") != nullptr;   // read per call (tests flip it)
  const size_t bytes = (size_t)rows * (size_t)n * 8;
  if (!pageable || no_stage || bytes < ((size_t)32 << 20)) {
    ADMM_CUDA(cudaMemcpy2DAsync(dst + r0, (size_t)ld * 8, D + r0, (size_t)ldD * 8, (size_t)rows * 8, (size_t)n,
                                cudaMemcpyHostToDevice, stream));
    return;
  }
  constexpr size_t kBufDoubles = (size_t)16 << 20;          // 128 MB per buffer
  if (!h->pin_buf[0] && !h->pin_failed) {
    for (int b = 0; b < admm_b200_handle::kPinBufs && !h->pin_failed; ++b) {
      if (cudaHostAlloc((void**)&h->pin_buf[b], kBufDoubles * 8, cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->pin_ev[b], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();                         // no pinned memory to be had (locked-memory limit): the driver's path still works
        h->pin_failed = true;
      }
    }
    if (h->pin_failed) {
      for (int b = 0; b < admm_b200_handle::kPinBufs; ++b) {
        if (h->pin_buf[b]) cudaFreeHost(h->pin_buf[b]);
        if (h->pin_ev[b]) cudaEventDestroy(h->pin_ev[b]);
        h->pin_buf[b] = nullptr; h->pin_ev[b] = nullptr;
      }
    } else {
      h->pin_cap = kBufDoubles;
    }
  }
  if (h->pin_failed || (size_t)n > h->pin_cap) {
    ADMM_CUDA(cudaMemcpy2DAsync(dst + r0, (size_t)ld * 8, D + r0, (size_t)ldD * 8, (size_t)rows * 8, (size_t)n,
                                cudaMemcpyHostToDevice, stream));
    return;
  }
  int64_t sub = std::max<int64_t>(1, (int64_t)(h->pin_cap / (size_t)n));
  if (sub >= 512) sub = sub / 512 * 512;                     // whole pages per column segment
  unsigned hw = std::max(2u, std::thread::hardware_concurrency());
  {
    cpu_set_t cpus;                                  // the cores this process may run on (containers, taskset)
    if (sched_getaffinity(0, sizeof(cpus), &cpus) == 0 && CPU_COUNT(&cpus) > 0) hw = std::max(2u, (unsigned)CPU_COUNT(&cpus));
  }
  const int nthreads = (int)std::min<unsigned>(8u, std::max(2u, hw / (2u * (unsigned)std::max(1, h->nranks))));
  for (int64_t s0 = 0; s0 < rows; s0 += sub) {
    const int64_t sr = std::min(sub, rows - s0);
    const int b = h->pin_next;
    h->pin_next = (h->pin_next + 1) % admm_b200_handle::kPinBufs;
    ADMM_CUDA(cudaEventSynchronize(h->pin_ev[b]));           // the DMA that last read this buffer is done
    double* stage = h->pin_buf[b];
    const double* src = D + r0 + s0;
    auto work = [=](int t) {
      const int64_t j0 = n * t / nthreads, j1 = n * (t + 1) / nthreads;
      for (int64_t j = j0; j < j1; ++j) copy_segment_nt(stage + j * sr, src + j * ldD, (size_t)sr);
#if defined(__x86_64__)
      _mm_sfence();
#endif
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    ADMM_CUDA(cudaMemcpy2DAsync(dst + r0 + s0, (size_t)ld * 8, stage, (size_t)sr * 8, (size_t)sr * 8, (size_t)n,
                                cudaMemcpyHostToDevice, stream));
    ADMM_CUDA(cudaEventRecord(h->pin_ev[b], stream));
  }
}

// copy `count` doubles from a host-or-device pointer into a device buffer (async on the stream)
static void copy_in(admm_b200_handle* h, double* dst, const double* src, int64_t count) {
  ADMM_CUDA(cudaMemcpyAsync(dst, src, (size_t)count * 8, cudaMemcpyDefault, h->stream));
}
static void copy_out(admm_b200_handle* h, double* dst, const double* src, int64_t count) {
  if (!dst || count <= 0) return;
  ADMM_CUDA(cudaMemcpyAsync(dst, src, (size_t)count * 8, cudaMemcpyDefault, h->stream));
}

static void ensure_tickets(admm_b200_handle* h, int64_t n) {
  if (n <= h->tickets_cap) return;
  if (h->tickets) cudaFree(h->tickets);
  h->tickets = nullptr;
  int64_t cap = std::max<int64_t>(n, 4096);
  ADMM_CUDA(cudaMalloc(&h->tickets, (size_t)cap * sizeof(unsigned)));
  ADMM_CUDA(cudaMemsetAsync(h->tickets, 0, (size_t)cap * sizeof(unsigned), h->stream));
  h->tickets_cap = cap;
}

// ---------------------------------------------------------------------------------------------
// GEMM launcher
// ---------------------------------------------------------------------------------------------
template <bool AK, bool BK, int VEC, int BN, int BM = 128>
static void gemm_launch_t(admm_b200_handle* h, const GemmArgs& g, dim3 grid) {
  static PerDevice configured_pd;
  size_t& configured = configured_pd(h->device);
  constexpr int smem = GemmCfg<BM>::SMEM_BYTES;
  if (!configured) {
    ADMM_CUDA(cudaFuncSetAttribute(gemm_f64_dmma_kernel<AK, BK, VEC, BN, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = 1;
  }
  gemm_f64_dmma_kernel<AK, BK, VEC, BN, BM><<<grid, GEMM_THREADS, smem, h->stream>>>(g);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
}

// ---- TMA descriptors (driver entry point fetched through the runtime: no -lcuda) ---------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encode_fn() {
  static TmapEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (TmapEncodeFn)p;
    else cudaGetLastError();
  }
  return fn;
}
// K-major operand tiles of a column-major matrix (rows = K, contiguous): box 16 (k) x 128 (columns), 128-byte swizzle
static bool make_kmajor_tmap(CUtensorMap* tm, const double* base, int64_t K, int64_t cols, int64_t ld) {
  TmapEncodeFn enc = tmap_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)cols};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)GT_BK, 128u};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct GemmOpt {
  int lower_only = 0;
  double diag_add = 0.0;
  int batch = 1;
  int64_t strideA = 0, strideB = 0, strideC = 0;
  int allow_splitk = 1;
  int a_lower = 0, b_lower = 0, a_upper = 0, b_upper = 0;
  int inplace = 0;        // C aliases A: only tilings whose CTAs read nothing another CTA writes (N <= one 128-wide tile column)
  int max_ctas = 0;       // > 0: at most this many CTAs, each walking several tiles (background work of the look-ahead Cholesky)
  int64_t k_chunk = 0;    // > 0: split K into chunks of this length whatever the tile count (short-lived CTAs: a background
                          //      product must not hold SMs for a millisecond while the chain waits for one)
  DBuf* ws = nullptr;     // split-K workspace to use instead of the handle's shared one (products on side streams)
};

static void gemm(admm_b200_handle* h, int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                 const GemmOpt& o = GemmOpt()) {
  if (M <= 0 || N <= 0) return;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.alpha = alpha; g.beta = beta; g.diag_add = o.diag_add;
  g.lower_only = o.lower_only;
  g.a_lower = o.a_lower; g.b_lower = o.b_lower; g.a_upper = o.a_upper; g.b_upper = o.b_upper;
  g.batch = o.batch; g.strideA = o.strideA; g.strideB = o.strideB; g.strideC = o.strideC;
  g.splits = 1; g.k_per_split = std::max<int64_t>(K, 1); g.ws = nullptr;
  g.bm = GEMM_BM; g.tri_skip = 0;
  g.persist = 0; g.tm = g.tn = 0; g.nz = 1;
  const bool skinny = (N <= 64) && !o.lower_only;   // 128 x 64 tiles for few right-hand sides
  // tall triangular op(A) x few right-hand sides (the lambda batch on a large factor): 256 x 64 tiles, uniform K chunks
  static const bool no_tall = getenv("ADMM_B200_NO_TALL_TRI") != nullptr;
  const bool tall_tri = skinny && !no_tall && o.batch == 1 && M >= 2048 && K >= 2048 && (o.a_lower || o.a_upper) && o.allow_splitk &&
                        (((uintptr_t)A & 15) == 0) && (((uintptr_t)B & 15) == 0) && (lda % 2 == 0) && (ldb % 2 == 0);
  // small outputs (the chain of the look-ahead Cholesky, factors of a few hundred columns): 64 x 64 tiles put four times
  // as many SMs to work on a product that would otherwise keep a handful of them busy for 4x as long
  static const bool no_small = getenv("ADMM_B200_NO_SMALL_TILES") != nullptr;
  const int64_t t128 = ((M + 127) / 128) * ((N + 127) / 128) * o.batch;
  const bool small = !no_small && !skinny && !tall_tri && !o.inplace && t128 <= 40 && M >= 64 && N >= 64 &&
                     (((uintptr_t)A & 15) == 0) && (((uintptr_t)B & 15) == 0) && (lda % 2 == 0) && (ldb % 2 == 0) &&
                     (o.strideA % 2 == 0) && (o.strideB % 2 == 0);
  // a rank's share of a lambda batch split over several GPUs: 32- or 16-wide tiles (K-major right-hand sides, 16-byte path)
  const int narrow = (tall_tri && transb == 0 && !getenv("ADMM_B200_NO_NARROW_TRI")) ? (N <= 16 ? 16 : (N <= 32 ? 32 : 0)) : 0;
  const int64_t bnsz = narrow ? narrow : ((skinny || small) ? 64 : GEMM_BN);
  const int64_t bmsz = tall_tri ? 256 : (small ? 64 : GEMM_BM);
  if (small) g.bm = 64;
  const int64_t tm = (M + bmsz - 1) / bmsz, tn = (N + bnsz - 1) / bnsz;
  int64_t tiles = o.lower_only ? tm * (tm + 1) / 2 : tm * tn;
  tiles *= o.batch;
  int64_t want_splits = 1;
  if (tall_tri) {
    static const int64_t chunk = getenv("ADMM_B200_TRI_CHUNK") ? atoll(getenv("ADMM_B200_TRI_CHUNK")) : 1024;
    want_splits = (K + chunk - 1) / chunk;
    g.bm = 256; g.tri_skip = want_splits > 1;
  } else if (o.allow_splitk && tiles < kNumSM && K >= 4096) {
    // few output tiles (n = 784 / 1024 Gram of a tall D): fill the machine twice over
    want_splits = std::min<int64_t>((2 * kNumSM + tiles - 1) / tiles, K / 1024);
  } else if (o.allow_splitk && tiles * 8 <= kNumSM && K >= 256) {
    // a handful of tiles and a short K (the n x nb x n products of the class / lambda batches on a small
    // factor): 7 CTAs marching through K one after the other is pure latency; split K ~64 wide
    want_splits = std::min<int64_t>(kNumSM / tiles, K / 64);
  } else if (o.allow_splitk && tiles >= kNumSM && K >= 16384 && !getenv("ADMM_B200_NO_TAIL_SPLIT")) {
    // wave quantisation: 2080 tiles on 148 SMs = 14.05 waves -> 15 (6.7% idle).  Splitting K by S
    // makes the work items S times smaller, so the idle tail shrinks to ~1/S of a tile time.
    auto waste = [&](int64_t S) { int64_t items = tiles * S; return (double)((items + kNumSM - 1) / kNumSM * kNumSM) / (double)items; };
    double best = waste(1);
    for (int64_t S = 2; S <= 4; ++S)
      if (waste(S) < best - 0.01 && (int64_t)o.batch * S * M * N * 8 <= (int64_t)4 << 30) { best = waste(S); want_splits = S; }
  }
  if (o.k_chunk > 0 && K > o.k_chunk && !tall_tri) want_splits = (K + o.k_chunk - 1) / o.k_chunk;
  if (want_splits > 1) {
    int64_t splits = want_splits;
    if (splits > 1) {
      g.k_per_split = round_up((K + splits - 1) / splits, GEMM_BK);
      g.splits = (int)((K + g.k_per_split - 1) / g.k_per_split);
      DBuf& wsb = o.ws ? *o.ws : h->gemm_ws;
      wsb.ensure((int64_t)o.batch * g.splits * M * N);
      g.ws = wsb.p;
    }
  }
  const bool vec_ok = (((uintptr_t)A & 15) == 0) && (((uintptr_t)B & 15) == 0) && (lda % 2 == 0) && (ldb % 2 == 0) &&
                      (o.strideA % 2 == 0) && (o.strideB % 2 == 0);
  dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)(o.batch * g.splits));
  if (o.max_ctas > 0 && tm * tn * o.batch * g.splits > o.max_ctas) {
    g.persist = o.max_ctas; g.tm = tm; g.tn = tn; g.nz = o.batch * g.splits;
    grid = dim3((unsigned)o.max_ctas, 1u, 1u);
  }
  // Gram-shaped products (both operands K-major, i.e. A'*B on column-major matrices) go through the TMA-fed kernel
  static const bool no_tma = getenv("ADMM_B200_NO_TMA") != nullptr;
  if (!no_tma && !g.persist && transa != 0 && transb == 0 && vec_ok && o.batch == 1 && !skinny && !tall_tri && !small && !o.a_lower && !o.b_lower &&
      !o.a_upper && !o.b_upper && M >= 128 && N >= 128 && K >= 64 && M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31)) {
    CUtensorMap tmA, tmB;
    if (make_kmajor_tmap(&tmA, A, K, M, lda) && make_kmajor_tmap(&tmB, B, K, N, ldb)) {
      static PerDevice configured_pd;
      size_t& configured = configured_pd(h->device);
      if (!configured) {
        ADMM_CUDA(cudaFuncSetAttribute(gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES));
        configured = 1;
      }
      GemmTmaArgs t;
      t.M = M; t.N = N; t.K = K; t.C = C; t.ldc = ldc; t.alpha = alpha; t.beta = beta; t.diag_add = o.diag_add;
      t.lower_only = o.lower_only; t.splits = g.splits; t.k_per_split = g.k_per_split; t.ws = g.ws;
      gemm_tma_kernel<<<grid, GT_THREADS, GT_SMEM_BYTES, h->stream>>>(tmA, tmB, t);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
      if (g.splits > 1) {
        int64_t total = M * N;
        dim3 rg((unsigned)std::min<int64_t>((total + 255) / 256, 4 * kNumSM), 1u);
        gemm_splitk_reduce_kernel<<<rg, 256, 0, h->stream>>>(g);
        ADMM_CUDA(cudaGetLastError());
        h->launches++;
      }
      return;
    }
  }
  const bool AK = transa != 0;  // 'T': op(A)[i,k] = A[k + i*lda]  (K contiguous)
  const bool BK = transb == 0;  // 'N': op(B)[k,j] = B[k + j*ldb]  (K contiguous)
#define ADMM_GEMM_CASE(a, b)                                     \
  if (AK == a && BK == b) {                                      \
    if (tall_tri && narrow == 16 && b) gemm_launch_t<a, true, 2, 16, 256>(h, g, grid);   \
    else if (tall_tri && narrow == 32 && b) gemm_launch_t<a, true, 2, 32, 256>(h, g, grid);   \
    else if (tall_tri) gemm_launch_t<a, b, 2, 64, 256>(h, g, grid);   \
    else if (small) gemm_launch_t<a, b, 2, 64, 64>(h, g, grid);  \
    else if (skinny) {                                           \
      if (vec_ok) gemm_launch_t<a, b, 2, 64>(h, g, grid);        \
      else gemm_launch_t<a, b, 1, 64>(h, g, grid);               \
    } else if (vec_ok) gemm_launch_t<a, b, 2, 128>(h, g, grid);  \
    else gemm_launch_t<a, b, 1, 128>(h, g, grid);                \
  }
  ADMM_GEMM_CASE(true, true)
  ADMM_GEMM_CASE(true, false)
  ADMM_GEMM_CASE(false, true)
  ADMM_GEMM_CASE(false, false)
#undef ADMM_GEMM_CASE
  if (g.splits > 1) {
    int64_t total = M * N;
    dim3 rg((unsigned)std::min<int64_t>((total + 255) / 256, 4 * kNumSM), (unsigned)o.batch);
    gemm_splitk_reduce_kernel<<<rg, 256, 0, h->stream>>>(g);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  }
}

static void symmetrize(admm_b200_handle* h, double* C, int64_t n, int64_t ldc) {
  unsigned t = (unsigned)((n + 31) / 32);
  symmetrize_lower_kernel<<<dim3(t, t), dim3(32, 8), 0, h->stream>>>(C, n, ldc);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
}

static void transpose(admm_b200_handle* h, const double* in, int64_t rows, int64_t cols, int64_t ldi, double* out,
                      int64_t ldo) {
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(in, rows, cols, ldi, out, ldo);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
}

// ---------------------------------------------------------------------------------------------
// blocked Cholesky (+ inverse factor by recursive doubling)
// ---------------------------------------------------------------------------------------------
// A: k x k, lower triangle holds the SPD matrix; on exit lower(A) = L, strict upper zero.
// W (k x k, ldw) zero-filled by the caller's allocation step here; on exit W = inv(L) when
// want_inverse, otherwise only its diagonal blocks are valid.
static void potrf_blocked_v1(admm_b200_handle* h, int64_t k, double* A, int64_t lda, double* W, int64_t ldw,
                          bool want_inverse) {
  static PerDevice configured_pd;
  size_t& configured = configured_pd(h->device);
  if (!configured) {
    ADMM_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_DIAG_SMEM));
    configured = 1;
  }
  ADMM_CUDA(cudaMemsetAsync(h->fail, 0, sizeof(int), h->stream));
  ADMM_CUDA(cudaMemset2DAsync(W, (size_t)ldw * 8, 0, (size_t)k * 8, (size_t)k, h->stream));
  cudaStream_t sA = h->stream, sB = h->stream2;
  ADMM_CUDA(cudaEventRecord(h->ev_la[0], sA));
  ADMM_CUDA(cudaStreamWaitEvent(sB, h->ev_la[0], 0));
  // two-level right-looking blocking: 512-wide outer panels (so the big trailing update runs with
  // K = 512 and near-Gram efficiency), 128-wide inner steps inside a panel.
  for (int64_t K0 = 0; K0 < k; K0 += CHOL_NBO) {
    const int64_t wb = std::min<int64_t>(CHOL_NBO, k - K0);
    for (int64_t k0 = K0; k0 < K0 + wb; k0 += CHOL_NB) {
      const int nb = (int)std::min<int64_t>(CHOL_NB, K0 + wb - k0);
      double* A11 = A + k0 + k0 * lda;
      double* W11 = W + k0 + k0 * ldw;
      potrf_diag_kernel<<<1, CHOL_DIAG_THREADS, CHOL_DIAG_SMEM, h->stream>>>(A11, lda, nb, W11, ldw, h->fail, (int)k0);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
      const int64_t rem = k - k0 - nb;
      if (rem <= 0) continue;
      double* A21 = A + (k0 + nb) + k0 * lda;
      // L21 = A21 * inv(L11)'.  In place: N = nb <= one tile column, so every CTA reads only the
      // rows it later writes, after its whole K loop.
      GemmOpt po;
      po.allow_splitk = 0;
      po.inplace = 1;
      gemm(h, 0, 1, rem, nb, nb, 1.0, A21, lda, W11, ldw, 0.0, A21, lda, po);
      // inner trailing update, remaining columns of this outer panel only
      const int64_t pc = K0 + wb - k0 - nb;
      if (pc > 0) {
        GemmOpt to;
        to.lower_only = 1;
        to.allow_splitk = 0;
        gemm(h, 0, 1, rem, pc, nb, -1.0, A21, lda, A21, lda, 1.0, A + (k0 + nb) + (k0 + nb) * lda, lda, to);
      }
    }
    const int64_t rem2 = k - K0 - wb;
    if (rem2 > 0) {  // A22 -= L21 * L21'  (lower tiles), K = wb
      // Look-ahead on a second stream: first the columns of the NEXT outer panel (so its latency-bound
      // factorisation chain can start), then the rest of the trailing matrix, which overlaps that chain.
      double* P = A + (K0 + wb) + K0 * lda;
      const int64_t K1 = K0 + wb, n1 = std::min<int64_t>(CHOL_NBO, rem2);
      GemmOpt to;
      to.lower_only = 1;
      to.allow_splitk = 0;
      ADMM_CUDA(cudaEventRecord(h->ev_la[0], sA));
      ADMM_CUDA(cudaStreamWaitEvent(sB, h->ev_la[0], 0));
      {
        StreamSwap on_b(h, sB);   // restores h->stream on every exit path (a throwing gemm must not leave the handle on stream2)
        gemm(h, 0, 1, rem2, n1, wb, -1.0, P, lda, P, lda, 1.0, A + K1 + K1 * lda, lda, to);
        ADMM_CUDA(cudaEventRecord(h->ev_la[1], sB));
        if (rem2 > n1) {
          double* P2 = P + n1;   // rows K1+n1.. of the panel
          gemm(h, 0, 1, rem2 - n1, rem2 - n1, wb, -1.0, P2, lda, P2, lda, 1.0, A + (K1 + n1) + (K1 + n1) * lda, lda, to);
        }
      }
      ADMM_CUDA(cudaStreamWaitEvent(sA, h->ev_la[1], 0));
    }
  }
  ADMM_CUDA(cudaEventRecord(h->ev_la[0], sB));
  ADMM_CUDA(cudaStreamWaitEvent(sA, h->ev_la[0], 0));
  {
    dim3 grid((unsigned)((k + 255) / 256), (unsigned)k);
    zero_upper_kernel<<<grid, 256, 0, h->stream>>>(A, k, lda);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  }
  int fail = 0;
  ADMM_CUDA(cudaMemcpyAsync(&fail, h->fail, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  ADMM_REQUIRE(fail == 0, ADMM_B200_ERR_NOTPOSDEF, "Matrix must be positive definite. (pivot %d is not positive)", fail);
  ADMM_CUDA(cudaEventRecord(h->evp[2], h->stream));
  if (!want_inverse) return;
  // inverse factor by recursive doubling: for [[A,0],[B,C]]: inv = [[Ai,0],[-Ci*B*Ai, Ci]]
  DBuf& T = h->scratch;  // T_p = B_p * inv(A_p), one b x b block per pair; sized once for all levels
  {
    int64_t need = 1;
    for (int64_t b = CHOL_NB; b < k; b *= 2) need = std::max(need, (k / (2 * b) + 1) * b * b);
    T.ensure(need);
  }
  for (int64_t b = CHOL_NB; b < k; b *= 2) {
    const int64_t npairs_full = k / (2 * b);  // pairs whose second block is a full b
    if (npairs_full > 0) {
      GemmOpt o1;
      o1.batch = (int)npairs_full;
      o1.strideA = 2 * b * (lda + 1);
      o1.strideB = 2 * b * (ldw + 1);
      o1.strideC = b * b;
      o1.allow_splitk = 0;
      o1.b_lower = 1;
      // T_p = B_p * Ai_p      (b x b), Ai_p lower triangular
      gemm(h, 0, 0, b, b, b, 1.0, A + b, lda, W, ldw, 0.0, T.p, b, o1);
      GemmOpt o2;
      o2.batch = (int)npairs_full;
      o2.strideA = 2 * b * (ldw + 1);
      o2.strideB = b * b;
      o2.strideC = 2 * b * (ldw + 1);
      o2.allow_splitk = 0;
      o2.a_lower = 1;
      // W21_p = -Ci_p * T_p, Ci_p lower triangular
      gemm(h, 0, 0, b, b, b, -1.0, W + b + b * ldw, ldw, T.p, b, 0.0, W + b, ldw, o2);
    }
    const int64_t start = npairs_full * 2 * b;
    const int64_t cb = k - start - b;  // ragged last pair: second block has cb rows, 0 < cb < b
    if (cb > 0) {
      GemmOpt o, oa;
      o.allow_splitk = oa.allow_splitk = 0;
      o.b_lower = 1;
      oa.a_lower = 1;
      double* Tl = T.p + npairs_full * b * b;
      gemm(h, 0, 0, cb, b, b, 1.0, A + (start + b) + start * lda, lda, W + start + start * ldw, ldw, 0.0, Tl, cb, o);
      gemm(h, 0, 0, cb, b, cb, -1.0, W + (start + b) + (start + b) * ldw, ldw, Tl, cb, 0.0,
           W + (start + b) + start * ldw, ldw, oa);
    }
  }
}

// inverse of a lower-triangular k x k matrix by recursive doubling, starting from inverted b0 x b0 diagonal
// blocks already in W: for [[A,0],[B,C]]: inv = [[Ai,0],[-Ci*B*Ai, Ci]] (two batched triangular DMMA GEMMs per level)
static void tri_inverse_doubling(admm_b200_handle* h, int64_t k, const double* A, int64_t lda, double* W, int64_t ldw,
                                 int64_t b0, DBuf& T) {
  for (int64_t b = b0; b < k; b *= 2) {
    const int64_t npairs_full = k / (2 * b);  // pairs whose second block is a full b
    if (npairs_full > 0) {
      GemmOpt o1;
      o1.batch = (int)npairs_full;
      o1.strideA = 2 * b * (lda + 1);
      o1.strideB = 2 * b * (ldw + 1);
      o1.strideC = b * b;
      o1.allow_splitk = 0;
      o1.b_lower = 1;
      gemm(h, 0, 0, b, b, b, 1.0, A + b, lda, W, ldw, 0.0, T.p, b, o1);           // T_p = B_p * Ai_p
      GemmOpt o2;
      o2.batch = (int)npairs_full;
      o2.strideA = 2 * b * (ldw + 1);
      o2.strideB = b * b;
      o2.strideC = 2 * b * (ldw + 1);
      o2.allow_splitk = 0;
      o2.a_lower = 1;
      gemm(h, 0, 0, b, b, b, -1.0, W + b + b * ldw, ldw, T.p, b, 0.0, W + b, ldw, o2);   // W21_p = -Ci_p * T_p
    }
    const int64_t start = npairs_full * 2 * b;
    const int64_t cb = k - start - b;  // ragged last pair: second block has cb rows, 0 < cb < b
    if (cb > 0) {
      GemmOpt o, oa;
      o.allow_splitk = oa.allow_splitk = 0;
      o.b_lower = 1;
      oa.a_lower = 1;
      double* Tl = T.p + npairs_full * b * b;
      gemm(h, 0, 0, cb, b, b, 1.0, A + (start + b) + start * lda, lda, W + start + start * ldw, ldw, 0.0, Tl, cb, o);
      gemm(h, 0, 0, cb, b, cb, -1.0, W + (start + b) + (start + b) * ldw, ldw, Tl, cb, 0.0,
           W + (start + b) + start * ldw, ldw, oa);
    }
  }
}
static int64_t doubling_scratch(int64_t k, int64_t b0) {
  int64_t need = 1;
  for (int64_t b = b0; b < k; b *= 2) need = std::max(need, (k / (2 * b) + 1) * b * b);
  return need;
}

static void potrf_diag_launch(admm_b200_handle* h, double* A11, int64_t lda, int nb, double* W11, int64_t ldw, int64_t k0) {
  static PerDevice configured_pd;
  size_t& configured = configured_pd(h->device);
  if (!configured) {
    ADMM_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_DIAG_SMEM));
    configured = 1;
  }
  potrf_diag_kernel<<<1, CHOL_DIAG_THREADS, CHOL_DIAG_SMEM, h->stream>>>(A11, lda, nb, W11, ldw, h->fail, (int)k0);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
}

// Cholesky AND full inverse of one wb x wb (<= 512) diagonal block, everything on h->stream, no host sync:
// 128-wide steps (diag kernel, in-place panel solve, trailing update inside the block), then two doubling levels.
static void potrf_block512(admm_b200_handle* h, double* A, int64_t lda, int64_t wb, double* W, int64_t ldw, int64_t pivot_base,
                           DBuf& T) {
  for (int64_t k0 = 0; k0 < wb; k0 += CHOL_NB) {
    const int nb = (int)std::min<int64_t>(CHOL_NB, wb - k0);
    double* A11 = A + k0 + k0 * lda;
    double* W11 = W + k0 + k0 * ldw;
    potrf_diag_launch(h, A11, lda, nb, W11, ldw, pivot_base + k0);
    const int64_t rem = wb - k0 - nb;
    if (rem <= 0) continue;
    double* A21 = A + (k0 + nb) + k0 * lda;
    GemmOpt po;
    po.allow_splitk = 0;
    po.inplace = 1;
    gemm(h, 0, 1, rem, nb, nb, 1.0, A21, lda, W11, ldw, 0.0, A21, lda, po);       // in place: N <= one tile column
    GemmOpt to;
    to.lower_only = 1;
    to.allow_splitk = 0;
    gemm(h, 0, 1, rem, rem, nb, -1.0, A21, lda, A21, lda, 1.0, A + (k0 + nb) + (k0 + nb) * lda, lda, to);
  }
  tri_inverse_doubling(h, wb, A, lda, W, ldw, CHOL_NB, T);
}

// Look-ahead blocked Cholesky (+ inverse factor).  Outer panels of 512 columns.  Three streams:
//   A (critical chain)  chol + inverse of the 512 x 512 diagonal block P; the block row right below it:
//                       L21_top = A21_top * inv(L_PP)', and the NEXT diagonal block -= L21_top * L21_top'
//   B (panel)           the rest of the panel solve L21 = A21 * inv(L_PP)' as ONE DMMA GEMM (K = 512), then the
//                       update of the next block column A[., P+1] -= L21 * L21_top'
//   C (bulk)            the big trailing SYRK A[P+2.., P+2..] -= L21 * L21' (K = 512, near Gram efficiency)
// so the latency-bound chain of diagonal kernels never waits for bulk flops of its own panel (only for the bulk
// update of the panel before), and the bulk GEMMs keep the other SMs busy meanwhile.
static void potrf_lookahead(admm_b200_handle* h, int64_t k, double* A, int64_t lda, double* W, int64_t ldw, bool want_inverse) {
  ADMM_CUDA(cudaMemsetAsync(h->fail, 0, sizeof(int), h->stream));
  ADMM_CUDA(cudaMemset2DAsync(W, (size_t)ldw * 8, 0, (size_t)k * 8, (size_t)k, h->stream));
  constexpr int64_t NBO = CHOL_NBO;
  const int64_t npan = (k + NBO - 1) / NBO;
  while ((int64_t)h->ev_pool.size() < 6 * npan + 3) {
    cudaEvent_t e;
    ADMM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->ev_pool.push_back(e);
  }
  auto ev = [&](int kind, int64_t P) { return h->ev_pool[(size_t)(kind * npan + P)]; };   // 0 W, 1 top, 2 T, 3 col P+1, 4 bulk, 5 col P+2
  cudaEvent_t ev_start = h->ev_pool[(size_t)(6 * npan)], ev_end = h->ev_pool[(size_t)(6 * npan + 1)];
  cudaEvent_t ev_inv = h->ev_pool[(size_t)(6 * npan + 2)];
  // Look-ahead depth 2 (default): panel P's update of block column P+2 is a separate GEMM on stream B and the bulk starts
  // at column P+3, so the chain at panel Q waits for the bulk of panel Q-3, not Q-2 -- the bulk of the early panels takes
  // longer than a chain step, and with depth 1 the chain stalled on it every panel (tools/chol_probe.py).
  static const int depth = getenv("ADMM_B200_CHOL_DEPTH1") ? 1 : 2;
  // The inverse factor rides one panel behind the factorisation on a fourth (low-priority) stream: block row P of
  // W = inv(L) is  W[P, :P] = -W_PP * (L[P, :P] * W[:P, :P]),  and L[P, :P] is final once panel P-1 has been solved,
  // so the big product runs while the chain factors diagonal block P and only the 512-wide product with W_PP comes
  // after it.  The n^3/3 flops of the inverse fill the SMs the latency-bound chain leaves idle instead of following it.
  static const bool inv_overlap = getenv("ADMM_B200_NO_INV_OVERLAP") == nullptr;
  // timing probe only (tools/chol_probe.py): leave the bulk trailing updates out to see the bare critical chain; the
  // factor is then WRONG and the pivot check is skipped
  static const bool probe_nobulk = getenv("ADMM_B200_CHOL_PROBE_NOBULK") != nullptr;
  const bool ride = want_inverse && inv_overlap;
  if (ride) {
    h->inv_ws.ensure(NBO * round_up(k, 2));
    const int64_t kmax = (npan - 1) * NBO;           // the last block row's product: sized once, no reallocation (= device sync) mid-way
    h->inv_ws2.ensure(((kmax + 1023) / 1024 + 1) * NBO * std::max<int64_t>(kmax, 1));
  }
  // Optional SM reserve for the chain (ADMM_B200_CHOL_RESERVE = number of SMs the background streams leave alone; the
  // background GEMMs then run as capped persistent grids, the cap split by the flops of the two streams at this panel).
  // Measured on B200 (profiles/r02_chol_probe.txt): it does NOT pay -- factor + inverse 19.5 ms without a reserve,
  // 23.1 / 24.7 / 26.4 / 28.6 ms with 12 / 24 / 36 / 48 SMs reserved: the chain waits on the bulk update of the panel
  // before (event 4), so slowing the background slows the chain more than free SMs speed it up.  Default: off.
  static const int reserve = getenv("ADMM_B200_CHOL_RESERVE") ? atoi(getenv("ADMM_B200_CHOL_RESERVE")) : 0;
  auto caps = [&](int64_t bulk_dim, int64_t inv_dim, int& cap_bulk, int& cap_inv) {
    cap_bulk = cap_inv = 0;
    if (reserve <= 0 || reserve >= kNumSM - 32) return;
    const int avail = kNumSM - reserve;
    const double b = (double)bulk_dim * (double)bulk_dim, v = ride ? (double)inv_dim * (double)inv_dim : 0.0;
    if (b + v <= 0.0) return;
    cap_bulk = (int)std::lround(avail * b / (b + v));
    cap_bulk = std::min(avail - 16, std::max(16, cap_bulk));
    cap_inv = avail - cap_bulk;
    if (v == 0.0) { cap_bulk = avail; cap_inv = 0; }
    if (b == 0.0) { cap_inv = avail; cap_bulk = 0; }
  };
  DBuf& T = h->scratch;
  T.ensure(std::max(doubling_scratch(k, NBO), doubling_scratch(NBO, CHOL_NB)));
  const int64_t lds = round_up(k, 2);
  h->chol_ws.ensure(lds * NBO + NBO * NBO);
  double* S2 = h->chol_ws.p;                 // bulk panel solve result (rows below the top block) x 512
  double* S1 = h->chol_ws.p + lds * NBO;     // top block solve result, 512 x 512
  cudaStream_t user = h->stream, sA = h->stream_hi, sB = h->stream2, sC = h->stream3, sD = h->stream4;
  ADMM_CUDA(cudaEventRecord(ev_start, user));
  ADMM_CUDA(cudaStreamWaitEvent(sA, ev_start, 0));
  ADMM_CUDA(cudaStreamWaitEvent(sB, ev_start, 0));
  ADMM_CUDA(cudaStreamWaitEvent(sC, ev_start, 0));
  ADMM_CUDA(cudaStreamWaitEvent(sD, ev_start, 0));
  {
  StreamSwap on_a(h, sA);                  // the chain: every launch below that does not name a stream goes to sA
  for (int64_t P = 0; P < npan; ++P) {
    const int64_t K0 = P * NBO, wb = std::min<int64_t>(NBO, k - K0), K1 = K0 + wb;
    double* APP = A + K0 + K0 * lda;
    double* WPP = W + K0 + K0 * ldw;
    // the diagonal block has received every update of the panels before: P-1's on stream A itself, and (depth 1) the bulk
    // of <= P-2, or (depth 2) the column-P+2 update of panel P-2 on stream B and the bulk of <= P-3
    if (depth == 1) {
      if (P >= 2) ADMM_CUDA(cudaStreamWaitEvent(sA, ev(4, P - 2), 0));
    } else {
      if (P >= 2) ADMM_CUDA(cudaStreamWaitEvent(sA, ev(5, P - 2), 0));
      if (P >= 3) ADMM_CUDA(cudaStreamWaitEvent(sA, ev(4, P - 3), 0));
    }
    if (ride && P >= 1) {
      // L[P, :P] is final: its last block came from the top solve of panel P-1 (stream A), the others from stream B
      StreamSwap on_d(h, sD);
      ADMM_CUDA(cudaStreamWaitEvent(sD, ev(1, P - 1), 0));
      if (P >= 2) ADMM_CUDA(cudaStreamWaitEvent(sD, ev(2, P - 2), 0));
      GemmOpt o1;
      o1.allow_splitk = 0;
      o1.b_lower = 1;
      {
        int cb, ci;
        caps(std::max<int64_t>(0, k - K0 - 2 * NBO), K0, cb, ci);   // the bulk running beside this product is panel P's
        o1.max_ctas = ci;
      }
      // K chunks of 1024: a 128 x 128 tile of this product with K = P * 512 would live up to a millisecond on its SM, and
      // the chain's kernels (high priority, but they need an EMPTY SM) queued behind such tiles: chain 11.2 -> 15.1 ms
      // with the unsplit product riding along (profiles/r02_chol_probe.txt)
      static const int64_t inv_chunk = getenv("ADMM_B200_INV_KCHUNK") ? atoll(getenv("ADMM_B200_INV_KCHUNK")) : 1024;
      o1.k_chunk = inv_chunk;
      o1.ws = &h->inv_ws2;
      gemm(h, 0, 0, wb, K0, K0, 1.0, A + K0, lda, W, ldw, 0.0, h->inv_ws.p, NBO, o1);
    }
    potrf_block512(h, APP, lda, wb, WPP, ldw, K0, T);
    ADMM_CUDA(cudaEventRecord(ev(0, P), sA));
    if (ride && P >= 1) {
      StreamSwap on_d(h, sD);
      ADMM_CUDA(cudaStreamWaitEvent(sD, ev(0, P), 0));
      GemmOpt o2;
      o2.allow_splitk = 0;
      o2.a_lower = 1;
      gemm(h, 0, 0, wb, K0, wb, -1.0, WPP, ldw, h->inv_ws.p, NBO, 0.0, W + K0, ldw, o2);
    }
    const int64_t rem = k - K1;
    if (rem <= 0) break;
    const int64_t n1 = std::min<int64_t>(NBO, rem), rem2 = rem - n1;
    const int64_t n2 = (depth == 2) ? std::min<int64_t>(NBO, rem2) : 0, rem3 = rem2 - n2;   // depth 2: block column P+2 apart
    const int64_t K2 = K1 + n1, K3 = K2 + n2;
    GemmOpt ts;                              // panel solve as a GEMM with the inverted diagonal block (upper-triangular op(B))
    ts.allow_splitk = 0;
    ts.b_upper = 1;
    // block row P+1 of block column P: updated by the column updates of panels P-1 / P-2 (stream B) and the bulk before
    if (P >= 1) ADMM_CUDA(cudaStreamWaitEvent(sA, ev(3, P - 1), 0));
    gemm(h, 0, 1, n1, wb, wb, 1.0, A + K1 + K0 * lda, lda, WPP, ldw, 0.0, S1, NBO, ts);
    ADMM_CUDA(cudaMemcpy2DAsync(A + K1 + K0 * lda, (size_t)lda * 8, S1, (size_t)NBO * 8, (size_t)n1 * 8, (size_t)wb,
                                cudaMemcpyDeviceToDevice, sA));
    {
      GemmOpt so;
      so.lower_only = 1;
      so.allow_splitk = 0;
      // next diagonal block: everything else that adds into it must be in before this rank-512 update lands on top
      if (depth == 1) {
        if (P >= 1) ADMM_CUDA(cudaStreamWaitEvent(sA, ev(4, P - 1), 0));
      } else {
        if (P >= 1) ADMM_CUDA(cudaStreamWaitEvent(sA, ev(5, P - 1), 0));
        if (P >= 2) ADMM_CUDA(cudaStreamWaitEvent(sA, ev(4, P - 2), 0));
      }
      gemm(h, 0, 1, n1, n1, wb, -1.0, A + K1 + K0 * lda, lda, A + K1 + K0 * lda, lda, 1.0, A + K1 + K1 * lda, lda, so);
    }
    ADMM_CUDA(cudaEventRecord(ev(1, P), sA));
    if (rem2 > 0) {
      double* A21b = A + K2 + K0 * lda;                  // rows below the top block, columns of panel P
      ADMM_CUDA(cudaStreamWaitEvent(sB, ev(0, P), 0));
      if (P >= depth + 1) ADMM_CUDA(cudaStreamWaitEvent(sB, ev(4, P - depth - 1), 0));
      {
        StreamSwap on_b(h, sB);
        gemm(h, 0, 1, rem2, wb, wb, 1.0, A21b, lda, WPP, ldw, 0.0, S2, lds, ts);
        ADMM_CUDA(cudaMemcpy2DAsync(A21b, (size_t)lda * 8, S2, (size_t)lds * 8, (size_t)rem2 * 8, (size_t)wb,
                                    cudaMemcpyDeviceToDevice, sB));
        ADMM_CUDA(cudaEventRecord(ev(2, P), sB));
        ADMM_CUDA(cudaStreamWaitEvent(sB, ev(1, P), 0));                 // L21_top
        if (P >= depth) ADMM_CUDA(cudaStreamWaitEvent(sB, ev(4, P - depth), 0));   // the last bulk that touched block column P+1
        GemmOpt co;
        co.allow_splitk = 0;
        gemm(h, 0, 1, rem2, n1, wb, -1.0, A21b, lda, A + K1 + K0 * lda, lda, 1.0, A + K2 + K1 * lda, lda, co);
        ADMM_CUDA(cudaEventRecord(ev(3, P), sB));
        if (depth == 2) {
          // block column P+2 (diagonal block included; its upper triangle is never read): after the bulk of panel P-1,
          // which adds into the same block column
          if (P >= 1) ADMM_CUDA(cudaStreamWaitEvent(sB, ev(4, P - 1), 0));
          gemm(h, 0, 1, rem2, n2, wb, -1.0, A21b, lda, A21b, lda, 1.0, A + K2 + K2 * lda, lda, co);
          ADMM_CUDA(cudaEventRecord(ev(5, P), sB));
        }
      }
      {
        StreamSwap on_c(h, sC);
        ADMM_CUDA(cudaStreamWaitEvent(sC, ev(2, P), 0));
        GemmOpt bo;
        bo.lower_only = 1;
        bo.allow_splitk = 0;
        {
          int cb, ci;
          caps(rem3, K1, cb, ci);                                    // the inverse row riding beside it is row P+1
          bo.max_ctas = cb;
        }
        if (!probe_nobulk && rem3 > 0) {
          // the rank-512 update in K pieces (ADMM_B200_BULK_PIECES): shorter-lived background CTAs free SMs for the chain sooner
          static const int pieces = getenv("ADMM_B200_BULK_PIECES") ? std::max(1, atoi(getenv("ADMM_B200_BULK_PIECES"))) : 1;
          const int64_t step = round_up((wb + pieces - 1) / pieces, 16);
          for (int64_t kk = 0; kk < wb; kk += step) {
            const int64_t kw = std::min(step, wb - kk);
            const double* Lp = A + K3 + (K0 + kk) * lda;
            gemm(h, 0, 1, rem3, rem3, kw, -1.0, Lp, lda, Lp, lda, 1.0, A + K3 + K3 * lda, lda, bo);
          }
        }
        ADMM_CUDA(cudaEventRecord(ev(4, P), sC));
      }
    } else {
      ADMM_CUDA(cudaEventRecord(ev(3, P), sA));
      ADMM_CUDA(cudaEventRecord(ev(4, P), sA));
      ADMM_CUDA(cudaEventRecord(ev(5, P), sA));
    }
  }
  }
  ADMM_CUDA(cudaEventRecord(ev_end, sB));
  ADMM_CUDA(cudaStreamWaitEvent(user, ev_end, 0));
  ADMM_CUDA(cudaEventRecord(ev_end, sC));
  ADMM_CUDA(cudaStreamWaitEvent(user, ev_end, 0));
  ADMM_CUDA(cudaEventRecord(ev_end, sA));
  ADMM_CUDA(cudaStreamWaitEvent(user, ev_end, 0));
  {
    dim3 grid((unsigned)((k + 255) / 256), (unsigned)k);
    zero_upper_kernel<<<grid, 256, 0, h->stream>>>(A, k, lda);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  }
  ADMM_CUDA(cudaEventRecord(h->evp[2], h->stream));          // factor done; what follows on the timeline is the inverse's tail
  ADMM_CUDA(cudaEventRecord(ev_inv, sD));
  ADMM_CUDA(cudaStreamWaitEvent(user, ev_inv, 0));
  int fail = 0;
  ADMM_CUDA(cudaMemcpyAsync(&fail, h->fail, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  ADMM_REQUIRE(fail == 0 || probe_nobulk, ADMM_B200_ERR_NOTPOSDEF, "Matrix must be positive definite. (pivot %d is not positive)", fail);
  if (!want_inverse || ride) return;
  tri_inverse_doubling(h, k, A, lda, W, ldw, NBO, T);
}

static void potrf_blocked(admm_b200_handle* h, int64_t k, double* A, int64_t lda, double* W, int64_t ldw, bool want_inverse) {
  static const bool old = getenv("ADMM_B200_CHOL_V1") != nullptr;
  if (!old && k > CHOL_NBO) potrf_lookahead(h, k, A, lda, W, ldw, want_inverse);
  else potrf_blocked_v1(h, k, A, lda, W, ldw, want_inverse);
}

// ---------------------------------------------------------------------------------------------
// streaming products
// ---------------------------------------------------------------------------------------------
// Work plan of one (mode, rows, cols) shape, built on the host once and kept on the device.
static ColdotPlan* coldot_plan(admm_b200_handle* h, int mode, int64_t rows, int64_t cols) {
  for (auto& e : h->plans)
    if (e->mode == mode && e->rows == rows && e->cols == cols) return e;
  ADMM_REQUIRE(rows < (1LL << 31) && cols < (1LL << 31), ADMM_B200_ERR_UNSUPPORTED, "coldot: dimension too large");
  ColdotPlan* pl = new ColdotPlan();
  pl->mode = mode; pl->rows = rows; pl->cols = cols;
  // row panels for tall rectangular matrices: enough virtual columns to balance 148 CTAs
  int64_t P = 1;
  int64_t per = rows;   // rows per panel, a multiple of COLDOT_ITEM
  if (mode == COLDOT_FULL && cols < 6000) {
    P = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((6000 + cols - 1) / cols, 32), rows / 2048));
    per = ((rows + P - 1) / P + COLDOT_ITEM - 1) / COLDOT_ITEM * COLDOT_ITEM;
    P = (rows + per - 1) / per;   // no empty panels after rounding
  }
  pl->panels = (int)P;
  const int64_t nv = P * cols;
  ADMM_REQUIRE(nv < (1LL << 31), ADMM_B200_ERR_UNSUPPORTED, "coldot: too many virtual columns");
  // virtual column vc = p*cols + j covers rows [lo, hi) of column j
  auto vrange = [&](int64_t vc, int64_t& j, int64_t& lo, int64_t& hi) {
    const int64_t p = vc / cols;
    j = vc % cols;
    if (mode == COLDOT_LOWER) { lo = j; hi = rows; }
    else if (mode == COLDOT_UPPER) { lo = 0; hi = j + 1; }
    else {
      lo = std::min(rows, p * per);
      hi = std::min(rows, (p + 1) * per);
    }
  };
  std::vector<int> order(nv), pos_item(nv + 1);
  // folded visiting order: triangular modes pair the longest remaining column with the shortest
  for (int64_t q = 0, lo = 0, hi = nv - 1; q < nv; ++q) {
    if (mode == COLDOT_FULL) order[q] = (int)q;
    else order[q] = (int)((q & 1) ? hi-- : lo++);
  }
  std::vector<ColdotItem> items;
  int64_t total = 0;
  for (int64_t q = 0; q < nv; ++q) {
    int64_t j, rlo, rhi;
    vrange(order[q], j, rlo, rhi);
    pos_item[q] = (int)items.size();
    // item boundaries on absolute multiples of COLDOT_ITEM so interior items start 16-byte aligned
    for (int64_t r = rlo; r < rhi;) {
      int64_t e = std::min<int64_t>(rhi, (r / COLDOT_ITEM + 1) * COLDOT_ITEM);
      items.push_back(ColdotItem{(int)j, (int)r, (int)(e - r), (int)q});
      total += e - r;
      r = e;
    }
    ADMM_REQUIRE(items.size() < (size_t)1 << 31, ADMM_B200_ERR_UNSUPPORTED, "coldot: too many items");
  }
  pos_item[nv] = (int)items.size();
  // enough CTAs that none owns more than ~256 (virtual) columns, at least ~64 KB of matrix per CTA otherwise
  const int grid = (int)std::min<int64_t>(kNumSM, std::max<int64_t>(std::max<int64_t>(1, total / 8192), (nv + 255) / 256));
  pl->grid = grid;
  // Equal COST split of the visiting sequence: a warp needs a few thousand cycles per item however
  // short it is (descriptor + DRAM latency + shuffle reduction), worth ~4.5 KB of streaming --
  // measured on B200 (profiles/r01_notes.md): an equal-area split of the natural column order left
  // the CTA owning the ~670 shortest columns running 1.7x longer than the rest.
  constexpr int64_t kItemFloorBytes = 4608;
  std::vector<int64_t> cost(nv + 1, 0);
  for (int64_t q = 0; q < nv; ++q) {
    int64_t cst = 0;
    for (int it = pos_item[q]; it < pos_item[q + 1]; ++it) cst += std::max<int64_t>((int64_t)items[it].nrows * 8, kItemFloorBytes);
    cost[q + 1] = cost[q] + cst;
  }
  std::vector<int> cta_pos(grid + 1);
  cta_pos[0] = 0;
  int64_t c = 0;
  int max_pos = 1;
  for (int b = 1; b <= grid; ++b) {
    const int64_t target = (int64_t)((double)cost[nv] * b / grid);
    while (c < nv && (b == grid || cost[c + 1] <= target)) ++c;
    cta_pos[b] = (int)c;
    max_pos = std::max(max_pos, cta_pos[b] - cta_pos[b - 1]);
  }
  pl->max_pos = max_pos;
  pl->nitems = (int)items.size();
  pl->npos = (int)nv;
  ADMM_CUDA(cudaMalloc(&pl->d_cta_pos, cta_pos.size() * sizeof(int)));
  ADMM_CUDA(cudaMalloc(&pl->d_pos_item, pos_item.size() * sizeof(int)));
  ADMM_CUDA(cudaMalloc(&pl->d_order, std::max<size_t>(order.size(), 1) * sizeof(int)));
  ADMM_CUDA(cudaMalloc(&pl->d_items, std::max<size_t>(items.size(), 1) * sizeof(ColdotItem)));
  // Upload ON THE HANDLE'S STREAM and wait for it.  A plain cudaMemcpy from pageable memory returns
  // once the data sits in the driver's staging buffer -- the DMA may still be in flight -- and it is
  // ordered only against the legacy default stream, not against the handle's non-blocking stream: the
  // first kernel using the plan then raced the upload (seen on B200 as a flaky
  // cudaErrorIllegalAddress on the first x-update after a setup).
  ADMM_CUDA(cudaMemcpyAsync(pl->d_cta_pos, cta_pos.data(), cta_pos.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  ADMM_CUDA(cudaMemcpyAsync(pl->d_pos_item, pos_item.data(), pos_item.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  ADMM_CUDA(cudaMemcpyAsync(pl->d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  ADMM_CUDA(cudaMemcpyAsync(pl->d_items, items.data(), items.size() * sizeof(ColdotItem), cudaMemcpyHostToDevice, h->stream));
  ADMM_CUDA(cudaStreamSynchronize(h->stream));   // the host vectors die at the end of this function
  h->plans.push_back(pl);
  return pl;
}

// rectangular D' * [v_k]: 256-row x 4-column items (gemvt.cuh)
static void gemvt_multi(admm_b200_handle* h, const double* M, int64_t ld, int64_t rows, int64_t cols, int nv,
                        const double* const* v, double* const* out, double scale, const double* addend, double addscale,
                        const int* done) {
  GemvtArgs a;
  a.M = M; a.ld = ld; a.rows = rows; a.cols = cols; a.done = done;
  a.ngroups = (cols + GEMVT_CG - 1) / GEMVT_CG;
  // enough (panel, column group) units to give every CTA ~8 of them, panels of at least 2048 rows
  int64_t P = 1;
  if (a.ngroups < 8 * kNumSM) P = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((8 * kNumSM + a.ngroups - 1) / a.ngroups, 64), rows / 2048));
  if (addend) P = 1;
  int64_t per = round_up((rows + P - 1) / P, GEMVT_ROWS);
  P = (rows + per - 1) / per;
  a.P = (int)P; a.per = per;
  const int64_t nunits = P * a.ngroups;
  const int grid = (int)std::min<int64_t>(kNumSM, nunits);
  a.units_per_cta = (nunits + grid - 1) / grid;
  const size_t smem = (size_t)a.units_per_cta * GEMVT_CG * nv * GEMVT_WARPS * 8;
  ADMM_REQUIRE(smem <= 200 * 1024, ADMM_B200_ERR_UNSUPPORTED, "gemvt: too many column groups per CTA");
  static PerDevice configured_pd[2];
  size_t& conf = configured_pd[nv == 3](h->device);
  if (smem > conf && smem > 48 * 1024) {
    if (nv == 1) ADMM_CUDA(cudaFuncSetAttribute(gemvt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else ADMM_CUDA(cudaFuncSetAttribute(gemvt_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  if (P == 1) {
    for (int k = 0; k < 3; ++k) { a.v[k] = v[k < nv ? k : 0]; a.out[k] = out[k < nv ? k : 0]; }
    a.scale = scale; a.addend = addend; a.addscale = addscale;
  } else {
    h->cd_ws.ensure(P * cols * nv);
    for (int k = 0; k < 3; ++k) { a.v[k] = v[k < nv ? k : 0]; a.out[k] = h->cd_ws.p + (int64_t)(k < nv ? k : 0) * P * cols; }
    a.scale = 1.0; a.addend = nullptr; a.addscale = 0.0;
  }
  if (nv == 1) gemvt_kernel<1><<<grid, GEMVT_THREADS, smem, h->stream>>>(a);
  else gemvt_kernel<3><<<grid, GEMVT_THREADS, smem, h->stream>>>(a);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
  if (P > 1) {
    PanelReduceArgs r;
    for (int k = 0; k < 3; ++k) { r.ws[k] = a.out[k]; r.out[k] = out[k < nv ? k : 0]; }
    r.nv = nv; r.panels = (int)P; r.cols = cols; r.scale = scale; r.done = done;
    panel_reduce_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, h->stream>>>(r);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  }
}

// out_k = scale * M' v_k (+ addscale*addend, NV == 1 only), k < nv
static void coldot_multi(admm_b200_handle* h, int mode, const double* M, int64_t ld, int64_t rows, int64_t cols, int nv,
                         const double* const* v, double* const* out, double scale, const double* addend,
                         double addscale, const int* done) {
  ADMM_REQUIRE(nv == 1 || nv == 3, ADMM_B200_ERR_INVALID, "coldot: nv must be 1 or 3");
  bool ok = (ld % 2 == 0) && (((uintptr_t)M & 15) == 0);
  for (int k = 0; k < nv; ++k) ok = ok && (((uintptr_t)v[k] & 15) == 0);
  ADMM_REQUIRE(ok, ADMM_B200_ERR_UNSUPPORTED,
               "coldot: matrix and vectors must be 16-byte aligned with an even leading dimension (ld=%lld)", (long long)ld);
  if (mode == COLDOT_FULL && rows >= 1024 && !getenv("ADMM_B200_NO_GEMVT")) {
    gemvt_multi(h, M, ld, rows, cols, nv, v, out, scale, addend, addscale, done);
    return;
  }
  const ColdotPlan* plan = coldot_plan(h, mode, rows, cols);
  const size_t smem = (size_t)plan->max_pos * nv * COLDOT_WARPS * 8;
  ADMM_REQUIRE(smem <= 200 * 1024, ADMM_B200_ERR_UNSUPPORTED, "coldot: too many columns per CTA (%d)", plan->max_pos);
  static PerDevice configured_pd[2];
  size_t& conf = configured_pd[nv == 3](h->device);
  if (smem > conf && smem > 48 * 1024) {
    if (nv == 1) ADMM_CUDA(cudaFuncSetAttribute(coldot_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else ADMM_CUDA(cudaFuncSetAttribute(coldot_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  ColdotArgs a;
  a.M = M; a.ld = ld; a.done = done;
  a.cta_pos = plan->d_cta_pos; a.pos_item = plan->d_pos_item; a.order = plan->d_order; a.items = plan->d_items;
  a.dbg_rows = (int)rows; a.dbg_cols = (int)cols; a.dbg_nitems = plan->nitems; a.dbg_npos = plan->npos;
  a.dbg = nullptr;
  static int* dbgbuf = nullptr;
  if (getenv("ADMM_B200_DEBUG")) {
    if (!dbgbuf) { ADMM_CUDA(cudaMalloc(&dbgbuf, 32)); ADMM_CUDA(cudaMemset(dbgbuf, 0, 32)); ADMM_CUDA(cudaDeviceSynchronize()); }
    a.dbg = dbgbuf;
  }
  const int P = plan->panels;
  if (P == 1) {
    for (int k = 0; k < 3; ++k) { a.v[k] = v[k < nv ? k : 0]; a.out[k] = out[k < nv ? k : 0]; }
    a.scale = scale; a.addend = addend; a.addscale = addscale;
  } else {
    ADMM_REQUIRE(addend == nullptr, ADMM_B200_ERR_UNSUPPORTED, "coldot: addend with row panels");
    h->cd_ws.ensure((int64_t)P * cols * nv);
    for (int k = 0; k < 3; ++k) { a.v[k] = v[k < nv ? k : 0]; a.out[k] = h->cd_ws.p + (int64_t)(k < nv ? k : 0) * P * cols; }
    a.scale = 1.0; a.addend = nullptr; a.addscale = 0.0;
  }
  if (nv == 1) coldot_kernel<1><<<plan->grid, COLDOT_THREADS, smem, h->stream>>>(a);
  else coldot_kernel<3><<<plan->grid, COLDOT_THREADS, smem, h->stream>>>(a);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
  if (a.dbg) {
    int hd[8];
    {
      cudaError_t e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) {
        fprintf(stderr, "COLDOT FAULT mode=%d rows=%lld cols=%lld nv=%d M=%p ld=%lld v0=%p out0=%p grid=%d smem=%zu items=%p n=%d | W=%p WT=%p L=%p y=%p t1=%p x=%p dts=%p s=%p D=%p\n",
                mode, (long long)rows, (long long)cols, nv, (const void*)M, (long long)ld, (const void*)v[0], (void*)out[0], plan->grid, smem,
                (void*)plan->d_items, plan->nitems, (void*)h->W.p, (void*)h->WT.p, (void*)h->L.p, (void*)h->y.p, (void*)h->t1.p, (void*)h->x.p,
                (void*)h->dts.p, (void*)h->s.p, (const void*)h->dD);
      }
    }
    ADMM_CUDA(cudaStreamSynchronize(h->stream));
    ADMM_CUDA(cudaMemcpy(hd, dbgbuf, 32, cudaMemcpyDeviceToHost));
    if (hd[0]) {
      fprintf(stderr, "COLDOT SELF-CHECK kind=%d mode=%d rows=%lld cols=%lld nv=%d: %d %d %d %d %d %d %d (nitems=%d npos=%d grid=%d)\n", hd[0],
              mode, (long long)rows, (long long)cols, nv, hd[1], hd[2], hd[3], hd[4], hd[5], hd[6], hd[7], plan->nitems, plan->npos, plan->grid);
      ADMM_CUDA(cudaMemset(dbgbuf, 0, 32));
    }
  }
  if (P > 1) {
    PanelReduceArgs r;
    for (int k = 0; k < 3; ++k) { r.ws[k] = a.out[k]; r.out[k] = out[k < nv ? k : 0]; }
    r.nv = nv; r.panels = P; r.cols = cols; r.scale = scale; r.done = done;
    panel_reduce_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, h->stream>>>(r);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  }
}

static void coldot(admm_b200_handle* h, int mode, const double* M, int64_t ld, int64_t rows, int64_t cols,
                   const double* v, double* out, double scale = 1.0, const double* addend = nullptr,
                   double addscale = 0.0, const int* done = nullptr) {
  const double* vs[1] = {v};
  double* os[1] = {out};
  coldot_multi(h, mode, M, ld, rows, cols, 1, vs, os, scale, addend, addscale, done);
}

static void gemvn(admm_b200_handle* h, const double* D, int64_t ld, int64_t m, int64_t n, const double* v,
                  double* out, double alpha = 1.0, double beta = 0.0, const double* w = nullptr,
                  const int* done = nullptr) {
  GemvNArgs a;
  a.D = D; a.ld = ld; a.m = m; a.n = n; a.v = v; a.out = out; a.alpha = alpha; a.beta = beta; a.w = w; a.done = done;
  const int64_t rb = (m + GEMVN_ROWS - 1) / GEMVN_ROWS;
  int64_t chunks = 1;
  if (rb < 2 * kNumSM) chunks = std::min<int64_t>((4 * kNumSM + rb - 1) / rb, std::max<int64_t>(1, n / 16));
  a.cols_per_chunk = (n + chunks - 1) / chunks;
  chunks = (n + a.cols_per_chunk - 1) / a.cols_per_chunk;
  if (chunks > 1) {
    h->gemv_ws.ensure(chunks * m);
    ensure_tickets(h, rb);
  }
  a.ws = h->gemv_ws.p;
  a.tickets = h->tickets;
  dim3 grid((unsigned)rb, (unsigned)chunks);
  const bool vec_ok = (((uintptr_t)D & 15) == 0) && (ld % 2 == 0);
  if (vec_ok) gemvn_kernel<2><<<grid, GEMVN_THREADS, 0, h->stream>>>(a);
  else gemvn_kernel<1><<<grid, GEMVN_THREADS, 0, h->stream>>>(a);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
}

// ---------------------------------------------------------------------------------------------
// one-pass x-update (symtri.cuh)
// ---------------------------------------------------------------------------------------------
// Columns [c_lo, c_hi) of WT: this rank's share (equal triangle AREA per rank) or all of them.
static void symtri_range(int64_t k, int rank, int nranks, int64_t& c_lo, int64_t& c_hi) {
  auto bound = [&](int g) { return (int64_t)llround((double)k * sqrt((double)g / (double)nranks)); };
  c_lo = (nranks > 1) ? bound(rank) : 0;
  c_hi = (nranks > 1) ? (rank == nranks - 1 ? k : bound(rank + 1)) : k;
}

static SymtriPlan* symtri_plan(admm_b200_handle* h, int64_t k, int64_t c_lo, int64_t c_hi) {
  for (auto& e : h->st_plans)
    if (e->k == k && e->c_lo == c_lo && e->c_hi == c_hi) return e;
  SymtriPlan* pl = new SymtriPlan();
  pl->k = k; pl->c_lo = c_lo; pl->c_hi = c_hi;
  // visiting order: longest remaining column, shortest remaining column, ... (their lengths sum to c_lo + c_hi + 1)
  std::vector<int> order;
  for (int64_t lo = c_lo, hi = c_hi - 1; lo <= hi;) {
    order.push_back((int)hi--);
    if (lo <= hi) order.push_back((int)lo++);
  }
  std::vector<SymtriEnt> ents;
  std::vector<int> round_ent;
  std::vector<int64_t> round_bytes;
  int used = 0, cnt = 0;
  for (int c : order) {
    const int len = c + 1, padded = (len + 1) & ~1;
    if (cnt == 0 || used + padded > ST_STAGE || cnt == ST_CPR) {
      round_ent.push_back((int)ents.size());
      round_bytes.push_back(0);
      used = 0; cnt = 0;
    }
    ents.push_back(SymtriEnt{c, used, len, 0});
    used += padded; ++cnt;
    round_bytes.back() += (int64_t)padded * 8;
  }
  round_ent.push_back((int)ents.size());
  const int nrounds = (int)round_bytes.size();
  // contiguous rounds per CTA, equal cost (bytes + a fixed cost per round for the two barriers)
  const int grid = std::max(1, std::min(kNumSM, nrounds));
  std::vector<int64_t> cost(nrounds + 1, 0);
  for (int r = 0; r < nrounds; ++r) cost[r + 1] = cost[r] + round_bytes[r] + 4096;
  std::vector<int> cta_round(grid + 1, 0);
  int c = 0;
  for (int b = 1; b <= grid; ++b) {
    const int64_t target = (int64_t)((double)cost[nrounds] * b / grid);
    while (c < nrounds && (b == grid || cost[c + 1] <= target)) ++c;
    cta_round[b] = c;
  }
  pl->grid = grid;
  ADMM_CUDA(cudaMalloc(&pl->d_cta_round, cta_round.size() * sizeof(int)));
  ADMM_CUDA(cudaMalloc(&pl->d_round_ent, round_ent.size() * sizeof(int)));
  ADMM_CUDA(cudaMalloc(&pl->d_ents, std::max<size_t>(ents.size(), 1) * sizeof(SymtriEnt)));
  ADMM_CUDA(cudaMemcpyAsync(pl->d_cta_round, cta_round.data(), cta_round.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  ADMM_CUDA(cudaMemcpyAsync(pl->d_round_ent, round_ent.data(), round_ent.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  ADMM_CUDA(cudaMemcpyAsync(pl->d_ents, ents.data(), ents.size() * sizeof(SymtriEnt), cudaMemcpyHostToDevice, h->stream));
  ADMM_CUDA(cudaStreamSynchronize(h->stream));   // the host vectors die here (same reason as coldot_plan)
  h->st_plans.push_back(pl);
  return pl;
}

static bool symtri_ok(const admm_b200_handle* h, int64_t k, int64_t ld, const double* WT) {
  static const bool off = getenv("ADMM_B200_NO_SYMTRI") != nullptr;
  // small factors stay on the two coldot passes: at k = 1500 (C1) the one-pass kernel's fixed cost (148 CTAs each loading
  // y and writing a partial x) makes the iteration 90 us against 79 us
  static const int64_t mink = getenv("ADMM_B200_SYMTRI_MINK") ? atoll(getenv("ADMM_B200_SYMTRI_MINK")) : 2048;
  return !off && k <= ST_MAXK && k >= std::max<int64_t>(2, mink) && (ld % 2 == 0) && (((uintptr_t)WT & 15) == 0);
}

// x = WT * (WT' * b) = W'(W b) in one pass over WT (upper triangular, columns contiguous).  shard: this rank
// takes an equal-area range of columns and the rank sums are allreduced (row-sharded lasso, large k).
static void symtri_solve(admm_b200_handle* h, const double* WT, int64_t ld, int64_t k, const double* b, double* x,
                         const int* done, bool shard) {
  static PerDevice configured_pd;
  size_t& configured = configured_pd(h->device);
  if (!configured) {
    ADMM_CUDA(cudaFuncSetAttribute(symtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    configured = 1;
  }
  int64_t c_lo, c_hi;
  symtri_range(k, shard ? h->rank : 0, shard ? h->nranks : 1, c_lo, c_hi);
  SymtriPlan* pl = symtri_plan(h, k, c_lo, c_hi);
  const int kpad = (int)round_up(k, 2);
  h->st_part.ensure((int64_t)pl->grid * kpad);
  SymtriArgs a;
  a.WT = WT; a.ld = ld; a.k = (int)k; a.kpad = kpad; a.y = b; a.xpart = h->st_part.p; a.done = done;
  static const int chunk = getenv("ADMM_B200_SYMTRI_CHUNK") ? std::max(64, atoi(getenv("ADMM_B200_SYMTRI_CHUNK")) & ~1) : (1 << 20);   // default: one copy per column (measured: 1024 / 2048 / 4096-double pieces 78 / 75 / 74 us, whole columns 71 us)
  a.chunk = chunk;
  static const int probe = getenv("ADMM_B200_SYMTRI_PROBE") ? 1 : 0;     // data movement only (timing experiment, wrong x)
  a.probe = probe;
  a.cta_round = pl->d_cta_round; a.round_ent = pl->d_round_ent; a.ents = pl->d_ents;
  symtri_kernel<<<pl->grid, ST_THREADS, ST_SMEM, h->stream>>>(a);
  ADMM_CUDA(cudaGetLastError());
  const int mail_on = (shard && h->p2p.ready && kpad <= P2P_LLCAP) ? 1 : 0;
  symtri_reduce_kernel<<<(unsigned)((k + 31) / 32), 256, 0, h->stream>>>(h->st_part.p, pl->grid, kpad, (int)k, x, 1.0, nullptr, 0.0, done,
                                                                         h->p2p.dev, mail_on);
  ADMM_CUDA(cudaGetLastError());
  h->launches += 2;
  if (mail_on) {          // the rank sums are already in every mailbox: one kernel adds them up into x
    p2p_ll_gather_kernel<<<(unsigned)((k + 255) / 256), 256, 0, h->stream>>>(h->p2p.dev, x, k, done, h->p2p.dev.ticket + 1);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  } else if (shard) {
    allreduce_sum(h, x, kpad, done);
  }
}

// x = L' \ (L \ b) with the cached factor (size k); second = the z-update factor of the model problem
static void factor_solve(admm_b200_handle* h, const double* b, double* tmp, double* x, int xsolve, const int* done,
                         bool second = false) {
  ADMM_REQUIRE(h->have_factor, ADMM_B200_ERR_STATE, "no cached factor: call a setup function first");
  if (second) {
    ADMM_REQUIRE(xsolve == ADMM_B200_XSOLVE_INVFACTOR && h->W2.p && h->WT2.p, ADMM_B200_ERR_UNSUPPORTED,
                 "the model problem's z-update is built for xsolve = INVFACTOR");
    if (symtri_ok(h, h->k, h->ldf, h->WT2.p)) {
      symtri_solve(h, h->WT2.p, h->ldf, h->k, b, x, done, false);
      return;
    }
    coldot(h, COLDOT_UPPER, h->WT2.p, h->ldf, h->k, h->k, b, tmp, 1.0, nullptr, 0.0, done);
    coldot(h, COLDOT_LOWER, h->W2.p, h->ldf, h->k, h->k, tmp, x, 1.0, nullptr, 0.0, done);
    return;
  }
  if (xsolve == ADMM_B200_XSOLVE_INVFACTOR) {
    ADMM_REQUIRE(h->have_inverse, ADMM_B200_ERR_STATE,
                 "xsolve = INVFACTOR but the setup was done with xsolve = SUBST (no inverse factor was built)");
    if (symtri_ok(h, h->k, h->ldf, h->WT.p)) {       // one pass over the triangle (symtri.cuh)
      symtri_solve(h, h->WT.p, h->ldf, h->k, b, x, done, h->xshard);
      return;
    }
    // t = W b : row i of W is column i of WT (rows 0..i)
    coldot(h, COLDOT_UPPER, h->WT.p, h->ldf, h->k, h->k, b, tmp, 1.0, nullptr, 0.0, done);
    // x = W' t : x_j = column j of W (rows j..k-1) . t
    coldot(h, COLDOT_LOWER, h->W.p, h->ldf, h->k, h->k, tmp, x, 1.0, nullptr, 0.0, done);
  } else if (xsolve == ADMM_B200_XSOLVE_SUBST) {
    // Blocked forward / back substitution on the factor L itself: x = L' \ (L \ b) as the reference writes it
    // (getProxOps.m:1200).  The inverted diagonal blocks are the diagonal blocks of W -- 512 wide when the look-ahead
    // Cholesky built them (k > 512: 16 dependent steps per solve at k = 8192), 128 wide otherwise -- so the critical
    // path is products only: y_P = W_PP y_P, then the panel below is streamed by all SMs (y[below] -= L[below, P] y_P);
    // backwards with the transposes.  Every entry of L is read once per solve.  The exact-semantics path: it is what
    // runs when the conditioning guard fires.
    ADMM_REQUIRE(h->W.p != nullptr, ADMM_B200_ERR_STATE, "xsolve = SUBST needs the inverted diagonal blocks");
    const int64_t k = h->k, ld = h->ldf, bs = h->subst_bs;
    double* y = tmp;
    h->subst_t.ensure(round_up(bs, 2) + 2);
    double* t = h->subst_t.p;
    vec_copy_kernel<<<(unsigned)((k + 255) / 256), 256, 0, h->stream>>>(y, b, k, done);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    auto diag_apply = [&](int64_t k0, int nb, int trans) {     // y_P <- W_PP y_P  or  W_PP' y_P
      const double* Wkk = h->W.p + k0 + k0 * ld;
      if (nb <= 128) {
        tri_block_mv_kernel<<<1, 128, 0, h->stream>>>(Wkk, ld, nb, y + k0, trans, done);
        ADMM_CUDA(cudaGetLastError());
        h->launches++;
        return;
      }
      if (!trans) gemvn(h, Wkk, ld, nb, nb, y + k0, t, 1.0, 0.0, nullptr, done);
      else coldot(h, COLDOT_FULL, Wkk, ld, nb, nb, y + k0, t, 1.0, nullptr, 0.0, done);
      vec_copy_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, h->stream>>>(y + k0, t, nb, done);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
    };
    for (int64_t k0 = 0; k0 < k; k0 += bs) {          // L y = b
      const int nb = (int)std::min<int64_t>(bs, k - k0);
      diag_apply(k0, nb, 0);
      const int64_t rem = k - k0 - nb;
      if (rem > 0)   // y[k0+nb:] -= L[k0+nb:, k0:k0+nb] * y_k
        gemvn(h, h->L.p + (k0 + nb) + k0 * ld, ld, rem, nb, y + k0, y + k0 + nb, -1.0, 1.0, y + k0 + nb, done);
    }
    const int64_t nblk = (k + bs - 1) / bs;
    for (int64_t bi = nblk - 1; bi >= 0; --bi) {           // L' x = y
      const int64_t k0 = bi * bs;
      const int nb = (int)std::min<int64_t>(bs, k - k0);
      diag_apply(k0, nb, 1);
      if (k0 > 0)    // y[0:k0] -= L[k0:k0+nb, 0:k0]' * x_k
        coldot(h, COLDOT_FULL, h->L.p + k0, ld, nb, k0, y + k0, y, -1.0, y, 1.0, done);
    }
    vec_copy_kernel<<<(unsigned)((k + 255) / 256), 256, 0, h->stream>>>(x, y, k, done);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  } else {
    ADMM_REQUIRE(false, ADMM_B200_ERR_INVALID, "unknown xsolve mode %d", xsolve);
  }
}

// ---------------------------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------------------------
static void stage_matrix(admm_b200_handle* h, int64_t m, int64_t n, const double* D, int64_t ldD) {
  if (is_device_ptr(D)) {
    h->dD = D;
    h->ldD = ldD;
    h->ownD.release();
  } else {
    int64_t ld = round_up(m, 2);
    h->ownD.ensure(ld * n);
    upload_rows(h, h->ownD.p, ld, D, ldD, 0, m, n, h->stream, is_pageable_host_ptr(D));
    h->dD = h->ownD.p;
    h->ldD = ld;
  }
  h->m = m;
  h->n = n;
}

// (max L_ii / min L_ii)^2 <= cond_2(A): a free lower bound of the condition number of the factored matrix
__global__ void diag_minmax_kernel(const double* __restrict__ L, int64_t ld, int64_t k, double* out2) {
  __shared__ double smin[32], smax[32];
  double lo = __longlong_as_double(0x7ff0000000000000LL), hi = 0.0;
  for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
    const double v = fabs(L[i + i * ld]);
    lo = fmin(lo, v); hi = fmax(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, smin[w]); hi = fmax(hi, smax[w]); }
    out2[0] = lo; out2[1] = hi;
  }
}

// The inverse-factor x-update (x = W'(W y), W = inv(L)) and the two substitutions the reference runs
// (getProxOps.m:1200) agree to ~1e-15 relative up to cond(A) ~ 1e8 (tests/test_gpu_conditioning.py; both are
// bounded by the error of the Cholesky factor itself).  Beyond the guard the factor is so ill-conditioned that
// no two implementations agree to 1e-9 any more; the engine then runs the reference's own formulation.
static double cond_guard() {
  static const double g = getenv("ADMM_B200_COND_GUARD") ? atof(getenv("ADMM_B200_COND_GUARD")) : 1e9;
  return g;
}

static void factor_current(admm_b200_handle* h, int64_t k, bool want_inverse) {
  // h->L holds the lower triangle of the SPD matrix
  h->W.ensure(h->ldf * k);
  potrf_blocked(h, k, h->L.p, h->ldf, h->W.p, h->ldf, want_inverse);
  if (want_inverse) {
    h->WT.ensure(h->ldf * k);
    ADMM_CUDA(cudaMemsetAsync(h->WT.p, 0, (size_t)h->ldf * k * 8, h->stream));   // rows k..ldf-1 are read as padding
    transpose(h, h->W.p, k, k, h->ldf, h->WT.p, h->ldf);
  }
  h->t2.ensure(2);
  diag_minmax_kernel<<<1, 1024, 0, h->stream>>>(h->L.p, h->ldf, k, h->t2.p);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
  double mm[2] = {1.0, 1.0};
  ADMM_CUDA(cudaMemcpyAsync(mm, h->t2.p, 16, cudaMemcpyDeviceToHost, h->stream));
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  h->diag_ratio = (mm[0] > 0.0) ? (mm[1] / mm[0]) * (mm[1] / mm[0]) : __builtin_inf();
  h->xsolve_eff = (want_inverse && h->diag_ratio <= cond_guard()) ? ADMM_B200_XSOLVE_INVFACTOR : ADMM_B200_XSOLVE_SUBST;
  h->k = k;
  h->have_factor = true;
  h->have_inverse = want_inverse;
  h->subst_bs = (k > CHOL_NBO && !getenv("ADMM_B200_CHOL_V1") && !getenv("ADMM_B200_SUBST_128")) ? CHOL_NBO : CHOL_NB;
}

// the x-update realisation a solve uses: the caller's choice, except that INVFACTOR gives way to SUBST when the
// conditioning guard fired at setup
static int eff_xsolve(const admm_b200_handle* h, const admm_b200_options& o) {
  return (o.xsolve == ADMM_B200_XSOLVE_INVFACTOR && h->xsolve_eff == ADMM_B200_XSOLVE_SUBST) ? ADMM_B200_XSOLVE_SUBST : o.xsolve;
}

// m = THIS rank's rows, m_total = rows of the whole problem.  m_total > m (row shards, one process per GPU):
// every rank forms D_g'D_g and D_g's_g on its rows, ONE allreduce sums the n x n Gram (+ D's), and every rank
// factors the same matrix -- the transpose reduction of unwrappedadmm.m:114-122 applied to lasso.m:160,168.
static void setup_lasso(admm_b200_handle* h, int64_t m, int64_t m_total, int64_t n, const double* D, int64_t ldD,
                        const double* s, double rho, int xsolve) {
  ADMM_REQUIRE(m > 0 && n > 0 && D && s && ldD >= m && m_total >= m, ADMM_B200_ERR_INVALID, "lasso: bad dimensions or null input");
  ADMM_REQUIRE(rho > 0, ADMM_B200_ERR_INVALID, "Argument options.rho is not a positive real number!");
  const bool sharded = h->nranks > 1 && m_total > m;
  ADMM_REQUIRE(m_total == m || h->nranks > 1, ADMM_B200_ERR_STATE, "lasso: row shards need a communicator (admm_b200_comm_init)");
  ADMM_REQUIRE(!sharded || m_total >= n, ADMM_B200_ERR_UNSUPPORTED,
               "lasso: only the tall problem (rows >= columns) is row-sharded; the fat one factors D*D', which couples all rows");
  ADMM_CUDA(cudaEventRecord(h->ev0, h->stream));
  h->have_factor = false; h->lasso_sharded = false; h->zero_cols = 0; h->xshard = false; h->have_Q = false;
  h->kind = ADMM_B200_LASSO;
  h->lasso_sharded = sharded;
  // a large factor on several GPUs: every rank streams its equal-area share of the inverse factor's columns and one
  // mailbox allreduce of the n-vector sums them (x = sum_i w_i (w_i . y) splits over i)
  h->xshard = sharded && h->p2p.ready && n >= 4096 && n <= ST_MAXK && round_up(n, 2) <= P2P_CAP && !getenv("ADMM_B200_NO_XSHARD");
  h->m_total = m_total;
  h->tall = (m_total >= n);
  h->rho_setup = rho;
  h->xsolve = xsolve;
  h->nA = h->nB = h->mc = n;
  h->s.ensure(round_up(m, 2));
  copy_in(h, h->s.p, s, m);
  h->dts.ensure(round_up(n, 2));
  const int64_t k = h->tall ? n : m;
  h->ldf = round_up(k, 16);
  h->L.ensure(h->ldf * k);
  if (sharded) {   // the whole buffer is summed over ranks: no uninitialised words (NaN patterns) in the padding
    ADMM_CUDA(cudaMemsetAsync(h->L.p, 0, (size_t)h->ldf * k * 8, h->stream));
    ADMM_CUDA(cudaMemsetAsync(h->dts.p, 0, (size_t)round_up(n, 2) * 8, h->stream));
  }
  GemmOpt o;
  o.lower_only = 1;
  constexpr int64_t kPanelRows = 16384;
  if (h->tall && !is_device_ptr(D) && m >= 4096 && m * n >= ((int64_t)32 << 20)) {
    // HOST matrix, tall: pipeline the upload with the Gram.  Row panels of D are copied on the second
    // stream while the previous panel's D_p'D_p (and D_p's_p) accumulate on the compute stream, so the
    // 4.3 GB H2D of config C2 hides behind the DMMA SYRK instead of preceding it.  The first panels are small
    // (2048, 2048, 4096, 8192 rows, then 16384 each): only the first panel's upload is exposed.
    const int64_t ld = round_up(m, 2);
    h->ownD.ensure(ld * n);
    h->dD = h->ownD.p; h->ldD = ld; h->m = m; h->n = n;
    cudaStream_t sC = h->stream, sX = h->stream2;
    ADMM_CUDA(cudaEventRecord(h->ev_la[0], sC));
    ADMM_CUDA(cudaStreamWaitEvent(sX, h->ev_la[0], 0));   // ownD may still be read by earlier work
    const bool pageable = is_pageable_host_ptr(D);
    std::vector<int64_t> start;
    {
      const int64_t ramp[4] = {2048, 2048, 4096, 8192};
      int64_t r = 0;
      for (int i = 0; r < m; ++i) {
        start.push_back(r);
        int64_t step = (i < 4) ? ramp[i] : kPanelRows;
        if (m - (r + step) < 1024) step = m - r;           // no sliver at the end
        r += step;
      }
      start.push_back(m);
    }
    const int64_t npanels = (int64_t)start.size() - 1;
    for (int64_t p = 0; p < npanels; ++p) {
      const int64_t r0 = start[p], rows = start[p + 1] - r0;
      upload_rows(h, h->ownD.p, ld, D, ldD, r0, rows, n, sX, pageable);
      ADMM_CUDA(cudaEventRecord(h->ev_la[p & 1], sX));
      ADMM_CUDA(cudaStreamWaitEvent(sC, h->ev_la[p & 1], 0));
      o.diag_add = (p == npanels - 1 && !sharded) ? rho : 0.0;
      gemm(h, 1, 0, n, n, rows, 1.0, h->ownD.p + r0, ld, h->ownD.p + r0, ld, p ? 1.0 : 0.0, h->L.p, h->ldf, o);
      coldot(h, COLDOT_FULL, h->ownD.p + r0, ld, rows, n, h->s.p + r0, h->dts.p, 1.0, p ? h->dts.p : nullptr, 1.0);
    }
  } else {
  stage_matrix(h, m, n, D, ldD);
  // Dts = D'*s  (lasso.m:160)
  coldot(h, COLDOT_FULL, h->dD, h->ldD, m, n, h->s.p, h->dts.p);
  if (h->tall) {  // chol(D'*D + rho*I)  (lasso.m:168)
    o.diag_add = sharded ? 0.0 : rho;
    gemm(h, 1, 0, n, n, m, 1.0, h->dD, h->ldD, h->dD, h->ldD, 0.0, h->L.p, h->ldf, o);
  } else {        // chol(1/rho*(D*D') + I)  (lasso.m:172)
    o.diag_add = 1.0;
    gemm(h, 0, 1, m, m, n, 1.0 / rho, h->dD, h->ldD, h->dD, h->ldD, 0.0, h->L.p, h->ldf, o);
  }
  }
  if (sharded) {
    // W = sum_g D_g'D_g, Dts = sum_g D_g's_g (unwrappedadmm.m:114-122), then + rho*I once.  The strict upper
    // tiles of L are never read, but the buffer travels whole: one NCCL ring/NVLS allreduce of ldf*n doubles.
    allreduce_sum(h, h->L.p, h->ldf * n);
    allreduce_sum(h, h->dts.p, round_up(n, 2));
    add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->L.p, h->ldf, n, rho);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  }
  ADMM_CUDA(cudaEventRecord(h->evp[1], h->stream));
  factor_current(h, k, xsolve == ADMM_B200_XSOLVE_INVFACTOR);
  ADMM_CUDA(cudaEventRecord(h->ev1, h->stream));
  ADMM_CUDA(cudaEventSynchronize(h->ev1));
  float ms = 0;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->setup_ms = ms;
  h->phase_ms[3] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->evp[1]));
  h->phase_ms[0] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->evp[1], h->evp[2]));
  h->phase_ms[1] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->evp[2], h->ev1));
  h->phase_ms[2] = ms;
  h->have_init = false;
  h->iter_ready = false;
  h->generation++;
}

// Basis pursuit (solvers/basispursuit.m:116-120): the reference caches the dense n x n projector
// P = I - D'(DD')\D and q = D'((DD')\s) (8 GiB at n = 32768).  The engine caches only
// L = chol(D*D') (m x m) and applies  x = P v + q = v - D'((DD') \ (D v - s))  per iteration.
static void setup_bp(admm_b200_handle* h, int64_t m, int64_t n, const double* D, int64_t ldD, const double* s) {
  ADMM_REQUIRE(m > 0 && n > 0 && D && s && ldD >= m, ADMM_B200_ERR_INVALID, "basispursuit: bad dimensions or null input");
  ADMM_REQUIRE(m < n, ADMM_B200_ERR_INVALID, "basispursuit: D must have fewer rows than columns");
  ADMM_CUDA(cudaEventRecord(h->ev0, h->stream));
  h->have_factor = false; h->lasso_sharded = false; h->zero_cols = 0; h->xshard = false; h->have_Q = false;
  stage_matrix(h, m, n, D, ldD);
  h->s.ensure(round_up(m, 2));
  copy_in(h, h->s.p, s, m);
  h->kind = ADMM_B200_BASISPURSUIT;
  h->tall = false;
  h->nA = h->nB = h->mc = n;
  h->ldf = round_up(m, 16);
  h->L.ensure(h->ldf * m);
  GemmOpt o;
  o.lower_only = 1;
  gemm(h, 0, 1, m, m, n, 1.0, h->dD, h->ldD, h->dD, h->ldD, 0.0, h->L.p, h->ldf, o);   // D*D'  (basispursuit.m:116)
  ADMM_CUDA(cudaEventRecord(h->evp[1], h->stream));
  factor_current(h, m, true);
  ADMM_CUDA(cudaEventRecord(h->ev1, h->stream));
  ADMM_CUDA(cudaEventSynchronize(h->ev1));
  float ms = 0;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->setup_ms = h->phase_ms[3] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->evp[1]));
  h->phase_ms[0] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->evp[1], h->evp[2]));
  h->phase_ms[1] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->evp[2], h->ev1));
  h->phase_ms[2] = ms;
  h->have_init = false;
  h->iter_ready = false;
  h->generation++;
}

// Quadratic objective 1/2 x'Px + q'x + r with a projection as z-prox (quadraticprogram.m 'bounded'
// branch, getProxOps.m:1441-1474): R = chol(P + rho*I) once, x = R \ (R' \ (rho*(z-u) - q)), which is
// the tall-lasso x-update with Dts := -q.  kind = PROX_BOX (min(ub,max(lb,x+u)), :1470-1474) or
// PROX_NONNEG (pos(x+u), the z-prox of :1378-1382 / :1422-1426).
static void setup_quadratic(admm_b200_handle* h, int kind, int64_t n, const double* P, int64_t ldP, const double* q,
                            double r, double rho, const double* lb, const double* ub) {
  ADMM_REQUIRE(kind == ADMM_B200_PROX_BOX || kind == ADMM_B200_PROX_NONNEG, ADMM_B200_ERR_INVALID,
               "setup_quadratic: kind must be PROX_BOX or PROX_NONNEG");
  ADMM_REQUIRE(n > 0 && P && q && ldP >= n, ADMM_B200_ERR_INVALID, "setup_quadratic: bad dimensions or null input");
  ADMM_REQUIRE(rho > 0, ADMM_B200_ERR_INVALID, "Argument options.rho is not a positive real number!");
  ADMM_REQUIRE(kind != ADMM_B200_PROX_BOX || (lb && ub), ADMM_B200_ERR_INVALID, "setup_quadratic: box bounds missing");
  ADMM_CUDA(cudaEventRecord(h->ev0, h->stream));
  h->have_factor = false; h->lasso_sharded = false; h->zero_cols = 0; h->xshard = false; h->have_Q = false;
  h->kind = kind;
  h->tall = true;
  h->m = h->n = n;
  h->nA = h->nB = h->mc = n;
  h->rho_setup = rho;
  h->qp_r = r;
  h->lambda = 0.0;
  const int64_t ldp = round_up(n, 2);
  h->Pfull.ensure(ldp * n);
  ADMM_CUDA(cudaMemcpy2DAsync(h->Pfull.p, (size_t)ldp * 8, P, (size_t)ldP * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDefault,
                              h->stream));
  h->dD = h->Pfull.p;
  h->ldD = ldp;
  h->ownD.release();
  h->dts.ensure(round_up(n, 2));
  copy_in(h, h->dts.p, q, n);
  if (kind == ADMM_B200_PROX_BOX) {
    h->lb.ensure(n); h->ub.ensure(n);
    copy_in(h, h->lb.p, lb, n);
    copy_in(h, h->ub.p, ub, n);
  }
  h->ldf = round_up(n, 16);
  h->L.ensure(h->ldf * n);
  ADMM_CUDA(cudaMemcpy2DAsync(h->L.p, (size_t)h->ldf * 8, h->Pfull.p, (size_t)ldp * 8, (size_t)n * 8, (size_t)n,
                              cudaMemcpyDeviceToDevice, h->stream));
  add_diag_negate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->L.p, h->ldf, n, rho, h->dts.p);   // P + rho*I, -q
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
  ADMM_CUDA(cudaEventRecord(h->evp[1], h->stream));
  factor_current(h, n, true);
  ADMM_CUDA(cudaEventRecord(h->ev1, h->stream));
  ADMM_CUDA(cudaEventSynchronize(h->ev1));
  float ms = 0;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->setup_ms = h->phase_ms[3] = ms;
  h->have_init = false;
  h->iter_ready = false;
  h->generation++;
}

// Model problem (solvers/model.m:119-146, getProxOps.m:55-110, 952-1012):
//   x = (PtP + rho I) \ (Ptr + rho (z - u)),   z = (QtQ + rho I) \ (Qts + rho (xhat + u)).
// Both Gram matrices stay on the device; (re)factoring for a rho is n^3/3 flop per matrix, no pass
// over P or Q.
static void model_factor(admm_b200_handle* h, double rho) {
  const int64_t n = h->n;
  ADMM_CUDA(cudaEventRecord(h->evp[1], h->stream));
  for (int which = 1; which >= 0; --which) {              // Q first, then P (which ends up in L / W / WT)
    const DBuf& G = which ? h->G2 : h->G1;
    h->L.ensure(h->ldf * n);
    ADMM_CUDA(cudaMemcpyAsync(h->L.p, G.p, (size_t)h->ldf * n * 8, cudaMemcpyDeviceToDevice, h->stream));
    add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->L.p, h->ldf, n, rho);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    factor_current(h, n, true);
    if (which) { std::swap(h->L, h->L2); std::swap(h->W, h->W2); std::swap(h->WT, h->WT2); }
  }
  h->rho_setup = rho;
}

static void setup_model(admm_b200_handle* h, int64_t m, int64_t n, const double* P, int64_t ldP, const double* Q,
                        int64_t ldQ, const double* r, const double* s, double rho) {
  ADMM_REQUIRE(m > 0 && n > 0 && P && Q && r && s && ldP >= m && ldQ >= m, ADMM_B200_ERR_INVALID,
               "model: bad dimensions or null input");
  ADMM_REQUIRE(rho > 0, ADMM_B200_ERR_INVALID, "Argument options.rho is not a positive real number!");
  ADMM_CUDA(cudaEventRecord(h->ev0, h->stream));
  h->have_factor = false; h->lasso_sharded = false; h->zero_cols = 0; h->xshard = false; h->have_Q = false;
  h->kind = ADMM_B200_MODEL;
  h->tall = true;
  h->xsolve = ADMM_B200_XSOLVE_INVFACTOR;
  h->nA = h->nB = h->mc = n;
  h->lambda = 0.0;
  stage_matrix(h, m, n, P, ldP);                          // h->dD (P), used in place when it is a device pointer
  const int64_t ldq = round_up(m, 2);
  h->Q2.ensure(ldq * n);
  ADMM_CUDA(cudaMemcpy2DAsync(h->Q2.p, (size_t)ldq * 8, Q, (size_t)ldQ * 8, (size_t)m * 8, (size_t)n, cudaMemcpyDefault,
                              h->stream));
  h->s.ensure(round_up(m, 2)); h->s2.ensure(round_up(m, 2));
  copy_in(h, h->s.p, r, m);
  copy_in(h, h->s2.p, s, m);
  h->dts.ensure(round_up(n, 2)); h->dts2.ensure(round_up(n, 2));
  coldot(h, COLDOT_FULL, h->dD, h->ldD, m, n, h->s.p, h->dts.p);        // Ptr = P'*r   (model.m:124)
  coldot(h, COLDOT_FULL, h->Q2.p, ldq, m, n, h->s2.p, h->dts2.p);       // Qts = Q'*s   (model.m:126)
  h->ldf = round_up(n, 16);
  h->G1.ensure(h->ldf * n); h->G2.ensure(h->ldf * n);
  GemmOpt o;
  o.lower_only = 1;
  gemm(h, 1, 0, n, n, m, 1.0, h->dD, h->ldD, h->dD, h->ldD, 0.0, h->G1.p, h->ldf, o);   // PtP (model.m:123)
  gemm(h, 1, 0, n, n, m, 1.0, h->Q2.p, ldq, h->Q2.p, ldq, 0.0, h->G2.p, h->ldf, o);     // QtQ (model.m:125)
  model_factor(h, rho);
  ADMM_CUDA(cudaEventRecord(h->ev1, h->stream));
  ADMM_CUDA(cudaEventSynchronize(h->ev1));
  float ms = 0;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->setup_ms = h->phase_ms[3] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->evp[1]));
  h->phase_ms[0] = ms;
  h->have_init = false;
  h->iter_ready = false;
  h->generation++;
}

// Total variation (solvers/totalvariation.m:122-164): only s and lambda are data; the operator D
// is implicit.  The pivot table depends on rho and is (re)built when the loop starts.
static void setup_tv(admm_b200_handle* h, int64_t n, const double* s, double lambda) {
  ADMM_REQUIRE(n > 0 && s, ADMM_B200_ERR_INVALID, "totalvariation: bad dimensions or null input");
  ADMM_REQUIRE(lambda >= 0, ADMM_B200_ERR_INVALID, "Given lambda parameter is not a nonnegative number!");
  h->have_factor = false; h->lasso_sharded = false; h->zero_cols = 0; h->xshard = false; h->have_Q = false;
  h->kind = ADMM_B200_TOTALVARIATION;
  h->m = h->n = n;
  h->nA = h->nB = h->mc = n;
  h->lambda = lambda;
  h->s.ensure(round_up(n, 2));
  copy_in(h, h->s.p, s, n);
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  h->setup_ms = 0.0;
  h->tv_rho = -1.0;
  h->have_init = false;
  h->iter_ready = false;
  h->generation++;
}

static void tv_prepare(admm_b200_handle* h, double rho) {
  const int64_t n = h->n;
  h->zz.ensure(2 * tv_stride(n));   // two halves at a 32-byte stride (256-bit loads)
  h->uu.ensure(2 * tv_stride(n));
  if (h->tv_rho == rho) return;
  // pivots of I + rho*D'D: delta_0 = 1 + rho, delta_i = 1 + 2 rho - rho^2/delta_{i-1}; contraction
  std::vector<double> inv;
  double d = 1.0 + rho;
  inv.push_back(1.0 / d);
  for (int64_t i = 1; i < n; ++i) {             // until the pivots reach their fixed point (or the end of the chain)
    const double dn = (1.0 + 2.0 * rho) - rho * rho / d;
    inv.push_back(1.0 / dn);
    if (dn == d) break;
    d = dn;
  }
  const double astar = rho * inv.back();       // forward/backward multiplier at the fixed point
  int64_t K = (astar > 0.0 && astar < 1.0) ? (int64_t)std::min(1e15, ceil(40.0 / -log(astar))) : 1;
  K = round_up(std::max<int64_t>(K, 16), 16);
  // a halo the windowed kernels cannot carry (rho >~ 2500): the exact chained solve takes over -- any rho runs,
  // as in the reference (getProxOps.m:1047)
  h->tv_exact = (4 * K > TvCfg<16>::SEG) || getenv("ADMM_B200_TV_EXACT") != nullptr;
  if (h->tv_exact) {
    const int64_t nseg = (n + TvCfg<16>::SEG - 1) / TvCfg<16>::SEG;
    h->tv_y.ensure(round_up(n, 2));
    h->tv_seg.ensure(3 * nseg);
    K = 16;
  }
  h->tvtab.ensure((int64_t)inv.size());
  ADMM_CUDA(cudaMemcpyAsync(h->tvtab.p, inv.data(), inv.size() * 8, cudaMemcpyHostToDevice, h->stream));
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  h->tv_ntab = (int)inv.size();
  h->tv_inv_star = inv.back();
  h->tv_halo = (int)K;
  h->tv_rho = rho;
}

// ---------------------------------------------------------------------------------------------
// NCCL (loaded lazily with dlopen so single-GPU use needs no NCCL at all)
// ---------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static void nccl_load() {
  if (g_nccl.lib) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  ADMM_REQUIRE(g_nccl.lib != nullptr, ADMM_B200_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  ADMM_REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.AllGather && g_nccl.CommDestroy, ADMM_B200_ERR_COMM,
               "libnccl is missing a required symbol");
}
#define ADMM_NCCL(call)                                                                          \
  do {                                                                                           \
    int _r = (call);                                                                             \
    ADMM_REQUIRE(_r == 0, ADMM_B200_ERR_COMM, "NCCL error %d at %s:%d: %s", _r, __FILE__, __LINE__, \
                 g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");                       \
  } while (0)

// ---- peer-memory mailboxes (p2p.cuh) ---------------------------------------------------------------
static void p2p_teardown(admm_b200_handle* h) {
  P2PState& P = h->p2p;
  for (int r = 0; r < P2P_MAXRANKS; ++r)
    if (P.opened[r]) { cudaIpcCloseMemHandle(P.opened[r]); P.opened[r] = nullptr; }
  if (P.local) cudaFree(P.local);
  if (P.dev.seq) cudaFree(P.dev.seq);
  P = P2PState();
}

struct P2PSlot { cudaIpcMemHandle_t hd; int ok; int pad[3]; };
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");

// Allocates and zeroes this rank's mailbox and exports its CUDA IPC handle (ok = 0 when IPC is unavailable or
// ADMM_B200_NO_P2P=1).  The stream is drained before returning: a peer may store into the mailbox as soon as
// it has the handle.
static P2PSlot p2p_alloc_local(admm_b200_handle* h) {
  P2PState& P = h->p2p;
  const int R = h->nranks;
  p2p_teardown(h);               // a single-rank local mailbox (ensure_mailbox) gives way to the mapped ones
  const bool want = !getenv("ADMM_B200_NO_P2P");
  P.bytes = p2p_mailbox_bytes(R);
  ADMM_CUDA(cudaMalloc(&P.local, P.bytes));
  ADMM_CUDA(cudaMemsetAsync(P.local, 0, P.bytes, h->stream));
  void* ctr = nullptr;                       // seq (8) | ticket (4) | ticket2 (4) | err (4)
  ADMM_CUDA(cudaMalloc(&ctr, 64));
  ADMM_CUDA(cudaMemsetAsync(ctr, 0, 64, h->stream));
  P.dev.seq = (unsigned long long*)ctr;
  P.dev.ticket = (unsigned*)((char*)ctr + 8);
  P.dev.err = (int*)((char*)ctr + 16);
  P.dev.rank = h->rank; P.dev.nranks = R; P.dev.cap = P2P_CAP;
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  P2PSlot mine{};
  mine.ok = 0;
  if (want && cudaIpcGetMemHandle(&mine.hd, P.local) == cudaSuccess) mine.ok = 1;
  else cudaGetLastError();
  return mine;
}

// Maps every peer's mailbox into this process; false when a handle is missing or cannot be opened.
static bool p2p_map_peers(admm_b200_handle* h, const P2PSlot* all) {
  P2PState& P = h->p2p;
  const int R = h->nranks;
  int ok = 1;
  for (int r = 0; r < R; ++r) ok &= all[r].ok;
  for (int r = 0; r < R && ok; ++r) {
    if (r == h->rank) { P.dev.mail[r] = (unsigned char*)P.local; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[r].hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    P.opened[r] = ptr;
    P.dev.mail[r] = (unsigned char*)ptr;
  }
  return ok != 0;
}

// Called by every rank right after the NCCL communicator exists: the IPC handles of the mailboxes travel through
// ncclAllGather.  When the peers cannot be mapped (no peer access between the devices, IPC disabled in the
// container, ADMM_B200_NO_P2P=1) the small allreduces stay on NCCL; admm_b200_comm_info reports which transport
// is in use.
static void p2p_setup(admm_b200_handle* h) {
  P2PState& P = h->p2p;
  const int R = h->nranks;
  if (R < 2 || R > P2P_MAXRANKS) return;
  const bool want = !getenv("ADMM_B200_NO_P2P");
  P2PSlot mine = p2p_alloc_local(h);
  P2PSlot* dall = nullptr;
  ADMM_CUDA(cudaMalloc(&dall, sizeof(P2PSlot) * (R + 1)));
  std::vector<P2PSlot> all(R);
  try {
    ADMM_CUDA(cudaMemcpyAsync(dall + R, &mine, sizeof(P2PSlot), cudaMemcpyHostToDevice, h->stream));
    ADMM_NCCL(g_nccl.AllGather(dall + R, dall, sizeof(P2PSlot), /*ncclInt8*/ 0, h->comm, h->stream));
    ADMM_CUDA(cudaMemcpyAsync(all.data(), dall, sizeof(P2PSlot) * R, cudaMemcpyDeviceToHost, h->stream));
    ADMM_CUDA(cudaStreamSynchronize(h->stream));
  } catch (...) {
    cudaFree(dall);
    throw;
  }
  const int ok = p2p_map_peers(h, all.data()) ? 1 : 0;
  // every rank must agree on the transport: min over ranks of `ok` (1 double through NCCL)
  double* dok = (double*)dall;
  double hok = ok;
  try {
    ADMM_CUDA(cudaMemcpyAsync(dok, &hok, 8, cudaMemcpyHostToDevice, h->stream));
    ADMM_NCCL(g_nccl.AllReduce(dok, dok, 1, /*ncclDouble*/ 8, /*ncclMin*/ 3, h->comm, h->stream));
    ADMM_CUDA(cudaMemcpyAsync(&hok, dok, 8, cudaMemcpyDeviceToHost, h->stream));
    ADMM_CUDA(cudaStreamSynchronize(h->stream));
  } catch (...) {
    cudaFree(dall);
    throw;
  }
  cudaFree(dall);
  P.ready = hok > 0.5;
  if (!P.ready && want && h->rank == 0)
    fprintf(stderr, "libadmm_b200: peer mailboxes could not be mapped (CUDA IPC); small allreduces stay on NCCL\n");
}

static void comm_destroy(admm_b200_handle* h) {
  p2p_teardown(h);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
  h->rank = 0;
  h->nranks = 1;
}

// in-place sum over ranks of `count` doubles on the handle's stream (no-op for a single rank).  Small
// messages go through the peer mailboxes (two tiny kernels, graph-capturable, bitwise identical result on
// every rank); large ones (the n x n Gram) through ncclAllReduce.
static void allreduce_sum(admm_b200_handle* h, double* buf, int64_t count, const int* done) {
  if (h->nranks <= 1) return;
  if (h->p2p.ready && count <= P2P_LLCAP && !getenv("ADMM_B200_P2P_FLAGS")) {
    // flag-in-data words: the writer just stores, the reader spins on the data itself (p2p.cuh)
    const int grid = (int)std::max<int64_t>(1, (count + 255) / 256);
    p2p_ll_push_kernel<<<std::min(grid, 64), 256, 0, h->stream>>>(h->p2p.dev, buf, count, done);
    ADMM_CUDA(cudaGetLastError());
    p2p_ll_gather_kernel<<<grid, 256, 0, h->stream>>>(h->p2p.dev, buf, count, done, h->p2p.dev.ticket + 1);
    ADMM_CUDA(cudaGetLastError());
    h->launches += 2;
    return;
  }
  if (h->p2p.ready && count <= P2P_CAP) {
    const int grid = (int)std::max<int64_t>(1, (count + 255) / 256);
    p2p_push_kernel<<<std::min(grid, 32), 256, 0, h->stream>>>(h->p2p.dev, buf, count, done);
    ADMM_CUDA(cudaGetLastError());
    p2p_wait_sum_kernel<<<grid, 256, 0, h->stream>>>(h->p2p.dev, buf, count, done, h->p2p.dev.ticket + 1);
    ADMM_CUDA(cudaGetLastError());
    h->launches += 2;
    return;
  }
  if (!h->comm) {
    // mailbox-only transport (admm_b200_comm_init_ipc): a large message goes through the plain slots in pieces
    ADMM_REQUIRE(h->p2p.ready, ADMM_B200_ERR_COMM, "allreduce: the handle has neither a communicator nor mapped mailboxes");
    for (int64_t off = 0; off < count; off += P2P_CAP) {
      const int64_t cnt = std::min<int64_t>(P2P_CAP, count - off);
      const int grid = (int)((cnt + 255) / 256);
      p2p_push_kernel<<<std::min(grid, 32), 256, 0, h->stream>>>(h->p2p.dev, buf + off, cnt, done);
      ADMM_CUDA(cudaGetLastError());
      p2p_wait_sum_kernel<<<grid, 256, 0, h->stream>>>(h->p2p.dev, buf + off, cnt, done, h->p2p.dev.ticket + 1);
      ADMM_CUDA(cudaGetLastError());
      h->launches += 2;
    }
    return;
  }
  ADMM_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, h->comm, h->stream));
}

// after a loop: a mailbox wait that timed out means a peer died or fell out of step
static void p2p_check(admm_b200_handle* h) {
  if (!h->p2p.ready) return;
  int err = 0;
  ADMM_CUDA(cudaMemcpyAsync(&err, h->p2p.dev.err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  ADMM_REQUIRE(err == 0, ADMM_B200_ERR_COMM, "peer-memory allreduce timed out: a rank stopped taking part in the exchange");
}

// ---------------------------------------------------------------------------------------------
// setup of the A = D problems: W = sum over ranks of D_g'*D_g, R = chol(W)
// (unwrappedadmm.m:96-123 nodepreprocessing; huberfit.m:166; lad.m:134)
// ---------------------------------------------------------------------------------------------
static void setup_unwrapped(admm_b200_handle* h, int kind, int64_t m_local, int64_t m_total, int64_t n, const double* D,
                            int64_t ldD, const double* aux, double C) {
  ADMM_REQUIRE(kind == ADMM_B200_SVM_HINGE || kind == ADMM_B200_SVM_01 || kind == ADMM_B200_HUBERFIT ||
                   kind == ADMM_B200_LAD, ADMM_B200_ERR_INVALID, "setup_unwrapped: kind %d is not an A = D problem", kind);
  ADMM_REQUIRE(m_local > 0 && n > 0 && D && aux && ldD >= m_local && m_total >= m_local, ADMM_B200_ERR_INVALID,
               "setup_unwrapped: bad dimensions or null input");
  if (kind == ADMM_B200_SVM_HINGE || kind == ADMM_B200_SVM_01)
    ADMM_REQUIRE(C >= 0, ADMM_B200_ERR_INVALID, "Given regularization parameter C is not a nonnegative number!");
  if (is_device_ptr(D)) h->ownD.release();      // a staged copy of an earlier problem's matrix: its cudaFree is not setup time
  ADMM_CUDA(cudaEventRecord(h->ev0, h->stream));
  h->have_factor = false; h->lasso_sharded = false; h->zero_cols = 0; h->xshard = false; h->have_Q = false;
  stage_matrix(h, m_local, n, D, ldD);
  h->aux.ensure(round_up(m_local, 2));
  copy_in(h, h->aux.p, aux, m_local);
  h->kind = kind;
  h->svmC = C;
  h->m_total = m_total;
  h->nA = n;
  h->nB = h->mc = m_local;
  h->ldf = round_up(n, 16);
  h->L.ensure(h->ldf * n);
  ADMM_CUDA(cudaMemsetAsync(h->L.p, 0, (size_t)h->ldf * n * 8, h->stream));
  GemmOpt o;
  o.lower_only = 1;
  gemm(h, 1, 0, n, n, m_local, 1.0, h->dD, h->ldD, h->dD, h->ldD, 0.0, h->L.p, h->ldf, o);
  allreduce_sum(h, h->L.p, h->ldf * n);
  // An all-zero column j of D (constant-zero pixels of MNIST, examples/mnistsvm.m) makes W = D'D singular: row and
  // column j of W vanish.  The reference's serial x-update is pinv(D)*(z-u) (unwrappedadmm.m:76-78,
  // linearsvm.m:185), whose minimum-norm solution has x_j = 0 exactly; putting W_jj = 1 gives the same x
  // (d_j = D(:,j)'(z-u) = 0), so the cached Cholesky reproduces pinv for this rank deficiency.
  {
    int* cnt = h->fail;
    ADMM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), h->stream));
    fix_zero_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->L.p, h->ldf, n, cnt);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    int zc = 0;
    ADMM_CUDA(cudaMemcpyAsync(&zc, cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ADMM_CUDA(cudaStreamSynchronize(h->stream));
    h->zero_cols = zc;
  }
  ADMM_CUDA(cudaEventRecord(h->evp[1], h->stream));
  try {
    factor_current(h, n, true);
  } catch (const ArgFail& f) {
    if (f.code != ADMM_B200_ERR_NOTPOSDEF) throw;
    // any other rank deficiency (collinear columns, fewer rows than columns): the reference's parfor path warns
    // "Matrix is singular to working precision" at W\d (unwrappedadmm.m:139) and its serial path takes pinv(D)
    std::string first = g_err;
    set_error("D'*D is singular beyond all-zero columns (%s): the x-update of unwrappedadmm.m needs pinv(D) "
              "(unwrappedadmm.m:76) for such a D; the engine reproduces pinv for all-zero columns only",
              first.c_str());
    throw;
  }
  ADMM_CUDA(cudaEventRecord(h->ev1, h->stream));
  ADMM_CUDA(cudaEventSynchronize(h->ev1));
  float ms = 0;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->setup_ms = h->phase_ms[3] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->evp[1]));
  h->phase_ms[0] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->evp[1], h->evp[2]));
  h->phase_ms[1] = ms;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->evp[2], h->ev1));
  h->phase_ms[2] = ms;
  h->have_init = false;
  h->iter_ready = false;
  h->generation++;
}

static bool is_unwrapped(int kind) {
  return kind == ADMM_B200_SVM_HINGE || kind == ADMM_B200_SVM_01 || kind == ADMM_B200_HUBERFIT || kind == ADMM_B200_LAD;
}
static int uw_kind(int kind) {
  return kind == ADMM_B200_SVM_HINGE ? UW_SVM_HINGE : kind == ADMM_B200_SVM_01 ? UW_SVM_01
         : kind == ADMM_B200_HUBERFIT ? UW_HUBER : UW_LAD;
}

// ---------------------------------------------------------------------------------------------
// the loop
// ---------------------------------------------------------------------------------------------
static LoopParams make_loop_params(admm_b200_handle* h, const admm_b200_options& o, int64_t maxiters, int raw) {
  LoopParams lp;
  lp.rho = o.rho; lp.relax = o.relax; lp.abstol = o.abstol; lp.reltol = o.reltol; lp.convtol = o.convtol;
  lp.hnormtol = o.hnormtol; lp.eps = 2.220446049250313e-16;
  lp.maxiters = maxiters;
  lp.domaxiters = o.domaxiters; lp.stopcond = o.stopcond; lp.nodualerror = o.nodualerror; lp.convtest = o.convtest;
  lp.objevals = o.objevals;
  lp.use_hnorm = (o.convtest || o.stopcond == ADMM_B200_STOP_HNORM || o.stopcond == ADMM_B200_STOP_BOTH) ? 1 : 0;
  lp.raw = raw;
  lp.alg = o.fast ? (o.fasttype ? 2 : 1) : 0;
  lp.nrestart = (o.restart > 0.0 && o.restart < 1.0) ? o.restart : 0.999;   // admm.m:287-290
  lp.dvaltol = o.dvaltol;
  double* hp = h->hist.p;
  lp.pnorm = hp; lp.dnorm = hp + h->hist_cap; lp.perr = hp + 2 * h->hist_cap; lp.derr = hp + 3 * h->hist_cap;
  lp.hn = hp + 4 * h->hist_cap; lp.obj = hp + 5 * h->hist_cap;
  lp.dvals = hp + 6 * h->hist_cap; lp.avals = hp + 7 * h->hist_cap; lp.rst = hp + 8 * h->hist_cap;
  return lp;
}

static int prox_grid(int64_t n) {
  return (int)std::min<int64_t>(2 * kNumSM, std::max<int64_t>(1, (n + PROX_THREADS * 4 - 1) / (PROX_THREADS * 4)));
}

static void alloc_iterates(admm_b200_handle* h) {
  const int64_t big = round_up(std::max(std::max(h->nA, h->nB), std::max(h->mc, std::max(h->m, h->n))), 2);
  h->x.ensure(round_up(h->nA, 2));
  h->z.ensure(round_up(h->nB, 2));
  h->u.ensure(round_up(h->mc, 2));
  h->y.ensure(big);
  h->t1.ensure(big);
  h->t2.ensure(big);
  h->partials.ensure(2 * kNumSM * 16);
  if (h->kind == ADMM_B200_MODEL) h->zsol.ensure(round_up(h->nB, 2));
  h->fv.ensure(round_up(h->nB, 2)); h->fuhat.ensure(round_up(h->mc, 2));
  h->fzprev.ensure(round_up(h->nB, 2)); h->fuprev.ensure(round_up(h->mc, 2));
  if (is_unwrapped(h->kind)) {
    const int64_t need = 3 * round_up(h->n, 2) + 16;
    if (h->cb.cap < need) {
      h->cb.ensure(need);
      ADMM_CUDA(cudaMemsetAsync(h->cb.p, 0, (size_t)need * 8, h->stream));
    }
    h->rvec.ensure(round_up(h->m, 2));
    h->dzvec.ensure(round_up(h->m, 2));
  }
}

static void load_init(admm_b200_handle* h) {
  if (h->have_init) {
    ADMM_CUDA(cudaMemcpyAsync(h->x.p, h->x0.p, (size_t)h->nA * 8, cudaMemcpyDeviceToDevice, h->stream));
    ADMM_CUDA(cudaMemcpyAsync(h->z.p, h->z0.p, (size_t)h->nB * 8, cudaMemcpyDeviceToDevice, h->stream));
    ADMM_CUDA(cudaMemcpyAsync(h->u.p, h->u0.p, (size_t)h->mc * 8, cudaMemcpyDeviceToDevice, h->stream));
  } else {
    ADMM_CUDA(cudaMemsetAsync(h->x.p, 0, (size_t)h->nA * 8, h->stream));
    ADMM_CUDA(cudaMemsetAsync(h->z.p, 0, (size_t)h->nB * 8, h->stream));
    ADMM_CUDA(cudaMemsetAsync(h->u.p, 0, (size_t)h->mc * 8, h->stream));
  }
  // v = z0, uhat = u0 (admm.m:268-269)
  ADMM_CUDA(cudaMemcpyAsync(h->fv.p, h->z.p, (size_t)h->nB * 8, cudaMemcpyDeviceToDevice, h->stream));
  ADMM_CUDA(cudaMemcpyAsync(h->fuhat.p, h->u.p, (size_t)h->mc * 8, cudaMemcpyDeviceToDevice, h->stream));
  ctl_init_kernel<<<1, 32, 0, h->stream>>>(h->ctl, 1);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
  h->iter_ready = true;
}

// total variation: the two z/u halves sit 32-byte aligned (256-bit loads of the fused kernel)
static inline int64_t tv_stride(int64_t n) { return round_up(n, 4); }
// CTA size of the fused iteration kernel, 0 when the halo is too wide for it (overhead > 25 %)
static int tv_fused_threads(const admm_b200_handle* h) {
  if (getenv("ADMM_B200_TV_UNFUSED") || h->tv_exact) return 0;
  static const int pref = getenv("ADMM_B200_TVF_T") ? atoi(getenv("ADMM_B200_TVF_T")) : 128;
  if (pref == 128 && 8 * h->tv_halo <= 128 * TVF_E) return 128;
  if (8 * h->tv_halo <= 256 * TVF_E) return 256;
  return 0;
}
static bool tv_fused_ok(const admm_b200_handle* h) { return tv_fused_threads(h) != 0; }
// One fused TV iteration reading half `par` (xonly: only materialise x from that half).
static void tv_fused_launch(admm_b200_handle* h, const admm_b200_options& o, const LoopParams& lp, int par, bool xonly,
                            bool history) {
  const int64_t n = h->n, st = tv_stride(n);
  const int T = tv_fused_threads(h);
  const int64_t TVF_SEG = (int64_t)T * TVF_E;
  TvFusedArgs a;
  a.n = n;
  a.hl = h->tv_halo + TVF_E; a.hr = h->tv_halo + TVF_E;   // >= halo + 1 / halo + 2, multiples of E (halo is one of 16)
  a.S = TVF_SEG - a.hl - a.hr;
  a.nseg = (n + a.S - 1) / a.S;
  a.s = h->s.p; a.z = h->zz.p + (int64_t)par * st; a.u = h->uu.p + (int64_t)par * st;
  a.znew = h->zz.p + (int64_t)(1 - par) * st; a.unew = h->uu.p + (int64_t)(1 - par) * st;
  a.x = h->x.p;
  a.rho = o.rho; a.lambda = h->lambda; a.invdelta = h->tvtab.p; a.inv_star = h->tv_inv_star; a.ntab = h->tv_ntab;
  a.c = o.rho * h->tv_inv_star;
  a.cp[0] = 1.0;
  for (int e = 1; e <= TVF_E; ++e) a.cp[e] = a.cp[e - 1] * a.c;
  a.pw[0] = a.cp[TVF_E];
  // fast segments: window [g, g + SEG) with g = seg*S - hl,  g - 1 >= ntab,  g >= 1,  g + SEG + 2 <= n
  a.seg_lo = (std::max<int64_t>(h->tv_ntab + 1, 1) + a.hl + a.S - 1) / a.S;
  a.seg_hi = (n - TVF_SEG - 2 + a.hl >= 0) ? std::min<int64_t>(a.nseg, (n - TVF_SEG - 2 + a.hl) / a.S + 1) : 0;
  if (a.seg_hi < a.seg_lo) a.seg_hi = a.seg_lo;
  for (int k = 1; k < 5; ++k) a.pw[k] = a.pw[k - 1] * a.pw[k - 1];
  a.aw = a.pw[4] * a.pw[4];
  a.xonly = xonly ? 1 : 0;
  a.ctl = h->ctl; a.lp = lp;
  a.xvals = history ? h->xvals.p : nullptr;
  a.zvals = history ? h->zvals.p : nullptr;
  a.uvals = history ? h->uvals.p : nullptr;
  const int grid = (int)std::min<int64_t>((512 / T) * kNumSM, a.nseg);
  h->partials.ensure((int64_t)grid * 8);
  a.partials = h->partials.p;
  if (T == 128) {
    if (o.relax == 1.0) tv_fused_kernel<128, true><<<grid, 128, 0, h->stream>>>(a);
    else tv_fused_kernel<128, false><<<grid, 128, 0, h->stream>>>(a);
  } else {
    if (o.relax == 1.0) tv_fused_kernel<256, true><<<grid, 256, 0, h->stream>>>(a);
    else tv_fused_kernel<256, false><<<grid, 256, 0, h->stream>>>(a);
  }
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
}

// Tile height of the single-pass A = D iteration (onepass.cuh), 0 when the two-pass kernels are used:
// wide matrices (the n-column tile does not fit shared memory at >= 16 rows), the fast variants (their
// restart decision sits between the two products), unaligned D, or problems too small to matter.
static int onepass_rows(const admm_b200_handle* h, const LoopParams& lp) {
  if (lp.alg != 0 || getenv("ADMM_B200_NO_ONEPASS")) return 0;
  const int64_t n = h->n, npad = round_up(n, 2);
  if (n > (int64_t)OP_MAXCOLS * OP_SLOTS) return 0;
  if ((((uintptr_t)h->dD) & 15) != 0 || (h->ldD % 2) != 0) return 0;
  if (!getenv("ADMM_B200_FORCE_ONEPASS") && (n < 128 || h->m * n < ((int64_t)1 << 21))) return 0;
  const size_t budget = 227 * 1024;
  if (OnepassCfg<32>::smem_bytes(n, npad) <= budget) return 32;
  if (OnepassCfg<24>::smem_bytes(n, npad) <= budget) return 24;
  if (OnepassCfg<16>::smem_bytes(n, npad) <= budget) return 16;
  return 0;
}

template <int R, int NCH>
static void onepass_launch_t(admm_b200_handle* h, const OnepassArgs& a, int grid) {
  static PerDevice conf_pd;
  size_t& conf = conf_pd(h->device);
  const size_t smem = OnepassCfg<R>::smem_bytes(a.uw.n, a.npad);
  if (smem > conf) {
    ADMM_CUDA(cudaFuncSetAttribute(uw_onepass_kernel<R, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  uw_onepass_kernel<R, NCH><<<grid, OP_THREADS, smem, h->stream>>>(a);
}
template <int R>
static void onepass_launch(admm_b200_handle* h, const OnepassArgs& a, int grid) {
  // 2 column chunks: measured on B200 against 1 (no overlap inside a tile) and 4 (more barriers):
  // C3 128 / 138 / 137 us per iteration, C4 8.60 / 9.92 / 9.83 ms (profiles/r01_notes.md)
  onepass_launch_t<R, 2>(h, a, grid);
}

// which: 0 whole iteration, 1 x-update only, 2 fused pass only
static void enqueue_iteration(admm_b200_handle* h, const admm_b200_options& o, const LoopParams& lp, int which,
                              bool history) {
  const int* done = &h->ctl->done;
  if (h->kind == ADMM_B200_TOTALVARIATION) {
    const int64_t n = h->n;
    static PerDevice configured_pd;
    size_t& configured = configured_pd(h->device);
    if (!configured) {
      ADMM_CUDA(cudaFuncSetAttribute(tv_solve_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TvCfg<8>::SMEM_BYTES));
      ADMM_CUDA(cudaFuncSetAttribute(tv_solve_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TvCfg<16>::SMEM_BYTES));
      configured = 1;
    }
    const bool fastmode = lp.alg != 0;            // fast / accelerated ADMM: x-update from (v, uhat), stepwise kernels
    if (which == 0 && !fastmode && tv_fused_ok(h)) {          // small halo: the whole iteration is one kernel (tv.cuh)
      tv_fused_launch(h, o, lp, h->tv_par, false, history);
      h->tv_par ^= 1;
      return;
    }
    const int64_t npad = tv_stride(n);
    const double* zc = h->zz.p + (int64_t)h->tv_par * npad;
    const double* uc = h->uu.p + (int64_t)h->tv_par * npad;
    const double* zsolve = fastmode ? h->fv.p : zc;       // admm.m:506: x = xminf(x, v, uhat, rho)
    const double* usolve = fastmode ? h->fuhat.p : uc;
    if (which != 2 && h->tv_exact) {
      // any rho: forward aggregates -> chain -> y + backward aggregates -> chain -> x (tv.cuh, tv_exact_kernel)
      static PerDevice conf_pd;
      size_t& conf = conf_pd(h->device);
      if (!conf) {
        ADMM_CUDA(cudaFuncSetAttribute(tv_exact_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TvCfg<16>::SMEM_BYTES));
        ADMM_CUDA(cudaFuncSetAttribute(tv_exact_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TvCfg<16>::SMEM_BYTES));
        ADMM_CUDA(cudaFuncSetAttribute(tv_exact_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TvCfg<16>::SMEM_BYTES));
        conf = 1;
      }
      const int64_t nseg = (n + TvCfg<16>::SEG - 1) / TvCfg<16>::SEG;
      TvExactArgs a;
      a.n = n; a.s = h->s.p; a.z = zsolve; a.u = usolve; a.y = h->tv_y.p; a.x = h->x.p; a.rho = o.rho; a.invdelta = h->tvtab.p;
      a.inv_star = h->tv_inv_star; a.ntab = h->tv_ntab; a.done = done;
      a.segA = h->tv_seg.p; a.segB = h->tv_seg.p + nseg; a.cin = h->tv_seg.p + 2 * nseg;
      double* cin = h->tv_seg.p + 2 * nseg;
      tv_exact_kernel<1><<<(unsigned)nseg, TV_THREADS, TvCfg<16>::SMEM_BYTES, h->stream>>>(a);
      tv_chain_kernel<<<1, 32, 0, h->stream>>>(a.segA, a.segB, nseg, cin, 0, done);
      tv_exact_kernel<2><<<(unsigned)nseg, TV_THREADS, TvCfg<16>::SMEM_BYTES, h->stream>>>(a);
      tv_chain_kernel<<<1, 32, 0, h->stream>>>(a.segA, a.segB, nseg, cin, 1, done);
      tv_exact_kernel<3><<<(unsigned)nseg, TV_THREADS, TvCfg<16>::SMEM_BYTES, h->stream>>>(a);
      ADMM_CUDA(cudaGetLastError());
      h->launches += 5;
    } else if (which != 2) {
      TvSolveArgs a;
      a.n = n; a.s = h->s.p; a.z = zsolve; a.u = usolve; a.x = h->x.p; a.rho = o.rho; a.invdelta = h->tvtab.p;
      a.inv_star = h->tv_inv_star; a.ntab = h->tv_ntab; a.halo = h->tv_halo; a.done = done;
      if (8 * h->tv_halo <= TvCfg<8>::SEG) {      // halo overhead <= 25 %: small segments, 3 CTAs per SM
        const int64_t S = TvCfg<8>::SEG - 2 * h->tv_halo;
        tv_solve_kernel<8><<<(unsigned)((n + S - 1) / S), TV_THREADS, TvCfg<8>::SMEM_BYTES, h->stream>>>(a);
      } else {
        const int64_t S = TvCfg<16>::SEG - 2 * h->tv_halo;
        tv_solve_kernel<16><<<(unsigned)((n + S - 1) / S), TV_THREADS, TvCfg<16>::SMEM_BYTES, h->stream>>>(a);
      }
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
    }
    if (which == 1) return;
    TvProxArgs p;
    p.n = n; p.x = h->x.p; p.s = h->s.p; p.z = zc; p.u = uc;
    p.znew = h->zz.p + (int64_t)(1 - h->tv_par) * npad;
    p.unew = h->uu.p + (int64_t)(1 - h->tv_par) * npad;
    p.lambda = h->lambda;
    const int grid = (int)std::min<int64_t>(2 * kNumSM, std::max<int64_t>(1, (n + TVP_THREADS * TVP_E - 1) / (TVP_THREADS * TVP_E)));
    h->partials.ensure((int64_t)grid * TVP_NRED);
    p.partials = h->partials.p; p.ctl = h->ctl; p.lp = lp;
    p.xvals = history ? h->xvals.p : nullptr;
    p.zvals = history ? h->zvals.p : nullptr;
    p.uvals = history ? h->uvals.p : nullptr;
    p.v = h->fv.p; p.uhat = h->fuhat.p;
    if (fastmode) tv_prox_kernel<true><<<grid, TVP_THREADS, 0, h->stream>>>(p);
    else tv_prox_kernel<false><<<grid, TVP_THREADS, 0, h->stream>>>(p);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    if (fastmode) {      // acceleration pass (admm.m:562-600) + scalar epilogue; z / u of the step before = the half just read
      TvAccelArgs b;
      b.n = n; b.z = p.znew; b.u = p.unew; b.zprev = zc; b.uprev = uc; b.v = h->fv.p; b.uhat = h->fuhat.p;
      b.partials = h->partials.p; b.ctl = h->ctl; b.lp = lp;
      tv_accel_kernel<<<grid, 256, 0, h->stream>>>(b);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
    }
    h->tv_par ^= 1;
  } else if (h->kind == ADMM_B200_BASISPURSUIT) {
    const int64_t n = h->n, m = h->m;
    if (which != 2) {
      // x = P(z-u) + q (getProxOps.m:1031) as v - D'((DD') \ (D v - s)), v = z - u in h->y
      gemvn(h, h->dD, h->ldD, m, n, h->y.p, h->t1.p, 1.0, -1.0, h->s.p, done);
      factor_solve(h, h->t1.p, h->t2.p, h->t1.p, eff_xsolve(h, o), done);
      coldot(h, COLDOT_FULL, h->dD, h->ldD, m, n, h->t1.p, h->x.p, -1.0, h->y.p, 1.0, done);
    }
    if (which == 1) return;
    ProxIdentArgs a;
    a.n = n; a.x = h->x.p; a.z = h->z.p; a.u = h->u.p; a.dts = nullptr; a.y = h->y.p;
    a.lb = a.ub = nullptr; a.zin = nullptr;
    a.thresh = 1.0 / o.rho;                      // getProxOps.m:142
    a.objscale = 1.0;                            // obj = norm(x,1), basispursuit.m:140
    a.kind = PROX_SOFT; a.next = NEXT_DIFF; a.obj_l1_of_x = 1;
    a.partials = h->partials.p; a.ctl = h->ctl; a.lp = lp;
    a.ld = 0; a.thresh_v = a.objscale_v = nullptr; a.hist_stride = 0; a.done_count = nullptr; a.xkeep = nullptr;
    a.xvals = history ? h->xvals.p : nullptr;
    a.zvals = history ? h->zvals.p : nullptr;
    a.uvals = history ? h->uvals.p : nullptr;
    a.v = h->fv.p; a.uhat = h->fuhat.p; a.zprev = h->fzprev.p; a.uprev = h->fuprev.p;
    prox_ident_kernel<<<prox_grid(n), PROX_THREADS, 0, h->stream>>>(a);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    if (lp.alg != 0) {   // acceleration pass (admm.m:562-600) + scalar epilogue
      AccelIdentArgs b;
      b.n = n; b.z = h->z.p; b.u = h->u.p; b.zprev = h->fzprev.p; b.uprev = h->fuprev.p; b.dts = a.dts;
      b.v = h->fv.p; b.uhat = h->fuhat.p; b.y = h->y.p; b.next = a.next;
      b.partials = h->partials.p; b.ctl = h->ctl; b.lp = lp;
      accel_ident_kernel<<<prox_grid(n), PROX_THREADS, 0, h->stream>>>(b);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
    }
  } else if (h->kind == ADMM_B200_LASSO || h->kind == ADMM_B200_PROX_BOX || h->kind == ADMM_B200_PROX_NONNEG) {
    const int64_t n = h->n, m = h->m;
    const bool qp = h->kind != ADMM_B200_LASSO;
    if (which != 2) {
      if (h->tall) {
        // x = U \ (L \ y)   (getProxOps.m:1200)
        factor_solve(h, h->y.p, h->t1.p, h->x.p, eff_xsolve(h, o), done);
      } else {
        // x = y/rho - D'*(U \ (L \ (D*y)))/rho^2   (getProxOps.m:1204)
        gemvn(h, h->dD, h->ldD, m, n, h->y.p, h->t1.p, 1.0, 0.0, nullptr, done);
        factor_solve(h, h->t1.p, h->t2.p, h->t1.p, eff_xsolve(h, o), done);
        coldot(h, COLDOT_FULL, h->dD, h->ldD, m, n, h->t1.p, h->x.p, -1.0 / (o.rho * o.rho), h->y.p, 1.0 / o.rho, done);
      }
    }
    if (which == 1) return;
    if (o.objevals && which == 0) {
      // obj = 1/2*norm(D*x - s)^2 + lambda*norm(z,1)   (lasso.m:227)  /  1/2*x'*P*x + q'*x + r
      gemvn(h, h->dD, h->ldD, m, n, h->x.p, h->t2.p, 1.0, 0.0, nullptr, done);
      if (qp) quad_obj_kernel<<<1, 1024, 0, h->stream>>>(h->x.p, h->t2.p, h->dts.p, h->qp_r, n, h->ctl);
      else half_sqdist_kernel<<<1, 1024, 0, h->stream>>>(h->t2.p, h->s.p, m, h->ctl);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
      if (h->lasso_sharded) allreduce_sum(h, &h->ctl->objpart, 1, done);   // 1/2*||D x - s||^2 over all row shards
    }
    ProxIdentArgs a;
    a.n = n; a.x = h->x.p; a.z = h->z.p; a.u = h->u.p; a.dts = h->dts.p; a.y = h->y.p;
    a.lb = qp ? h->lb.p : nullptr; a.zin = nullptr;
    a.ub = qp ? h->ub.p : nullptr;
    a.thresh = h->lambda / o.rho;
    a.objscale = qp ? 0.0 : h->lambda;
    a.kind = h->kind == ADMM_B200_PROX_BOX ? PROX_BOX : (h->kind == ADMM_B200_PROX_NONNEG ? PROX_NONNEG : PROX_SOFT);
    a.next = NEXT_LASSO; a.obj_l1_of_x = 0;
    a.partials = h->partials.p; a.ctl = h->ctl; a.lp = lp;
    a.ld = 0; a.thresh_v = a.objscale_v = nullptr; a.hist_stride = 0; a.done_count = nullptr; a.xkeep = nullptr;
    a.xvals = history ? h->xvals.p : nullptr;
    a.zvals = history ? h->zvals.p : nullptr;
    a.uvals = history ? h->uvals.p : nullptr;
    a.v = h->fv.p; a.uhat = h->fuhat.p; a.zprev = h->fzprev.p; a.uprev = h->fuprev.p;
    prox_ident_kernel<<<prox_grid(n), PROX_THREADS, 0, h->stream>>>(a);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    if (lp.alg != 0) {   // acceleration pass (admm.m:562-600) + scalar epilogue
      AccelIdentArgs b;
      b.n = n; b.z = h->z.p; b.u = h->u.p; b.zprev = h->fzprev.p; b.uprev = h->fuprev.p; b.dts = a.dts;
      b.v = h->fv.p; b.uhat = h->fuhat.p; b.y = h->y.p; b.next = a.next;
      b.partials = h->partials.p; b.ctl = h->ctl; b.lp = lp;
      accel_ident_kernel<<<prox_grid(n), PROX_THREADS, 0, h->stream>>>(b);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
    }
  } else if (h->kind == ADMM_B200_MODEL) {
    const int64_t n = h->n, m = h->m;
    const bool fastmode = lp.alg != 0;
    if (which != 2) factor_solve(h, h->y.p, h->t1.p, h->x.p, ADMM_B200_XSOLVE_INVFACTOR, done);       // xminModel, getProxOps.m:973
    if (which == 1) return;
    // zminModel (getProxOps.m:1011): z = (QtQ + rho I) \ (Qts + rho*(x + u)); x is Axhat when relaxed
    // (admm.m:521), u is uhat for the fast variants (admm.m:508)
    model_rhs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, h->x.p, h->z.p, fastmode ? h->fuhat.p : h->u.p,
                                                                         h->dts2.p, o.rho, o.relax, h->t2.p, h->ctl);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    factor_solve(h, h->t2.p, h->t1.p, h->zsol.p, ADMM_B200_XSOLVE_INVFACTOR, done, true);
    if (o.objevals && which == 0) {
      // obj = 1/2*norm(P*x - r)^2 + 1/2*norm(Q*z - s)^2   (model.m:139-140)
      gemvn(h, h->dD, h->ldD, m, n, h->x.p, h->t2.p, 1.0, 0.0, nullptr, done);
      half_sqdist_kernel<<<1, 1024, 0, h->stream>>>(h->t2.p, h->s.p, m, h->ctl, 0);
      gemvn(h, h->Q2.p, round_up(m, 2), m, n, h->zsol.p, h->t2.p, 1.0, 0.0, nullptr, done);
      half_sqdist_kernel<<<1, 1024, 0, h->stream>>>(h->t2.p, h->s2.p, m, h->ctl, 1);
      ADMM_CUDA(cudaGetLastError());
      h->launches += 2;
    }
    ProxIdentArgs a;
    a.n = n; a.x = h->x.p; a.z = h->z.p; a.u = h->u.p; a.dts = h->dts.p; a.y = h->y.p;
    a.lb = a.ub = nullptr; a.zin = h->zsol.p;
    a.thresh = 0.0; a.objscale = 0.0;
    a.kind = PROX_GIVEN; a.next = NEXT_LASSO; a.obj_l1_of_x = 0;
    a.partials = h->partials.p; a.ctl = h->ctl; a.lp = lp;
    a.ld = 0; a.thresh_v = a.objscale_v = nullptr; a.hist_stride = 0; a.done_count = nullptr; a.xkeep = nullptr;
    a.xvals = history ? h->xvals.p : nullptr;
    a.zvals = history ? h->zvals.p : nullptr;
    a.uvals = history ? h->uvals.p : nullptr;
    a.v = h->fv.p; a.uhat = h->fuhat.p; a.zprev = h->fzprev.p; a.uprev = h->fuprev.p;
    prox_ident_kernel<<<prox_grid(n), PROX_THREADS, 0, h->stream>>>(a);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    if (fastmode) {   // acceleration pass (admm.m:562-600) + scalar epilogue
      AccelIdentArgs b;
      b.n = n; b.z = h->z.p; b.u = h->u.p; b.zprev = h->fzprev.p; b.uprev = h->fuprev.p; b.dts = a.dts;
      b.v = h->fv.p; b.uhat = h->fuhat.p; b.y = h->y.p; b.next = a.next;
      b.partials = h->partials.p; b.ctl = h->ctl; b.lp = lp;
      accel_ident_kernel<<<prox_grid(n), PROX_THREADS, 0, h->stream>>>(b);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
    }
  } else if (is_unwrapped(h->kind)) {
    const int64_t n = h->n, m = h->m, npad = round_up(n, 2);
    const int nv = o.nodualerror ? 1 : 3;
    double* d = h->cb.p;
    double* scal = h->cb.p + (int64_t)nv * npad;
    if (which != 2) {
      // x = W \ d (unwrappedadmm.m:139) / Rt \ (R \ d) (getProxOps.m:1514), d summed over ranks
      factor_solve(h, d, h->t1.p, h->x.p, eff_xsolve(h, o), done);
    }
    if (which == 1) return;
    UwArgs a;
    a.D = h->dD; a.ld = h->ldD; a.m = m; a.n = n; a.x = h->x.p; a.z = h->z.p; a.u = h->u.p; a.aux = h->aux.p;
    a.rvec = h->rvec.p; a.dzvec = (nv == 3) ? h->dzvec.p : nullptr;
    a.rho = o.rho; a.relax = o.relax; a.C = h->svmC; a.kind = uw_kind(h->kind);
    const int64_t rb = (m + UW_ROWS - 1) / UW_ROWS;
    int64_t chunks = 1;
    if (rb < 2 * kNumSM) chunks = std::min<int64_t>((4 * kNumSM + rb - 1) / rb, std::max<int64_t>(1, n / 16));
    a.cols_per_chunk = (n + chunks - 1) / chunks;
    chunks = (n + a.cols_per_chunk - 1) / a.cols_per_chunk;
    if (chunks > 1) {
      h->gemv_ws.ensure(chunks * m);
      ensure_tickets(h, rb);
    }
    a.ws = h->gemv_ws.p; a.tickets = h->tickets;
    h->uw_partials.ensure(rb * UW_NRED);
    a.partials = h->uw_partials.p; a.grid_ticket = h->grid_ticket; a.scalars = scal; a.ctl = h->ctl;
    a.zvals = history ? h->zvals.p : nullptr;
    a.uvals = history ? h->uvals.p : nullptr;
    a.alg = lp.alg; a.v = h->fv.p; a.uhat = h->fuhat.p; a.zprev = h->fzprev.p; a.uprev = h->fuprev.p;
    if (const int R = onepass_rows(h, lp)) {
      // D read once: D_g*x, the prox and D_g'*[rhs, dz, u] from one shared-memory tile (onepass.cuh)
      OnepassArgs op;
      op.uw = a; op.nv = nv; op.npad = npad;
      // (A 2-CTA cluster form with two tile buffers per CTA was parity-green but measured SLOWER on B200 -- C3 143 vs
      // 128 us, C4 10.1 vs 8.8 ms per iteration: the kernel is bound by the LSU / shared-memory pipe, not by missing
      // overlap, profiles/r01_notes.md -- and was removed in round 2.)
      op.ntiles = (m + R - 1) / R;
      const int grid1 = (int)std::min<int64_t>(kNumSM, op.ntiles), ndparts = 2 * grid1;
      h->op_dpart.ensure((int64_t)ndparts * nv * npad);
      h->uw_partials.ensure((int64_t)grid1 * UW_NRED);
      op.dpart = h->op_dpart.p; op.partials = h->uw_partials.p;
      if (R == 32) onepass_launch<32>(h, op, grid1);
      else if (R == 24) onepass_launch<24>(h, op, grid1);
      else onepass_launch<16>(h, op, grid1);
      ADMM_CUDA(cudaGetLastError());
      // row-sharded runs with mapped mailboxes: the finish kernel stores this rank's sums into every peer (NVLink)
      // and the epilogue adds the ranks up -- no library collective, no extra launches
      const int64_t msg_count = (int64_t)nv * npad + UW_NRED;
      const int use_mail = (h->nranks > 1 && h->p2p.ready && msg_count <= P2P_CAP) ? 1 : 0;
      uw_onepass_finish_kernel<<<(unsigned)((nv * n + UW_NRED + 7) / 8), 256, 0, h->stream>>>(h->op_dpart.p, ndparts, grid1, nv, n, npad, d,
                                                                                        h->uw_partials.p, scal, h->ctl, h->p2p.dev, use_mail);
      ADMM_CUDA(cudaGetLastError());
      h->launches += 2;
      if (!use_mail) allreduce_sum(h, d, msg_count, done);
      UwEpiArgs e;
      e.n = n; e.x = h->x.p; e.dzv = (nv == 3) ? d + npad : nullptr; e.duv = (nv == 3) ? d + 2 * npad : nullptr;
      e.scalars = scal; e.m_total = (double)h->m_total; e.kind = a.kind; e.C = h->svmC; e.ctl = h->ctl; e.lp = lp;
      e.xvals = history ? h->xvals.p : nullptr;
      e.mail = h->p2p.dev; e.use_mail = use_mail; e.msg = d; e.msg_count = msg_count;
      uw_epilogue_kernel<<<1, 256, 0, h->stream>>>(e);
      ADMM_CUDA(cudaGetLastError());
      h->launches++;
      return;
    }
    dim3 grid((unsigned)rb, (unsigned)chunks);
    const bool vec_ok = (((uintptr_t)h->dD & 15) == 0) && (h->ldD % 2 == 0);
    if (vec_ok) uw_gemv_prox_kernel<2><<<grid, UW_THREADS, 0, h->stream>>>(a);
    else uw_gemv_prox_kernel<1><<<grid, UW_THREADS, 0, h->stream>>>(a);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    if (lp.alg != 0) {
      // fast variants: the restart decision needs the rank-summed ||u-uhat||^2, ||z-v||^2 BEFORE the
      // acceleration pass, so the scalars travel in their own (tiny) allreduce
      allreduce_sum(h, scal, UW_NRED, done);
      uw_accel_decide_kernel<<<1, 1, 0, h->stream>>>(h->ctl, lp, scal);
      ADMM_CUDA(cudaGetLastError());
      uw_accel_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(m, h->z.p, h->u.p, h->fzprev.p, h->fuprev.p,
                                                                          h->aux.p, a.kind, h->fv.p, h->fuhat.p, h->rvec.p,
                                                                          (nv == 3) ? h->dzvec.p : nullptr, h->ctl);
      ADMM_CUDA(cudaGetLastError());
      h->launches += 2;
    }
    // D_g' * [rhs, z - zprev (or z - v), u] in one pass
    const double* vs[3] = {h->rvec.p, h->dzvec.p, h->u.p};
    double* os[3] = {d, d + npad, d + 2 * npad};
    coldot_multi(h, COLDOT_FULL, h->dD, h->ldD, m, n, nv, vs, os, 1.0, nullptr, 0.0, done);
    allreduce_sum(h, d, (int64_t)nv * npad + (lp.alg == 0 ? UW_NRED : 0), done);
    UwEpiArgs e;
    e.n = n; e.x = h->x.p; e.dzv = (nv == 3) ? d + npad : nullptr; e.duv = (nv == 3) ? d + 2 * npad : nullptr;
    e.scalars = scal; e.m_total = (double)h->m_total; e.kind = a.kind; e.C = h->svmC; e.ctl = h->ctl; e.lp = lp;
    e.xvals = history ? h->xvals.p : nullptr;
    e.mail = h->p2p.dev; e.use_mail = 0; e.msg = nullptr; e.msg_count = 0;
    uw_epilogue_kernel<<<1, 256, 0, h->stream>>>(e);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  } else {
    ADMM_REQUIRE(false, ADMM_B200_ERR_UNSUPPORTED, "problem kind %d has no iteration built yet", h->kind);
  }
}

static void enqueue_first_rhs(admm_b200_handle* h, const admm_b200_options& o) {
  if (h->kind == ADMM_B200_LASSO || h->kind == ADMM_B200_BASISPURSUIT || h->kind == ADMM_B200_PROX_BOX ||
      h->kind == ADMM_B200_PROX_NONNEG || h->kind == ADMM_B200_MODEL) {
    const int64_t n = h->n;
    const bool bp = h->kind == ADMM_B200_BASISPURSUIT;
    first_rhs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, h->z.p, h->u.p, bp ? nullptr : h->dts.p, o.rho,
                                                                         bp ? NEXT_DIFF : NEXT_LASSO, h->y.p);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  } else if (h->kind == ADMM_B200_TOTALVARIATION) {
    const int64_t n = h->n;
    tv_prepare(h, o.rho);
    h->tv_par = 0;
    ADMM_CUDA(cudaMemcpyAsync(h->zz.p, h->z.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
    ADMM_CUDA(cudaMemcpyAsync(h->uu.p, h->u.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
  } else if (is_unwrapped(h->kind)) {
    const int64_t n = h->n, m = h->m;
    uw_first_rhs_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(m, h->z.p, h->u.p, h->aux.p, uw_kind(h->kind),
                                                                            h->rvec.p);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
    coldot(h, COLDOT_FULL, h->dD, h->ldD, m, n, h->rvec.p, h->cb.p);
    allreduce_sum(h, h->cb.p, round_up(n, 2));
  }
}

static void validate_options(admm_b200_handle* h, const admm_b200_options& o) {
  ADMM_REQUIRE(h->kind != 0, ADMM_B200_ERR_STATE, "no problem set up on this handle");
  ADMM_REQUIRE(o.rho > 0, ADMM_B200_ERR_INVALID, "Argument options.rho is not a positive real number!");
  ADMM_REQUIRE(o.stopcond >= 0 && o.stopcond <= 2, ADMM_B200_ERR_INVALID, "invalid stopcond %d", o.stopcond);
  if (h->kind == ADMM_B200_SVM_HINGE || h->kind == ADMM_B200_SVM_01)
    ADMM_REQUIRE(o.relax == 1.0, ADMM_B200_ERR_INVALID,
                 "Inner matrix dimensions must agree. (linearsvm with relax ~= 1: the reference's zminLinearSVM "
                 "multiplies D (m x n) by the relaxed m-vector, getProxOps.m:1088, admm.m:521)");
  if (h->kind == ADMM_B200_MODEL && o.rho != h->rho_setup) model_factor(h, o.rho);   // getProxOps.m:967-970
  if (h->kind == ADMM_B200_LASSO || h->kind == ADMM_B200_PROX_BOX || h->kind == ADMM_B200_PROX_NONNEG)
    ADMM_REQUIRE(o.rho == h->rho_setup, ADMM_B200_ERR_INVALID,
                 "options.rho (%g) differs from the rho the factor was built with (%g); redo the setup", o.rho,
                 h->rho_setup);
}

static void prepare_loop(admm_b200_handle* h, const admm_b200_options& o, int64_t maxiters, bool history) {
  alloc_iterates(h);
  if (maxiters > h->hist_cap) {
    h->hist.release();
    h->hist.ensure(9 * maxiters);
    h->hist_cap = maxiters;
  }
  if (history) {
    const int64_t need = std::max(std::max(h->nA, h->nB), h->mc) * maxiters;
    ADMM_REQUIRE(need * 8 * 3 < (int64_t)60e9, ADMM_B200_ERR_UNSUPPORTED,
                 "history of %lld iterations x %lld entries does not fit; set options.history = 0",
                 (long long)maxiters, (long long)std::max(std::max(h->nA, h->nB), h->mc));
    h->xvals.ensure(h->nA * maxiters);
    h->zvals.ensure(h->nB * maxiters);
    h->uvals.ensure(h->mc * maxiters);
  }
}

// ---------------------------------------------------------------------------------------------
// Burst runner.  The host enqueues check_every iterations, then reads the device stop flag.  The
// kernels of an iteration have identical arguments every time (iteration counters, history slots
// and stop state live on the device), so from the third full burst on the burst is ONE CUDA graph
// launch: small problems (C1 lasso, the L2-resident shards of C3) are launch-bound otherwise
// (~3 us of host work per kernel against 6-12 us kernels).  The first bursts always run eagerly: the first
// performs every lazy allocation, plan upload and cudaFuncSetAttribute, none of which may happen
// inside a stream capture.  `period`: a graph may only hold a multiple of `period` iterations (the
// total-variation loop alternates two buffer halves on the host side).
// ---------------------------------------------------------------------------------------------
template <class EnqueueOne, class Finished>
static void run_bursts(admm_b200_handle* h, const admm_b200_options& o, int64_t N, int period, EnqueueOne&& enqueue_one,
                       Finished&& finished) {
  const int check = std::max(1, o.check_every);
  // Row-sharded handles stay eager: with the per-iteration ncclAllReduce inside the graph the 2-GPU SVM loop
  // measured SLOWER (133 vs 115 us per iteration, profiles/r01_notes.md); ADMM_B200_GRAPH_NCCL=1 overrides.
  const bool graph_ok = o.graph && check % period == 0 && !getenv("ADMM_B200_DEBUG") && !getenv("ADMM_B200_NO_GRAPH") &&
                        (h->nranks == 1 || h->p2p.ready || getenv("ADMM_B200_GRAPH_NCCL"));
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t graph_launches = 0;
  auto drop = [&]() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    exec = nullptr; graph = nullptr;
  };
  try {
    int64_t enq = 0, bursts = 0;
    while (true) {
      const int64_t burst = std::min<int64_t>(check, N - enq);
      if (graph_ok && bursts >= 2 && burst == check) {   // capture + instantiate (~0.3 ms) only pays for long loops
        if (!exec) {
          const int64_t before = h->launches;
          ADMM_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
          try {
            for (int64_t c = 0; c < burst; ++c) enqueue_one();
          } catch (...) {
            cudaGraph_t dead = nullptr;
            cudaStreamEndCapture(h->stream, &dead);
            if (dead) cudaGraphDestroy(dead);
            throw;
          }
          ADMM_CUDA(cudaStreamEndCapture(h->stream, &graph));
          ADMM_CUDA(cudaGraphInstantiate(&exec, graph, 0));
          graph_launches = h->launches - before;
          h->launches = before;
        }
        ADMM_CUDA(cudaGraphLaunch(exec, h->stream));
        h->launches += graph_launches;
        h->graph_replays++;
      } else {
        for (int64_t c = 0; c < burst; ++c) enqueue_one();
      }
      enq += burst;
      ++bursts;
      if (finished() || enq >= N) break;
    }
  } catch (...) {
    drop();
    throw;
  }
  drop();
}

// ---------------------------------------------------------------------------------------------
// persistent A = D iteration (persist.cuh)
// ---------------------------------------------------------------------------------------------
// a single rank runs the same kernel on a LOCAL mailbox (no IPC); p2p_setup replaces it when ranks attach
static void ensure_mailbox(admm_b200_handle* h) {
  P2PState& P = h->p2p;
  if (P.local) return;
  P.bytes = p2p_mailbox_bytes(1);
  ADMM_CUDA(cudaMalloc(&P.local, P.bytes));
  ADMM_CUDA(cudaMemsetAsync(P.local, 0, P.bytes, h->stream));
  void* ctr = nullptr;
  ADMM_CUDA(cudaMalloc(&ctr, 64));
  ADMM_CUDA(cudaMemsetAsync(ctr, 0, 64, h->stream));
  P.dev.seq = (unsigned long long*)ctr;
  P.dev.ticket = (unsigned*)((char*)ctr + 8);
  P.dev.err = (int*)((char*)ctr + 16);
  P.dev.rank = 0; P.dev.nranks = 1; P.dev.cap = P2P_CAP;
  P.dev.mail[0] = (unsigned char*)P.local;
  P.ready = false;
}

// Q_g = D_g * inv(R)' on this rank's rows: one DMMA GEMM per setup (m x n x n), kept next to D
static void ensure_Q(admm_b200_handle* h) {
  if (h->have_Q) return;
  const int64_t n = h->n, m = h->m;
  h->ldq = round_up(m, 2);
  h->Qm.ensure(h->ldq * n);
  GemmOpt g;
  gemm(h, 0, 1, m, n, n, 1.0, h->dD, h->ldD, h->W.p, h->ldf, 0.0, h->Qm.p, h->ldq, g);
  h->have_Q = true;
}

static bool persist_ok(const admm_b200_handle* h, const admm_b200_options& o, const LoopParams& lp, bool history) {
  if (getenv("ADMM_B200_NO_PERSIST")) return false;
  if (!is_unwrapped(h->kind) || lp.alg != 0 || !o.nodualerror || o.objevals || history || lp.raw) return false;
  if (!h->have_inverse || h->xsolve_eff != ADMM_B200_XSOLVE_INVFACTOR || o.xsolve != ADMM_B200_XSOLVE_INVFACTOR) return false;
  if (h->nranks > 1 && !h->p2p.ready) return false;
  if (round_up(h->n, 2) + UW_NRED > P2P_LLCAP) return false;
  return onepass_rows(h, lp) != 0;
}

template <int R>
static void persist_launch(admm_b200_handle* h, PersistArgs& a, int grid) {
  static PerDevice conf_pd;
  size_t& conf = conf_pd(h->device);
  const size_t smem = OnepassCfg<R>::smem_bytes(a.uw.n, a.npad);
  if (smem > conf) {
    ADMM_CUDA(cudaFuncSetAttribute(uw_persist_kernel<R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  void* args[] = {&a};
  ADMM_CUDA(cudaLaunchCooperativeKernel((const void*)uw_persist_kernel<R, 2>, dim3(grid), dim3(OP_THREADS), args, smem, h->stream));
  h->launches++;
}

// The whole loop of a persist_ok problem.  On return h->x / z / u hold the final iterates and h->h_ctl the loop state.
static void run_persist(admm_b200_handle* h, const admm_b200_options& o, const LoopParams& lp, int64_t N) {
  const int64_t n = h->n, m = h->m, npad = round_up(n, 2);
  if (h->nranks == 1) ensure_mailbox(h);
  ensure_Q(h);
  h->tcur.ensure(npad + 2); h->tlast.ensure(npad + 2);
  // t of the first iteration: sum_g Q_g' r0_g, r0 = z0 - u0 (unwrappedadmm.m:133) or s + z0 - u0 (getProxOps.m:1514)
  uw_first_rhs_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(m, h->z.p, h->u.p, h->aux.p, uw_kind(h->kind), h->rvec.p);
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
  ADMM_CUDA(cudaMemsetAsync(h->tcur.p, 0, (size_t)(npad + 2) * 8, h->stream));
  coldot(h, COLDOT_FULL, h->Qm.p, h->ldq, m, n, h->rvec.p, h->tcur.p);
  allreduce_sum(h, h->tcur.p, npad);
  ADMM_CUDA(cudaMemcpyAsync(h->tlast.p, h->tcur.p, (size_t)npad * 8, cudaMemcpyDeviceToDevice, h->stream));

  // Tile height: the tallest that fits shared memory streams best (onepass.cuh), but a burst is as long as its
  // busiest CTA -- ceil(tiles / CTAs) tiles of R rows -- so a slightly lower tile that divides the rows evenly wins
  // (7500 rows per rank at 8 GPUs: R = 32 gives 2 x 32 row slots per CTA for 50.7 rows of work, R = 28 gives 2 x 28).
  int R = onepass_rows(h, lp);
  {
    const size_t budget = 227 * 1024;
    double best = 1e300;
    int bestR = R;
    const int cand[5] = {32, 28, 24, 20, 16};
    for (int c : cand) {
      if (c > R) continue;
      const size_t smem = (size_t)(n * (c + 2) + npad + (OP_THREADS / (c / 2)) * c + 3 * c + (OP_THREADS / 32) * UW_NRED) * 8;
      if (smem > budget) continue;
      const int64_t nt = (m + c - 1) / c;
      const int64_t g = std::min<int64_t>(kNumSM, nt);
      const double cost = (double)((nt + g - 1) / g) * c * (1.0 + 0.10 * (32 - c) / 16.0);
      if (cost < best - 1e-9) { best = cost; bestR = c; }
    }
    R = bestR;
  }
  PersistArgs a;
  UwArgs& u = a.uw;
  u.D = h->Qm.p; u.ld = h->ldq; u.m = m; u.n = n; u.x = nullptr; u.z = h->z.p; u.u = h->u.p; u.aux = h->aux.p;
  u.rvec = nullptr; u.dzvec = nullptr; u.rho = o.rho; u.relax = o.relax; u.C = h->svmC; u.kind = uw_kind(h->kind);
  u.cols_per_chunk = 0; u.ws = nullptr; u.tickets = nullptr; u.partials = nullptr; u.grid_ticket = nullptr; u.scalars = nullptr;
  u.ctl = h->ctl; u.zvals = u.uvals = nullptr; u.alg = 0; u.v = u.uhat = nullptr; u.zprev = u.uprev = nullptr;
  a.ntiles = (m + R - 1) / R; a.npad = npad;
  const int grid = (int)std::min<int64_t>(kNumSM, a.ntiles);
  h->op_dpart.ensure((int64_t)2 * grid * npad);
  h->uw_partials.ensure((int64_t)grid * UW_NRED);
  a.dpart = h->op_dpart.p; a.partials = h->uw_partials.p; a.tcur = h->tcur.p; a.tlast = h->tlast.p;
  a.mail = h->p2p.dev; a.ctl = h->ctl; a.lp = lp; a.m_total = (double)h->m_total;
  a.prof = nullptr;
  static const bool want_prof = getenv("ADMM_B200_PERSIST_PROF") != nullptr;
  if (want_prof) {
    ADMM_CUDA(cudaMalloc(&a.prof, (size_t)grid * 8 * sizeof(long long)));
    ADMM_CUDA(cudaMemsetAsync(a.prof, 0, (size_t)grid * 8 * sizeof(long long), h->stream));
  }
  const int check = std::max(1, o.check_every);
  int64_t enq = 0;
  while (true) {
    a.burst = (int)std::min<int64_t>(check, N - enq);
    if (R == 32) persist_launch<32>(h, a, grid);
    else if (R == 28) persist_launch<28>(h, a, grid);
    else if (R == 24) persist_launch<24>(h, a, grid);
    else if (R == 20) persist_launch<20>(h, a, grid);
    else persist_launch<16>(h, a, grid);
    enq += a.burst;
    ADMM_CUDA(cudaMemcpyAsync(h->h_ctl, h->ctl, sizeof(LoopCtl), cudaMemcpyDeviceToHost, h->stream));
    ADMM_CUDA(cudaStreamSynchronize(h->stream));
    if (h->h_ctl->done != 0 || enq >= N) break;
  }
  if (a.prof) {
    std::vector<long long> hp((size_t)grid * 8);
    ADMM_CUDA(cudaMemcpy(hp.data(), a.prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(a.prof);
    const double its = std::max<double>(1.0, (double)h->h_ctl->it);
    const char* names[6] = {"D-phase", "grid.sync", "R sums+stores", "(unused)", "values arrive", "stop tests"};
    char line[256];
    std::string txt;                     // one write per rank: the ranks' reports must not interleave
    snprintf(line, sizeof line, "persist profile (rank %d of %d, %lld rows here, %d CTAs, %.0f iterations): cycles per iteration, mean / max over CTAs\n",
             h->rank, h->nranks, (long long)h->m, grid, its);
    txt += line;
    for (int k = 0; k < 6; ++k) {
      double mean = 0.0, mx = 0.0;
      for (int c = 0; c < grid; ++c) { const double v = (double)hp[(size_t)c * 8 + k] / its; mean += v; mx = std::max(mx, v); }
      snprintf(line, sizeof line, "  [r%d] %-14s %9.0f / %9.0f\n", h->rank, names[k], mean / grid, mx);
      txt += line;
    }
    fputs(txt.c_str(), stderr);
  }
  // x of the last iteration: x = inv(R)' t = W' t, x_j = column j of W (rows j..n-1) . t
  coldot(h, COLDOT_LOWER, h->W.p, h->ldf, n, n, h->tlast.p, h->x.p);
}

static void solve(admm_b200_handle* h, const admm_b200_options& o, admm_b200_result* res) {
  validate_options(h, o);
  int64_t N = o.maxiters > 0 ? o.maxiters : 1000;  // admm.m:334-339
  const bool history = o.history && res && res->xvals && res->zvals && res->uvals;
  prepare_loop(h, o, N, history);
  LoopParams lp = make_loop_params(h, o, N, 0);
  ADMM_CUDA(cudaEventRecord(h->ev0, h->stream));
  load_init(h);
  if (persist_ok(h, o, lp, history)) {
    run_persist(h, o, lp, N);
  } else {
  enqueue_first_rhs(h, o);
  run_bursts(
      h, o, N, h->kind == ADMM_B200_TOTALVARIATION ? 2 : 1, [&]() { enqueue_iteration(h, o, lp, 0, history); },
      [&]() {
        ADMM_CUDA(cudaMemcpyAsync(h->h_ctl, h->ctl, sizeof(LoopCtl), cudaMemcpyDeviceToHost, h->stream));
        ADMM_CUDA(cudaStreamSynchronize(h->stream));
        return h->h_ctl->done != 0;
      });
  }
  ADMM_CUDA(cudaEventRecord(h->ev1, h->stream));
  ADMM_CUDA(cudaEventSynchronize(h->ev1));
  p2p_check(h);
  float ms = 0;
  ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (h->kind == ADMM_B200_TOTALVARIATION) {   // the last iteration that ran wrote half (steps mod 2)
    const int64_t half = h->h_ctl->it % 2, n = h->n;
    ADMM_CUDA(cudaMemcpyAsync(h->z.p, h->zz.p + half * tv_stride(n), (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
    ADMM_CUDA(cudaMemcpyAsync(h->u.p, h->uu.p + half * tv_stride(n), (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
    h->tv_par = (int)half;
    // the fused iteration keeps x in registers: materialise the x of the last iteration from the
    // half it read (untouched since: later launches exit on ctl->done)
    if (lp.alg == 0 && tv_fused_ok(h) && h->h_ctl->it >= 1) tv_fused_launch(h, o, lp, (int)(1 - half), true, false);
  }
  if (!res) return;
  const int64_t steps = h->h_ctl->it;
  res->steps = steps;
  res->status = h->h_ctl->status;
  res->setup_ms = h->setup_ms;
  res->loop_ms = ms;
  res->objopt = NAN;
  const double* hp = h->hist.p;
  if (lp.alg != 2) {
    copy_out(h, res->pnorm, hp, steps);
    copy_out(h, res->dnorm, hp + h->hist_cap, steps);
    copy_out(h, res->perr, hp + 2 * h->hist_cap, steps);
    copy_out(h, res->derr, hp + 3 * h->hist_cap, steps);
  }
  if (lp.alg == 2) {
    copy_out(h, res->dvals, hp + 6 * h->hist_cap, steps);
    copy_out(h, res->restarted, hp + 8 * h->hist_cap, steps);
  }
  if (lp.alg != 0) copy_out(h, res->avals, hp + 7 * h->hist_cap, steps);
  if (lp.use_hnorm) copy_out(h, res->hnormsq, hp + 4 * h->hist_cap, steps);
  if (o.objevals) {
    copy_out(h, res->objevals, hp + 5 * h->hist_cap, steps);
    if (steps > 0) ADMM_CUDA(cudaMemcpyAsync(&res->objopt, hp + 5 * h->hist_cap + steps - 1, 8, cudaMemcpyDeviceToHost, h->stream));
  }
  copy_out(h, res->xopt, h->x.p, h->nA);
  copy_out(h, res->zopt, h->z.p, h->nB);
  copy_out(h, res->uopt, h->u.p, h->mc);
  if (history) {
    copy_out(h, res->xvals, h->xvals.p, h->nA * steps);
    copy_out(h, res->zvals, h->zvals.p, h->nB * steps);
    copy_out(h, res->uvals, h->uvals.p, h->mc * steps);
  }
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
}

// ---------------------------------------------------------------------------------------------
// regularisation path: nb lambda values on one cached factor (BASELINE.json configs[1]).  The two
// triangular solves become two triangular DMMA GEMMs with nb right-hand sides (multi-RHS TRSM on
// the inverse factor); every column runs admm.m's iteration with its own lambda and stop test.
// ---------------------------------------------------------------------------------------------
struct BatchOut {
  int64_t* steps; int32_t* status;
  double *xopt, *zopt, *uopt;                   // n x nb (ld = n)
  double *pnorm, *dnorm, *perr, *derr, *objevals;  // maxiters x nb (ld = maxiters), may be NULL
  double* loop_ms;
};

static void solve_lasso_batch(admm_b200_handle* h, const admm_b200_options& o, int64_t nb, const double* lambdas,
                              const BatchOut& out) {
  validate_options(h, o);
  ADMM_REQUIRE(h->kind == ADMM_B200_LASSO && h->tall, ADMM_B200_ERR_UNSUPPORTED,
               "lasso batch: only the tall (m >= n) lasso is built");
  ADMM_REQUIRE(nb >= 1 && nb <= 4096 && lambdas, ADMM_B200_ERR_INVALID, "lasso batch: bad batch size or null lambdas");
  ADMM_REQUIRE(!o.objevals, ADMM_B200_ERR_UNSUPPORTED, "lasso batch: objevals is not built for a batch");
  const int64_t n = h->n, ld = round_up(n, 2);
  const int64_t N = o.maxiters > 0 ? o.maxiters : 1000;
  DBuf X, Z, U, Y, T, XK, hist, thr, osc, part;
  LoopCtl* ctl = nullptr;
  int* done_count = nullptr;
  std::vector<double> hthr(nb), hosc(nb);
  for (int64_t j = 0; j < nb; ++j) {
    ADMM_REQUIRE(lambdas[j] >= 0, ADMM_B200_ERR_INVALID, "Argument lambda is not a nonnegative real number!");
    hthr[j] = lambdas[j] / o.rho;
    hosc[j] = lambdas[j];
  }
  auto cleanup = [&]() {
    X.release(); Z.release(); U.release(); Y.release(); T.release(); XK.release(); hist.release(); thr.release(); osc.release();
    part.release();
    if (ctl) cudaFree(ctl);
    if (done_count) cudaFree(done_count);
  };
  try {
    X.ensure(ld * nb); Z.ensure(ld * nb); U.ensure(ld * nb); Y.ensure(ld * nb); T.ensure(ld * nb); XK.ensure(ld * nb);
    hist.ensure(9 * N * nb); thr.ensure(nb); osc.ensure(nb);
    const int pg = prox_grid(n);
    part.ensure((int64_t)pg * PROX_NRED * nb);
    ADMM_CUDA(cudaMalloc(&ctl, sizeof(LoopCtl) * nb));
    ADMM_CUDA(cudaMalloc(&done_count, sizeof(int)));
    ctl_init_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, h->stream>>>(ctl, (int)nb);
    ADMM_CUDA(cudaGetLastError());
    ADMM_CUDA(cudaMemsetAsync(done_count, 0, sizeof(int), h->stream));
    ADMM_CUDA(cudaMemcpyAsync(thr.p, hthr.data(), nb * 8, cudaMemcpyHostToDevice, h->stream));
    ADMM_CUDA(cudaMemcpyAsync(osc.p, hosc.data(), nb * 8, cudaMemcpyHostToDevice, h->stream));
    ADMM_CUDA(cudaEventRecord(h->ev0, h->stream));
    // x0 = z0 = u0 = 0 (admm.m:252-254) for every column  ->  y0 = Dts
    ADMM_CUDA(cudaMemsetAsync(X.p, 0, (size_t)ld * nb * 8, h->stream));
    ADMM_CUDA(cudaMemsetAsync(Z.p, 0, (size_t)ld * nb * 8, h->stream));
    ADMM_CUDA(cudaMemsetAsync(U.p, 0, (size_t)ld * nb * 8, h->stream));
    for (int64_t j = 0; j < nb; ++j)
      ADMM_CUDA(cudaMemcpyAsync(Y.p + j * ld, h->dts.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
    LoopParams lp = make_loop_params(h, o, N, 0);
    lp.pnorm = hist.p; lp.dnorm = hist.p + N * nb; lp.perr = hist.p + 2 * N * nb; lp.derr = hist.p + 3 * N * nb;
    lp.hn = hist.p + 4 * N * nb; lp.obj = hist.p + 5 * N * nb;
    lp.dvals = hist.p + 6 * N * nb; lp.avals = hist.p + 7 * N * nb; lp.rst = hist.p + 8 * N * nb;
    ADMM_REQUIRE(lp.alg == 0, ADMM_B200_ERR_UNSUPPORTED, "lasso batch: options.fast is not built for a batch");
    int hdone = 0;
    run_bursts(
        h, o, N, 1,
        [&]() {
        GemmOpt g1;           // T = W * Y, W lower triangular
        g1.a_lower = 1;
        gemm(h, 0, 0, n, nb, n, 1.0, h->W.p, h->ldf, Y.p, ld, 0.0, T.p, ld, g1);
        GemmOpt g2;           // X = W' * T, W' upper triangular (operand read K-major from W)
        g2.a_upper = 1;
        gemm(h, 1, 0, n, nb, n, 1.0, h->W.p, h->ldf, T.p, ld, 0.0, X.p, ld, g2);
        ProxIdentArgs a;
        a.n = n; a.x = X.p; a.z = Z.p; a.u = U.p; a.dts = h->dts.p; a.y = Y.p;
        a.lb = a.ub = nullptr; a.zin = nullptr;
        a.thresh = 0.0; a.objscale = 0.0;
        a.kind = PROX_SOFT; a.next = NEXT_LASSO; a.obj_l1_of_x = 0;
        a.partials = part.p; a.ctl = ctl; a.lp = lp;
        a.xvals = a.zvals = a.uvals = nullptr;
        a.ld = ld; a.thresh_v = thr.p; a.objscale_v = osc.p; a.hist_stride = N; a.done_count = done_count; a.xkeep = XK.p;
        a.v = a.uhat = nullptr; a.zprev = a.uprev = nullptr;
        prox_ident_kernel<<<dim3(pg, (unsigned)nb), PROX_THREADS, 0, h->stream>>>(a);
        ADMM_CUDA(cudaGetLastError());
        h->launches++;
        },
        [&]() {
          ADMM_CUDA(cudaMemcpyAsync(&hdone, done_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
          ADMM_CUDA(cudaStreamSynchronize(h->stream));
          return hdone >= nb;
        });
    ADMM_CUDA(cudaEventRecord(h->ev1, h->stream));
    ADMM_CUDA(cudaEventSynchronize(h->ev1));
    float ms = 0;
    ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (out.loop_ms) *out.loop_ms = ms;
    std::vector<LoopCtl> hc(nb);
    ADMM_CUDA(cudaMemcpy(hc.data(), ctl, sizeof(LoopCtl) * nb, cudaMemcpyDeviceToHost));
    for (int64_t j = 0; j < nb; ++j) {
      if (out.steps) out.steps[j] = hc[j].it;
      if (out.status) out.status[j] = hc[j].status;
    }
    auto mat_out = [&](double* dst, const double* src) {
      if (dst) ADMM_CUDA(cudaMemcpy2DAsync(dst, (size_t)n * 8, src, (size_t)ld * 8, (size_t)n * 8, (size_t)nb, cudaMemcpyDefault, h->stream));
    };
    mat_out(out.xopt, XK.p); mat_out(out.zopt, Z.p); mat_out(out.uopt, U.p);
    copy_out(h, out.pnorm, hist.p, N * nb);
    copy_out(h, out.dnorm, hist.p + N * nb, N * nb);
    copy_out(h, out.perr, hist.p + 2 * N * nb, N * nb);
    copy_out(h, out.derr, hist.p + 3 * N * nb, N * nb);
    ADMM_CUDA(cudaStreamSynchronize(h->stream));
  } catch (...) {
    cudaStreamSynchronize(h->stream);
    cleanup();
    throw;
  }
  cleanup();
}

// ---------------------------------------------------------------------------------------------
// class batch of A = D problems sharing D (one-vs-all linear SVM, examples/mnistsvm.m:121-156)
// ---------------------------------------------------------------------------------------------
template <int NV>
static void gemvt_strided_t(admm_b200_handle* h, GemvtArgs& a, int grid, size_t smem) {
  static PerDevice conf_pd;
  size_t& conf = conf_pd(h->device);
  if (smem > conf && smem > 48 * 1024) {
    ADMM_CUDA(cudaFuncSetAttribute(gemvt_kernel<NV, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  gemvt_kernel<NV, 256><<<grid, 256, smem, h->stream>>>(a);
}

// out[k*ostride + j] = sum_r M[r + j*ld] * v[k*vstride + r], k < nvt (nvt in {2,4,8,10,16}; only nv are real)
static void gemvt_strided(admm_b200_handle* h, const double* M, int64_t ld, int64_t rows, int64_t cols, int nvt,
                          const double* vbase, int64_t vstride, double* obase, int64_t ostride) {
  GemvtArgs a;
  a.M = M; a.ld = ld; a.rows = rows; a.cols = cols; a.done = nullptr;
  a.ngroups = (cols + GEMVT_CG - 1) / GEMVT_CG;
  int64_t P = 1;
  if (a.ngroups < 8 * kNumSM) P = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((8 * kNumSM + a.ngroups - 1) / a.ngroups, 64), rows / 2048));
  int64_t per = round_up((rows + P - 1) / P, GEMVT_ROWS);
  P = (rows + per - 1) / per;
  a.P = (int)P; a.per = per;
  const int64_t nunits = P * a.ngroups;
  const int grid = (int)std::min<int64_t>(kNumSM, nunits);
  a.units_per_cta = (nunits + grid - 1) / grid;
  const size_t smem = (size_t)a.units_per_cta * GEMVT_CG * nvt * (256 / 32) * 8;
  ADMM_REQUIRE(smem <= 200 * 1024, ADMM_B200_ERR_UNSUPPORTED, "gemvt: too many column groups per CTA");
  for (int k = 0; k < 3; ++k) { a.v[k] = vbase; a.out[k] = obase; }
  a.vbase = vbase; a.vstride = vstride;
  a.scale = 1.0; a.addend = nullptr; a.addscale = 0.0;
  if (P == 1) { a.obase = obase; a.ostride = ostride; }
  else { h->cd_ws.ensure(P * cols * nvt); a.obase = h->cd_ws.p; a.ostride = P * cols; }
  switch (nvt) {
    case 2: gemvt_strided_t<2>(h, a, grid, smem); break;
    case 4: gemvt_strided_t<4>(h, a, grid, smem); break;
    case 8: gemvt_strided_t<8>(h, a, grid, smem); break;
    case 10: gemvt_strided_t<10>(h, a, grid, smem); break;
    default: gemvt_strided_t<16>(h, a, grid, smem); break;
  }
  ADMM_CUDA(cudaGetLastError());
  h->launches++;
  if (P > 1) {
    panel_reduce_strided_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)nvt), 256, 0, h->stream>>>(
        h->cd_ws.p, obase, ostride, nvt, (int)P, cols, 1.0);
    ADMM_CUDA(cudaGetLastError());
    h->launches++;
  }
}

struct UwBatchOut {
  int64_t* steps; int32_t* status;
  double *xopt, *zopt, *uopt;            // n x nb, m_local x nb, m_local x nb (ld = n / m_local)
  double *pnorm, *perr, *objevals;       // maxiters x nb or NULL
  double* loop_ms;
};

template <int NB>
static void uwb_pass1_t(admm_b200_handle* h, const UwbArgs& a, dim3 grid, size_t smem) {
  static PerDevice configured_pd;
  size_t& configured = configured_pd(h->device);
  if (!configured) {
    ADMM_CUDA(cudaFuncSetAttribute(uwb_gemm_prox_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = 1;
  }
  uwb_gemm_prox_kernel<NB><<<grid, UW_THREADS, smem, h->stream>>>(a);
}

template <int NBT>
static void persist_batch_launch(admm_b200_handle* h, PersistBatchArgs& a, int grid) {
  static PerDevice conf_pd;
  size_t& conf = conf_pd(h->device);
  const size_t smem = PersistBatchCfg<NBT>::smem_bytes(a.n, a.npad);
  if (smem > conf) {
    ADMM_CUDA(cudaFuncSetAttribute(uwb_persist_kernel<NBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  void* args[] = {&a};
  ADMM_CUDA(cudaLaunchCooperativeKernel((const void*)uwb_persist_kernel<NBT>, dim3(grid), dim3(OP_THREADS), args, smem, h->stream));
  h->launches++;
}

static size_t persist_batch_smem(int nbt, int64_t n, int64_t npad) {
  switch (nbt) {
    case 2: return PersistBatchCfg<2>::smem_bytes(n, npad);
    case 4: return PersistBatchCfg<4>::smem_bytes(n, npad);
    case 8: return PersistBatchCfg<8>::smem_bytes(n, npad);
    default: return PersistBatchCfg<10>::smem_bytes(n, npad);
  }
}

// the class batch as one persistent kernel per burst (persist_batch.cuh)
static bool persist_batch_ok(const admm_b200_handle* h, const admm_b200_options& o, int nbt, int64_t cbs) {
  if (getenv("ADMM_B200_NO_PERSIST") || getenv("ADMM_B200_NO_PERSIST_BATCH")) return false;
  if (nbt > 10 || o.objevals || !h->have_inverse || h->xsolve_eff != ADMM_B200_XSOLVE_INVFACTOR) return false;
  if (h->nranks > 1 && !h->p2p.ready) return false;
  if (h->n > (int64_t)PB_MAXN || (h->n % 4) != 0 || (int64_t)nbt * cbs > P2P_LLCAP) return false;
  if ((((uintptr_t)h->dD) & 15) != 0 || (h->ldD % 2) != 0) return false;
  return persist_batch_smem(nbt, h->n, round_up(h->n, 2)) <= (size_t)227 * 1024;
}

static void solve_unwrapped_batch(admm_b200_handle* h, const admm_b200_options& o, int64_t nb, const double* AUX,
                                  int64_t ldaux, const double* X0, const double* Z0, const double* U0,
                                  const UwBatchOut& out) {
  validate_options(h, o);
  ADMM_REQUIRE(is_unwrapped(h->kind) && h->have_inverse, ADMM_B200_ERR_STATE, "unwrapped batch: call admm_b200_setup_unwrapped first");
  ADMM_REQUIRE(nb >= 1 && nb <= 16 && AUX && ldaux >= h->m, ADMM_B200_ERR_INVALID, "unwrapped batch: 1 <= nb <= 16 label / target columns");
  ADMM_REQUIRE(o.nodualerror && o.relax == 1.0 && !o.fast && !o.convtest, ADMM_B200_ERR_UNSUPPORTED,
               "unwrapped batch: built for the unwrappedadmm.m configuration (nodualerror = 1, relax = 1, no fast / convtest)");
  const int64_t n = h->n, m = h->m, npad = round_up(n, 2), mpad = round_up(m, 2);
  const int64_t N = o.maxiters > 0 ? o.maxiters : 1000;
  const int nbt = nb <= 2 ? 2 : nb <= 4 ? 4 : nb <= 8 ? 8 : nb <= 10 ? 10 : 16;   // template width
  const int64_t cbs = npad + 16;
  DBuf X, XK, T, Z, U, R, A, CB, hist, part, ws;
  LoopCtl* ctl = nullptr;
  int* done_count = nullptr;
  auto cleanup = [&]() {
    DBuf* bs[] = {&X, &XK, &T, &Z, &U, &R, &A, &CB, &hist, &part, &ws};
    for (DBuf* b : bs) b->release();
    if (ctl) cudaFree(ctl);
    if (done_count) cudaFree(done_count);
  };
  try {
    X.ensure(npad * nbt); XK.ensure(npad * nbt); T.ensure(npad * nbt);
    Z.ensure(mpad * nbt); U.ensure(mpad * nbt); R.ensure(mpad * nbt); A.ensure(mpad * nbt);
    CB.ensure(cbs * nbt); hist.ensure(9 * N * nbt);
    ADMM_CUDA(cudaMalloc(&ctl, sizeof(LoopCtl) * nbt));
    ADMM_CUDA(cudaMalloc(&done_count, sizeof(int)));
    cudaStream_t st = h->stream;
    for (DBuf* b : {&X, &XK, &T, &Z, &U, &R, &A, &CB}) ADMM_CUDA(cudaMemsetAsync(b->p, 0, (size_t)b->cap * 8, st));
    ADMM_CUDA(cudaMemsetAsync(done_count, 0, sizeof(int), st));
    ctl_init_kernel<<<1, 32, 0, st>>>(ctl, nbt);
    ADMM_CUDA(cudaGetLastError());
    auto mat_in = [&](double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows) {
      if (src) ADMM_CUDA(cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)rows * 8, (size_t)nb, cudaMemcpyDefault, st));
    };
    mat_in(A.p, mpad, AUX, ldaux, m);
    mat_in(X.p, npad, X0, n, n);
    mat_in(Z.p, mpad, Z0, m, m);
    mat_in(U.p, mpad, U0, m, m);
    LoopParams lp = make_loop_params(h, o, N, 0);
    lp.pnorm = hist.p; lp.dnorm = hist.p + N * nbt; lp.perr = hist.p + 2 * N * nbt; lp.derr = hist.p + 3 * N * nbt;
    lp.hn = hist.p + 4 * N * nbt; lp.obj = hist.p + 5 * N * nbt;
    lp.dvals = hist.p + 6 * N * nbt; lp.avals = hist.p + 7 * N * nbt; lp.rst = hist.p + 8 * N * nbt;
    if (persist_batch_ok(h, o, nbt, cbs)) {
      // ---- persistent path: Q = D inv(R)', t_c = sum_g Q_g' r_{g,c}; no x inside the loop ----
      if (h->nranks == 1) ensure_mailbox(h);
      ensure_Q(h);
      DBuf TL;
      try {
        TL.ensure(cbs * nbt);
        ADMM_CUDA(cudaMemsetAsync(TL.p, 0, (size_t)cbs * nbt * 8, st));
        ADMM_CUDA(cudaEventRecord(h->ev0, st));
        uwb_first_rhs_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)nb), 256, 0, st>>>(m, (int)nb, mpad, Z.p, U.p, A.p, uw_kind(h->kind), R.p);
        ADMM_CUDA(cudaGetLastError());
        gemvt_strided(h, h->Qm.p, h->ldq, m, n, nbt, R.p, mpad, CB.p, cbs);     // t_c of the first iteration
        allreduce_sum(h, CB.p, cbs * nbt);
        ADMM_CUDA(cudaMemcpyAsync(TL.p, CB.p, (size_t)cbs * nbt * 8, cudaMemcpyDeviceToDevice, st));
        PersistBatchArgs pa;
        pa.Q = h->Qm.p; pa.ld = h->ldq; pa.m = m; pa.n = n; pa.npad = npad;
        pa.Z = Z.p; pa.U = U.p; pa.AUX = A.p; pa.ldm = mpad; pa.nb = (int)nb;
        pa.rho = o.rho; pa.C = h->svmC; pa.kind = uw_kind(h->kind);
        pa.ntiles = (m + PB_R - 1) / PB_R;
        const int grid = (int)std::min<int64_t>(kNumSM, pa.ntiles);
        h->op_dpart.ensure((int64_t)grid * nbt * npad);
        h->uw_partials.ensure((int64_t)grid * nbt * UW_NRED);
        pa.dpart = h->op_dpart.p; pa.partials = h->uw_partials.p; pa.tcur = CB.p; pa.tlast = TL.p; pa.cbs = cbs;
        pa.mail = h->p2p.dev; pa.ctl = ctl; pa.lp = lp; pa.hist_stride = N; pa.done_count = done_count;
        pa.m_total = (double)h->m_total;
        pa.prof = nullptr;
        static const bool want_prof = getenv("ADMM_B200_PERSIST_PROF") != nullptr;
        if (want_prof) {
          ADMM_CUDA(cudaMalloc(&pa.prof, (size_t)grid * 8 * sizeof(long long)));
          ADMM_CUDA(cudaMemsetAsync(pa.prof, 0, (size_t)grid * 8 * sizeof(long long), st));
        }
        const int check = std::max(1, o.check_every);
        int64_t enq = 0;
        int hdone = 0;
        while (true) {
          pa.burst = (int)std::min<int64_t>(check, N - enq);
          switch (nbt) {
            case 2: persist_batch_launch<2>(h, pa, grid); break;
            case 4: persist_batch_launch<4>(h, pa, grid); break;
            case 8: persist_batch_launch<8>(h, pa, grid); break;
            default: persist_batch_launch<10>(h, pa, grid); break;
          }
          enq += pa.burst;
          ADMM_CUDA(cudaMemcpyAsync(&hdone, done_count, sizeof(int), cudaMemcpyDeviceToHost, st));
          ADMM_CUDA(cudaStreamSynchronize(st));
          if (hdone >= nb || enq >= N) break;
        }
        if (pa.prof) {
          std::vector<long long> hp((size_t)grid * 8);
          ADMM_CUDA(cudaMemcpy(hp.data(), pa.prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost));
          cudaFree(pa.prof);
          const double its = std::max<double>(1.0, (double)enq);
          const char* names[7] = {"tile wait + T*t", "shuffles", "proxes", "T'*r", "writes + barrier", "owner sums + exchange", "barrier 2 + stop"};
          char line[256];
          std::string txt;               // one write per rank: the ranks' reports must not interleave
          snprintf(line, sizeof line, "batch persist profile (rank %d of %d, %lld rows here, %d CTAs, %d classes, %.0f iterations): cycles per iteration, mean / max over CTAs\n",
                   h->rank, h->nranks, (long long)h->m, grid, (int)nb, its);
          txt += line;
          for (int k = 0; k < 7; ++k) {
            double mean = 0.0, mx = 0.0;
            for (int c = 0; c < grid; ++c) { const double v = (double)hp[(size_t)c * 8 + k] / its; mean += v; mx = std::max(mx, v); }
            snprintf(line, sizeof line, "  [r%d] %-22s %9.0f / %9.0f\n", h->rank, names[k], mean / grid, mx);
            txt += line;
          }
          fputs(txt.c_str(), stderr);
        }
        // x_c of the last iteration each class ran: X = inv(R)' TLAST = W' TLAST
        GemmOpt gx; gx.a_upper = 1;
        gemm(h, 1, 0, n, nb, n, 1.0, h->W.p, h->ldf, TL.p, cbs, 0.0, XK.p, npad, gx);
      } catch (...) {
        cudaStreamSynchronize(st);
        TL.release();
        throw;
      }
      cudaStreamSynchronize(st);
      TL.release();
    } else {
    // pass-1 launch geometry
    const int64_t rb = (m + UW_ROWS - 1) / UW_ROWS;
    // columns are swept in 128-wide chunks INSIDE a CTA; the grid is split over columns (partials through
    // a workspace) only when there are too few row blocks to fill the machine
    int64_t chunks = 1;
    if (rb < 3 * kNumSM) chunks = std::min<int64_t>((3 * kNumSM + rb - 1) / rb, std::max<int64_t>(1, n / 128));
    const int64_t cpc = (n + chunks - 1) / chunks;
    chunks = (n + cpc - 1) / cpc;
    part.ensure(rb * nb * UW_NRED);
    if (chunks > 1) { ws.ensure(chunks * nb * m); ensure_tickets(h, rb); }
    UwbArgs a;
    a.D = h->dD; a.ld = h->ldD; a.m = m; a.n = n; a.X = X.p; a.ldx = npad; a.Z = Z.p; a.U = U.p; a.R = R.p; a.AUX = A.p;
    a.ldm = mpad; a.nb = (int)nb; a.rho = o.rho; a.C = h->svmC; a.kind = uw_kind(h->kind); a.cols_per_chunk = cpc;
    a.ws = ws.p; a.tickets = h->tickets; a.partials = part.p; a.grid_ticket = h->grid_ticket;
    a.cb = CB.p; a.cb_stride = cbs; a.scal_off = npad; a.ctl = ctl;
    const size_t smem1 = (size_t)128 * (nbt + (nbt & 1)) * 8;   // one 128-column chunk of X, classes in pairs
    UwbEpiArgs e;
    e.n = n; e.X = X.p; e.ldx = npad; e.XK = XK.p; e.cb = CB.p; e.cb_stride = cbs; e.scal_off = npad;
    e.m_total = (double)h->m_total; e.kind = a.kind; e.C = h->svmC; e.ctl = ctl; e.lp = lp; e.hist_stride = N;
    e.done_count = done_count;
    ADMM_CUDA(cudaEventRecord(h->ev0, st));
    // rhs of the first x-update: d_k = D'(z0_k - u0_k)
    uwb_first_rhs_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)nb), 256, 0, st>>>(m, (int)nb, mpad, Z.p, U.p, A.p, a.kind, R.p);
    ADMM_CUDA(cudaGetLastError());
    gemvt_strided(h, h->dD, h->ldD, m, n, nbt, R.p, mpad, CB.p, cbs);
    allreduce_sum(h, CB.p, cbs * nbt);
    int hdone = 0;
    run_bursts(
        h, o, N, 1,
        [&]() {
        GemmOpt g1; g1.a_lower = 1;      // T = W * [d_1 .. d_nb]
        gemm(h, 0, 0, n, nb, n, 1.0, h->W.p, h->ldf, CB.p, cbs, 0.0, T.p, npad, g1);
        GemmOpt g2; g2.a_upper = 1;      // X = W' * T
        gemm(h, 1, 0, n, nb, n, 1.0, h->W.p, h->ldf, T.p, npad, 0.0, X.p, npad, g2);
        dim3 grid((unsigned)rb, (unsigned)chunks);
        switch (nbt) {
          case 2: uwb_pass1_t<2>(h, a, grid, smem1); break;
          case 4: uwb_pass1_t<4>(h, a, grid, smem1); break;
          case 8: uwb_pass1_t<8>(h, a, grid, smem1); break;
          case 10: uwb_pass1_t<10>(h, a, grid, smem1); break;
          default: uwb_pass1_t<16>(h, a, grid, smem1); break;
        }
        ADMM_CUDA(cudaGetLastError());
        h->launches++;
        gemvt_strided(h, h->dD, h->ldD, m, n, nbt, R.p, mpad, CB.p, cbs);
        allreduce_sum(h, CB.p, cbs * nbt);
        uwb_epilogue_kernel<<<(unsigned)nb, 256, 0, st>>>(e);
        ADMM_CUDA(cudaGetLastError());
        h->launches++;
        },
        [&]() {
          ADMM_CUDA(cudaMemcpyAsync(&hdone, done_count, sizeof(int), cudaMemcpyDeviceToHost, st));
          ADMM_CUDA(cudaStreamSynchronize(st));
          return hdone >= nb;
        });
    }
    ADMM_CUDA(cudaEventRecord(h->ev1, st));
    ADMM_CUDA(cudaEventSynchronize(h->ev1));
    p2p_check(h);
    float ms = 0;
    ADMM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (out.loop_ms) *out.loop_ms = ms;
    std::vector<LoopCtl> hc(nb);
    ADMM_CUDA(cudaMemcpy(hc.data(), ctl, sizeof(LoopCtl) * nb, cudaMemcpyDeviceToHost));
    for (int64_t j = 0; j < nb; ++j) {
      if (out.steps) out.steps[j] = hc[j].it;
      if (out.status) out.status[j] = hc[j].status;
    }
    auto mat_out = [&](double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows) {
      if (dst) ADMM_CUDA(cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)rows * 8, (size_t)nb, cudaMemcpyDefault, st));
    };
    mat_out(out.xopt, n, XK.p, npad, n);
    mat_out(out.zopt, m, Z.p, mpad, m);
    mat_out(out.uopt, m, U.p, mpad, m);
    copy_out(h, out.pnorm, hist.p, N * nb);
    copy_out(h, out.perr, hist.p + 2 * N * nbt, N * nb);
    if (o.objevals) copy_out(h, out.objevals, hist.p + 5 * N * nbt, N * nb);
    ADMM_CUDA(cudaStreamSynchronize(st));
  } catch (...) {
    cudaStreamSynchronize(h->stream);
    cleanup();
    throw;
  }
  cleanup();
}

}  // namespace admmb200

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
#define ADMM_API_BEGIN try {
#define ADMM_API_END                                                        \
  }                                                                         \
  catch (const admmb200::CudaFail&) { return ADMM_B200_ERR_CUDA; }          \
  catch (const admmb200::ArgFail& f) { return f.code; }                     \
  catch (const std::exception& e) {                                         \
    admmb200::set_error("internal error: %s", e.what());                    \
    return ADMM_B200_ERR_CUDA;                                              \
  }                                                                         \
  return ADMM_B200_OK;

extern "C" {

int admm_b200_version(void) { return ADMM_B200_VERSION; }
const char* admm_b200_last_error(void) { return admmb200::g_err; }

void admm_b200_default_options(admm_b200_options* o) {
  if (!o) return;
  o->rho = 1.0; o->relax = 1.0; o->abstol = 1e-5; o->reltol = 1e-3; o->convtol = 1e-10; o->hnormtol = 1e-6;
  o->maxiters = 1000; o->domaxiters = 0; o->stopcond = ADMM_B200_STOP_STANDARD; o->nodualerror = 0;
  o->convtest = 0; o->objevals = 0; o->history = 1; o->xsolve = ADMM_B200_XSOLVE_INVFACTOR; o->check_every = 8;
  o->fast = 0; o->fasttype = 1; o->restart = 0.999; o->dvaltol = 1e-8;
  o->graph = 1; o->reserved = 0;
}

int admm_b200_create(int device, admm_b200_handle** out) {
  ADMM_API_BEGIN
  ADMM_REQUIRE(out != nullptr, ADMM_B200_ERR_INVALID, "null output pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    set_error("no usable CUDA device (%s); libadmm_b200 has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return ADMM_B200_ERR_CUDA;
  }
  ADMM_REQUIRE(device >= 0 && device < count, ADMM_B200_ERR_INVALID, "device %d out of range (0..%d)", device, count - 1);
  ADMM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ADMM_CUDA(cudaGetDeviceProperties(&prop, device));
  ADMM_REQUIRE(prop.major == 10, ADMM_B200_ERR_UNSUPPORTED,
               "device %d is sm_%d%d; libadmm_b200 is built for sm_100a (Blackwell B200) only", device, prop.major,
               prop.minor);
  admm_b200_handle* h = new admm_b200_handle();
  h->device = device;
  ADMM_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  {
    // The look-ahead Cholesky runs its latency-bound chain of small kernels on a HIGH priority stream and the bulk
    // trailing updates on a LOW priority one: without priorities the block scheduler dispatches kernels in launch
    // order, so a small chain kernel would wait behind every queued CTA of a big SYRK and nothing would overlap.
    int least = 0, greatest = 0;
    ADMM_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    ADMM_CUDA(cudaStreamCreateWithPriority(&h->stream_hi, cudaStreamNonBlocking, greatest));
    ADMM_CUDA(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, (least + greatest) / 2));
    ADMM_CUDA(cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, least));
    ADMM_CUDA(cudaStreamCreateWithPriority(&h->stream4, cudaStreamNonBlocking, least));
  }
  for (auto& e : h->ev_la) ADMM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  h->stream = h->own_stream;
  ADMM_CUDA(cudaEventCreate(&h->ev0));
  ADMM_CUDA(cudaEventCreate(&h->ev1));
  for (auto& e : h->evp) ADMM_CUDA(cudaEventCreate(&e));
  ADMM_CUDA(cudaMalloc(&h->ctl, sizeof(LoopCtl)));
  ADMM_CUDA(cudaMemset(h->ctl, 0, sizeof(LoopCtl)));
  ADMM_CUDA(cudaMalloc(&h->fail, sizeof(int)));
  ADMM_CUDA(cudaMalloc(&h->grid_ticket, sizeof(unsigned)));
  ADMM_CUDA(cudaMemset(h->grid_ticket, 0, sizeof(unsigned)));
  ADMM_CUDA(cudaMallocHost(&h->h_ctl, sizeof(LoopCtl)));
  ADMM_CUDA(cudaDeviceSynchronize());   // the memsets above ran on the legacy stream; the handle's stream is non-blocking
  *out = h;
  ADMM_API_END
}

int admm_b200_destroy(admm_b200_handle* h) {
  ADMM_API_BEGIN
  if (!h) return ADMM_B200_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  DBuf* bufs[] = {&h->ownD, &h->s, &h->dts, &h->L, &h->W, &h->WT, &h->x, &h->z, &h->u, &h->y, &h->t1, &h->t2,
                  &h->x0, &h->z0, &h->u0, &h->partials, &h->hist, &h->xvals, &h->zvals, &h->uvals, &h->gemm_ws,
                  &h->gemv_ws, &h->scratch, &h->cd_ws, &h->aux, &h->rvec, &h->dzvec, &h->cb, &h->uw_partials, &h->zz, &h->uu, &h->tvtab, &h->Pfull, &h->lb, &h->ub, &h->fv, &h->fuhat, &h->fzprev, &h->fuprev,
                  &h->op_dpart, &h->Q2, &h->s2, &h->dts2, &h->G1, &h->G2, &h->L2, &h->W2, &h->WT2, &h->zsol};
  for (DBuf* b : bufs) b->release();
  for (ColdotPlan* p : h->plans) {
    cudaFree(p->d_cta_pos); cudaFree(p->d_pos_item); cudaFree(p->d_order); cudaFree(p->d_items);
    delete p;
  }
  for (SymtriPlan* p : h->st_plans) {
    cudaFree(p->d_cta_round); cudaFree(p->d_round_ent); cudaFree(p->d_ents);
    delete p;
  }
  h->st_part.release(); h->Qm.release(); h->tcur.release(); h->tlast.release();
  p2p_teardown(h);
  if (h->tickets) cudaFree(h->tickets);
  if (h->grid_ticket) cudaFree(h->grid_ticket);
  comm_destroy(h);
  if (h->ctl) cudaFree(h->ctl);
  if (h->fail) cudaFree(h->fail);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (auto& e : h->evp) if (e) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream3) cudaStreamDestroy(h->stream3);
  if (h->stream4) cudaStreamDestroy(h->stream4);
  for (int b = 0; b < admm_b200_handle::kPinBufs; ++b) {
    if (h->pin_buf[b]) cudaFreeHost(h->pin_buf[b]);
    if (h->pin_ev[b]) cudaEventDestroy(h->pin_ev[b]);
  }
  if (h->stream_hi) cudaStreamDestroy(h->stream_hi);
  for (auto& e : h->ev_pool) cudaEventDestroy(e);
  h->chol_ws.release();
  for (auto& e : h->ev_la) if (e) cudaEventDestroy(e);
  delete h;
  ADMM_API_END
}

int admm_b200_set_stream(admm_b200_handle* h, void* cuda_stream) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  ADMM_API_END
}

int admm_b200_synchronize(admm_b200_handle* h) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  ADMM_API_END
}

int admm_b200_setup_lasso(admm_b200_handle* h, int64_t m, int64_t n, const double* D, int64_t ldD, const double* s,
                          double rho, int32_t xsolve) {
  ADMM_API_BEGIN
  check_handle(h);
  setup_lasso(h, m, m, n, D, ldD, s, rho, xsolve);
  ADMM_API_END
}

int admm_b200_setup_lasso_sharded(admm_b200_handle* h, int64_t m_local, int64_t m_total, int64_t n, const double* D,
                                  int64_t ldD, const double* s, double rho, int32_t xsolve) {
  ADMM_API_BEGIN
  check_handle(h);
  setup_lasso(h, m_local, m_total, n, D, ldD, s, rho, xsolve);
  ADMM_API_END
}

int admm_b200_setup_unwrapped(admm_b200_handle* h, int32_t kind, int64_t m_local, int64_t m_total, int64_t n,
                              const double* D, int64_t ldD, const double* aux, double C) {
  ADMM_API_BEGIN
  check_handle(h);
  setup_unwrapped(h, kind, m_local, m_total, n, D, ldD, aux, C);
  ADMM_API_END
}

int admm_b200_setup_basispursuit(admm_b200_handle* h, int64_t m, int64_t n, const double* D, int64_t ldD,
                                 const double* s) {
  ADMM_API_BEGIN
  check_handle(h);
  setup_bp(h, m, n, D, ldD, s);
  ADMM_API_END
}

int admm_b200_setup_totalvariation(admm_b200_handle* h, int64_t n, const double* s, double lambda) {
  ADMM_API_BEGIN
  check_handle(h);
  setup_tv(h, n, s, lambda);
  ADMM_API_END
}

int admm_b200_setup_quadratic(admm_b200_handle* h, int32_t kind, int64_t n, const double* P, int64_t ldP, const double* q,
                              double r, double rho, const double* lb, const double* ub) {
  ADMM_API_BEGIN
  check_handle(h);
  setup_quadratic(h, kind, n, P, ldP, q, r, rho, lb, ub);
  ADMM_API_END
}

int admm_b200_setup_model(admm_b200_handle* h, int64_t m, int64_t n, const double* P, int64_t ldP, const double* Q,
                          int64_t ldQ, const double* r, const double* s, double rho) {
  ADMM_API_BEGIN
  admmb200::check_handle(h);
  admmb200::setup_model(h, m, n, P, ldP, Q, ldQ, r, s, rho);
  ADMM_API_END
}

int admm_b200_get_unique_id(void* out128) {
  ADMM_API_BEGIN
  ADMM_REQUIRE(out128 != nullptr, ADMM_B200_ERR_INVALID, "null output");
  nccl_load();
  NcclId id;
  ADMM_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  ADMM_API_END
}

int admm_b200_comm_init(admm_b200_handle* h, int rank, int nranks, const void* unique_id128) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks && unique_id128, ADMM_B200_ERR_INVALID, "comm_init: bad arguments");
  comm_destroy(h);
  if (nranks == 1) return ADMM_B200_OK;
  nccl_load();
  NcclId id;
  memcpy(&id, unique_id128, sizeof(id));
  void* comm = nullptr;
  ADMM_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  h->comm = comm;
  h->rank = rank;
  h->nranks = nranks;
  p2p_setup(h);     // also performs the first collectives, so NCCL's lazy channel setup is not inside a timed setup
  ADMM_API_END
}

int admm_b200_comm_ipc_export(admm_b200_handle* h, int rank, int nranks, void* out_handle64) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(nranks >= 2 && nranks <= admmb200::P2P_MAXRANKS && rank >= 0 && rank < nranks && out_handle64, ADMM_B200_ERR_INVALID,
               "comm_ipc_export: bad arguments (2 <= nranks <= %d)", admmb200::P2P_MAXRANKS);
  comm_destroy(h);
  h->rank = rank;
  h->nranks = nranks;
  const P2PSlot mine = p2p_alloc_local(h);
  if (!mine.ok) {
    comm_destroy(h);
    ADMM_REQUIRE(false, ADMM_B200_ERR_COMM, "comm_ipc_export: CUDA IPC is not available for the mailbox (or ADMM_B200_NO_P2P is set)");
  }
  memcpy(out_handle64, &mine.hd, 64);
  ADMM_API_END
}

int admm_b200_comm_ipc_attach(admm_b200_handle* h, const void* handles) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(handles && h->nranks >= 2 && h->p2p.local && !h->comm && !h->p2p.ready, ADMM_B200_ERR_STATE,
               "comm_ipc_attach: call admm_b200_comm_ipc_export on this handle first");
  std::vector<P2PSlot> all(h->nranks);
  for (int r = 0; r < h->nranks; ++r) {
    memcpy(&all[r].hd, (const char*)handles + 64 * r, 64);
    all[r].ok = 1;
  }
  if (!p2p_map_peers(h, all.data())) {
    comm_destroy(h);
    ADMM_REQUIRE(false, ADMM_B200_ERR_COMM, "comm_ipc_attach: a peer's mailbox could not be mapped (cudaIpcOpenMemHandle)");
  }
  h->p2p.ready = true;
  ADMM_API_END
}

int admm_b200_comm_destroy(admm_b200_handle* h) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  comm_destroy(h);
  ADMM_API_END
}

int admm_b200_allreduce(admm_b200_handle* h, double* buf, int64_t count) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(buf && count >= 0, ADMM_B200_ERR_INVALID, "allreduce: bad arguments");
  if (admmb200::is_device_ptr(buf)) {
    allreduce_sum(h, buf, count);
  } else {
    h->scratch.ensure(count);
    copy_in(h, h->scratch.p, buf, count);
    allreduce_sum(h, h->scratch.p, count);
    copy_out(h, buf, h->scratch.p, count);
  }
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  ADMM_API_END
}

int admm_b200_set_lambda(admm_b200_handle* h, double lambda) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(lambda >= 0, ADMM_B200_ERR_INVALID, "Argument lambda is not a nonnegative real number!");
  h->lambda = lambda;
  ADMM_API_END
}

int admm_b200_set_init(admm_b200_handle* h, const double* x0, const double* z0, const double* u0) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(h->kind != 0, ADMM_B200_ERR_STATE, "no problem set up on this handle");
  if (!x0 && !z0 && !u0) {
    h->have_init = false;
    return ADMM_B200_OK;
  }
  h->x0.ensure(h->nA); h->z0.ensure(h->nB); h->u0.ensure(h->mc);
  if (x0) copy_in(h, h->x0.p, x0, h->nA); else ADMM_CUDA(cudaMemsetAsync(h->x0.p, 0, (size_t)h->nA * 8, h->stream));
  if (z0) copy_in(h, h->z0.p, z0, h->nB); else ADMM_CUDA(cudaMemsetAsync(h->z0.p, 0, (size_t)h->nB * 8, h->stream));
  if (u0) copy_in(h, h->u0.p, u0, h->mc); else ADMM_CUDA(cudaMemsetAsync(h->u0.p, 0, (size_t)h->mc * 8, h->stream));
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  h->have_init = true;
  ADMM_API_END
}

int admm_b200_solve(admm_b200_handle* h, const admm_b200_options* opts, admm_b200_result* res) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(opts != nullptr, ADMM_B200_ERR_INVALID, "Given options is not a struct! At least pass empty struct!");
  solve(h, *opts, res);
  ADMM_API_END
}

int admm_b200_solve_lasso_batch(admm_b200_handle* h, const admm_b200_options* opts, int64_t nb, const double* lambdas,
                                int64_t* steps, int32_t* status, double* xopt, double* zopt, double* uopt,
                                double* pnorm, double* dnorm, double* perr, double* derr, double* loop_ms) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(opts != nullptr, ADMM_B200_ERR_INVALID, "Given options is not a struct! At least pass empty struct!");
  BatchOut out{steps, status, xopt, zopt, uopt, pnorm, dnorm, perr, derr, nullptr, loop_ms};
  solve_lasso_batch(h, *opts, nb, lambdas, out);
  ADMM_API_END
}

int admm_b200_solve_unwrapped_batch(admm_b200_handle* h, const admm_b200_options* opts, int64_t nb, const double* aux,
                                    int64_t ldaux, const double* X0, const double* Z0, const double* U0, int64_t* steps,
                                    int32_t* status, double* xopt, double* zopt, double* uopt, double* pnorm, double* perr,
                                    double* objevals, double* loop_ms) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(opts != nullptr, ADMM_B200_ERR_INVALID, "Given options is not a struct! At least pass empty struct!");
  UwBatchOut out{steps, status, xopt, zopt, uopt, pnorm, perr, objevals, loop_ms};
  solve_unwrapped_batch(h, *opts, nb, aux, ldaux, X0, Z0, U0, out);
  ADMM_API_END
}

int admm_b200_get_dims(admm_b200_handle* h, int64_t* nA, int64_t* nB, int64_t* m) {
  ADMM_API_BEGIN
  check_handle(h);
  if (nA) *nA = h->nA;
  if (nB) *nB = h->nB;
  if (m) *m = h->mc;
  ADMM_API_END
}

int admm_b200_get_factor(admm_b200_handle* h, double* L, int64_t ldL, int64_t* k) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(h->have_factor, ADMM_B200_ERR_STATE, "no cached factor: call a setup function first");
  if (k) *k = h->k;
  if (L) {
    ADMM_REQUIRE(ldL >= h->k, ADMM_B200_ERR_INVALID, "ldL too small");
    ADMM_CUDA(cudaMemcpy2DAsync(L, (size_t)ldL * 8, h->L.p, (size_t)h->ldf * 8, (size_t)h->k * 8, (size_t)h->k,
                                cudaMemcpyDefault, h->stream));
    ADMM_CUDA(cudaStreamSynchronize(h->stream));
  }
  ADMM_API_END
}

// stage a host-or-device matrix into a handle-owned scratch (returns device pointer + ld)
namespace {
struct Staged {
  admmb200::DBuf buf;
  const double* p = nullptr;
  int64_t ld = 0;
  ~Staged() { buf.release(); }
};
void stage_any(admm_b200_handle* h, Staged& st, const double* A, int64_t rows, int64_t cols, int64_t lda) {
  if (admmb200::is_device_ptr(A)) {
    st.p = A;
    st.ld = lda;
    return;
  }
  st.ld = admmb200::round_up(rows, 2);
  st.buf.ensure(st.ld * std::max<int64_t>(cols, 1));
  ADMM_CUDA(cudaMemcpy2DAsync(st.buf.p, (size_t)st.ld * 8, A, (size_t)lda * 8, (size_t)rows * 8, (size_t)cols,
                              cudaMemcpyHostToDevice, h->stream));
  st.p = st.buf.p;
}
void unstage(admm_b200_handle* h, Staged& st, double* A, int64_t rows, int64_t cols, int64_t lda) {
  if (st.p == A) return;
  ADMM_CUDA(cudaMemcpy2DAsync(A, (size_t)lda * 8, st.p, (size_t)st.ld * 8, (size_t)rows * 8, (size_t)cols,
                              cudaMemcpyDeviceToHost, h->stream));
}
}  // namespace

int admm_b200_dgemm(admm_b200_handle* h, int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha,
                    const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                    int lower_only) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(M >= 0 && N >= 0 && K >= 0 && A && B && C, ADMM_B200_ERR_INVALID, "dgemm: bad arguments");
  Staged sa, sb, sc;
  stage_any(h, sa, A, transa ? K : M, transa ? M : K, lda);
  stage_any(h, sb, B, transb ? N : K, transb ? K : N, ldb);
  stage_any(h, sc, C, M, N, ldc);
  GemmOpt o;
  o.lower_only = lower_only;
  gemm(h, transa, transb, M, N, K, alpha, sa.p, sa.ld, sb.p, sb.ld, beta, const_cast<double*>(sc.p), sc.ld, o);
  unstage(h, sc, C, M, N, ldc);
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  ADMM_API_END
}

int admm_b200_gram(admm_b200_handle* h, int trans, int64_t m, int64_t n, const double* D, int64_t ldD, double scale,
                   double shift, double* G, int64_t ldG) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(m > 0 && n > 0 && D && G, ADMM_B200_ERR_INVALID, "gram: bad arguments");
  const int64_t k = trans ? n : m;
  Staged sd, sg;
  stage_any(h, sd, D, m, n, ldD);
  const bool gdev = admmb200::is_device_ptr(G);
  if (gdev) { sg.p = G; sg.ld = ldG; }
  else { sg.ld = admmb200::round_up(k, 2); sg.buf.ensure(sg.ld * k); sg.p = sg.buf.p; }
  GemmOpt o;
  o.lower_only = 1;
  o.diag_add = shift;
  if (trans) gemm(h, 1, 0, n, n, m, scale, sd.p, sd.ld, sd.p, sd.ld, 0.0, const_cast<double*>(sg.p), sg.ld, o);
  else gemm(h, 0, 1, m, m, n, scale, sd.p, sd.ld, sd.p, sd.ld, 0.0, const_cast<double*>(sg.p), sg.ld, o);
  symmetrize(h, const_cast<double*>(sg.p), k, sg.ld);
  if (!gdev) unstage(h, sg, G, k, k, ldG);
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  ADMM_API_END
}

int admm_b200_potrf(admm_b200_handle* h, int64_t k, double* A, int64_t lda, double* Winv, int64_t ldw) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(k > 0 && A && lda >= k, ADMM_B200_ERR_INVALID, "potrf: bad arguments");
  Staged sa;
  stage_any(h, sa, A, k, k, lda);
  admmb200::DBuf w;
  const int64_t ldwi = admmb200::round_up(k, 16);
  w.ensure(ldwi * k);
  try {
    potrf_blocked(h, k, const_cast<double*>(sa.p), sa.ld, w.p, ldwi, Winv != nullptr);
  } catch (...) {
    w.release();
    throw;
  }
  unstage(h, sa, A, k, k, lda);
  if (Winv) {
    ADMM_CUDA(cudaMemcpy2DAsync(Winv, (size_t)ldw * 8, w.p, (size_t)ldwi * 8, (size_t)k * 8, (size_t)k,
                                cudaMemcpyDefault, h->stream));
  }
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  w.release();
  ADMM_API_END
}

int admm_b200_factor_solve(admm_b200_handle* h, const double* b, double* x, int32_t xsolve) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(h->have_factor && b && x, ADMM_B200_ERR_STATE, "factor_solve: no cached factor or null argument");
  alloc_iterates(h);
  h->t1.ensure(admmb200::round_up(h->k, 2));
  h->t2.ensure(admmb200::round_up(h->k, 2));
  h->y.ensure(admmb200::round_up(h->k, 2));
  copy_in(h, h->y.p, b, h->k);
  factor_solve(h, h->y.p, h->t1.p, h->t2.p, xsolve, nullptr);
  copy_out(h, x, h->t2.p, h->k);
  ADMM_CUDA(cudaStreamSynchronize(h->stream));
  ADMM_API_END
}

int admm_b200_iterate_raw(admm_b200_handle* h, const admm_b200_options* opts, int which, int reps) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(opts && reps >= 0 && which >= 0 && which <= 2, ADMM_B200_ERR_INVALID, "iterate_raw: bad arguments");
  validate_options(h, *opts);
  prepare_loop(h, *opts, std::max<int64_t>(h->hist_cap, 16), false);
  LoopParams lp = make_loop_params(h, *opts, h->hist_cap, 1);
  admm_b200_options o = *opts;
  o.objevals = 0;
  lp.objevals = 0;
  if (!h->iter_ready) {
    load_init(h);
    enqueue_first_rhs(h, o);
  }
  ADMM_CUDA(cudaMemsetAsync(&h->ctl->done, 0, sizeof(int), h->stream));
  for (int r = 0; r < reps; ++r) enqueue_iteration(h, o, lp, which, false);
  ADMM_API_END
}

int64_t admm_b200_launch_count(admm_b200_handle* h) { return h ? h->launches : 0; }
int64_t admm_b200_graph_replays(admm_b200_handle* h) { return h ? h->graph_replays : 0; }

int admm_b200_get_setup_phases(admm_b200_handle* h, double* out4) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(out4 != nullptr, ADMM_B200_ERR_INVALID, "null output");
  for (int i = 0; i < 4; ++i) out4[i] = h->phase_ms[i];
  ADMM_API_END
}

int admm_b200_get_info(admm_b200_handle* h, admm_b200_info* out) {
  ADMM_API_BEGIN
  check_handle(h);
  ADMM_REQUIRE(out != nullptr, ADMM_B200_ERR_INVALID, "null output");
  out->generation = h->generation;
  out->zero_cols = h->zero_cols;
  out->diag_ratio = h->diag_ratio;
  out->xsolve_effective = h->xsolve_eff;
  out->p2p_ready = h->p2p.ready ? 1 : 0;
  out->nranks = h->nranks;
  out->rank = h->rank;
  ADMM_API_END
}

int admm_b200_slicemaker(int64_t len, int64_t workers, int64_t* out) {
  ADMM_API_BEGIN
  ADMM_REQUIRE(len >= 0 && workers > 0 && out, ADMM_B200_ERR_INVALID, "slicemaker: bad arguments");
  const int64_t rem = len % workers, base = len / workers;  // errorcheck.m:249-259
  for (int64_t w = 0; w < workers; ++w) out[w] = base + (w < rem ? 1 : 0);
  ADMM_API_END
}

}  // extern "C"
