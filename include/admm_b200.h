/*
 * admm_b200.h -- C-ABI of libadmm_b200.so, a Blackwell (sm_100a) FP64 engine for the hot path of
 * PeterSutor/ADMM-Project: the scaled-dual ADMM loop of admm.m, the getProxOps.m proximal
 * operators and the one-time setup of solvers/<name>.m.
 *
 * This is the drop-in boundary.  Everything is plain C: column-major
 * `double*`, `int64_t` sizes, `int` status codes.  No torch / C++ types cross it.
 *
 * Conventions
 *   - Every matrix is column-major FP64 (MATLAB layout) with an explicit leading dimension.
 *   - A pointer argument may be a HOST pointer or a DEVICE pointer of the handle's GPU; the
 *     library asks the CUDA runtime which (cudaPointerGetAttributes).  Host data is staged to
 *     the device by the library, device data is used in place (D) or copied device-to-device.
 *   - The caller owns every buffer it passes; the library owns every device buffer it creates
 *     and frees them in admm_b200_destroy.
 *   - A handle is not thread-safe; use one host thread per handle.  One handle drives one GPU
 *     (one process per GPU for the row-sharded solvers, admm_b200_comm_init).
 *   - Functions return ADMM_B200_OK (0) or a negative-free error code; the text of the last
 *     error of the calling thread is returned by admm_b200_last_error().  The reference reports
 *     errors with MATLAB error(...) strings (no codes); the host-side mirror
 *     turns a non-zero status into an exception carrying that text.
 *   - There is NO CPU fallback: without a usable CUDA device admm_b200_create fails.
 *
 * Each entry point cites the reference interface it replaces (file:line in PeterSutor/ADMM-Project).
 */
#ifndef ADMM_B200_H
#define ADMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADMM_B200_VERSION 200 /* 0.2.0 */

/* ---- status codes ------------------------------------------------------------------------- */
enum {
  ADMM_B200_OK = 0,
  ADMM_B200_ERR_INVALID = 1,     /* bad argument (the reference's errorcheck.m / error(...) cases) */
  ADMM_B200_ERR_CUDA = 2,        /* CUDA runtime failure, no device, out of memory */
  ADMM_B200_ERR_STATE = 3,       /* call order (solve before setup, ...) */
  ADMM_B200_ERR_NOTPOSDEF = 4,   /* chol(...) failed: matrix is not positive definite */
  ADMM_B200_ERR_COMM = 5,        /* NCCL failure / NCCL not loadable */
  ADMM_B200_ERR_UNSUPPORTED = 6  /* a combination the engine does not implement (fails loudly) */
};

/* ---- problem kinds: the in-scope entries of the getProxOps.m switch (getProxOps.m:52-917) - */
enum {
  ADMM_B200_LASSO = 1,          /* getProxOps.m:311-456, xminLASSO :1192-1206, soft threshold :933-938 */
  ADMM_B200_BASISPURSUIT = 2,   /* getProxOps.m:126-142, xminBasisPursuit :1027-1032 */
  ADMM_B200_TOTALVARIATION = 3, /* getProxOps.m:172-199, xminTotalVariation :1044-1048 */
  ADMM_B200_SVM_HINGE = 4,      /* getProxOps.m:256-309, zminLinearSVM :1084-1103 (loss ~= '01') */
  ADMM_B200_SVM_01 = 5,         /* same, loss == '01' -> minz01 :1158-1180 */
  ADMM_B200_HUBERFIT = 6,       /* getProxOps.m:890-912, zminHuberSoftThresholding :1529-1539 */
  ADMM_B200_LAD = 7,            /* getProxOps.m:780-810, xminLAD :1511-1515 */
  ADMM_B200_PROX_NONNEG = 8,    /* z-prox of linearprogram / quadraticprogram: pos(x+u) :1378-1382 */
  ADMM_B200_PROX_BOX = 9,       /* z-prox of bounded QP: min(ub,max(lb,x+u)) :1470-1474 */
  ADMM_B200_MODEL = 10          /* getProxOps.m:55-110, xminModel :952-974, zminModel :989-1012 */
};

/* ---- stop conditions (admm.m:706-722) ------------------------------------------------------ */
enum { ADMM_B200_STOP_STANDARD = 0, ADMM_B200_STOP_HNORM = 1, ADMM_B200_STOP_BOTH = 2 };

/* ---- why the loop ended --------------------------------------------------------------------- */
enum {
  ADMM_B200_RUNNING = 0,
  ADMM_B200_CONVERGED_STD = 1,   /* admm.m:710-713 */
  ADMM_B200_CONVERGED_HNORM = 2, /* admm.m:719-722 */
  ADMM_B200_MAXITERS = 3,        /* loop ran out, admm.m:496 */
  ADMM_B200_DIVERGED_RETURN = 4, /* H-norm monotonicity test fired, the reference `return`s, admm.m:692-700 */
  ADMM_B200_CONVERGED_DVAL = 5   /* accelerated ADMM: abs(d - dprev) <= dvaltol*dprev, admm.m:706-707 */
};

/* ---- x-update realisation ------------------------------------------------------------------- */
enum {
  ADMM_B200_XSOLVE_INVFACTOR = 0, /* L\ and L'\ applied as products with the cached inverse factor
                                     (block-inverse triangular solve with one block); no dependency
                                     chain, two streaming passes over one triangle each */
  ADMM_B200_XSOLVE_SUBST = 1      /* blocked forward/back substitution on the cached factor L */
};

typedef struct admm_b200_handle admm_b200_handle;

/* The options struct of admm.m (defaults admm.m:51-76; setopt admm.m:780-971), as plain data.
 * Fill with admm_b200_default_options() first. */
typedef struct admm_b200_options {
  double rho;         /* admm.m:53  default 1.0 */
  double relax;       /* admm.m:56  default 1.0 (over-relaxation alpha) */
  double abstol;      /* admm.m:67  default 1e-5 */
  double reltol;      /* admm.m:68  default 1e-3 */
  double convtol;     /* admm.m:64  default 1e-10 */
  double hnormtol;    /* admm.m:69  default 1e-6 */
  int64_t maxiters;   /* admm.m:54  default 1000; <=0 becomes 1000 (admm.m:334-339) */
  int32_t domaxiters; /* admm.m:55 */
  int32_t stopcond;   /* admm.m:65  ADMM_B200_STOP_* */
  int32_t nodualerror;/* admm.m:66 */
  int32_t convtest;   /* admm.m:63 */
  int32_t objevals;   /* admm.m:62  evaluate the solver's objective every iteration */
  int32_t history;    /* extension: 1 = record xvals/zvals/uvals (admm.m:608-610), 0 = do not */
  int32_t xsolve;     /* ADMM_B200_XSOLVE_* */
  int32_t check_every;/* iterations enqueued between host polls of the device stop flag (>=1) */
  int32_t fast;       /* admm.m:59  fast / accelerated ADMM of Goldstein et al. (admm.m:267-298, 562-600) */
  int32_t fasttype;   /* admm.m:60  1 = 'weak' (default): accelerated ADMM with restart (alg 2); 0 = fast ADMM (alg 1) */
  double restart;     /* admm.m:287 options.restart, default 0.999 (values outside (0,1) become 0.999) */
  double dvaltol;     /* admm.m:291 options.dvaltol, default 1e-8 */
  int32_t graph;      /* extension, default 1: after the first two bursts of check_every iterations, replay each
                         further burst as one CUDA graph launch (same kernels, same arguments, same results);
                         single-rank handles only (row-sharded loops stay eager) */
  int32_t reserved;
} admm_b200_options;

/* What admm.m returns in `results` (admm.m:603-610, 618-658, 682, 746-767).  Every pointer is
 * caller-owned and may be NULL (then that output is skipped).  Per-iteration arrays must hold
 * `maxiters` doubles; xvals/zvals/uvals hold n x maxiters column-major. */
typedef struct admm_b200_result {
  int64_t steps;      /* results.steps */
  int32_t status;     /* ADMM_B200_CONVERGED_* / MAXITERS / DIVERGED_RETURN */
  int32_t reserved;
  double objopt;      /* results.objopt (NaN unless objevals) */
  double setup_ms;    /* device time of the one-time setup (Gram, Cholesky, inverse factor) */
  double loop_ms;     /* device time of the iteration loop */
  double *xopt, *zopt, *uopt;                   /* lengths nA, nB, m */
  double *pnorm, *dnorm, *perr, *derr;          /* results.pnorm/dnorm/perr/derr */
  double *hnormsq;                              /* results.Hnormsq */
  double *objevals;                             /* results.objevals */
  double *xvals, *zvals, *uvals;                /* results.xvals/zvals/uvals */
  double *dvals, *avals, *restarted;            /* fast modes: results.dvals / avals / restarted (admm.m:586-599) */
} admm_b200_result;

/* ---- library ------------------------------------------------------------------------------- */
int admm_b200_version(void);
const char* admm_b200_last_error(void);
void admm_b200_default_options(admm_b200_options* o);          /* admm.m:51-76 */

/* ---- handle -------------------------------------------------------------------------------- */
int admm_b200_create(int device, admm_b200_handle** out);
int admm_b200_destroy(admm_b200_handle* h);
/* Run the handle's work on a caller-supplied cudaStream_t (e.g. torch's current stream) so the
 * caller's CUDA events bracket it.  NULL restores the handle's own stream. */
int admm_b200_set_stream(admm_b200_handle* h, void* cuda_stream);
int admm_b200_synchronize(admm_b200_handle* h);

/* ---- one-time setup: solvers/<name>.m ------------------------------------------------------------ */
/* solvers/lasso.m:159-192: Dts = D'*s; L = chol(D'D + rho I,'lower') (m >= n) or
 * chol(DD'/rho + I,'lower') (m < n).  D is m x n, leading dimension ldD. */
int admm_b200_setup_lasso(admm_b200_handle* h, int64_t m, int64_t n, const double* D, int64_t ldD,
                          const double* s, double rho, int32_t xsolve);
/* The same setup from ROW SHARDS, one process per GPU (admm_b200_comm_init first): D / s hold THIS rank's
 * m_local rows (errorcheck.m:249-259 partition), m_total is the global row count (m_total >= n).  Every rank
 * forms D_g'D_g and D_g's_g, ONE allreduce sums them -- the transpose reduction of unwrappedadmm.m:114-122
 * applied to lasso.m:160,168 -- and every rank factors the same n x n matrix; the iterations then run
 * replicated (x, z, u are n-vectors), the objective 1/2*||D x - s||^2 is summed over the shards. */
int admm_b200_setup_lasso_sharded(admm_b200_handle* h, int64_t m_local, int64_t m_total, int64_t n, const double* D,
                                  int64_t ldD, const double* s, double rho, int32_t xsolve);
/* The A = D problems (constraint D*x - z = c): linear SVM by unwrapped ADMM / transpose reduction
 * (solvers/unwrappedadmm.m:43-141, linearsvm.m:154-243; c = 0, aux = labels ell, C = regularisation),
 * Huber fitting and least absolute deviations (huberfit.m:155-183, lad.m:123-151; c = aux = s).
 * kind is ADMM_B200_SVM_HINGE / SVM_01 / HUBERFIT / LAD.  D holds THIS rank's m_local rows
 * (errorcheck.m:249-259 partition, admm_b200_slicemaker); m_total is the global row count.
 * W = sum over ranks of D_g'*D_g (unwrappedadmm.m:114-122) is allreduced when a communicator is
 * attached, then R = chol(W,'lower') is cached on every rank (huberfit.m:166, lad.m:134).
 * All-zero columns of D (constant-zero MNIST pixels, examples/mnistsvm.m) get x_j = 0, which is what the
 * reference's serial x-update pinv(D)*(z-u) (unwrappedadmm.m:76-78, linearsvm.m:185) returns; any other rank
 * deficiency fails with ADMM_B200_ERR_NOTPOSDEF and a text that says so. */
int admm_b200_setup_unwrapped(admm_b200_handle* h, int32_t kind, int64_t m_local, int64_t m_total, int64_t n,
                              const double* D, int64_t ldD, const double* aux, double C);

/* solvers/basispursuit.m:116-120 (D is m x n with m < n): caches L = chol(D*D','lower') instead of
 * the reference's dense n x n projector P and applies x = P(z-u) + q (getProxOps.m:1031) as
 * v - D'((DD') \ (D v - s)). */
int admm_b200_setup_basispursuit(admm_b200_handle* h, int64_t m, int64_t n, const double* D, int64_t ldD,
                                 const double* s);

/* solvers/totalvariation.m:122-164: 1-D total variation denoising of s (length n) with the implicit
 * difference operator D = spdiags([1 -1],0:1,n,n).  The x-update (getProxOps.m:1047) is a constant
 * tridiagonal SPD solve run as two block-parallel scans; the sparse matrix is never formed. */
int admm_b200_setup_totalvariation(admm_b200_handle* h, int64_t n, const double* s, double lambda);

/* Quadratic objective 1/2 x'Px + q'x + r (P symmetric positive semidefinite, n x n) with a projection
 * as z-prox -- the stand-alone prox catalogue entries of the reference on a cached-Cholesky x-update:
 *   kind = ADMM_B200_PROX_BOX    quadraticprogram.m 'bounded' (:210-216), getProxOps.m:1441-1474:
 *                                R = chol(P + rho*I), x = R \ (R' \ (rho*(z-u) - q)), z = min(ub,max(lb,x+u))
 *   kind = ADMM_B200_PROX_NONNEG same x-update with z = pos(x+u) (getProxOps.m:1378-1382, 1422-1426);
 *                                lb / ub may be NULL. */
int admm_b200_setup_quadratic(admm_b200_handle* h, int32_t kind, int64_t n, const double* P, int64_t ldP, const double* q,
                              double r, double rho, const double* lb, const double* ub);

/* solvers/model.m:119-146 -- minimise 1/2||P x - r||^2 + 1/2||Q x - s||^2 (P, Q are m x n) as f(x) + g(z),
 * x - z = 0.  PtP = P'P, QtQ = Q'Q, Ptr = P'r, Qts = Q's are formed on the device (model.m:123-128);
 * the reference's prox operators re-add rho to the diagonals and re-solve the dense systems with `\`
 * in EVERY iteration (getProxOps.m:967-973, 1004-1011: `rhoprev` is never updated).  The engine keeps
 * both Gram matrices and caches chol(PtP + rho I), chol(QtQ + rho I); a later solve with a different
 * options.rho only re-adds the diagonal and refactors (examples/stepsizetesting.m sweeps rho). */
int admm_b200_setup_model(admm_b200_handle* h, int64_t m, int64_t n, const double* P, int64_t ldP, const double* Q,
                          int64_t ldQ, const double* r, const double* s, double rho);

/* Row-sharded runs, one process per GPU.  Rank 0 calls admm_b200_get_unique_id (128 bytes, an
 * ncclUniqueId), the host side broadcasts it (torch.distributed / MPI / a file), every rank calls
 * admm_b200_comm_init before the setup.  Replaces the PCT worker pool of the reference
 * (gcp / parfor, admm.m:343-408, unwrappedadmm.m:45-74).  nranks == 1 detaches. */
int admm_b200_get_unique_id(void* out128);
int admm_b200_comm_init(admm_b200_handle* h, int rank, int nranks, const void* unique_id128);
int admm_b200_comm_destroy(admm_b200_handle* h);
/* The same worker pool WITHOUT NCCL (mailbox-only transport): every rank exports the CUDA IPC handle of its
 * mailbox (64 bytes), the host side all-gathers the handles in rank order (nranks x 64 bytes) and every rank
 * attaches them.  All sums over ranks -- the Gram in pieces of 32768 doubles too -- then go through the mailboxes.
 * It is what lets several ranks share ONE device (NCCL refuses two ranks on a GPU): the reference's parfor runs
 * its slices on however many workers the pool has (admm.m:343-408), and the 2-rank parity tests run on a 1-GPU
 * box this way (the ranks' kernels are time-sliced by the driver; the waits are bounded, p2p.cuh).  2..8 ranks. */
int admm_b200_comm_ipc_export(admm_b200_handle* h, int rank, int nranks, void* out_handle64);
int admm_b200_comm_ipc_attach(admm_b200_handle* h, const void* handles_nranks_x_64);
/* comm_init also maps every peer's MAILBOX (CUDA IPC over NVLink): messages of up to 32768 doubles -- the one
 * exchange per iteration of the row-sharded loops -- are summed by a one-shot peer-memory allreduce inside the
 * engine's own kernels (p2p.cuh) instead of a library collective; larger ones (the n x n Gram) use ncclAllReduce.
 * In-place sum over ranks of `count` doubles (host or device pointer); no-op for a single rank. */
int admm_b200_allreduce(admm_b200_handle* h, double* buf, int64_t count);

/* getProxOps.m:455 -- lambda of the soft threshold; may be changed between solves without
 * redoing the setup (the factor does not depend on lambda). */
int admm_b200_set_lambda(admm_b200_handle* h, double lambda);
/* options.x0 / z0 / u0 (admm.m:252-254); NULL means zeros. */
int admm_b200_set_init(admm_b200_handle* h, const double* x0, const double* z0, const double* u0);

/* ---- the loop: admm.m:496-767 ---------------------------------------------------------------- */
int admm_b200_solve(admm_b200_handle* h, const admm_b200_options* opts, admm_b200_result* res);

/* Regularisation path on one cached factor (BASELINE.json configs[1]: "a batch of 64 lambda values
 * as multi-RHS TRSM"): nb independent lasso problems sharing D, s, rho, started from zeros.  Column
 * j runs admm.m:496-767 with lambda[j] and its own stop test; a stopped column is frozen.  The
 * x-update of all columns is two triangular FP64 DMMA GEMMs on the cached inverse factor.
 * steps/status: nb entries; xopt/zopt/uopt: n x nb column-major; pnorm..derr: maxiters x nb
 * (column j at offset j*maxiters) or NULL.  Tall lasso (m >= n) only. */
int admm_b200_solve_lasso_batch(admm_b200_handle* h, const admm_b200_options* opts, int64_t nb, const double* lambdas,
                                int64_t* steps, int32_t* status, double* xopt, double* zopt, double* uopt,
                                double* pnorm, double* dnorm, double* perr, double* derr, double* loop_ms);

/* nb (<= 16) A = D problems that share the D of the last admm_b200_setup_unwrapped -- the one-vs-all
 * classifiers of examples/mnistsvm.m:121-156 (BASELINE.json configs[2]) -- advanced together: D is
 * swept twice per iteration for ALL classes.  aux: m_local x nb labels / targets (this rank's rows),
 * X0 / Z0 / U0: n x nb, m_local x nb, m_local x nb initial iterates (NULL = zeros; unwrappedadmm.m:87-89
 * draws them with rand).  Options as forced by unwrappedadmm.m:90-92 (nodualerror = 1, relax = 1).
 * Each class runs admm.m's loop with its own stop test; a stopped class is frozen.  Row-sharded runs
 * exchange ONE allreduce of nb x [D'r ; scalars] per iteration.  pnorm / perr / objevals: maxiters x nb. */
int admm_b200_solve_unwrapped_batch(admm_b200_handle* h, const admm_b200_options* opts, int64_t nb, const double* aux,
                                    int64_t ldaux, const double* X0, const double* Z0, const double* U0, int64_t* steps,
                                    int32_t* status, double* xopt, double* zopt, double* uopt, double* pnorm, double* perr,
                                    double* objevals, double* loop_ms);

/* Sizes of the current problem: nA (x), nB (z), m (u) -- admm.m:74-76. */
int admm_b200_get_dims(admm_b200_handle* h, int64_t* nA, int64_t* nB, int64_t* m);
/* Lower Cholesky factor of the current setup (k x k, column-major, ldL >= k) for setup parity
 * tests (lasso.m:168,172; huberfit.m:166; lad.m:134). */
int admm_b200_get_factor(admm_b200_handle* h, double* L, int64_t ldL, int64_t* k);

/* ---- building blocks (each is one hand-written kernel family; exposed for parity tests and
 *      per-kernel roofline measurement) -------------------------------------------------------- */
/* C = alpha*op(A)*op(B) + beta*C on FP64 DMMA tiles.  transa/transb: 0 = 'N', 1 = 'T'.
 * lower_only != 0 computes only the tiles on/below the diagonal (SYRK use: D'*D lasso.m:168). */
int admm_b200_dgemm(admm_b200_handle* h, int transa, int transb, int64_t M, int64_t N, int64_t K,
                    double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                    double beta, double* C, int64_t ldc, int lower_only);
/* G = D'*D + shift*I (trans = 1, k = n) or G = scale*D*D' + shift*I (trans = 0, k = m); full
 * symmetric k x k result.  lasso.m:168,172; unwrappedadmm.m:115. */
int admm_b200_gram(admm_b200_handle* h, int trans, int64_t m, int64_t n, const double* D, int64_t ldD,
                   double scale, double shift, double* G, int64_t ldG);
/* In-place lower Cholesky of the k x k matrix A (chol(A,'lower'), lasso.m:168); the strict
 * upper triangle is zeroed.  Winv (may be NULL) receives inv(L), k x k. */
int admm_b200_potrf(admm_b200_handle* h, int64_t k, double* A, int64_t lda, double* Winv, int64_t ldw);
/* x = L' \ (L \ b) with the factor cached by the last setup (getProxOps.m:1200); b, x length k. */
int admm_b200_factor_solve(admm_b200_handle* h, const double* b, double* x, int32_t xsolve);
/* Run `reps` x-updates + fused z/u/residual passes of the current problem back to back on the
 * handle's stream without the stop test (kernel timing; `which`: 0 = whole iteration,
 * 1 = x-update only, 2 = fused prox pass only). */
int admm_b200_iterate_raw(admm_b200_handle* h, const admm_b200_options* opts, int which, int reps);
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t admm_b200_launch_count(admm_b200_handle* h);
/* Number of CUDA graph launches the iteration loops of this handle have issued (options.graph). */
int64_t admm_b200_graph_replays(admm_b200_handle* h);

/* Device time (ms) of the last setup by phase: [0] Gram (+ D's), [1] Cholesky, [2] inverse factor,
 * [3] total.  (The reference only has tic/toc: results.solverruntime, lasso.m:117,243.) */
int admm_b200_get_setup_phases(admm_b200_handle* h, double* out4);

/* State of the last setup and of the communicator (host mirror, tests). */
typedef struct admm_b200_info {
  int64_t generation;       /* bumped by every successful setup: a (minx, minz) pair made for an older setup is stale */
  int64_t zero_cols;        /* A = D problems: all-zero columns of D handled as pinv does (x_j = 0) */
  double diag_ratio;        /* (max L_ii / min L_ii)^2 of the cached factor, a lower bound of cond_2 */
  int32_t xsolve_effective; /* ADMM_B200_XSOLVE_*: SUBST when diag_ratio exceeded the guard (1e9, env ADMM_B200_COND_GUARD) */
  int32_t p2p_ready;        /* 1: small allreduces run over the peer mailboxes, 0: over NCCL */
  int32_t nranks, rank;
} admm_b200_info;
int admm_b200_get_info(admm_b200_handle* h, admm_b200_info* out);

/* errorcheck.m:216-267 slicemaker, balanced case (slices == 0, :249-259): out[w] = rows of
 * worker w; this is the row partition of the multi-GPU solvers. */
int admm_b200_slicemaker(int64_t len, int64_t workers, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* ADMM_B200_H */
