/*
 * admm_b200_mex.c -- MATLAB MEX / GNU Octave (mkoctfile --mex) gateway onto libadmm_b200.so.
 * It owns no numerics: it marshals mxArray <-> the C-ABI of include/admm_b200.h.
 *
 *   h   = admm_b200_mex('create', device)
 *         admm_b200_mex('destroy', h)
 *         admm_b200_mex('setup_lasso', h, D, s, rho)               solvers/lasso.m:159-176
 *         admm_b200_mex('setup_unwrapped', h, kind, D, aux, C)     unwrappedadmm.m:96-123, huberfit.m:166, lad.m:134
 *         admm_b200_mex('setup_basispursuit', h, D, s)             basispursuit.m:116-120
 *         admm_b200_mex('setup_totalvariation', h, s, lambda)      totalvariation.m:127-131
 *         admm_b200_mex('setup_model', h, P, Q, r, s, rho)         model.m:119-146
 *         admm_b200_mex('setup_quadratic', h, kind, P, q, r, rho, lb, ub)   quadraticprogram.m:210-216 (kind 9 = box, 8 = nonneg)
 *         admm_b200_mex('set_lambda', h, lambda)                   getProxOps.m:455
 *         admm_b200_mex('set_init', h, x0, z0, u0)                 admm.m:252-254 ([] = zeros)
 *   id  = admm_b200_mex('unique_id')                               uint8(1,128), made by rank 0 (replaces the PCT pool, admm.m:343-408)
 *         admm_b200_mex('comm_init', h, rank, nranks, id)          one MATLAB process per GPU; see solvers/b200_comm.m
 *         admm_b200_mex('comm_destroy', h)
 *   sl  = admm_b200_mex('slicemaker', len, workers)                errorcheck.m:249-259 (balanced row blocks)
 *         admm_b200_mex('setup_lasso_sharded', h, Drows, srows, rho, m_total)
 *         admm_b200_mex('setup_unwrapped', h, kind, Drows, auxrows, C, m_total)   (m_total optional: row-sharded form)
 *   res = admm_b200_mex('solve', h, opts)                          admm.m:496-767
 *   res = admm_b200_mex('solve_lasso_batch', h, opts, lambdas)
 *
 * opts is a struct with the numeric fields of admm_b200_options (missing fields keep the defaults
 * of admm.m:51-76); res carries the fields of admm.m's results struct that the engine fills.
 * Errors become MATLAB errors with the text of admm_b200_last_error().
 *
 * Build:  mex -I../include admm_b200_mex.c -L../admm_project_b200 -ladmm_b200
 *    or:  mkoctfile --mex -I../include admm_b200_mex.c -L../admm_project_b200 -ladmm_b200
 * NOT run in this repository's image (no MATLAB / Octave); `make -C matlab check` only compiles it
 * against matlab/stub/mex.h.
 */
#include <math.h>
#include <string.h>

#include "admm_b200.h"
#include "mex.h"

static void check(int status) {
  if (status != ADMM_B200_OK) mexErrMsgIdAndTxt("admm_b200:engine", "%s", admm_b200_last_error());
}

static admm_b200_handle* get_handle(const mxArray* a) {
  if (mxGetNumberOfElements(a) != 1) mexErrMsgIdAndTxt("admm_b200:handle", "invalid engine handle");
  return (admm_b200_handle*)(uintptr_t)(*(uint64_t*)mxGetData(a));
}

static const double* dense(const mxArray* a, const char* what) {
  if (!mxIsDouble(a) || mxIsComplex(a) || mxIsSparse(a))
    mexErrMsgIdAndTxt("admm_b200:type", "%s must be a full real double array", what);
  return mxGetPr(a);
}

static double field(const mxArray* s, const char* name, double dflt) {
  const mxArray* f = mxIsStruct(s) ? mxGetField(s, 0, name) : NULL;
  return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

static void read_options(const mxArray* s, admm_b200_options* o) {
  admm_b200_default_options(o);
  o->rho = field(s, "rho", o->rho);
  o->relax = field(s, "relax", o->relax);
  o->abstol = field(s, "abstol", o->abstol);
  o->reltol = field(s, "reltol", o->reltol);
  o->convtol = field(s, "convtol", o->convtol);
  o->hnormtol = field(s, "hnormtol", o->hnormtol);
  o->maxiters = (int64_t)field(s, "maxiters", (double)o->maxiters);
  o->domaxiters = (int32_t)field(s, "domaxiters", 0);
  o->stopcond = (int32_t)field(s, "stopcond", 0);
  o->nodualerror = (int32_t)field(s, "nodualerror", 0);
  o->convtest = (int32_t)field(s, "convtest", 0);
  o->objevals = (int32_t)field(s, "objevals", 0);
  o->history = (int32_t)field(s, "history", 1);
  o->xsolve = (int32_t)field(s, "xsolve", 0);
  o->check_every = (int32_t)field(s, "check_every", 8);
  o->fast = (int32_t)field(s, "fast", 0);
  o->fasttype = (int32_t)field(s, "fasttype", 1);   /* 1 = 'weak' (accelerated, restart), 0 = fast ADMM */
  o->restart = field(s, "restart", 0.999);
  o->dvaltol = field(s, "dvaltol", 1e-8);
  o->graph = (int32_t)field(s, "graph", 1);
}

static mxArray* put(mxArray* s, const char* name, mwSize m, mwSize n) {
  mxArray* a = mxCreateDoubleMatrix(m, n, mxREAL);
  mxAddField(s, name);
  mxSetField(s, 0, name, a);
  return a;
}

static void cmd_solve(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  admm_b200_handle* h = get_handle(prhs[1]);
  admm_b200_options o;
  admm_b200_result r;
  int64_t nA, nB, m, N, k;
  mxArray* s = mxCreateStructMatrix(1, 1, 0, NULL);
  mxArray *px, *pz, *pu, *pn, *dn, *pe, *de, *hn, *ob, *dv, *av, *rs, *xv = NULL, *zv = NULL, *uv = NULL;
  (void)nlhs;
  if (nrhs < 3) mexErrMsgIdAndTxt("admm_b200:args", "solve needs (h, opts)");
  read_options(prhs[2], &o);
  check(admm_b200_get_dims(h, &nA, &nB, &m));
  N = o.maxiters > 0 ? o.maxiters : 1000;
  memset(&r, 0, sizeof(r));
  px = put(s, "xopt", nA, 1); pz = put(s, "zopt", nB, 1); pu = put(s, "uopt", m, 1);
  pn = put(s, "pnorm", 1, N); dn = put(s, "dnorm", 1, N); pe = put(s, "perr", 1, N); de = put(s, "derr", 1, N);
  hn = put(s, "Hnormsq", 1, N); ob = put(s, "objevals", 1, N);
  dv = put(s, "dvals", 1, N); av = put(s, "avals", 1, N); rs = put(s, "restarted", 1, N);
  r.xopt = mxGetPr(px); r.zopt = mxGetPr(pz); r.uopt = mxGetPr(pu);
  r.pnorm = mxGetPr(pn); r.dnorm = mxGetPr(dn); r.perr = mxGetPr(pe); r.derr = mxGetPr(de);
  r.hnormsq = mxGetPr(hn); r.objevals = mxGetPr(ob);
  r.dvals = mxGetPr(dv); r.avals = mxGetPr(av); r.restarted = mxGetPr(rs);
  if (o.history) {
    xv = put(s, "xvals", nA, N); zv = put(s, "zvals", nB, N); uv = put(s, "uvals", m, N);
    r.xvals = mxGetPr(xv); r.zvals = mxGetPr(zv); r.uvals = mxGetPr(uv);
  }
  check(admm_b200_solve(h, &o, &r));
  k = r.steps;   /* the .m wrapper trims the per-iteration arrays to 1:steps */
  mxAddField(s, "steps");     mxSetField(s, 0, "steps", mxCreateDoubleScalar((double)k));
  mxAddField(s, "status");    mxSetField(s, 0, "status", mxCreateDoubleScalar((double)r.status));
  mxAddField(s, "objopt");    mxSetField(s, 0, "objopt", mxCreateDoubleScalar(r.objopt));
  mxAddField(s, "setup_ms");  mxSetField(s, 0, "setup_ms", mxCreateDoubleScalar(r.setup_ms));
  mxAddField(s, "loop_ms");   mxSetField(s, 0, "loop_ms", mxCreateDoubleScalar(r.loop_ms));
  plhs[0] = s;
}

static void cmd_batch(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  admm_b200_handle* h = get_handle(prhs[1]);
  admm_b200_options o;
  int64_t nA, nB, m, nb, N, j;
  mxArray* s = mxCreateStructMatrix(1, 1, 0, NULL);
  mxArray *px, *pz, *pu, *pn, *dn, *pe, *de, *st, *ss;
  int64_t* steps;
  int32_t* status;
  double ms = 0.0;
  (void)nlhs;
  if (nrhs < 4) mexErrMsgIdAndTxt("admm_b200:args", "solve_lasso_batch needs (h, opts, lambdas)");
  read_options(prhs[2], &o);
  check(admm_b200_get_dims(h, &nA, &nB, &m));
  nb = (int64_t)mxGetNumberOfElements(prhs[3]);
  N = o.maxiters > 0 ? o.maxiters : 1000;
  px = put(s, "xopt", nA, nb); pz = put(s, "zopt", nB, nb); pu = put(s, "uopt", m, nb);
  pn = put(s, "pnorm", N, nb); dn = put(s, "dnorm", N, nb); pe = put(s, "perr", N, nb); de = put(s, "derr", N, nb);
  st = mxCreateNumericMatrix(nb, 1, mxUINT64_CLASS, mxREAL);
  ss = mxCreateNumericMatrix(nb, 1, mxUINT64_CLASS, mxREAL);
  steps = (int64_t*)mxGetData(st);
  status = (int32_t*)mxGetData(ss);
  check(admm_b200_solve_lasso_batch(h, &o, nb, dense(prhs[3], "lambdas"), steps, status, mxGetPr(px), mxGetPr(pz),
                                    mxGetPr(pu), mxGetPr(pn), mxGetPr(dn), mxGetPr(pe), mxGetPr(de), &ms));
  {
    mxArray* sd = put(s, "steps", nb, 1);
    for (j = 0; j < nb; ++j) mxGetPr(sd)[j] = (double)steps[j];
  }
  mxAddField(s, "loop_ms");
  mxSetField(s, 0, "loop_ms", mxCreateDoubleScalar(ms));
  plhs[0] = s;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  char* cmd;
  if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("admm_b200:args", "first argument must be a command string");
  cmd = mxArrayToString(prhs[0]);
  if (!strcmp(cmd, "create")) {
    admm_b200_handle* h = NULL;
    check(admm_b200_create(nrhs > 1 ? (int)mxGetScalar(prhs[1]) : 0, &h));
    plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)h;
    mexLock();
  } else if (!strcmp(cmd, "destroy")) {
    check(admm_b200_destroy(get_handle(prhs[1])));
  } else if (!strcmp(cmd, "setup_lasso")) {              /* (h, D, s, rho) */
    check(admm_b200_setup_lasso(get_handle(prhs[1]), (int64_t)mxGetM(prhs[2]), (int64_t)mxGetN(prhs[2]),
                                dense(prhs[2], "D"), (int64_t)mxGetM(prhs[2]), dense(prhs[3], "s"),
                                mxGetScalar(prhs[4]), ADMM_B200_XSOLVE_INVFACTOR));
  } else if (!strcmp(cmd, "setup_unwrapped")) {          /* (h, kind, D, aux, C [, m_total]) */
    int64_t m = (int64_t)mxGetM(prhs[3]);
    int64_t mt = nrhs > 6 ? (int64_t)mxGetScalar(prhs[6]) : m;      /* row-sharded: D holds this rank's rows of an mt-row matrix */
    check(admm_b200_setup_unwrapped(get_handle(prhs[1]), (int32_t)mxGetScalar(prhs[2]), m, mt, (int64_t)mxGetN(prhs[3]),
                                    dense(prhs[3], "D"), m, dense(prhs[4], "ell/s"), nrhs > 5 ? mxGetScalar(prhs[5]) : 0.0));
  } else if (!strcmp(cmd, "setup_lasso_sharded")) {      /* (h, Drows, srows, rho, m_total) */
    int64_t m = (int64_t)mxGetM(prhs[2]);
    check(admm_b200_setup_lasso_sharded(get_handle(prhs[1]), m, (int64_t)mxGetScalar(prhs[5]), (int64_t)mxGetN(prhs[2]),
                                        dense(prhs[2], "D"), m, dense(prhs[3], "s"), mxGetScalar(prhs[4]),
                                        ADMM_B200_XSOLVE_INVFACTOR));
  } else if (!strcmp(cmd, "unique_id")) {                /* rank 0: the 128-byte id every rank needs for comm_init */
    plhs[0] = mxCreateNumericMatrix(1, 128, mxUINT8_CLASS, mxREAL);
    check(admm_b200_get_unique_id(mxGetData(plhs[0])));
  } else if (!strcmp(cmd, "comm_init")) {                /* (h, rank, nranks, id) */
    if (nrhs < 5 || mxGetNumberOfElements(prhs[4]) != 128 || mxGetClassID(prhs[4]) != mxUINT8_CLASS)
      mexErrMsgIdAndTxt("admm_b200:args", "comm_init needs (h, rank, nranks, id) with id = uint8(1,128) from 'unique_id'");
    check(admm_b200_comm_init(get_handle(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), mxGetData(prhs[4])));
  } else if (!strcmp(cmd, "comm_destroy")) {
    check(admm_b200_comm_destroy(get_handle(prhs[1])));
  } else if (!strcmp(cmd, "slicemaker")) {               /* (len, workers) -> row counts, errorcheck.m:249-259 */
    int64_t w = (int64_t)mxGetScalar(prhs[2]), i;
    int64_t* tmp;
    if (w < 1 || w > 65536) mexErrMsgIdAndTxt("admm_b200:args", "slicemaker: workers out of range");
    tmp = (int64_t*)mxMalloc((size_t)w * sizeof(int64_t));
    check(admm_b200_slicemaker((int64_t)mxGetScalar(prhs[1]), w, tmp));
    plhs[0] = mxCreateDoubleMatrix(1, (mwSize)w, mxREAL);
    for (i = 0; i < w; ++i) mxGetPr(plhs[0])[i] = (double)tmp[i];
    mxFree(tmp);
  } else if (!strcmp(cmd, "setup_basispursuit")) {       /* (h, D, s) */
    check(admm_b200_setup_basispursuit(get_handle(prhs[1]), (int64_t)mxGetM(prhs[2]), (int64_t)mxGetN(prhs[2]),
                                       dense(prhs[2], "D"), (int64_t)mxGetM(prhs[2]), dense(prhs[3], "s")));
  } else if (!strcmp(cmd, "setup_totalvariation")) {     /* (h, s, lambda) */
    check(admm_b200_setup_totalvariation(get_handle(prhs[1]), (int64_t)mxGetNumberOfElements(prhs[2]),
                                         dense(prhs[2], "s"), mxGetScalar(prhs[3])));
  } else if (!strcmp(cmd, "setup_model")) {              /* (h, P, Q, r, s, rho) */
    int64_t m = (int64_t)mxGetM(prhs[2]);
    check(admm_b200_setup_model(get_handle(prhs[1]), m, (int64_t)mxGetN(prhs[2]), dense(prhs[2], "P"), m,
                                dense(prhs[3], "Q"), (int64_t)mxGetM(prhs[3]), dense(prhs[4], "r"), dense(prhs[5], "s"),
                                nrhs > 6 ? mxGetScalar(prhs[6]) : 1.0));
  } else if (!strcmp(cmd, "setup_quadratic")) {          /* (h, kind, P, q, r, rho, lb, ub) */
    int64_t n = (int64_t)mxGetM(prhs[3]);
    const double* lb = (nrhs > 7 && !mxIsEmpty(prhs[7])) ? dense(prhs[7], "lb") : NULL;
    const double* ub = (nrhs > 8 && !mxIsEmpty(prhs[8])) ? dense(prhs[8], "ub") : NULL;
    check(admm_b200_setup_quadratic(get_handle(prhs[1]), (int32_t)mxGetScalar(prhs[2]), n, dense(prhs[3], "P"), n,
                                    dense(prhs[4], "q"), mxGetScalar(prhs[5]), mxGetScalar(prhs[6]), lb, ub));
  } else if (!strcmp(cmd, "set_lambda")) {
    check(admm_b200_set_lambda(get_handle(prhs[1]), mxGetScalar(prhs[2])));
  } else if (!strcmp(cmd, "set_init")) {                 /* (h, x0, z0, u0), [] = zeros */
    const double* v[3] = {NULL, NULL, NULL};
    int i;
    for (i = 0; i < 3 && i + 2 < nrhs; ++i)
      if (!mxIsEmpty(prhs[i + 2])) v[i] = dense(prhs[i + 2], "x0/z0/u0");
    check(admm_b200_set_init(get_handle(prhs[1]), v[0], v[1], v[2]));
  } else if (!strcmp(cmd, "solve")) {
    cmd_solve(nlhs, plhs, nrhs, prhs);
  } else if (!strcmp(cmd, "solve_lasso_batch")) {
    cmd_batch(nlhs, plhs, nrhs, prhs);
  } else {
    mexErrMsgIdAndTxt("admm_b200:cmd", "unknown command '%s'", cmd);
  }
  mxFree(cmd);
}
