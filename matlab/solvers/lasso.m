function results = lasso(D, s, lambda, options)
% LASSO  Drop-in for solvers/lasso.m:77 (serial path).  Dts, chol(D'D + rho*I) or chol(DD'/rho + I)
% (lasso.m:159-176) are built on the GPU.  UNTESTED HERE; tested twin: admm_project_b200/solvers/lasso.py.
t = tic;
if ~isstruct(options), error('Given options is not a struct! At least pass empty struct!'); end
if ~(isscalar(lambda) && isreal(lambda) && lambda >= 0), error('Argument lambda is not a nonnegative real number!'); end
if size(D, 1) ~= numel(s), error('The number of rows in argument D do not match size of s!'); end
rho = 1.0; if isfield(options, 'rho'), rho = options.rho; end
if ~(rho > 0), error('Argument options.rho is not a positive real number!'); end
if isfield(options, 'parallel') && any(strcmp(options.parallel, {'both', 'zming', 'xminf'}))
    error('admm_b200: the parfor consensus LASSO branch (lasso.m:193-224) is out of scope.');
end
h = b200_engine(options);
[m, n] = size(D); s = s(:);
[~, world, lo, hi] = b200_comm(h, m);                  % one MATLAB per GPU: Gram of this rank's rows, ONE allreduce (lasso.m:160,168)
if world > 1 && m >= n
    admm_b200_mex('setup_lasso_sharded', h, D(lo:hi, :), s(lo:hi), rho, m);
else
    admm_b200_mex('setup_lasso', h, D, s, rho);
end
args = struct('h', h, 'lambda', lambda, 'm', size(D, 1), 'n', n, 'rho', rho, 'parallel', 0);
[minx, minz] = getproxops('LASSO', args);
options.A = 1; options.At = 1; options.B = -1; options.c = 0; options.m = n; options.nA = n; options.nB = n;
options.parallel = 'none';
results = admm(minx, minz, options);
results.solverruntime = toc(t);
end
