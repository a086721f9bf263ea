function results = huberfit(D, s, options)
% HUBERFIT  Drop-in for solvers/huberfit.m.  R = chol(D'*D,'lower') is built on the GPU.  UNTESTED HERE;
% tested twin: admm_project_b200/solvers/robustfit.py.
t = tic;
if ~isstruct(options), error('Given options is not a struct! At least pass empty struct!'); end
if size(D, 1) ~= numel(s), error('The number of rows in argument D do not match size of s!'); end
[m, n] = size(D);
h = b200_engine(options);
[~, world, lo, hi] = b200_comm(h, m);                  % one MATLAB per GPU: this rank keeps rows lo:hi (errorcheck.m:249-259)
s = s(:);
if world > 1, D = D(lo:hi, :); s = s(lo:hi); end
args = struct('h', h, 'D', D, 's', s, 'm_total', m);
[minx, minz] = getproxops('huberfit', args);
options.A = 1; options.B = -1; options.c = s; options.m = m; options.nA = n; options.nB = m;   % A = D lives on the device
results = admm(minx, minz, options);
results.solverruntime = toc(t);
end
