function results = linearsvm(D, ell, C, options)
% LINEARSVM  Drop-in for solvers/linearsvm.m:92.  UNTESTED HERE; twin: admm_project_b200/solvers/linearsvm.py.
t = tic;
if ~isstruct(options), error('Given options is not a struct! At least pass empty struct!'); end
if ~(isscalar(C) && isreal(C) && C >= 0), error('Given regularization parameter C is not a nonnegative number!'); end
if size(D, 1) ~= numel(ell), error('Product ell*D is not possible; sizes incompatible!'); end
loss = 'hinge'; if isfield(options, 'lossfunction'), loss = options.lossfunction; end
h = b200_engine(options); m = size(D, 1); ell = ell(:);
[~, world, lo, hi] = b200_comm(h, m);                  % one MATLAB per GPU: this rank keeps rows lo:hi (errorcheck.m:249-259)
if world > 1, D = D(lo:hi, :); ell = ell(lo:hi); options.b200_rows = [lo hi m]; end
args = struct('h', h, 'D', D, 'ell', ell, 'C', C, 'lossfunction', loss, 'm_total', m);
[~, minz] = getproxops('LinearSVM', args);
results = unwrappedadmm(minz, D, options);
results.solverruntime = toc(t);
end
