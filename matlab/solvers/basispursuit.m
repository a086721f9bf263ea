function results = basispursuit(D, s, options)
% BASISPURSUIT  Drop-in for solvers/basispursuit.m:52; chol(D*D') replaces the dense projector.  UNTESTED HERE.
t = tic;
[mD, nD] = size(D); ms = numel(s);
if mD == nD && mD == ms, error('Square matrix problem Dx = s; don''t need Basis Pursuit to solve this!');
elseif mD > nD && mD == ms, error('Overdetermined system Dx = s, as D has more rows thancolumns; use Unwrapped ADMM solver for efficiency, instead.');
elseif mD ~= ms, error('The number of rows in matrix D must match the number of rows in signal vector s!'); end
if ~isstruct(options), error('Given options is not a struct! At least pass empty struct!'); end
args = struct('h', b200_engine(options), 'D', D, 's', s(:));
[minx, minz] = getproxops('BasisPursuit', args);
options.A = 1; options.B = -1; options.c = 0; options.m = nD; options.nA = nD; options.nB = nD;
results = admm(minx, minz, options);
results.solverruntime = toc(t);
end
