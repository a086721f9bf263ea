function results = quadraticprogram(P, q, r, cons1, cons2, options)
% QUADRATICPROGRAM  Drop-in for the 'bounded' branch of solvers/quadraticprogram.m:99-257 (lb <= x <= ub):
% cached chol(P + rho*I) x-update + box projection on the device.  UNTESTED HERE (no MATLAB / Octave in the image).
t = tic;
if ~isstruct(options), error('Given options is not a struct! At least pass empty struct!'); end
nP = size(P, 1); q = q(:);
if numel(q) ~= nP, error('The dimensions of square matrix P and vector q do not match!'); end
if ~(isvector(cons1) && isvector(cons2))
    error('admm_b200: quadraticprogram ''standard'' (dense KKT solve every iteration) is outside the engine''s hot path.');
end
c1 = cons1(:); c2 = cons2(:);
if numel(c1) ~= numel(c2), error('Lengths of lower and upper bound constraints on solution x do not match!'); end
if numel(c1) ~= nP, error('Bound vectors do not match predicted length of solution x!'); end
if isequal(max(c1, c2), c1), tmp = c1; c1 = c2; c2 = tmp;               % quadraticprogram.m:312-319
elseif ~isequal(max(c1, c2), c2), error('Given constraint variables do not specify an upper and lower bound on solution x!'); end
rho = 1; if isfield(options, 'rho'), rho = options.rho; end
args = struct('h', b200_engine(options), 'P', P, 'q', q, 'r', r, 'lb', c1, 'ub', c2, 'rho', rho, 'n', nP, ...
              'constraint', 'bounded');
[minx, minz] = getproxops('QuadraticProgram', args);
options.A = 1; options.B = -1; options.c = 0; options.m = nP; options.nA = nP; options.nB = nP;
options.obj = 'engine';                                    % 1/2*x''*P*x + q''*x + r, quadraticprogram.m:237
results = admm(minx, minz, options);
results.solverruntime = toc(t);
end
