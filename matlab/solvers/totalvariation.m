function results = totalvariation(s, lambda, options)
% TOTALVARIATION  Drop-in for solvers/totalvariation.m:62; D stays implicit on the device.  UNTESTED HERE.
t = tic;
if ~isstruct(options), error('Given options is not a struct! At least pass empty struct!'); end
if ~(isscalar(lambda) && isreal(lambda) && lambda >= 0), error('Given lambda parameter is not a nonnegative number!'); end
n = numel(s);
args = struct('h', b200_engine(options), 's', s(:), 'lambda', lambda);
[xmin, zmin] = getproxops('TotalVariation', args);
options.A = 1; options.At = 1; options.B = -1; options.nB = n; options.c = 0; options.m = n;
results = admm(xmin, zmin, options);
results.solverruntime = toc(t);
end
