function results = model(P, Q, r, s, options)
% MODEL  Drop-in for solvers/model.m:47; both Gram matrices and Cholesky factors live on the device.  UNTESTED HERE.
t = tic;
if ~ismatrix(P), error('Argument P is not a matrix!'); end
if ~ismatrix(Q), error('Argument Q is not a matrix!'); end
if ~isvector(r), error('Argument r is not a vector!'); end
if ~isvector(s), error('Argument s is not a vector!'); end
[mP, nP] = size(P); [mQ, nQ] = size(Q); r = r(:); s = s(:);
if mP ~= mQ, error('Number of rows in P do not match number of rows in Q!');
elseif nP ~= nQ, error('Number of columns in P do not match number of columns in Q!');
elseif mP ~= numel(r), error('Number of rows in P does not match length of vector r!');
elseif mQ ~= numel(s), error('Number of rows in Q does not match length of vector s!'); end
if ~isstruct(options), error('Given options argument is not a struct! Please check your arguments and try again.'); end
rho = 1; if isfield(options, 'rho'), rho = options.rho; end
args = struct('h', b200_engine(options), 'P', P, 'Q', Q, 'r', r, 's', s, 'n', nP, 'rho', rho);
[minx, minz] = getproxops('Model', args);
options.A = 1; options.B = -1; options.c = 0; options.m = nP; options.nA = nP; options.nB = nP;
options.obj = 'engine';                                    % 1/2*norm(P*x-r)^2 + 1/2*norm(Q*z-s)^2, model.m:139-140
results = admm(minx, minz, options);
results.solverruntime = toc(t);
end
