function results = unwrappedadmm(zming, D, options)
% UNWRAPPEDADMM  Drop-in for solvers/unwrappedadmm.m:1.  The x-update is always the transpose
% reduction W \ D'(z-u) with one cached Cholesky of W = D'D (unwrappedadmm.m:96-141).  UNTESTED HERE.
t = tic;
[m, n] = size(D);
rows = [1 m m]; if isfield(options, 'b200_rows'), rows = options.b200_rows; m = rows(3); end   % row-sharded: D holds rows(1):rows(2) of m
xminf = zming;                                   % same descriptor; admm() only needs the engine handle
options.A = 1; options.At = 1; options.B = -1; options.nB = m; options.c = 0; options.m = m;   % D lives on the device
options.x0 = rand(n, 1); options.z0 = rand(m, 1); options.u0 = rand(m, 1);                      % :87-89 (same draw on every rank)
options.z0 = options.z0(rows(1):rows(2)); options.u0 = options.u0(rows(1):rows(2));
options.maxiters = 1000; options.stopcond = 'both'; options.nodualerror = 1;                    % :90-92
options.parallel = 'none';
results = admm(xminf, zming, options);
results.solverruntime = toc(t);
end
