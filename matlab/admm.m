function results = admm(xminf, zming, options)
% ADMM  Drop-in for the reference's admm(xminf, zming, options) (admm.m:24) on libadmm_b200.
%   xminf / zming must be the handles returned by this directory's getproxops: they wrap an engine
%   descriptor, and the WHOLE loop (admm.m:496-767) runs on the GPU through admm_b200_mex('solve').
%   Arbitrary function handles raise an error: the engine has no CPU path.
%   UNTESTED IN THIS REPOSITORY'S IMAGE (no MATLAB / Octave there); the Python mirror
%   admm_project_b200/admm.py is the tested twin of this file.
if ~isstruct(options)
    error('Given options is not a struct! At least pass empty struct!');
end
dx = b200_descriptor(xminf); dz = b200_descriptor(zming);
if isempty(dx) || isempty(dz) || dx.h ~= dz.h
    error('admm_b200: xminf/zming must come from getproxops (device-resident operators; no CPU path).');
end
if getopt(options, 'adaptive', 0)
    error('admm_b200: options.adaptive is not built (unfinished experiment in the reference, admm.m:724-741).');
end
for f = {'altu', 'specialnorms'}          % admm.m:556-558, 612-616: host handles evaluated inside every iteration
    if isfield(options, f{1}) && isa(options.(f{1}), 'function_handle')
        error('admm_b200: options.%s is a host function handle evaluated in every iteration; the device loop cannot call back.', f{1});
    end
end
stopnames = {'standard', 'hnorm', 'both'};
o = struct();
o.rho = getopt(options, 'rho', 1.0);            o.relax = getopt(options, 'relax', 1);
N = getopt(options, 'maxiters', 1000);          if N <= 0, N = 1000; end
o.maxiters = ceil(N);                           o.domaxiters = getopt(options, 'domaxiters', 0);
o.objevals = double(getopt(options, 'objevals', 0) ~= 0);
o.convtest = getopt(options, 'convtest', 0);    o.convtol = getopt(options, 'convtol', 1e-10);
sc = getopt(options, 'stopcond', 'standard');   k = find(strcmp(sc, stopnames));
if isempty(k), k = 1; o.domaxiters = 1; end,    o.stopcond = k - 1;
o.nodualerror = getopt(options, 'nodualerror', 0);
o.abstol = getopt(options, 'abstol', 1e-5);     o.reltol = getopt(options, 'reltol', 1e-3);
if isfield(options, 'Hnormtol'), o.hnormtol = options.Hreltol; else, o.hnormtol = 1e-6; end  % admm.m:927-928 quirk
o.history = getopt(options, 'history', 1);
o.fast = double(getopt(options, 'fast', 0) ~= 0);                       % admm.m:59-60, 267-298
o.fasttype = double(strcmp(getopt(options, 'fasttype', 'weak'), 'weak'));
o.restart = getopt(options, 'restart', 0.999); if o.restart <= 0 || o.restart >= 1, o.restart = 0.999; end
o.dvaltol = getopt(options, 'dvaltol', 1e-8);
alg = o.fast*(1 + o.fasttype);
for f = {'A', 'B'}
    if ~isfield(options, f{1}), error('Must specify a matrix %s in constraint Ax + Bz = c!', f{1}); end
end
admm_b200_mex('set_init', dx.h, getopt(options, 'x0', []), getopt(options, 'z0', []), getopt(options, 'u0', []));
t = tic;
if isfield(options, 'preprocess') && isa(options.preprocess, 'function_handle'), options.preprocess(); end   % admm.m:473-476
r = admm_b200_mex('solve', dx.h, o);
k = r.steps;
if alg == 2     % accelerated ADMM records d, alpha and restarts instead of residual norms (admm.m:586-599)
    results = struct('dvals', r.dvals(1:k), 'restarted', r.restarted(1:k), 'dvaltol', o.dvaltol);
else
    results = struct('pnorm', r.pnorm(1:k), 'dnorm', r.dnorm(1:k), 'perr', r.perr(1:k), 'derr', r.derr(1:k));
end
if alg > 0, results.avals = r.avals(1:k); end
usesH = o.convtest || any(strcmp(sc, {'hnorm', 'both'}));
if usesH, results.Hnormsq = r.Hnormsq(1:k); results.Hnormtol = o.hnormtol; end
if o.objevals, results.objevals = r.objevals(1:k); end
if o.history
    results.xvals = r.xvals(:, 1:k); results.zvals = r.zvals(:, 1:k); results.uvals = r.uvals(:, 1:k);
    if usesH, results.wvals = [results.xvals; results.zvals; o.rho*results.uvals]; end
end
if r.status == 4      % admm.m:692-700: message and early return, steps/xopt/... stay unset
    disp('ADMM seems to not be converging! Please check that your proximal operators are correct!');
    return;
end
results.steps = k; results.xopt = r.xopt; results.zopt = r.zopt; results.uopt = r.uopt;
if o.objevals, results.objopt = r.objopt; end
results.runtime = toc(t);
results.options = options;
end

function v = getopt(s, name, dflt)
if isfield(s, name), v = s.(name); else, v = dflt; end
end

function d = b200_descriptor(f)
d = [];
if isa(f, 'function_handle')
    w = functions(f);
    if isfield(w, 'workspace') && ~isempty(w.workspace) && isfield(w.workspace{1}, 'desc')
        d = w.workspace{1}.desc;
    end
end
end
