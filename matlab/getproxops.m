function [minx, minz, extra] = getproxops(problem, args)
% GETPROXOPS  Drop-in for getProxOps.m:13 (callers spell it lowercase, e.g. lasso.m:192).
%   Returns handles that carry an engine descriptor instead of MATLAB closures over L/U/D; admm.m of
%   this directory recognises them.  args.h is the engine handle made by the solver wrapper.
%   UNTESTED HERE (no MATLAB / Octave in the image); tested twin: admm_project_b200/getproxops.py.
extra = struct();
if ~ischar(problem)
    error('Given problem argument is not a string specifying for which problem proximal operators are needed!');
end
if ~isstruct(args)
    error('Given struct args is not a struct containing arguments needed for proximal operators for the given problem!');
end
problem = lower(problem);
switch problem
    case 'lasso'
        admm_b200_mex('set_lambda', args.h, args.lambda);
    case 'basispursuit'
        admm_b200_mex('setup_basispursuit', args.h, args.D, args.s);
    case 'totalvariation'
        admm_b200_mex('setup_totalvariation', args.h, args.s, args.lambda);
    case 'linearsvm'
        kind = 4 + strcmp(args.lossfunction, '01');        % only the exact string '01' is the 0-1 loss (getProxOps.m:1094)
        admm_b200_mex('setup_unwrapped', args.h, kind, args.D, args.ell, args.C, b200_mtotal(args));
    case 'huberfit'
        admm_b200_mex('setup_unwrapped', args.h, 6, args.D, args.s, 0, b200_mtotal(args));
    case 'lad'
        admm_b200_mex('setup_unwrapped', args.h, 7, args.D, args.s, 0, b200_mtotal(args));
    case 'model'                                           % P, Q, r, s instead of PtP .. Qts (model.m:123-128): Grams on the device
        rho = 1; if isfield(args, 'rho'), rho = args.rho; end
        admm_b200_mex('setup_model', args.h, args.P, args.Q, args.r, args.s, rho);
    case 'quadraticprogram'                                % only the 'bounded' form (getProxOps.m:1441-1474) is on the device
        if ~isfield(args, 'constraint') || ~strcmp(args.constraint, 'bounded')
            error('admm_b200: quadraticprogram ''standard'' (dense KKT solve every iteration) is outside the engine''s hot path.');
        end
        r = 0; if isfield(args, 'r'), r = args.r; end
        admm_b200_mex('setup_quadratic', args.h, 9, args.P, args.q, r, args.rho, args.lb, args.ub);
    case {'linearprogram', 'covarianceselection'}
        error('admm_b200: problem ''%s'' is outside the engine''s hot path.', problem);
    otherwise
        error('Invalid input for problem - given string is not a solver!');
end
desc = struct('h', args.h, 'problem', problem);
minx = @(x, z, u, rho) b200_nohost(desc);
minz = @(x, z, u, rho) b200_nohost(desc);
end

function mt = b200_mtotal(args)      % global row count of a row-sharded A = D problem (solvers/b200_comm.m); default: all rows here
mt = size(args.D, 1); if isfield(args, 'm_total'), mt = args.m_total; end
end

function b200_nohost(desc) %#ok<INUSD>
error('admm_b200: device-resident proximal operator; it is evaluated inside admm() on the GPU.');
end
