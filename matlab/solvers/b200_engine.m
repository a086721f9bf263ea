function h = b200_engine(options)
% One engine handle per MATLAB session (options.engine overrides).
persistent H
if isfield(options, 'engine'), h = options.engine; return; end
if isempty(H), dev = 0; if isfield(options, 'device'), dev = options.device; end, H = admm_b200_mex('create', dev); end
h = H;
end
