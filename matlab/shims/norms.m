function y = norms(x, p, dim)
% CVX shim: column (dim = 1) or row (dim = 2) p-norms.
if nargin < 2, p = 2; end
if nargin < 3, dim = 1; end
y = sum(abs(x).^p, dim).^(1/p);
end
