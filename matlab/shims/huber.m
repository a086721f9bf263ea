function y = huber(x, M)
% CVX shim: x.^2 for |x| <= M, 2*M*|x| - M^2 otherwise.
if nargin < 2, M = 1; end
a = abs(x);
y = a.^2;
k = a > M;
y(k) = 2*M*a(k) - M^2;
end
