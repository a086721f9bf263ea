function y = pos(x)
% CVX shim: max(x, 0).
y = max(x, 0);
end
