function y = subplus(x)
% Curve Fitting Toolbox shim: max(x, 0).
y = max(x, 0);
end
