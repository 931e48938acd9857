"""Fit of the inner polynomial of the tanh-form GELU used by the FFN-up epilogue (bert.cu gelu_erf):
minimise max |x/2 (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) - x/2 (1 + erf(x/sqrt 2))| over [-8, 8]."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

x = np.linspace(-8, 8, 40001)
g = 0.5 * x * (1 + erf(x / np.sqrt(2)))


def err(c):
    x2 = x * x
    return np.abs(0.5 * x * (1 + np.tanh(x * (c[0] + x2 * (c[1] + x2 * c[2])))) - g).max()


c = np.array([np.sqrt(2 / np.pi), np.sqrt(2 / np.pi) * 0.044715, 0.0])
for _ in range(6):
    c = minimize(err, c, method="Nelder-Mead", options=dict(xatol=1e-12, fatol=1e-12, maxiter=40000, maxfev=40000)).x
print("coefficients", list(c), "max abs error", err(c))
