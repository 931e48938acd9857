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


def fit_poly(C=4.0, d=6):
    """`--poly`: the MUFU-free form (bert.cu gelu_poly2): z = clip(x / 2C + 1/2, 0, 1) - 1/2,
    Phi(x) - 1/2 = z Q(z^2), Q of degree d, minimax in |x| (Phi error); evaluated in fp32 like the kernel."""
    z = np.cos(np.pi * (np.arange(8001) + 0.5) / 8001) * 0.5
    xs = z * 2 * C
    u4 = z * z * 4.0
    A = np.stack([z * u4 ** k for k in range(d + 1)], 1)
    y = 0.5 * (1 + erf(xs / np.sqrt(2))) - 0.5
    w = np.ones_like(z)
    for _ in range(200):
        c, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None)
        e = np.abs(A @ c - y) * np.maximum(np.abs(xs), 0.5)
        w = w * (1 + 2 * e / e.max())
        w /= w.mean()
    c32 = (c * 4.0 ** np.arange(d + 1)).astype(np.float32)  # coefficients in u = z^2
    xx = np.linspace(-12, 12, 400001).astype(np.float32)
    zz = np.clip(xx * np.float32(1 / (2 * C)) + np.float32(0.5), 0, 1).astype(np.float32) - np.float32(0.5)
    uu = (zz * zz).astype(np.float32)
    q = np.full_like(uu, c32[d])
    for k in range(d - 1, -1, -1):
        q = (q * uu + c32[k]).astype(np.float32)
    got = (xx * (zz * q + np.float32(0.5)).astype(np.float32)).astype(np.float32)
    ref = xx.astype(np.float64) * 0.5 * (1 + erf(xx.astype(np.float64) / np.sqrt(2)))
    print("poly coefficients (u^0 ..)", [float(v) for v in c32], "max abs error (fp32 evaluation)", np.abs(got - ref).max())


if __name__ == "__main__":
    import sys

    if "--poly" in sys.argv:
        fit_poly()
