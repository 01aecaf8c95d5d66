"""Fit of the GELU evaluation used in the fc1 epilogue (k1_qv_lora_fwd_2cta.cu::gelu_erf).

erf-GELU(v) = v * Phi(v) = v / (1 + exp(-2 g(v))),  g(v) = atanh(erf(v / sqrt 2))  (exact identity)
g is fitted by v * (c0 + c1 v^2 + c2 v^4), minimising max |v*Phi_fit(v) - v*Phi(v)| over [-8, 8] (Nelder-Mead).
"""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

v = np.linspace(-8, 8, 160001)
gelu = v * 0.5 * (1 + erf(v / np.sqrt(2)))


def approx(c, v):
    vc = np.clip(v, -10, 10)
    v2 = vc * vc
    return v / (1 + np.exp(-2 * vc * (c[0] + c[1] * v2 + c[2] * v2 * v2)))


def err(c):
    return np.max(np.abs(approx(c, v) - gelu))


c = np.array([np.sqrt(2 / np.pi), 0.044715 * np.sqrt(2 / np.pi), 0.0])
for _ in range(3):
    c = minimize(err, c, method="Nelder-Mead", options=dict(xatol=1e-13, fatol=1e-15, maxiter=40000, maxfev=40000)).x
print("c =", c.tolist(), "max abs error =", err(c))
w = np.linspace(-40, 40, 80001)
print("max abs error on [-40, 40] =", np.max(np.abs(approx(c, w) - w * 0.5 * (1 + erf(w / np.sqrt(2))))))
