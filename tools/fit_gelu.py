#!/usr/bin/env python
"""Fit of the polynomial in csrc/tc_common.cuh::gelu_fast: erf(t/sqrt2) = 1 - 2^P(t), P(0) = 0, degree 7, t in [0, 5.7].
Weighted (Lawson-iterated) least squares on the GELU error 0.5*t*erfc(t/sqrt2)*ln2*dP; prints coefficients and errors."""
import numpy as np
from scipy.special import erf, erfc
T, deg = 5.7, 7
t = np.cos(np.linspace(0, np.pi, 4001)) * 0.5 * T + 0.5 * T
t = t[t > 0]
f = np.log2(erfc(t / np.sqrt(2.0)))
w = 0.5 * t * erfc(t / np.sqrt(2)) * np.log(2) + 1e-12
A = np.vstack([t ** k for k in range(1, deg + 1)]).T
lw = np.ones_like(t)
for _ in range(60):
    W = w * lw
    coef, *_ = np.linalg.lstsq(A * W[:, None], f * W, rcond=None)
    err = np.abs((A @ coef - f) * w)
    lw = lw * (err / err.mean() + 1e-3)
    lw /= lw.mean()
c32 = coef.astype(np.float32)
x = np.linspace(-8, 8, 400001).astype(np.float32)
tt = np.minimum(np.abs(x), np.float32(T))
p = np.zeros_like(tt)
for k in range(deg - 1, -1, -1):
    p = (p * tt + c32[k]).astype(np.float32)
p = (p * tt).astype(np.float32)
e = np.exp2(p.astype(np.float64)).astype(np.float32)
h = np.float32(0.5) * x
g = (h * np.copysign(np.float32(1) - e, x) + h).astype(np.float32)
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("coefficients t^1..t^7:", ", ".join("%.9ef" % c for c in c32))
print("max |gelu_fast - gelu| on [-8,8] (fp32 evaluation): %.3e" % np.abs(g - ref).max())
