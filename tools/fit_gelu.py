"""Fit of the branch-free erf used by gelu_fast() in csrc/common.cuh:
erf(|x|/sqrt 2) ~= 1 - 2^(-q(|x|)), q = a*(c1 + a*(c2 + ... )) on 0 <= a <= 5.6.
Prints the float32 coefficients and the resulting absolute GELU error."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf, erfc

DEG = 5
a = np.linspace(0, 5.6, 20001)
t = a / np.sqrt(2)
V = np.stack([a ** k for k in range(1, DEG + 1)], 1)
w = erfc(t) + 1e-7
c0 = np.linalg.lstsq(V * w[:, None], -np.log2(np.maximum(erfc(t), 1e-300)) * w, rcond=None)[0]
c = least_squares(lambda c: (1 - np.exp2(-(V @ c))) - erf(t), c0, xtol=1e-15, ftol=1e-15, gtol=1e-15).x
c32 = c.astype(np.float32)
q = np.zeros_like(a, dtype=np.float32)
for k in range(DEG - 1, -1, -1):
    q = q * a.astype(np.float32) + c32[k]
q = q * a.astype(np.float32)
approx = 1 - np.exp2(-q.astype(np.float64))
print("coefficients c1..c%d:" % DEG, [float(x) for x in c32])
print("max |erf err|  :", np.abs(approx - erf(t)).max())
print("max |gelu err| :", np.abs(0.5 * a * (approx - erf(t))).max())
