#!/usr/bin/env python
"""Coefficients of the erfc form used by the fc1 epilogue (csrc/vit_gemm.cu: gelu_erf):
erfc(a) = 2^(-a q(a)) on [0, 4], q a polynomial fitted by least squares weighted with d(erf)/dq, checked in emulated fp32."""
import numpy as np
from scipy.special import erf, erfc

a = np.linspace(1e-6, 4.0, 20001)
y = -np.log2(erfc(a)) / a
w = erfc(a) * np.log(2) * a
for deg in (4, 5, 6):
    co = np.polynomial.chebyshev.Chebyshev.fit(a, y, deg, w=w).convert(kind=np.polynomial.Polynomial).coef.astype(np.float32)
    af = a.astype(np.float32)
    q = np.zeros_like(af) + co[-1]
    for k in range(deg - 1, -1, -1):
        q = (q * af + co[k]).astype(np.float32)
    e = (np.float32(1) - np.exp2((-af * q).astype(np.float32)).astype(np.float32)).astype(np.float32)
    print(deg, "max |erf error|", np.abs(e.astype(np.float64) - erf(a)).max(), [float(c) for c in co])
