#!/usr/bin/env python
"""Degree-3 polynomial for 2^f on [0, 1): the fit behind the experiment of taking every second exponential of the attention
kernel off the MUFU pipe (floor by a round-down add of 1.5 x 2^23, 2^frac by this polynomial, the integer part added into the
exponent field).  Measured neutral on B200 -- the phase is not MUFU-bound -- and not in the kernel
(profiles/r2_vit_attn_steps.txt, step 7).  Minimax in relative error by Remez-style iteratively re-weighted least squares, then the error of the
float32 Horner evaluation the kernel performs."""
import numpy as np

f = np.linspace(0.0, 1.0, 20001)
y = 2.0 ** f
w = np.ones_like(f)
V = np.vander(f, 4, increasing=True)
for _ in range(200):
    c, *_ = np.linalg.lstsq(V * (w / y)[:, None], w, rcond=None)
    e = np.abs(V @ c / y - 1.0)
    w *= 1.0 + 0.5 * e / e.max()
c32 = c.astype(np.float32)
x = np.float32(f[:-1])
p = np.float32(c32[3])
for k in (2, 1, 0):
    p = np.float32(p * x + c32[k])
rel = np.abs(p.astype(np.float64) / 2.0 ** x.astype(np.float64) - 1.0)
print("coefficients c0..c3:", ", ".join(repr(float(v)) for v in c32))
print(f"max relative error (float32 Horner): {rel.max():.3e}  (bf16 rounding of P: 3.9e-03)")
