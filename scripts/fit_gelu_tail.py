#!/usr/bin/env python
"""Fit of p(u) = log2 Q(u), Q(u) = 0.5 erfc(u / sqrt 2), on [0, c] (Chebyshev basis, Lawson-reweighted least squares ->
near-minimax), evaluated in fp32 Horner form the way csrc/gemm_tcgen05.cu::gelu_tail2 does: gelu(x) = max(x, 0) - u Q(u).
The kernel uses the c = 6, degree 6 row."""
import numpy as np, math
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erfc
def cheb_to_mono(coef, lo, hi):
    poly_x = C.cheb2poly(coef)
    a, b = 2 / (hi - lo), -(hi + lo) / (hi - lo)
    res = np.zeros(len(poly_x)); cur = np.array([1.0]); lin = np.array([b, a])
    for ci in poly_x:
        res[:len(cur)] += ci * cur
        cur = P.polymul(cur, lin)
    return res
def fit_exp(c, deg):
    n = 6000; k = np.arange(n)
    u = 0.5 * c * (1 - np.cos(np.pi * (k + 0.5) / n))
    p = np.log2(0.5 * erfc(u / math.sqrt(2)))
    x = 2 * u / c - 1
    # error in gelu = u*Q*ln2*dp  -> weight u*Q (abs error) ; relative error of tail = ln2*dp -> want uniform dp mostly
    w = np.ones_like(u)
    ww = w.copy()
    for _ in range(80):
        coef = C.chebfit(x, p, deg, w=ww)
        err = np.abs(C.chebval(x, coef) - p)
        ww = ww * (0.5 + err / err.max())
    return cheb_to_mono(coef, 0.0, c)
def eval_f32(co, v, c):
    v = v.astype(np.float32)
    u = np.minimum(np.abs(v), np.float32(c))
    acc = np.float32(co[-1]) * np.ones_like(u)
    for ci in co[-2::-1]:
        acc = acc * u + np.float32(ci)
    q = np.exp2(acc).astype(np.float32)
    return np.maximum(v, 0) - u * q, q
vv = np.linspace(-12, 12, 2000001)
Q = 0.5 * erfc(np.abs(vv) / math.sqrt(2))
true_gelu = vv * 0.5 * erfc(-vv / math.sqrt(2))
for c, deg in ((5.0, 5), (5.0, 6), (5.0, 7), (6.0, 6), (6.0, 7), (6.0, 8), (7.0, 7), (7.0,8)):
    co = fit_exp(c, deg)
    g, q = eval_f32(co, vv, c)
    m = np.abs(vv) <= c
    relq = (np.abs(q - Q) / Q)[m].max()
    e_abs = np.abs(g - true_gelu).max()
    rel = (np.abs(g - true_gelu) / np.maximum(np.abs(true_gelu), 1e-6)).max()
    print("c=%.1f deg=%d  max rel dQ (|v|<=c)=%.2e  max|dGELU|=%.2e  max rel GELU (floor 1e-6)=%.2e" % (c, deg, relq, e_abs, rel), co.tolist() if deg in (6,7) and c==6.0 else "")
