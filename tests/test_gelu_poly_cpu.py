"""The GELU of the GEMM epilogues (csrc/gemm_tcgen05.cu::gelu_tail2) restated in numpy fp32 with the kernel's OWN constants
(parsed from the source): gelu(x) = max(x, 0) - u Q(u), Q(u) = 2^p(u), u = min(|x|, 6), against the exact erf GELU of
transformers' `gelu` (x Phi(x)).  Pins the accuracy the kernel's comment claims without a GPU."""
import math
import os
import re

import numpy as np
from scipy.special import erfc

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kb-ner_b200", "csrc", "gemm_tcgen05.cu")


def _coefficients():
    body = open(SRC).read()
    body = body[body.index("void gelu_tail2("):body.index("void gelu_erf2(")]
    # Horner order in the source: highest degree first, each constant appears twice inside f2_pack(c, c)
    vals = [float(m) for m in re.findall(r"f2_pack\((-?[0-9.]+(?:e-?[0-9]+)?)f,", body)]
    assert len(vals) == 7, vals
    return vals


def _tail_fp32(x):
    c = [np.float32(v) for v in _coefficients()]
    u = np.minimum(np.abs(x), np.float32(6.0)).astype(np.float32)
    p = u * c[0] + c[1]
    for k in c[2:]:
        p = (p * u + k).astype(np.float32)
    return u, np.exp2(p.astype(np.float32)).astype(np.float32)


def test_gelu_from_the_upper_tail_polynomial():
    x = np.linspace(-12.0, 12.0, 1200001).astype(np.float32)
    u, q = _tail_fp32(x)
    gelu = (np.maximum(x, np.float32(0)) - u * q).astype(np.float32)
    xd = x.astype(np.float64)
    exact = xd * 0.5 * erfc(-xd / math.sqrt(2.0))
    assert float(np.abs(gelu - exact).max()) <= 1e-5                      # comment in the kernel: <= 6.7e-6
    inside = np.abs(xd) <= 6.0
    q_exact = 0.5 * erfc(np.abs(xd) / math.sqrt(2.0))
    assert float((np.abs(q - q_exact) / q_exact)[inside].max()) <= 1e-4   # relative error of the tail probability
    # far below the bf16 output's half ulp (2^-9 relative) wherever the output is not negligible
    big = np.abs(exact) >= 1e-3
    assert float((np.abs(gelu - exact) / np.abs(exact))[big].max()) <= 2.0 ** -12


def test_gelu_gradient_from_the_same_tail():
    x = np.linspace(-10.0, 10.0, 400001).astype(np.float32)
    u, q = _tail_fp32(x)
    phi_big = np.where(x >= 0, np.float32(1) - q, q).astype(np.float32)
    e = np.exp2((x * np.float32(-0.7213475204444817)) * x).astype(np.float32)
    grad = (x * np.float32(0.3989422804014327) * e + phi_big).astype(np.float32)
    xd = x.astype(np.float64)
    exact = 0.5 * erfc(-xd / math.sqrt(2.0)) + xd * np.exp(-0.5 * xd * xd) / math.sqrt(2.0 * math.pi)
    assert float(np.abs(grad - exact).max()) <= 3e-5
