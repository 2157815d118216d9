#!/usr/bin/env python
"""256 x 256 vs 256 x 128 tiles of the tcgen05 GEMM (KBNER_GEMM_TN is read once per process: run once per width).
Shapes: the inference GEMMs (M = 16384) and the fine-tuning forward / dgrad ones (M = 4096)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops


def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


tn = os.environ.get("KBNER_GEMM_TN", "auto")
for M, N, K, epi in ((16384, 3072, 1024, 0), (16384, 4096, 1024, 1), (16384, 1024, 1024, 3), (16384, 1024, 4096, 3),
                     (4096, 3072, 1024, 0), (4096, 4096, 1024, 1), (4096, 1024, 4096, 3), (4096, 1024, 1024, 3), (4096, 4096, 1024, 0)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    c = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi == 3 else torch.bfloat16)
    us = timeit(lambda: ops.gemm_bf16_tn(a, b, None if epi == 3 else bias, epilogue=epi, out=c))
    print(json.dumps({"tn": tn, "M": M, "N": N, "K": K, "epi": epi, "us": round(us, 1), "tflops": round(2e-6 * M * N * K / us, 1)}), flush=True)
