#!/usr/bin/env python
"""Where does the time of the short-K / fp32-output GEMM go?  Sweeps K, the epilogue and the number of tile rounds at
N=1024 (attention-out / FFN-down shapes).  Prints one JSON line per case."""
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


N = int(os.environ.get("N", "1024"))
for M in (4736, 9472, 16384, 18944):          # 74 / 148 / 256 / 296 cluster tiles at N=1024
    for K in (64, 1024, 4096):
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda")
        resid = torch.randn(M, N, device="cuda").bfloat16()
        row = {"M": M, "N": N, "K": K, "tiles": ((M + 255) // 256) * ((N + 255) // 256)}
        for epi, name in ((0, "bias_bf16"), (3, "none_f32"), (2, "resid_f32")):
            c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if epi == 0 else torch.float32)
            us = timeit(lambda: ops.gemm_bf16_tn(a, b, bias if epi != 3 else None, resid if epi == 2 else None, epilogue=epi, out=c))
            row[name + "_us"] = round(us, 1)
        row["cublas_us"] = round(timeit(lambda: torch.matmul(a, b.t())), 1)
        print(json.dumps(row), flush=True)
