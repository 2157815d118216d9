#!/usr/bin/env python
"""Per-shape throughput of the tcgen05 GEMM (the four encoder shapes at M=16384) next to cuBLAS (torch.matmul)
on the same box.  KBNER_GEMM selects the implementation.  Prints JSON."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops

def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps

out = {"impl": os.environ.get("KBNER_GEMM", "1"), "rows": []}
M = int(os.environ.get("M", "16384"))
only = os.environ.get("SHAPES")
for name, N, K, epi in (("qkv", 3072, 1024, 0), ("attn_out", 1024, 1024, 2), ("ffn_up", 4096, 1024, 1), ("ffn_down", 1024, 4096, 2), ("plain_f32", 4096, 1024, 3)):
    if only and name not in only.split(","):
        continue
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda").bfloat16()
    c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
    ms = timeit(lambda: ops.gemm_bf16_tn(a, b, bias if epi != 3 else None, resid if epi == 2 else None, epilogue=epi, out=c), 30)
    ms_cublas = timeit(lambda: torch.matmul(a, b.t()), 30)
    fl = 2.0 * M * N * K
    out["rows"].append({"shape": name, "M": M, "N": N, "K": K, "epi": epi, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1),
                        "cublas_ms": round(ms_cublas, 4), "cublas_tflops": round(fl / ms_cublas / 1e9, 1)})
print(json.dumps(out))
