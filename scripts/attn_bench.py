#!/usr/bin/env python
"""Attention forward / backward alone at the bench shape (R=32, S=512, 16 heads), with a cuBLAS GEMM timed next to it as a
box-speed reference (boxes differ by a few percent under the power cap)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops


def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


R, S, heads = int(os.environ.get("R", "32")), 512, 16
H = heads * 64
qkv = torch.randn(R * S, 3 * H, device="cuda").bfloat16()
key_len = torch.full((R,), S, dtype=torch.int32, device="cuda")
out = torch.empty(R * S, H, device="cuda", dtype=torch.bfloat16)
row = {"attn_fwd_us": round(timeit(lambda: ops.attention_fwd(qkv, key_len, R, S, heads, out=out)), 1)}
seed = torch.tensor([1, 2], dtype=torch.int32, device="cuda")
o, lse = ops.attention_fwd(qkv, key_len, R, S, heads, want_lse=True)
row["attn_fwd_dropout_us"] = round(timeit(lambda: ops.attention_fwd(qkv, key_len, R, S, heads, out=out, drop=(seed, 3, 0.1))), 1)
do = torch.randn(R * S, H, device="cuda").bfloat16()
dqkv = torch.empty(R * S, 3 * H, device="cuda", dtype=torch.bfloat16)
ws = (torch.empty((R, heads, S), dtype=torch.float32, device="cuda"), torch.empty((R * S, H), dtype=torch.float32, device="cuda"))
row["attn_bwd_us"] = round(timeit(lambda: ops.attention_bwd(qkv, o, do, lse, key_len, R, S, heads, dqkv=dqkv, workspace=ws), 20), 1)
od, lsed = ops.attention_fwd(qkv, key_len, R, S, heads, want_lse=True, drop=(seed, 3, 0.1))
row["attn_bwd_dropout_us"] = round(timeit(lambda: ops.attention_bwd(qkv, od, do, lsed, key_len, R, S, heads, dqkv=dqkv, workspace=ws,
                                                                 drop=(seed, 3, 0.1)), 20), 1)
a = torch.randn(16384, 1024, device="cuda").bfloat16()
b = torch.randn(3072, 1024, device="cuda").bfloat16()
row["cublas_qkv_us"] = round(timeit(lambda: torch.matmul(a, b.t())), 1)
row["own_qkv_us"] = round(timeit(lambda: ops.gemm_bf16_tn(a, b, None, epilogue=3)), 1)
print(json.dumps(row))
