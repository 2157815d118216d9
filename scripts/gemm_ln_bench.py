#!/usr/bin/env python
"""Fused GEMM + bias + residual + LayerNorm vs the unfused pair (GEMM fp32 out, LayerNorm with bias + residual) and cuBLAS."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops, _lib


def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


for M, N, K in ((16384, 1024, 1024), (16384, 1024, 4096), (4096, 1024, 1024), (16384, 1024, 64), (16384, 512, 1024), (16384, 256, 1024)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias, gamma, beta = torch.randn(N, device="cuda"), torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda").bfloat16()
    y = torch.empty(M, N, device="cuda", dtype=torch.float32)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    row = {"M": M, "N": N, "K": K}
    lws = ops.gemm_ln_workspace(M, N, "cuda")
    ops._GEMM_LN_IMPL = "grid"
    row["fused_us"] = round(timeit(lambda: ops.gemm_ln(a, w, bias, resid, gamma, beta, 1e-5, out=out, ws=lws)), 1)
    row["fused_noresid_us"] = round(timeit(lambda: ops.gemm_ln(a, w, bias, None, gamma, beta, 1e-5, out=out, ws=lws)), 1)
    y_grid = out.clone()
    ops._GEMM_LN_IMPL = "cluster"
    row["fused_cluster_us"] = round(timeit(lambda: ops.gemm_ln(a, w, bias, resid, gamma, beta, 1e-5, out=out)), 1)
    ops.gemm_ln(a, w, bias, None, gamma, beta, 1e-5, out=out)
    row["grid_vs_cluster_max_diff"] = float((out.float() - y_grid.float()).abs().max())
    ops._GEMM_LN_IMPL = "grid"
    row["gemm_f32_us"] = round(timeit(lambda: ops.gemm_bf16_tn(a, w, None, epilogue=3, out=y)), 1)
    row["ln_us"] = round(timeit(lambda: ops.layernorm_fwd(y, gamma, beta, 1e-5, out=out, bias=bias, resid=resid)), 1)

    def pair():
        ops.gemm_bf16_tn(a, w, None, epilogue=3, out=y)
        ops.layernorm_fwd(y, gamma, beta, 1e-5, out=out, bias=bias, resid=resid)
    row["pair_us"] = round(timeit(pair), 1)
    row["gemm_bf16_us"] = round(timeit(lambda: ops.gemm_bf16_tn(a, w, bias, epilogue=0, out=out)), 1)
    row["resident_clusters"] = _lib.load().kbner_gemm_ln_resident_clusters(N)
    print(json.dumps(row), flush=True)
