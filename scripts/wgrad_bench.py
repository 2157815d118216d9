#!/usr/bin/env python
"""The four weight-gradient GEMMs of an encoder layer at the fine-tuning shape (4096 tokens): dW[out, in] += dY^T . X with
both operands read MN-major in place, fp32 reduce-add epilogue.  KBNER_GEMM_STREAMK=0 selects round 1's split-K."""
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


M = 4096
mode = "split-K" if os.environ.get("KBNER_GEMM_STREAMK") == "0" else "stream-K"
tot = 0.0
for name, n_out, n_in in (("qkv", 3072, 1024), ("attn_out", 1024, 1024), ("ffn_up", 4096, 1024), ("ffn_down", 1024, 4096)):
    dy = (torch.randn(M, n_out, device="cuda") * 0.1).bfloat16()
    x = (torch.randn(M, n_in, device="cuda") * 0.5).bfloat16()
    dw = torch.zeros(n_out, n_in, device="cuda")
    ops.gemm_bf16(dy, x, n_out, n_in, M, ops.EPI_ACCUM_F32, out=dw, a_mn=True, b_mn=True)
    ref = dy.float().t() @ x.float()
    err = float((dw - ref).abs().max() / ref.abs().max())
    us = timeit(lambda: ops.gemm_bf16(dy, x, n_out, n_in, M, ops.EPI_ACCUM_F32, out=dw, a_mn=True, b_mn=True))
    tot += us
    print(json.dumps({"mode": mode, "wgrad": name, "out": n_out, "in": n_in, "tokens": M, "us": round(us, 1),
                      "tflops": round(2e-6 * M * n_out * n_in / us, 1), "max_err_over_max": err}), flush=True)
print(json.dumps({"mode": mode, "layer_total_us": round(tot, 1)}))

# the same four gradients as ONE grouped launch
probs = []
for name, n_out, n_in in (("ffn_down", 1024, 4096), ("ffn_up", 4096, 1024), ("attn_out", 1024, 1024), ("qkv", 3072, 1024)):
    probs.append(((torch.randn(M, n_out, device="cuda") * 0.1).bfloat16(), (torch.randn(M, n_in, device="cuda") * 0.5).bfloat16(),
                  torch.zeros(n_out, n_in, device="cuda")))
us = timeit(lambda: ops.gemm_wgrad_group(probs))
fl = sum(2.0 * M * p[0].shape[1] * p[1].shape[1] for p in probs)
print(json.dumps({"mode": "grouped", "layer_total_us": round(us, 1), "tflops": round(fl / us / 1e6, 1)}))
