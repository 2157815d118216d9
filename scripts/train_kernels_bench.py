#!/usr/bin/env python
"""CUDA-event timing of the HBM-bound kernels of the fine-tuning micro-step at its real shapes (8 x 512 tokens, H = 1024):
LayerNorm backward (fused bias / residual / dropout), bias-gradient column sums, tag-projection backward, and the dgrad
GEMM through the FFN GELU (DGELU epilogue) next to the forward FFN-up GEMM of the same shape.  L2 is flushed between
iterations.  Prints one JSON object with microseconds and GB/s on algorithmic bytes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


REPS, WARM = int(os.environ.get("REPS", "9")), int(os.environ.get("WARM", "3"))   # REPS=1 WARM=0: one launch each (ncu)


def timeit(fn, reps=REPS):
    for _ in range(WARM):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


out = {}
M, H, F, L = 4096, 1024, 4096, 13
x = torch.randn(M, H, device=dev, generator=g)
dout = torch.randn(M, H, device=dev, generator=g)
resid = torch.randn(M, H, device=dev, generator=g).bfloat16()
dres = torch.randn(M, H, device=dev, generator=g).bfloat16()
bias = torch.randn(H, device=dev, generator=g)
gamma = torch.randn(H, device=dev, generator=g)
mean = x.mean(1).contiguous()
rstd = (1.0 / x.std(1)).contiguous()
dgamma, dbeta, dxsum = (torch.zeros(H, device=dev) for _ in range(3))
seed = torch.tensor([1, 2], dtype=torch.int32, device=dev)
o1 = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
o2 = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
us = timeit(lambda: ops.layernorm_bwd(x, dout, gamma, mean, rstd, dgamma, dbeta, out=o1, dxsum=dxsum, bias=bias, resid=resid,
                                      dres=dres, drop=(seed, 2, 0.1), out_masked=o2))
by = M * H * (4 + 4 + 2 + 2 + 2 + 2)
out["layernorm_bwd_fused_dropout"] = {"us": round(us, 2), "bytes": by, "GBps": round(by / us / 1e3, 1)}
for N in (3072, 4096):
    dY = torch.randn(M, N, device=dev, generator=g).bfloat16()
    db = torch.zeros(N, device=dev)
    us = timeit(lambda: ops.colsum_bf16(dY, db))
    out["colsum_%d" % N] = {"us": round(us, 2), "bytes": M * N * 2, "GBps": round(M * N * 2 / us / 1e3, 1)}
B, T, S = 8, 510, 512
hidden = torch.randn(B * S, H, device=dev, generator=g).bfloat16()
row_of = torch.arange(B, dtype=torch.int32, device=dev)
first = (torch.arange(T, dtype=torch.int32, device=dev)[None, :] + 1).repeat(B, 1).contiguous()
W = torch.randn(L, H, device=dev, generator=g) * 0.05
dlog = torch.randn(B, T, L, device=dev, generator=g)
d_hidden = torch.zeros(B * S, H, device=dev)
dW = torch.zeros(L, H, device=dev)
dbt = torch.zeros(L, device=dev)
us = timeit(lambda: ops.gather_tagproj_bwd(hidden, row_of, first, W, dlog, S, d_hidden, dW, dbt))
by = B * T * (H * 2 + H * 4 + L * 4)
out["gather_tagproj_bwd"] = {"us": round(us, 2), "bytes": by, "GBps": round(by / us / 1e3, 1)}
A = torch.randn(M, H, device=dev, generator=g).bfloat16()         # dY of FFN-down [tokens, H]
W2 = (torch.randn(H, F, device=dev, generator=g) * 0.02).bfloat16()  # FFN-down weight [out=H, in=F] read MN-major
aux = torch.randn(M, F, device=dev, generator=g).bfloat16()
o = torch.empty(M, F, dtype=torch.bfloat16, device=dev)
us = timeit(lambda: ops.gemm_bf16(A, W2, M, F, H, ops.EPI_DGELU_BF16, aux=aux, out=o, b_mn=True))
out["gemm_dgelu_4096x4096x1024"] = {"us": round(us, 2), "TFLOPs": round(2.0 * M * F * H / us / 1e6, 1)}
W1 = (torch.randn(F, H, device=dev, generator=g) * 0.02).bfloat16()
b1 = torch.randn(F, device=dev, generator=g)
us = timeit(lambda: ops.gemm_bf16(A, W1, M, F, H, ops.EPI_BIAS_GELU, bias=b1, aux_out=aux, out=o))
out["gemm_gelu_fwd_4096x4096x1024"] = {"us": round(us, 2), "TFLOPs": round(2.0 * M * F * H / us / 1e6, 1)}
us = timeit(lambda: ops.gemm_bf16(A, W1, M, F, H, ops.EPI_BIAS, bias=b1, out=o))
out["gemm_bias_4096x4096x1024"] = {"us": round(us, 2), "TFLOPs": round(2.0 * M * F * H / us / 1e6, 1)}
print(json.dumps(out))
