#!/usr/bin/env python
"""gather + tag projection alone at the bench shape (32 x 510 words, H = 1024, L = 13): HBM roofline = 2H B in + 4L B out per word."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops

R, S, H, L = 32, 512, 1024, 13
T = S - 2
hidden = torch.randn(R * S, H, device="cuda").bfloat16()
row_of = torch.arange(R, dtype=torch.int32, device="cuda")
first_idx = (torch.arange(T, dtype=torch.int32, device="cuda") + 1).repeat(R, 1).contiguous()
W = torch.randn(L, H, device="cuda") * 0.02
b = torch.randn(L, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(13):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); out = ops.gather_tagproj_fwd(hidden, row_of, first_idx, W, b, S); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e) * 1e3)
us = sorted(ts[3:])[len(ts[3:]) // 2]
ref = hidden.view(R, S, H)[:, 1:T + 1].float() @ W.t() + b
err = (out - ref).abs().max().item()
byts = R * T * (2 * H + 4 * L)
print(json.dumps({"tagproj_us": round(us, 1), "GBps": round(byts / us / 1e3, 1), "max_abs_err": err}))
