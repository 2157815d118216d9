#!/usr/bin/env python
"""Bisect harness: runs attention fwd+bwd for one shape in THIS process, printing progress; the shell wrapper runs each
case under `timeout`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops
R, S, heads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lens = [int(x) for x in sys.argv[4].split(",")]
H = heads * 64
g = torch.Generator(device="cuda").manual_seed(1)
qkv = torch.randn(R * S, 3 * H, device="cuda", generator=g).bfloat16()
kl = torch.tensor(lens, dtype=torch.int32, device="cuda")
out, lse = ops.attention_fwd(qkv, kl, R, S, heads, want_lse=True)
torch.cuda.synchronize(); print("fwd ok", flush=True)
d_out = torch.randn(R * S, H, device="cuda", generator=g).bfloat16()
dqkv = ops.attention_bwd(qkv, out, d_out, lse, kl, R, S, heads)
torch.cuda.synchronize(); print("bwd ok", float(dqkv.float().abs().sum()), flush=True)
