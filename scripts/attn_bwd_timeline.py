#!/usr/bin/env python
"""Debug-build only (KBNER_EXTRA_NVCC_FLAGS=-DKBNER_ATTN_BWD_DEBUG): clock64 stamps of the attention backward's MMA warp
(0 P/dS seen, 1 dV/dK/dQ issued, 2 next S/dP issued) and compute warps 0 / 7 (0 top, 1 S/dP ready, 2 P/dS written,
3 dV/dK/dQ retired, 4 dQ read out; block nqb: 0 = dK/dV stored) for CTAs 0 and 300 at the fine-tuning shape, with dropout."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops, _lib
R, S, heads = 8, 512, 16
H = heads * 64
qkv = torch.randn(R * S, 3 * H, device="cuda").bfloat16()
key_len = torch.full((R,), S, dtype=torch.int32, device="cuda")
seed = torch.tensor([1, 2], dtype=torch.int32, device="cuda")
drop = (seed, 3, 0.1) if os.environ.get("DROP", "1") == "1" else None
o, lse = ops.attention_fwd(qkv, key_len, R, S, heads, want_lse=True, drop=drop)
do = torch.randn(R * S, H, device="cuda").bfloat16()
dqkv = torch.empty(R * S, 3 * H, device="cuda", dtype=torch.bfloat16)
ws = (torch.empty((R, heads, S), dtype=torch.float32, device="cuda"), torch.empty((R * S, H), dtype=torch.float32, device="cuda"))
for _ in range(3):
    ops.attention_bwd(qkv, o, do, lse, key_len, R, S, heads, dqkv=dqkv, workspace=ws, drop=drop)
torch.cuda.synchronize()
n = 2 * 3 * 6 * 6
buf = (ctypes.c_ulonglong * n)()
lib = _lib.load()
lib.kbner_attention_bwd_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.kbner_attention_bwd_debug_read(buf, n)
v = list(buf)
res = {"rc": rc, "dropout": drop is not None}
for c in range(2):
    vals = [x for x in v[c * 108:(c + 1) * 108] if x]
    base = min(vals)
    for role, name in enumerate(("mma", "compute_w0", "compute_w7")):
        res["cta%d_%s" % (c, name)] = [[int(v[((c * 3 + role) * 6 + j) * 6 + k] - base) if v[((c * 3 + role) * 6 + j) * 6 + k] else None
                                        for k in range(5)] for j in range(5)]
print(json.dumps(res))
