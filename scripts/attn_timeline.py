#!/usr/bin/env python
"""Debug-build only (KBNER_EXTRA_NVCC_FLAGS=-DKBNER_ATTN_DEBUG): per-key-block clock64 stamps of the attention forward's
MMA warp and two softmax warps for CTAs (0,0,0) and (0,0,20); prints cycle offsets relative to the CTA's first stamp."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops, _lib
R, S, heads = 32, 512, 16
H = heads * 64
qkv = torch.randn(R * S, 3 * H, device="cuda").bfloat16()
key_len = torch.full((R,), S, dtype=torch.int32, device="cuda")
out = torch.empty(R * S, H, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention_fwd(qkv, key_len, R, S, heads, out=out)
torch.cuda.synchronize()
n = 2 * 3 * 8 * 4
buf = (ctypes.c_ulonglong * n)()
lib = _lib.load()
lib.kbner_attention_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.kbner_attention_debug_read(buf, n)
v = list(buf)
res = {"rc": rc}
for c in range(2):
    base = min(x for x in v[c * 96:(c + 1) * 96] if x)
    for role, name in enumerate(("mma[p_seen,v_ready,s_issued,-]", "softmax_w0[top,s_ready,max_done,p_arrived]", "softmax_w7")):
        res["cta%d_%s" % (c, name)] = [[int(v[((c * 3 + role) * 8 + j) * 4 + k] - base) if v[((c * 3 + role) * 8 + j) * 4 + k] else None
                                        for k in range(4)] for j in range(8)]
print(json.dumps(res))
