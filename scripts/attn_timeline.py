#!/usr/bin/env python
"""Debug-build only (KBNER_EXTRA_NVCC_FLAGS=-DKBNER_ATTN_DEBUG): per-key-block clock64 stamps of the attention forward's
MMA warp and two softmax warps for CTAs 0 and 150 (first 24 key blocks = 3 work items); prints cycle offsets relative to the CTA's first stamp."""
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
NB = 24
n = 2 * 3 * NB * 4
buf = (ctypes.c_ulonglong * n)()
lib = _lib.load()
lib.kbner_attention_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.kbner_attention_debug_read(buf, n)
v = list(buf)
res = {"rc": rc}
for c in range(2):
    base = min(x for x in v[c * 3 * NB * 4:(c + 1) * 3 * NB * 4] if x)
    for role, name in enumerate(("mma[p_seen,v_ready,pv_issued,s_issued]", "softmax_w0[top,s_ready,max_done,p_arrived]", "softmax_w7")):
        res["cta%d_%s" % (c, name)] = [[int(v[((c * 3 + role) * NB + j) * 4 + k] - base) if v[((c * 3 + role) * NB + j) * 4 + k] else None
                                        for k in range(4)] for j in range(NB)]
print(json.dumps(res))
