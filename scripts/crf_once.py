#!/usr/bin/env python
"""Launches each CRF kernel twice (Viterbi, log-partition with alpha, gradient) at 4096 x 512 x 13 -- the target of
`ncu -k regex:crf_ -c 6` (scripts/gpu_final.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from kbner_b200 import ops
B, T, L = int(os.environ.get("B", "4096")), 512, 13
rng = np.random.RandomState(0)
trans = rng.randn(L, L).astype(np.float32)
trans[L - 2, :] = -1e12
trans[:, L - 1] = -1e12
trans = torch.from_numpy(trans).cuda()
emis = torch.randn(B, T, L, device="cuda") * 3
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
tags = torch.randint(1, L - 2, (B, T), device="cuda", dtype=torch.int32)
w = torch.full((B,), 1.0 / B, device="cuda")
for _ in range(2):
    ops.crf_viterbi(emis, trans, lens, lens, L - 2, L - 1)
for _ in range(2):
    _, _, alpha = ops.crf_nll_fwd(emis, tags, trans, lens, L - 2, L - 1, want_alpha=True)
for _ in range(2):
    ops.crf_nll_bwd(emis, tags, trans, lens, alpha, w, L - 2, L - 1)
torch.cuda.synchronize()
