#!/usr/bin/env python
"""CPU only: how far does the bf16 number FORMAT alone move the XLM-R-large hidden state from the fp32 oracle?
oracle.encoder_forward_bf16_points restates the rounding points of the inference kernels in torch; the rel-L2 distance to
the fp32 oracle on the same seeded weights / ids is the floor any bf16-operand implementation sits on (DESIGN.md section 5
puts the kernels' measured distance next to it).  Prints one JSON object."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import encoder_oracle as E

torch.set_num_threads(os.cpu_count() or 8)
cfg = dict(E.XLMR_LARGE)
layers = int(os.environ.get("LAYERS", cfg["layers"]))
cfg["layers"] = layers
cfg["vocab"] = 5000                       # the vocabulary size does not enter the arithmetic
params = E.init_params(cfg, seed=1234)
g = torch.Generator().manual_seed(7)
R, S = int(os.environ.get("R", "2")), 512
ids = torch.randint(3, cfg["vocab"], (R, S), generator=g)
ids[:, 0] = 0
key_len = torch.tensor([S, 300][:R] + [S] * max(0, R - 2))
for rr in range(R):
    ids[rr, key_len[rr] - 1] = 2
    ids[rr, key_len[rr]:] = 0
t0 = time.time()
with torch.no_grad():
    ref = E.encoder_forward(params, ids, key_len, cfg)
    emu = E.encoder_forward_bf16_points(params, ids, key_len, cfg)
out = {"layers": layers, "R": R, "S": S, "seconds": round(time.time() - t0, 1)}
num = den = 0.0
for rr in range(R):
    n = int(key_len[rr])
    a, b = emu[rr, :n], ref[rr, :n]
    num += float(((a - b) ** 2).sum()); den += float((b ** 2).sum())
out["rel_l2_bf16_points_vs_fp32"] = (num / den) ** 0.5
out["max_abs_over_max_abs"] = float((emu - ref)[0].abs().max() / ref[0].abs().max())
print(json.dumps(out))
