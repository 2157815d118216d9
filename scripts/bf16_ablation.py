#!/usr/bin/env python
"""CPU only: which bf16 rounding points of the inference path cost how much distance to the fp32 oracle?
Switches (all on = oracle.encoder_forward_bf16_points): W weights, X LayerNorm output used as GEMM operand, R the same
LayerNorm output used as RESIDUAL (off = fp32 residual stream), Q fused QKV output, P probabilities, C attention output,
G GELU output.  Prints one JSON line per variant: rel-L2 of the 24-layer hidden state and of 13-tag logits."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.nn.functional as F
import encoder_oracle as E

torch.set_num_threads(os.cpu_count() or 8)


def fwd(params, ids, key_len, cfg, on):
    rr = lambda t, k: t.to(torch.bfloat16).to(torch.float32) if k in on else t
    H, heads, NL, eps = cfg["hidden"], cfg["heads"], cfg["layers"], cfg.get("eps", 1e-5)
    R, S = ids.shape
    d = H // heads
    x = (params["embeddings.word_embeddings.weight"][ids] + params["embeddings.token_type_embeddings.weight"][0][None, None, :]
         + params["embeddings.position_embeddings.weight"][E.position_ids(ids, cfg.get("pad_id", 1))])
    x = F.layer_norm(x, (H,), params["embeddings.LayerNorm.weight"], params["embeddings.LayerNorm.bias"], eps)
    kmask = torch.arange(S)[None, :] < key_len[:, None]
    for i in range(NL):
        pre = "encoder.layer.%d." % i
        lin = lambda t, nm: F.linear(t, rr(params[pre + nm + ".weight"], "W"), params[pre + nm + ".bias"])
        xo, xr = rr(x, "X"), rr(x, "R")
        q = rr(lin(xo, "attention.self.query"), "Q").view(R, S, heads, d).transpose(1, 2)
        k = rr(lin(xo, "attention.self.key"), "Q").view(R, S, heads, d).transpose(1, 2)
        v = rr(lin(xo, "attention.self.value"), "Q").view(R, S, heads, d).transpose(1, 2)
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(d)
        sc = sc.masked_fill(~kmask[:, None, None, :], float("-inf"))
        p = torch.exp(sc - sc.amax(-1, keepdim=True))
        ctx = rr((rr(p, "P") @ v) / p.sum(-1, keepdim=True), "C").transpose(1, 2).reshape(R, S, H)
        x = F.layer_norm(lin(ctx, "attention.output.dense") + xr, (H,), params[pre + "attention.output.LayerNorm.weight"],
                         params[pre + "attention.output.LayerNorm.bias"], eps)
        xo, xr = rr(x, "X"), rr(x, "R")
        h = rr(F.gelu(lin(xo, "intermediate.dense")), "G")
        x = F.layer_norm(lin(h, "output.dense") + xr, (H,), params[pre + "output.LayerNorm.weight"],
                         params[pre + "output.LayerNorm.bias"], eps)
    return x


cfg = dict(E.XLMR_LARGE)
cfg["layers"] = int(os.environ.get("LAYERS", 24))
cfg["vocab"] = 5000
params = E.init_params(cfg, seed=1234)
g = torch.Generator().manual_seed(7)
S = 512
ids = torch.randint(3, cfg["vocab"], (1, S), generator=g)
ids[:, 0], ids[:, -1] = 0, 2
key_len = torch.tensor([S])
Wt = torch.randn(13, cfg["hidden"], generator=g) * 0.02
with torch.no_grad():
    ref = fwd(params, ids, key_len, cfg, "")
    for on in os.environ.get("VARIANTS", "WXRQPCG,WXQPCG,W,XR,X,R,Q,P,C,G,XQPCG,WX").split(","):
        t0 = time.time()
        out = fwd(params, ids, key_len, cfg, on)
        rel = float((out - ref).norm() / ref.norm())
        lo, lr = out @ Wt.t(), ref @ Wt.t()
        print(json.dumps({"rounded": on, "hidden_rel_l2": rel, "logits_rel_l2": float((lo - lr).norm() / lr.norm()),
                          "max_over_max": float((out - ref).abs().max() / ref.abs().max()), "s": round(time.time() - t0, 1)}), flush=True)
