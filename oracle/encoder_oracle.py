"""TEST INFRASTRUCTURE ONLY -- fp32 restatement of the XLM-R encoder arithmetic.

The reference calls a third-party module that is NOT under /root/reference:
``transformers==3.0.0`` (requirements.txt:30), ``XLMRobertaModel.forward`` from
``flair/embeddings.py:3269`` (construction :2951-2953).  This file restates its published
post-LN BERT algorithm in plain torch fp32 ops -- embeddings (word + type + position, position
ids = cumsum(ids != pad) * (ids != pad) + pad), LayerNorm(eps), 24 x [QKV, softmax(QK^T/sqrt(d)
+ key mask) V, out-proj + residual + LN, GELU(erf) FFN + residual + LN] -- over a state dict with
the HF parameter names.  It is pinned against the installed ``transformers`` 5.5
``XLMRobertaModel`` (eager attention) in tests/test_oracle_encoder.py, and against the
reference's first-sub-token pooling (flair/embeddings.py:3288-3345) by construction.
Parity note: transformers 3.0.0 itself is absent here => "parity unpinned" at that boundary
(SURVEY.md section 8(c)); the pin is the same architecture in transformers 5.5.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math

import torch
import torch.nn.functional as F


def init_params(cfg, seed=1234, dtype=torch.float32):
    """HF-style init: N(0, 0.02) Linear / Embedding, LN gamma=1 beta=0."""
    g = torch.Generator().manual_seed(seed)
    H, F_, V, P, NL = cfg["hidden"], cfg["ffn"], cfg["vocab"], cfg["max_pos"], cfg["layers"]

    def n(*shape):
        return (torch.randn(*shape, generator=g) * 0.02).to(dtype)

    p = {
        "embeddings.word_embeddings.weight": n(V, H),
        "embeddings.position_embeddings.weight": n(P, H),
        "embeddings.token_type_embeddings.weight": n(1, H),
        "embeddings.LayerNorm.weight": torch.ones(H, dtype=dtype),
        "embeddings.LayerNorm.bias": torch.zeros(H, dtype=dtype),
    }
    p["embeddings.word_embeddings.weight"][cfg.get("pad_id", 1)].zero_()
    p["embeddings.position_embeddings.weight"][cfg.get("pad_id", 1)].zero_()
    for i in range(NL):
        pre = "encoder.layer.%d." % i
        for nm, (o, k) in {"attention.self.query": (H, H), "attention.self.key": (H, H),
                           "attention.self.value": (H, H), "attention.output.dense": (H, H),
                           "intermediate.dense": (F_, H), "output.dense": (H, F_)}.items():
            p[pre + nm + ".weight"] = n(o, k)
            p[pre + nm + ".bias"] = n(o)          # non-zero biases so the bias path is exercised
        for nm in ("attention.output.LayerNorm", "output.LayerNorm"):
            p[pre + nm + ".weight"] = 1.0 + n(H)
            p[pre + nm + ".bias"] = n(H)
    return p


def position_ids(ids, pad_id=1):
    mask = (ids != pad_id).to(torch.int64)
    return torch.cumsum(mask, dim=1) * mask + pad_id


def encoder_forward(params, ids, key_len, cfg, all_layers=False, masks=None):
    """ids [R,S] int64, key_len [R] -> last hidden state [R,S,H] fp32 (or list of all 1+NL states).
    masks (training-mode dropout, transformers modeling_bert.py: BertEmbeddings.dropout, BertSelfAttention.dropout,
    BertSelfOutput.dropout, BertOutput.dropout): optional dict of multiplicative masks (0 or 1/(1-p)) with keys
    "emb" [R,S,H], ("attn", i) [R,heads,S,S], ("attn_out", i) [R,S,H], ("ffn_out", i) [R,S,H]; torch.nn.Dropout draws
    them from torch's generator, the tests feed the masks the CUDA path generated."""
    masks = masks or {}
    mk = lambda t, key: t * masks[key] if key in masks else t
    H, heads, NL, eps = cfg["hidden"], cfg["heads"], cfg["layers"], cfg.get("eps", 1e-5)
    pad = cfg.get("pad_id", 1)
    R, S = ids.shape
    d = H // heads
    x = (params["embeddings.word_embeddings.weight"][ids]
         + params["embeddings.token_type_embeddings.weight"][0][None, None, :]
         + params["embeddings.position_embeddings.weight"][position_ids(ids, pad)])
    x = mk(F.layer_norm(x, (H,), params["embeddings.LayerNorm.weight"], params["embeddings.LayerNorm.bias"], eps), "emb")
    kmask = torch.arange(S, device=ids.device)[None, :] < key_len[:, None].to(ids.device)
    states = [x]
    for i in range(NL):
        pre = "encoder.layer.%d." % i
        lin = lambda t, nm: F.linear(t, params[pre + nm + ".weight"], params[pre + nm + ".bias"])
        q = lin(x, "attention.self.query").view(R, S, heads, d).transpose(1, 2)
        k = lin(x, "attention.self.key").view(R, S, heads, d).transpose(1, 2)
        v = lin(x, "attention.self.value").view(R, S, heads, d).transpose(1, 2)
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(d)
        sc = sc.masked_fill(~kmask[:, None, None, :], torch.finfo(sc.dtype).min)
        ctx = (mk(torch.softmax(sc, -1), ("attn", i)) @ v).transpose(1, 2).reshape(R, S, H)
        x = F.layer_norm(mk(lin(ctx, "attention.output.dense"), ("attn_out", i)) + x, (H,),
                         params[pre + "attention.output.LayerNorm.weight"],
                         params[pre + "attention.output.LayerNorm.bias"], eps)
        h = F.gelu(lin(x, "intermediate.dense"))                   # erf GELU ("gelu")
        x = F.layer_norm(mk(lin(h, "output.dense"), ("ffn_out", i)) + x, (H,), params[pre + "output.LayerNorm.weight"],
                         params[pre + "output.LayerNorm.bias"], eps)
        states.append(x)
    return states if all_layers else x


def encoder_forward_bf16_points(params, ids, key_len, cfg):
    """The same arithmetic with the ROUNDING POINTS of the B200 inference path restated in torch: every GEMM operand is
    bf16 (weights and activations), accumulation and everything between a GEMM and the next store is fp32, and a value is
    rounded to bf16 exactly where the kernels store bf16 -- embedding LayerNorm output, fused QKV output, the unnormalised
    probabilities exp(s - rowmax) that feed P.V, the attention output, GELU(FFN-up) and both LayerNorm outputs.  Not a
    second oracle: it exists to separate the cost of the number FORMAT from kernel error (DESIGN.md section 5 reports
    |this - fp32 oracle| next to |kernels - fp32 oracle|).  Returns the last hidden state [R,S,H] (bf16 values in fp32)."""
    r = lambda t: t.to(torch.bfloat16).to(torch.float32)
    H, heads, NL, eps = cfg["hidden"], cfg["heads"], cfg["layers"], cfg.get("eps", 1e-5)
    pad = cfg.get("pad_id", 1)
    R, S = ids.shape
    d = H // heads
    x = (params["embeddings.word_embeddings.weight"][ids]
         + params["embeddings.token_type_embeddings.weight"][0][None, None, :]
         + params["embeddings.position_embeddings.weight"][position_ids(ids, pad)])
    x = r(F.layer_norm(x, (H,), params["embeddings.LayerNorm.weight"], params["embeddings.LayerNorm.bias"], eps))
    kmask = torch.arange(S, device=ids.device)[None, :] < key_len[:, None].to(ids.device)
    for i in range(NL):
        pre = "encoder.layer.%d." % i
        lin = lambda t, nm: F.linear(t, r(params[pre + nm + ".weight"]), params[pre + nm + ".bias"])
        q = r(lin(x, "attention.self.query")).view(R, S, heads, d).transpose(1, 2)
        k = r(lin(x, "attention.self.key")).view(R, S, heads, d).transpose(1, 2)
        v = r(lin(x, "attention.self.value")).view(R, S, heads, d).transpose(1, 2)
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(d)
        sc = sc.masked_fill(~kmask[:, None, None, :], float("-inf"))
        p = torch.exp(sc - sc.amax(-1, keepdim=True))                     # fp32, row sum taken before the rounding
        ctx = r((r(p) @ v) / p.sum(-1, keepdim=True)).transpose(1, 2).reshape(R, S, H)
        x = r(F.layer_norm(lin(ctx, "attention.output.dense") + x, (H,), params[pre + "attention.output.LayerNorm.weight"],
                           params[pre + "attention.output.LayerNorm.bias"], eps))
        h = r(F.gelu(lin(x, "intermediate.dense")))
        x = r(F.layer_norm(lin(h, "output.dense") + x, (H,), params[pre + "output.LayerNorm.weight"],
                           params[pre + "output.LayerNorm.bias"], eps))
    return x


def first_subtoken_pool(hidden, row_of, first_idx):
    """flair/embeddings.py:3300-3345 ('first' pooling): word t of sentence b = hidden[row_of[b], first_idx[b,t]];
    first_idx < 0 (word with 0 sub-tokens / padding) -> zero vector (:3306-3308)."""
    B, T = first_idx.shape
    out = hidden.new_zeros((B, T, hidden.shape[-1]))
    for b in range(B):
        for t in range(T):
            if first_idx[b, t] >= 0:
                out[b, t] = hidden[row_of[b], first_idx[b, t]]
    return out


def tagger_logits(params, hidden, row_of, first_idx, W, bias):
    """sequence_tagger_model.py:909-1027 in eval mode: Linear over the pooled word vectors."""
    return F.linear(first_subtoken_pool(hidden, row_of, first_idx), W, bias)


XLMR_LARGE = dict(hidden=1024, heads=16, ffn=4096, layers=24, vocab=250002, max_pos=514, eps=1e-5, pad_id=1)
XLMR_BASE = dict(hidden=768, heads=12, ffn=3072, layers=12, vocab=250002, max_pos=514, eps=1e-5, pad_id=1)
