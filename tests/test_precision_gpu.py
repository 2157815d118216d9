"""GPU: the floating-point parity row stated as numbers next to BASELINE.json's tolerance.

north_star: "logits and CRF loss within 1e-3 relative in bf16" against the reference's fp32 arithmetic (the transformers
module called at /root/reference/flair/embeddings.py:3269 + sequence_tagger_model.py:1027, :2499-2506).  What is asserted:

  precision "bf16x3"      logits rel-L2 <= 1e-3, CRF loss rel err <= 1e-3          -- the tolerance north_star states
  precision "bf16-res32"  hidden rel-L2 <= 1e-2 and >= 10 % below the plain bf16 figure (fp32 residual stream only)
  precision "bf16"        hidden rel-L2 <= 1.3e-2 vs fp32 (the number FORMAT: bf16 weights alone cost 6.4e-3,
                          scripts/bf16_ablation.py) and no further from fp32 than the bf16-rounding-point restatement of
                          the same arithmetic (oracle.encoder_forward_bf16_points): the kernels sit ON the format's floor
plus Viterbi tag agreement with the fp32 path in every mode, the kernels the modes add, and the backward pass at
BASELINE configs[2] shape (24 layers, 8 x 512) against autograd through the fp32 oracle.
"""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LARGE = dict(vocab_size=250002, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
             max_position_embeddings=514)
NORTH_STAR_TOL = 1e-3


@pytest.fixture(autouse=True)
def _fp32_oracle_math():
    """The oracle runs in torch on the GPU: make sure its matmuls are real fp32, not TF32."""
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _large(seed=11, small_vocab=True):
    from test_api_gpu import _models
    kw = dict(LARGE)
    if small_vocab:
        kw["vocab_size"] = 5000             # the vocabulary size does not enter the arithmetic
    return _models(kw, 13, seed=seed)


def _batch(n_full, n_short, seed=5):
    from kbner_b200.data import BatchedData, Sentence
    rnd = random.Random(seed)
    sents = [Sentence(tokens=["w%03x" % rnd.randrange(4096) for _ in range(510)]) for _ in range(n_full)]
    sents += [Sentence(tokens=["w%03x" % rnd.randrange(4096) for _ in range(rnd.randint(40, 400))]) for _ in range(n_short)]
    return BatchedData(sents)


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_kernels_sit_on_the_bf16_rounding_floor():
    """Kernels vs the torch restatement that rounds to bf16 at exactly the kernels' store points
    (oracle.encoder_forward_bf16_points).  ONE layer: the two agree tightly -- same arithmetic, different summation order.
    24 layers: two bf16 evaluations of the same function decorrelate (a value that rounds the other way moves by one bf16
    ulp = 4e-3 relative, and 96 GEMMs amplify that), so the meaningful statement is that both sit at the SAME distance from
    fp32: the kernels add nothing measurable on top of the number format."""
    import encoder_oracle as E
    from test_api_gpu import _models
    out = {}
    for layers in (1, 24):
        kw = dict(LARGE, vocab_size=5000, num_hidden_layers=layers)
        tagger, emb, params, ocfg = _models(kw, 13, seed=11)
        batch = _batch(1, 1)
        ids, key_len, row_of, first_idx, lengths, S = emb.build_batch(batch)
        dp = {k: v.cuda() for k, v in params.items()}
        with torch.no_grad():
            hid = emb.model.forward_hidden(ids.cuda(), key_len.cuda()).float().view(2, S, -1).clone()
            emu = E.encoder_forward_bf16_points(dp, ids.long().cuda(), key_len.long().cuda(), ocfg)
            ref = E.encoder_forward(dp, ids.long().cuda(), key_len.long().cuda(), ocfg)
        n0, n1 = int(key_len[0]), int(key_len[1])
        cat = lambda t: torch.cat([t[0, :n0], t[1, :n1]])
        out[layers] = dict(kernels_vs_restatement=_rel(cat(hid), cat(emu)), kernels_vs_fp32=_rel(cat(hid), cat(ref)),
                           restatement_vs_fp32=_rel(cat(emu), cat(ref)))
        del tagger, emb, dp
    print("kernels / bf16-point restatement / fp32 oracle:", out)
    one, deep = out[1], out[24]
    # one layer, measured: 1.8e-3 to the restatement (rounding-boundary flips: ex2.approx / polynomial-erf / summation
    # order move a pre-rounding value by ~1e-6, which flips ~0.1 % of the bf16 roundings by a whole ulp), 3.2e-3 to fp32
    assert one["kernels_vs_restatement"] < 2.5e-3 and one["kernels_vs_fp32"] < 4e-3, out
    assert one["kernels_vs_restatement"] < 0.7 * one["kernels_vs_fp32"], out        # closer to its restatement than to fp32
    assert abs(one["kernels_vs_fp32"] - one["restatement_vs_fp32"]) < 0.1 * one["restatement_vs_fp32"], out
    assert deep["kernels_vs_fp32"] < 1.3e-2 and deep["restatement_vs_fp32"] < 1.3e-2, out
    assert deep["kernels_vs_fp32"] < 1.1 * deep["restatement_vs_fp32"], out         # the kernels sit on the format's floor


def test_precision_modes_meet_the_stated_tolerances():
    """hidden / logits / CRF loss / Viterbi tags of every precision mode against the fp32 oracle, configs[1] row shape."""
    import crf_oracle as O
    import encoder_oracle as E
    tagger, emb, params, ocfg = _large(seed=12)
    batch = _batch(3, 1, seed=8)
    ids, key_len, row_of, first_idx, lengths, S = emb.build_batch(batch)
    B, T = first_idx.shape
    dp = {k: v.cuda() for k, v in params.items()}
    W, bvec = tagger.linear.weight.detach().float(), tagger.linear.bias.detach().float()
    trans = tagger.transitions.detach().cpu().numpy()
    with torch.no_grad():
        ref_h = E.encoder_forward(dp, ids.long().cuda(), key_len.long().cuda(), ocfg)
        flat = ref_h.reshape(-1, ref_h.shape[-1])
        idx = row_of.long()[:, None] * S + first_idx.long().clamp(min=0)
        x = flat[idx.cuda()] * (first_idx >= 0).float().cuda()[..., None]
        ref_logits = x @ W.t() + bvec
    lens = np.array(lengths, np.int32)
    rng = np.random.RandomState(3)
    legal = [i for i in range(13) if i not in (0, tagger.x_idx, tagger.start_idx, tagger.stop_idx)]
    gold = np.zeros((B, T), np.int32)
    for b, n in enumerate(lengths):
        gold[b, :n] = rng.choice(legal, n)
    keep = (np.arange(T)[None, :] < lens[:, None]).astype(np.uint8)
    ref_loss = float(O.crf_loss(ref_logits.cpu().numpy(), gold, trans, keep, start=tagger.start_idx, stop=tagger.stop_idx))
    ref_tags, _ = O.viterbi(ref_logits.cpu().numpy(), trans, lens, start=tagger.start_idx, stop=tagger.stop_idx, x_idx=tagger.x_idx)
    for s, row in zip(batch, gold):
        s.ner_tags = torch.from_numpy(row[:len(s.tokens)].copy())
    valid = torch.from_numpy(keep.astype(bool)).cuda()
    res = {}
    for mode in ("bf16", "bf16-res32", "bf16x3"):
        emb.model.set_precision(mode)
        with torch.no_grad():
            for _ in range(2):                                   # second call replays the captured graph
                batch.features = {}
                feats = tagger.forward(batch)
            hid = batch.features[emb.name].hidden.float().view(-1, S, ref_h.shape[-1])
            h_rel = max(_rel(hid[r, :int(key_len[r])], ref_h[r, :int(key_len[r])]) for r in range(hid.shape[0]))
            l_rel = _rel(feats[valid], ref_logits[valid])
            loss = float(tagger._calculate_loss(feats, batch, tagger.mask))
            tags, _ = tagger._decode_batch(feats)
        agree = float((torch.from_numpy(ref_tags).cuda()[valid] == tags[valid]).float().mean())
        res[mode] = dict(hidden=h_rel, logits=l_rel, loss_rel=abs(loss - ref_loss) / abs(ref_loss), tag_agreement=agree)
    emb.model.set_precision("bf16")
    print("precision modes vs fp32 oracle (24 layers, 4 windows):", res)
    x3, r32, fast = res["bf16x3"], res["bf16-res32"], res["bf16"]
    assert x3["logits"] <= NORTH_STAR_TOL and x3["loss_rel"] <= NORTH_STAR_TOL and x3["hidden"] <= NORTH_STAR_TOL, res
    assert x3["tag_agreement"] >= 0.995, res
    assert r32["hidden"] <= 1.0e-2 and r32["hidden"] < 0.9 * fast["hidden"], res
    assert fast["hidden"] <= 1.3e-2 and fast["logits"] <= 2e-2 and fast["loss_rel"] <= 5e-3, res
    assert fast["tag_agreement"] >= 0.95 and r32["tag_agreement"] >= 0.95, res


def test_split_kernels():
    """The kernels the precision modes add, one by one, against torch."""
    from kbner_b200 import ops
    torch.manual_seed(0)
    M, H, F = 300, 1024, 4096
    dev = "cuda"
    hi_lo = lambda t3, n: t3[:, :n].float() + t3[:, n:2 * n].float()
    # LayerNorm with fp32 residual, fp32 + split outputs
    x, res = torch.randn(M, H, device=dev) * 3, torch.randn(M, H, device=dev)
    bias, g, b = torch.randn(H, device=dev), 1 + 0.1 * torch.randn(H, device=dev), 0.1 * torch.randn(H, device=dev)
    want = torch.nn.functional.layer_norm(x + bias + res, (H,), g, b, 1e-5)
    y3, y32 = torch.empty(M, 3 * H, dtype=torch.bfloat16, device=dev), torch.empty(M, H, device=dev)
    ops.layernorm_fwd_res32(x, g, b, 1e-5, out=y3, out32=y32, bias=bias, resid=res, split=True)
    assert (y32 - want).abs().max().item() < 2e-5
    assert torch.equal(y3[:, :H], y32.bfloat16()) and torch.equal(y3[:, :H], y3[:, 2 * H:])
    assert (hi_lo(y3, H) - y32).abs().max().item() <= 2 ** -15 * y32.abs().max().item()
    y1 = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
    ops.layernorm_fwd_res32(x, g, b, 1e-5, out=y1, out32=None, bias=None, resid=None, split=False)
    assert torch.equal(y1, torch.nn.functional.layer_norm(x, (H,), g, b, 1e-5).bfloat16()) or \
        (y1.float() - torch.nn.functional.layer_norm(x, (H,), g, b, 1e-5)).abs().max().item() < 4e-2
    # bias + erf-GELU + split
    z, bf = torch.randn(M, F, device=dev) * 2, torch.randn(F, device=dev)
    h3 = torch.empty(M, 3 * F, dtype=torch.bfloat16, device=dev)
    ops.bias_gelu_split(z, bf, h3)
    wantg = torch.nn.functional.gelu(z + bf)
    assert (hi_lo(h3, F) - wantg).abs().max().item() < 1e-5 * max(1.0, wantg.abs().max().item()) + 2e-6
    assert torch.equal(h3[:, :F], h3[:, 2 * F:])
    # the K-concatenated GEMM: [x_hi | x_lo | x_hi] . [W_hi | W_hi | W_lo]^T ~ fp32 x . W^T
    a, w = torch.randn(M, H, device=dev), torch.randn(512, H, device=dev) * 0.05
    a_hi = a.bfloat16(); a_lo = (a - a_hi.float()).bfloat16()
    w_hi = w.bfloat16(); w_lo = (w - w_hi.float()).bfloat16()
    c = ops.gemm_bf16_tn(torch.cat([a_hi, a_lo, a_hi], 1).contiguous(), torch.cat([w_hi, w_hi, w_lo], 1).contiguous(),
                         epilogue=ops.EPI_NONE_F32)
    exact = (a.double() @ w.double().t()).float()
    plain = ops.gemm_bf16_tn(a_hi.contiguous(), w_hi.contiguous(), epilogue=ops.EPI_NONE_F32)
    e3, e1 = _rel(c, exact), _rel(plain, exact)
    print("bf16x3 GEMM rel-L2 %.2e (plain bf16 %.2e)" % (e3, e1))
    assert e3 < 3e-5 and e1 > 20 * e3
    # embedding + LayerNorm: fp32 copy and split rows agree with the plain kernel
    ids = torch.randint(3, 500, (3, 64), device=dev, dtype=torch.int32)
    we, pe, te = torch.randn(500, H, device=dev) * 0.02, torch.randn(70, H, device=dev) * 0.02, torch.randn(H, device=dev) * 0.02
    base = ops.embed_ln_fwd(ids, we, pe, te, g, b, 1e-5, 1)
    o32 = torch.empty(3 * 64, H, device=dev)
    o3 = ops.embed_ln_fwd(ids, we, pe, te, g, b, 1e-5, 1, out32=o32, split=True)
    assert torch.equal(o3[:, :H], base) and torch.equal(o32.bfloat16(), base) and torch.equal(o3[:, 2 * H:], base)
    assert (hi_lo(o3, H) - o32).abs().max().item() <= 2 ** -15 * o32.abs().max().item()
    # tag projection over an fp32 hidden state
    hid = torch.randn(2 * 64, H, device=dev)
    row_of = torch.tensor([0, 1], dtype=torch.int32, device=dev)
    fi = torch.randint(-1, 64, (2, 40), dtype=torch.int32, device=dev)
    Wt, bt = torch.randn(13, H, device=dev) * 0.1, torch.randn(13, device=dev)
    lg = ops.gather_tagproj_fwd(hid, row_of, fi, Wt, bt, 64)
    idx = (row_of.long()[:, None] * 64 + fi.long().clamp(min=0))
    wantl = (hid[idx] * (fi >= 0).float()[..., None]) @ Wt.t() + bt
    assert (lg - wantl).abs().max().item() < 2e-4


def test_attention_split_output_and_row_stride():
    """attention_fwd(split=True): rows [ hi | lo | hi ]; hi is bit-identical to the plain output, hi + lo is closer to fp32."""
    from kbner_b200 import ops
    torch.manual_seed(1)
    R, S, heads = 3, 320, 4                      # S not a multiple of 128: ragged query block and ragged key block
    H = heads * 64
    qkv = (torch.randn(R * S, 3 * H, device="cuda") * 0.7).bfloat16()
    key_len = torch.tensor([320, 200, 65], dtype=torch.int32, device="cuda")
    plain = ops.attention_fwd(qkv, key_len, R, S, heads)
    buf = torch.full((R * S, 3 * H), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.attention_fwd(qkv, key_len, R, S, heads, out=buf, split=True)
    assert torch.equal(buf[:, :H], plain) and torch.equal(buf[:, 2 * H:], plain)
    q, k, v = [t.float().view(R, S, heads, 64).transpose(1, 2) for t in qkv.split(H, dim=1)]
    sc = (q @ k.transpose(-1, -2)) / 8.0
    kmask = torch.arange(S, device="cuda")[None, :] < key_len[:, None]
    sc = sc.masked_fill(~kmask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(R * S, H)
    e_hi = (buf[:, :H].float() - ref).abs().max().item()
    e_sum = (buf[:, :H].float() + buf[:, H:2 * H].float() - ref).abs().max().item()
    print("attention output max abs err: hi %.2e, hi+lo %.2e" % (e_hi, e_sum))
    assert e_sum < e_hi and e_sum < 8e-3        # what remains is P's bf16 rounding inside the kernel


def test_arena_optimizer_kernels():
    """adamw_step with a bf16 gradient buffer and a bf16 shadow, pack_bf16, sumsq over bf16."""
    from kbner_b200 import ops
    torch.manual_seed(2)
    n, ns = 4096 * 33 + 8, 4096 * 10
    p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda") * 1e-2
    m = torch.randn(n, device="cuda") * 1e-3; v = torch.rand(n, device="cuda") * 1e-4
    gb = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    ops.pack_bf16(g, gb, scale=0.5)
    assert torch.equal(gb, (g * 0.5).bfloat16())
    s1, s2 = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    ops.sumsq(gb, s1); ops.sumsq(gb.float(), s2)
    assert abs(s1.item() - s2.item()) <= 1e-5 * s2.item()
    lr, b1, b2, eps, wd, step, gs = 3e-4, 0.9, 0.999, 1e-6, 0.01, 3, 0.25
    pd, gd, md, vd = p.double(), gb.double() * gs, m.double(), v.double()
    md = b1 * md + (1 - b1) * gd
    vd = b2 * vd + (1 - b2) * gd * gd
    pd = pd - lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step) * md / (vd.sqrt() + eps)
    pd = pd - lr * wd * pd
    shadow = torch.zeros(ns, dtype=torch.bfloat16, device="cuda")
    ops.adamw_step(p, gb, m, v, lr, b1, b2, eps, wd, step, gscale_host=gs, shadow=shadow)
    assert (p.double() - pd).abs().max().item() < 1e-6 and (m.double() - md).abs().max().item() < 1e-8
    assert torch.equal(shadow, p[:ns].bfloat16())


def test_backward_parity_at_configs2_shape():
    """BASELINE configs[2]: 24 layers, 8 x 512 sub-tokens.  loss.backward() through the hand-written backward (split-K wgrad
    at M = 4096 included) against torch autograd through the fp32 oracle, fed with the same d(loss)/d(logits).  Run twice:
    with and without the attention output's rounding residual in the backward's D = rowsum(dO * O)."""
    import encoder_oracle as E
    from kbner_b200.data import BatchedData, Sentence
    tagger, emb, params, ocfg = _large(seed=21, small_vocab=False)
    emb.fine_tune, emb.static_embeddings = True, False
    tagger.train()
    emb.train()
    tagger.use_word_dropout = 0.0
    rnd = random.Random(4)
    sents = [Sentence(tokens=["w%03x" % rnd.randrange(4096) for _ in range(510)]) for _ in range(8)]
    d = tagger.tag_dictionary
    rng = np.random.RandomState(1)
    legal = [i for i in range(len(d)) if i not in (0, tagger.x_idx, tagger.start_idx, tagger.stop_idx)]
    for s in sents:
        for tok in s.tokens:
            tok.add_tag("ner", d.get_item_for_index(legal[rng.randint(len(legal))]))
    batch = BatchedData(sents)
    enc = emb.model
    enc.ensure_arena()
    ids, key_len, row_of, first_idx, lengths, S = emb.build_batch(batch)
    assert tuple(ids.shape) == (8, 512)
    own = dict(enc.named_parameters())
    report = {}
    op = None
    for residual in (False, True):
        enc._ctx_residual = residual
        enc.arena.zero_grad()
        batch.features = {}
        feats = tagger.forward(batch)
        feats.retain_grad()
        loss = tagger._calculate_loss(feats, batch, tagger.mask)
        loss.backward()
        if op is None:                     # the oracle's gradients, once (identical d_logits both times: same forward)
            d_logits = feats.grad.detach().clone()
            op = {k: v.cuda().clone().requires_grad_(True) for k, v in params.items()}
            hidden = E.encoder_forward(op, ids.long().cuda(), key_len.long().cuda(), ocfg)
            flat = hidden.reshape(-1, hidden.shape[-1])
            idx = row_of.long()[:, None] * S + first_idx.long().clamp(min=0)
            x = flat[idx.cuda()] * (first_idx >= 0).float().cuda()[..., None]
            (x @ tagger.linear.weight.detach().t()).backward(d_logits)
            del hidden, flat, x
        rel, tiny, num, den = {}, [], 0.0, 0.0
        for name, ref in op.items():
            if ref.grad is None:
                continue
            denom = ref.grad.norm().item()
            got = own[name].grad
            num += float(((got - ref.grad) ** 2).sum())
            den += denom ** 2
            if denom < 1e-7 * max(1.0, ref.numel() ** 0.5):
                tiny.append(name)                  # e.g. key biases: the exact gradient is 0 (softmax is shift-invariant)
                continue
            rel[name] = ((got - ref.grad).norm() / denom).item()
        vals = sorted(rel.values())
        report[residual] = dict(n=len(vals), median=vals[len(vals) // 2], p90=vals[int(len(vals) * 0.9)], max=vals[-1],
                                whole_gradient=(num / den) ** 0.5, worst=sorted(rel.items(), key=lambda kv: -kv[1])[:3],
                                zero_gradient_tensors=len(tiny))
    print("configs[2] backward vs fp32 autograd, per-tensor rel-L2 {without / with the attention-output residual in D}:", report)
    r = report[True]
    # bf16 operands through 24 layers forward AND backward: bounds = the measured format cost with headroom, written next
    # to north_star's 1e-3 (which is stated for logits / loss, not for gradients)
    # measured: whole gradient 8.7e-3, median 8.9e-3, p90 1.1e-2, max 1.4e-2 (without the residual the top layer's
    # dW_q / dW_k sit at 1.1e-1: near-uniform attention makes dP - D a difference of nearly equal numbers)
    assert r["whole_gradient"] < 1.5e-2 and r["median"] < 1.5e-2 and r["p90"] < 2e-2 and r["max"] < 3e-2, report
    assert r["max"] < 0.5 * report[False]["max"], report


def test_hf_saved_model_hidden_states_match_transformers(tmp_path):
    """A directory written by transformers' own save_pretrained (safetensors) -> from_pretrained -> the kernels' hidden
    state against the SAME transformers module's forward (eager attention, fp32) on the GPU."""
    import transformers
    from kbner_b200.encoder import XLMRobertaEncoderB200
    torch.manual_seed(0)
    cfg = transformers.XLMRobertaConfig(vocab_size=300, hidden_size=256, num_hidden_layers=3, num_attention_heads=4,
                                        intermediate_size=512, max_position_embeddings=130, type_vocab_size=1, pad_token_id=1,
                                        layer_norm_eps=1e-5, attn_implementation="eager")
    hf = transformers.XLMRobertaModel(cfg).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.02)
    hf.save_pretrained(str(tmp_path / "hf"))
    enc = XLMRobertaEncoderB200.from_pretrained(str(tmp_path / "hf")).cuda()
    ids = torch.randint(3, 300, (3, 128))
    ids[:, 0] = 0
    lens = [128, 77, 20]
    mask = torch.zeros(3, 128, dtype=torch.long)
    for r, n in enumerate(lens):
        ids[r, n - 1] = 2
        ids[r, n:] = 0                      # the reference pads with id 0 (embeddings.py:3247-3251)
        mask[r, :n] = 1
    with torch.no_grad():
        want = hf.cuda()(input_ids=ids.cuda(), attention_mask=mask.cuda()).last_hidden_state
        res = {}
        for mode in ("bf16", "bf16x3"):
            enc.set_precision(mode)
            got = enc(ids.cuda(), attention_mask=mask.cuda())[0]
            res[mode] = max(_rel(got[r, :n], want[r, :n]) for r, n in enumerate(lens))
    print("vs transformers.XLMRobertaModel:", res)
    assert res["bf16"] < 1e-2 and res["bf16x3"] < NORTH_STAR_TOL, res


@pytest.mark.parametrize("clip", [None, 1.0])
def test_row_skipping_optimizer_matches_the_dense_passes(clip):
    """ParamArena.enable_row_skipping: clip norm / AdamW / zero_grad over the marked rows of an embedding table only.  Without
    clipping the parameters are BIT-identical to the dense passes (same arithmetic per element; untouched rows have g = m = v =
    0); with clipping the two norms differ in summation order only (<= 1 ulp of the coefficient)."""
    from kbner_b200 import ops
    from kbner_b200.encoder import ParamArena
    from kbner_b200.optim import FusedAdamW
    torch.manual_seed(4)
    V, H = 500, 256

    def make():
        g = torch.Generator(device="cuda").manual_seed(9)
        ps = [torch.nn.Parameter(torch.randn(40, 8, device="cuda", generator=g)),
              torch.nn.Parameter(torch.randn(V, H, device="cuda", generator=g) * 0.02),
              torch.nn.Parameter(torch.randn(16, device="cuda", generator=g))]
        ar = ParamArena(ps)
        return ps, ar, FusedAdamW([{"arena": ar, "lr": 1e-3}], max_grad_norm=clip)
    (pd, ad, od), (ps_, as_, os_) = make(), make()
    as_.enable_row_skipping(ps_[1])
    gen = torch.Generator(device="cuda").manual_seed(1)
    for step in range(6):
        ids = torch.randint(0, V, (3, 11), device="cuda", generator=gen).to(torch.int32)
        rows = torch.randn(33, H, device="cuda", generator=gen)
        g0, g2 = torch.randn(40, 8, device="cuda", generator=gen), torch.randn(16, device="cuda", generator=gen)
        for ps, ar in ((pd, ad), (ps_, as_)):
            ps[0].grad.copy_(g0)
            ps[2].grad.copy_(g2)
            ps[1].grad.index_add_(0, ids.view(-1).long(), rows)           # what embed_ln_bwd does: scatter-add per token
        as_.mark_touched(ids)
        if step == 4:                                                    # a step in which the table receives no gradient at all:
            for ps in (pd, ps_):                                         # rows touched earlier keep moving on their momentum
                ps[1].grad.zero_()
        od.step(grad_scale=0.5)
        os_.step(grad_scale=0.5)
        if clip is None:
            assert torch.equal(ad.flat, as_.flat), step
        else:
            assert torch.allclose(ad.flat, as_.flat, rtol=2e-6, atol=1e-9), step
        od.zero_grad()
        os_.zero_grad()
        assert float(as_.grad.abs().max()) == 0.0 and float(ad.grad.abs().max()) == 0.0
    touched = as_.row_table["touched"]
    assert 0 < int(touched.sum()) < V                                    # some rows were never embedded: they never moved
    never = ~touched.bool()
    assert torch.equal(ps_[1].detach()[never], pd[1].detach()[never])
