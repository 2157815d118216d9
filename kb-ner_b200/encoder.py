"""XLM-R(-large/-base) encoder forward on the sm_100a kernels.

Replaces the third-party ``transformers==3.0.0`` ``XLMRobertaModel.forward`` the reference calls at
``/root/reference/flair/embeddings.py:3269`` (construction :2951-2953).  Parameter names follow the HF
state dict (``embeddings.word_embeddings.weight`` ... ``encoder.layer.N.output.LayerNorm.bias``) so that
HF checkpoints load unchanged and the reference trainer's name-based LR groups
(``flair/trainers/finetune_trainer.py:552-553``) see the same names.

fp32 master parameters; bf16 compute copies of the Linear weights (Q|K|V fused into one [3H,H] matrix);
fp32 accumulation, fp32 pre-LayerNorm sums, fp32 LayerNorm / softmax statistics, bf16 activations.
Only the requested final hidden state is produced (the reference's ``torch.stack`` of all 25 layer
outputs, embeddings.py:3275, and the unused pooler are not computed -- SURVEY E7/E8).
"""
import torch

from . import ops


def _lib_note_launches(n):
    from . import _lib
    _lib.load().kbner_add_launches(int(n))


class _Holder(torch.nn.Module):
    """Bare container so parameter names nest like the HF module tree."""


def _lin(out_f, in_f, std=0.02):
    m = _Holder()
    m.weight = torch.nn.Parameter(torch.randn(out_f, in_f) * std)
    m.bias = torch.nn.Parameter(torch.zeros(out_f))
    return m


def _ln(h):
    m = _Holder()
    m.weight = torch.nn.Parameter(torch.ones(h))
    m.bias = torch.nn.Parameter(torch.zeros(h))
    return m


class EncoderConfig:
    def __init__(self, vocab_size=250002, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                 intermediate_size=4096, max_position_embeddings=514, layer_norm_eps=1e-5, pad_token_id=1,
                 type_vocab_size=1, name="xlm-roberta-large", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 **_unused):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.max_position_embeddings = max_position_embeddings
        self.layer_norm_eps = layer_norm_eps
        self.pad_token_id = pad_token_id
        self.type_vocab_size = type_vocab_size
        self.output_hidden_states = True
        self.name = name
        # transformers XLMRobertaConfig defaults; active only in the fine-tuning forward of a module in train() mode
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob

    @classmethod
    def xlmr_large(cls, **kw):
        return cls(**kw)

    @classmethod
    def xlmr_base(cls, **kw):
        return cls(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                   name="xlm-roberta-base", **kw)

    def to_dict(self):
        return dict(self.__dict__)


class XLMRobertaEncoderB200(torch.nn.Module):
    def __init__(self, config: EncoderConfig):
        super().__init__()
        c = self.config = config
        if c.hidden_size != c.num_attention_heads * 64:
            raise ValueError("attention kernel is built for head dim 64 (XLM-R base / large)")
        H, F = c.hidden_size, c.intermediate_size
        self.embeddings = _Holder()
        self.embeddings.word_embeddings = _Holder()
        self.embeddings.word_embeddings.weight = torch.nn.Parameter(torch.randn(c.vocab_size, H) * 0.02)
        self.embeddings.position_embeddings = _Holder()
        self.embeddings.position_embeddings.weight = torch.nn.Parameter(torch.randn(c.max_position_embeddings, H) * 0.02)
        self.embeddings.token_type_embeddings = _Holder()
        self.embeddings.token_type_embeddings.weight = torch.nn.Parameter(torch.randn(c.type_vocab_size, H) * 0.02)
        self.embeddings.LayerNorm = _ln(H)
        self.encoder = _Holder()
        self.encoder.layer = torch.nn.ModuleList()
        for _ in range(c.num_hidden_layers):
            lyr = _Holder()
            lyr.attention = _Holder()
            lyr.attention.self = _Holder()
            lyr.attention.self.query = _lin(H, H)
            lyr.attention.self.key = _lin(H, H)
            lyr.attention.self.value = _lin(H, H)
            lyr.attention.output = _Holder()
            lyr.attention.output.dense = _lin(H, H)
            lyr.attention.output.LayerNorm = _ln(H)
            lyr.intermediate = _Holder()
            lyr.intermediate.dense = _lin(F, H)
            lyr.output = _Holder()
            lyr.output.dense = _lin(H, F)
            lyr.output.LayerNorm = _ln(H)
            self.encoder.layer.append(lyr)
        self._compute = None          # bf16 / fused compute copies, built by sync_compute_weights()
        self._ws = {}
        self._graphs = {}
        self._tgraphs = {}
        import os
        self._use_graphs = os.environ.get("KBNER_GRAPHS", "1") != "0"
        self._fuse_ln = os.environ.get("KBNER_FUSE_LN", "1") != "0"

    # ---- checkpointing: only parameters travel; graphs, workspaces, compute copies and the arena are rebuilt ----
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_compute", "arena", "_drop_seed", "_drop_seed_buf"):
            state[k] = None
        for k in ("_ws", "_graphs", "_tgraphs"):
            state[k] = {}
        state["_compute_static"] = False
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        # parameters arrive as views of the saved arena storage: give each its own storage again
        for p in self.parameters():
            p.data = p.data.clone()

    # ---- weights ---------------------------------------------------------------------------------
    def load_hf_state_dict(self, sd):
        own = self.state_dict()
        sd = {k: v for k, v in sd.items() if k in own}
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError("missing encoder parameters: %s" % missing[:4])
        self.load_state_dict(sd)
        self._compute = None

    @torch.no_grad()
    def sync_compute_weights(self):
        """(Re)build the bf16 compute copies from the fp32 masters (after load / optimizer step)."""
        layers = []
        for lyr in self.encoder.layer:
            a = lyr.attention
            layers.append(dict(
                wqkv=torch.cat([a.self.query.weight, a.self.key.weight, a.self.value.weight], 0).bfloat16().contiguous(),
                bqkv=torch.cat([a.self.query.bias, a.self.key.bias, a.self.value.bias], 0).float().contiguous(),
                wo=a.output.dense.weight.bfloat16().contiguous(), bo=a.output.dense.bias.float().contiguous(),
                g1=a.output.LayerNorm.weight.float().contiguous(), b1=a.output.LayerNorm.bias.float().contiguous(),
                w1=lyr.intermediate.dense.weight.bfloat16().contiguous(), bi=lyr.intermediate.dense.bias.float().contiguous(),
                w2=lyr.output.dense.weight.bfloat16().contiguous(), b2=lyr.output.dense.bias.float().contiguous(),
                g2=lyr.output.LayerNorm.weight.float().contiguous(), bb2=lyr.output.LayerNorm.bias.float().contiguous()))
        self._compute = layers

    def _workspace(self, M, dev):
        key = (M, str(dev))
        ws = self._ws.get(key)
        if ws is None:
            H, F = self.config.hidden_size, self.config.intermediate_size
            bf, f32 = torch.bfloat16, torch.float32
            ws = dict(x0=torch.empty((M, H), dtype=bf, device=dev), x1=torch.empty((M, H), dtype=bf, device=dev),
                      qkv=torch.empty((M, 3 * H), dtype=bf, device=dev), ctx=torch.empty((M, H), dtype=bf, device=dev),
                      y=torch.empty((M, H), dtype=f32, device=dev), h=torch.empty((M, F), dtype=bf, device=dev))
            self._ws = {key: ws}      # keep one shape resident
        return ws

    # ---- forward -----------------------------------------------------------------------------------
    @torch.no_grad()
    def _forward_hidden_eager(self, ids, key_len):
        c = self.config
        R, S = ids.shape
        M = R * S
        ws = self._workspace(M, ids.device)
        e = self.embeddings
        x, xn = ws["x0"], ws["x1"]
        ops.embed_ln_fwd(ids, e.word_embeddings.weight, e.position_embeddings.weight,
                         e.token_type_embeddings.weight[0], e.LayerNorm.weight, e.LayerNorm.bias,
                         c.layer_norm_eps, c.pad_token_id, out=x)
        # attention-output and FFN-down projections: bias + residual + LayerNorm fused into the GEMM epilogue over a
        # thread-block cluster that owns full rows (csrc/gemm_ln_tcgen05.cu); hidden sizes it is not built for, or
        # KBNER_FUSE_LN=0, take the GEMM (fp32 out) + LayerNorm-with-bias-and-residual pair instead
        fuse = self._fuse_ln and c.hidden_size in (256, 512, 768, 1024)
        y, h, ctx = ws["y"], ws["h"], ws["ctx"]
        for w in self._compute:
            ops.gemm_bf16_tn(x, w["wqkv"], w["bqkv"], epilogue=ops.EPI_BIAS, out=ws["qkv"])
            ops.attention_fwd(ws["qkv"], key_len, R, S, c.num_attention_heads, out=ctx)
            if fuse:
                ops.gemm_ln(ctx, w["wo"], w["bo"], x, w["g1"], w["b1"], c.layer_norm_eps, out=xn)
            else:
                ops.gemm_bf16_tn(ctx, w["wo"], None, epilogue=ops.EPI_NONE_F32, out=y)
                ops.layernorm_fwd(y, w["g1"], w["b1"], c.layer_norm_eps, out=xn, bias=w["bo"], resid=x)
            ops.gemm_bf16_tn(xn, w["w1"], w["bi"], epilogue=ops.EPI_BIAS_GELU, out=h)
            if fuse:
                ops.gemm_ln(h, w["w2"], w["b2"], xn, w["g2"], w["bb2"], c.layer_norm_eps, out=x)
            else:
                ops.gemm_bf16_tn(h, w["w2"], None, epilogue=ops.EPI_NONE_F32, out=y)
                ops.layernorm_fwd(y, w["g2"], w["bb2"], c.layer_norm_eps, out=x, bias=w["b2"], resid=xn)
        return x

    @torch.no_grad()
    def forward_hidden(self, ids, key_len):
        """ids [R,S] int32 (cuda), key_len [R] int32 -> last hidden state [R*S, H] bf16.
        The returned tensor aliases an internal workspace buffer (valid until the next call).

        The 1 + 7*layers launches of one shape are captured into a CUDA graph on the second call with that shape and
        replayed afterwards (static id / length buffers, static workspace): the forward of a batch is one graph launch,
        which takes ~170 ctypes round trips per batch off the host's critical path.  KBNER_GRAPHS=0 disables it."""
        if self._compute is None:
            self.sync_compute_weights()
            self._graphs = {}
        R, S = ids.shape
        key = (R, S, str(ids.device), id(self._compute))
        st = self._graphs.get(key) if self._use_graphs else None
        if st is not None and st["graph"] is not None:
            st["ids"].copy_(ids, non_blocking=True)
            st["key_len"].copy_(key_len, non_blocking=True)
            st["graph"].replay()
            _lib_note_launches(st["launches"])
            return st["out"]
        if not self._use_graphs:
            return self._forward_hidden_eager(ids, key_len)
        if st is None:                                   # first call with this shape: eager (also warms every kernel up)
            self._graphs = {k: v for k, v in self._graphs.items() if k[3] == id(self._compute)}
            self._graphs[key] = {"graph": None, "calls": 1}
            return self._forward_hidden_eager(ids, key_len)
        # second call: capture
        from . import _lib
        st["ids"], st["key_len"] = ids.clone(), key_len.clone()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(g):
            st["out"] = self._forward_hidden_eager(st["ids"], st["key_len"])
        st["launches"] = _lib.launch_count() - l0
        st["graph"] = g
        g.replay()
        _lib_note_launches(st["launches"])
        return st["out"]

    def forward(self, input_ids, attention_mask=None, **_unused):
        """HF-like call: returns (sequence_output [R,S,H] fp32,) -- the contract used at embeddings.py:3269,
        restricted to the last layer."""
        ids = input_ids.to(torch.int32).contiguous()
        if attention_mask is None:
            key_len = torch.full((ids.shape[0],), ids.shape[1], dtype=torch.int32, device=ids.device)
        else:
            key_len = attention_mask.to(torch.int32).sum(1).to(torch.int32).contiguous()
        h = self.forward_hidden(ids, key_len)
        return (h.float().view(ids.shape[0], ids.shape[1], -1),)

    # the HF surface the reference touches: train.py:208-209,260-261; finetune_trainer.py:1297-1298
    def save_pretrained(self, path):
        import json
        import os
        os.makedirs(path, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(path, "pytorch_model.bin"))
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.config.to_dict(), f, indent=1)

    @classmethod
    def from_pretrained(cls, path, **kw):
        import json
        import os
        with open(os.path.join(path, "config.json")) as f:
            cfg = json.load(f)
        m = cls(EncoderConfig(**{**cfg, **kw}))
        m.load_hf_state_dict(torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu"))
        return m


# =====================================================================================================
# Fine-tuning path: flat parameter / gradient arenas, forward that keeps what the backward needs, and the
# hand-written backward (no autograd inside the encoder).  Semantics = what `loss.backward()` does to the
# transformers module in the reference (flair/trainers/finetune_trainer.py:939-957) with dropout disabled.
# =====================================================================================================
class ParamArena:
    """All parameters of a module as views into ONE flat fp32 buffer (and their .grad into another), in a caller-
    chosen order.  One buffer = one fused optimizer launch, one gradient-norm launch, contiguous NCCL buckets; the
    order lets Q|K|V weights (and biases) sit next to each other so the fused [3H,H] projection is a plain view."""

    def __init__(self, params):
        params = list(params)
        dev = params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]           # keep every view 16-byte aligned
        self.numel = sum(sizes)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.offsets = {}
        off = 0
        for p, n in zip(params, sizes):
            self.flat[off:off + p.numel()].copy_(p.detach().reshape(-1).float())
            p.data = self.flat[off:off + p.numel()].view(p.shape)
            p.grad = self.grad[off:off + p.numel()].view(p.shape)
            self.offsets[id(p)] = off
            off += n

    def view(self, p_first, shape, grad=False):
        """A [shape] view starting at parameter p_first (used for the fused Q|K|V weight / bias)."""
        off = self.offsets[id(p_first)]
        n = 1
        for s in shape:
            n *= s
        return (self.grad if grad else self.flat)[off:off + n].view(shape)

    def zero_grad(self):
        self.grad.zero_()


def _arena_order(enc):
    out = []
    for lyr in enc.encoder.layer:
        a = lyr.attention
        out += [a.self.query.weight, a.self.key.weight, a.self.value.weight,
                a.self.query.bias, a.self.key.bias, a.self.value.bias,
                a.output.dense.weight, a.output.dense.bias, a.output.LayerNorm.weight, a.output.LayerNorm.bias,
                lyr.intermediate.dense.weight, lyr.intermediate.dense.bias,
                lyr.output.dense.weight, lyr.output.dense.bias, lyr.output.LayerNorm.weight, lyr.output.LayerNorm.bias]
    e = enc.embeddings
    out += [e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight,
            e.LayerNorm.weight, e.LayerNorm.bias]
    return out


def _ensure_arena(self):
    if getattr(self, "arena", None) is None:
        self.arena = ParamArena(_arena_order(self))
        self._compute = None
    return self.arena


@torch.no_grad()
def _sync_compute_weights_arena(self):
    """bf16 compute copies straight from the arena (Q|K|V are one contiguous [3H,H] region: no concatenation).
    The copies live in STATIC buffers that are refreshed in place after every optimizer step, so CUDA graphs that
    captured their addresses stay valid."""
    ar = self.arena
    H = self.config.hidden_size
    if getattr(self, "_compute_static", False) and self._compute is not None:
        for lyr, w in zip(self.encoder.layer, self._compute):
            a = lyr.attention
            w["wqkv"].copy_(ar.view(a.self.query.weight, (3 * H, H)))
            w["wo"].copy_(a.output.dense.weight)
            w["w1"].copy_(lyr.intermediate.dense.weight)
            w["w2"].copy_(lyr.output.dense.weight)
        return
    layers = []
    for lyr in self.encoder.layer:
        a = lyr.attention
        layers.append(dict(
            wqkv=ar.view(a.self.query.weight, (3 * H, H)).bfloat16(), bqkv=ar.view(a.self.query.bias, (3 * H,)),
            wo=a.output.dense.weight.bfloat16(), bo=a.output.dense.bias.data,
            g1=a.output.LayerNorm.weight.data, b1=a.output.LayerNorm.bias.data,
            w1=lyr.intermediate.dense.weight.bfloat16(), bi=lyr.intermediate.dense.bias.data,
            w2=lyr.output.dense.weight.bfloat16(), b2=lyr.output.dense.bias.data,
            g2=lyr.output.LayerNorm.weight.data, bb2=lyr.output.LayerNorm.bias.data))
    self._compute = layers
    self._compute_static = True
    self._graphs = {}
    self._tgraphs = {}


def _dropout_sites(self, li):
    """(attention-probability, attention-output, FFN-output) dropout descriptors of layer li, or Nones when off."""
    seed = getattr(self, "_drop_seed", None)
    if seed is None:
        return None, None, None
    c = self.config
    pa, ph = c.attention_probs_dropout_prob, c.hidden_dropout_prob
    return ((seed, 4 * li + 0, pa) if pa > 0 else None, (seed, 4 * li + 1, ph) if ph > 0 else None,
            (seed, 4 * li + 2, ph) if ph > 0 else None)


@torch.no_grad()
def _forward_train_eager(self, ids, key_len):
    c = self.config
    R, S = ids.shape
    M, H, F = R * S, c.hidden_size, c.intermediate_size
    dev = ids.device
    bf = torch.bfloat16
    e = self.embeddings
    x = ops.embed_ln_fwd(ids, e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight[0],
                         e.LayerNorm.weight, e.LayerNorm.bias, c.layer_norm_eps, c.pad_token_id)
    drop_on = getattr(self, "_drop_seed", None) is not None
    if drop_on and c.hidden_dropout_prob > 0:
        ops.dropout_apply(x, (self._drop_seed, 4 * len(self._compute), c.hidden_dropout_prob))
    saved = {"ids": ids, "key_len": key_len, "R": R, "S": S, "layers": [], "dropout": drop_on}
    for li, w in enumerate(self._compute):
        d_attn, d_h1, d_h2 = _dropout_sites(self, li)
        qkv = ops.gemm_bf16(x, w["wqkv"], M, 3 * H, H, ops.EPI_BIAS, bias=w["bqkv"])
        ctx, lse = ops.attention_fwd(qkv, key_len, R, S, c.num_attention_heads, want_lse=True, drop=d_attn)
        y1 = ops.gemm_bf16(ctx, w["wo"], M, H, H, ops.EPI_NONE_F32)
        x1, mean1, rstd1 = ops.layernorm_fwd(y1, w["g1"], w["b1"], c.layer_norm_eps, save_stats=True, bias=w["bo"], resid=x,
                                             drop=d_h1)
        hpre = torch.empty((M, F), dtype=bf, device=dev)
        h = ops.gemm_bf16(x1, w["w1"], M, F, H, ops.EPI_BIAS_GELU, bias=w["bi"], aux_out=hpre)
        y2 = ops.gemm_bf16(h, w["w2"], M, H, F, ops.EPI_NONE_F32)
        xo, mean2, rstd2 = ops.layernorm_fwd(y2, w["g2"], w["bb2"], c.layer_norm_eps, save_stats=True, bias=w["b2"], resid=x1,
                                             drop=d_h2)
        saved["layers"].append((x, qkv, lse, ctx, y1, mean1, rstd1, x1, hpre, h, y2, mean2, rstd2))
        x = xo
    return x, saved


@torch.no_grad()
def _forward_train(self, ids, key_len):
    """Forward that keeps the activations the backward needs.  Returns (hidden [R*S,H] bf16, saved).
    Like forward_hidden, the launches of one (R, S) shape are captured into a CUDA graph on the second call and
    replayed afterwards; the saved activations then live in the graph's private pool (valid until the next replay)."""
    _ensure_arena(self)
    if self._compute is None or not getattr(self, "_compute_static", False):
        self._compute = None
        _sync_compute_weights_arena(self)
    # dropout (transformers: hidden 0.1, attention probabilities 0.1) is active in train() mode, like torch.nn.Dropout.
    # The two seed words live in a static device buffer (the captured graphs read it) and are refreshed from torch's CUDA
    # generator before every forward, so torch.manual_seed() makes a run reproducible; the backward reuses them.
    c = self.config
    drop_on = bool(self.training and (c.hidden_dropout_prob > 0 or c.attention_probs_dropout_prob > 0))
    if drop_on:
        if getattr(self, "_drop_seed_buf", None) is None or self._drop_seed_buf.device != ids.device:
            self._drop_seed_buf = torch.zeros(2, dtype=torch.int32, device=ids.device)
        self._drop_seed_buf.copy_(torch.randint(0, 2 ** 31 - 1, (2,), dtype=torch.int32, device=ids.device))
        self._drop_seed = self._drop_seed_buf
    else:
        self._drop_seed = None
    if not self._use_graphs:
        return _forward_train_eager(self, ids, key_len)
    R, S = ids.shape
    key = (R, S, str(ids.device), drop_on)
    st = self._tgraphs.get(key)
    if st is None:
        self._tgraphs[key] = {"fwd": None, "bwd": None}
        return _forward_train_eager(self, ids, key_len)
    if st["fwd"] is None:
        from . import _lib
        st["ids"], st["key_len"] = ids.clone(), key_len.clone()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(g):
            st["out"], st["saved"] = _forward_train_eager(self, st["ids"], st["key_len"])
        st["fwd_launches"] = _lib.launch_count() - l0
        st["saved"]["graph_state"] = st
        st["fwd"] = g
        st["pool"] = g.pool()
    else:
        st["ids"].copy_(ids, non_blocking=True)
        st["key_len"].copy_(key_len, non_blocking=True)
    st["fwd"].replay()
    _lib_note_launches(st["fwd_launches"])
    return st["out"], st["saved"]


@torch.no_grad()
def _backward_workspace(self, saved, dev):
    c = self.config
    R, S = saved["R"], saved["S"]
    return (torch.empty((R, c.num_attention_heads, S), dtype=torch.float32, device=dev),
            torch.empty((R * S, c.hidden_size), dtype=torch.float32, device=dev))


def _backward_layers(self, saved, dout, dres, li_hi, li_lo, ws):
    """Backward through layers li_hi .. li_lo (descending).  `dout` (fp32) + `dres` (bf16, or None) is the gradient w.r.t.
    the output of layer li_hi: the dgrad GEMMs keep a plain fp32 epilogue and the gradient that arrives over the residual
    connection is added by the LayerNorm backward pass.  Returns the pair for layer li_lo - 1."""
    c = self.config
    ar = self.arena
    R, S = saved["R"], saved["S"]
    M, H, F = R * S, c.hidden_size, c.intermediate_size
    heads = c.num_attention_heads
    key_len = saved["key_len"]
    for li in range(li_hi, li_lo - 1, -1):
        lyr, w = self.encoder.layer[li], self._compute[li]
        a = lyr.attention
        d_attn, d_h1, d_h2 = _dropout_sites(self, li) if saved.get("dropout") else (None, None, None)
        x, qkv, lse, ctx, y1, mean1, rstd1, x1, hpre, h, y2, mean2, rstd2 = saved["layers"][li]
        # ---- FFN block ------------------------------------------------------------------------------
        dz2 = ops.layernorm_bwd(y2, dout, w["g2"], mean2, rstd2, lyr.output.LayerNorm.weight.grad, lyr.output.LayerNorm.bias.grad,
                                dxsum=lyr.output.dense.bias.grad, bias=w["b2"], resid=x1, dres=dres, drop=d_h2)
        dz2, dy2 = dz2 if isinstance(dz2, tuple) else (dz2, dz2)      # (residual path, through the dropout mask)
        ops.gemm_bf16(dy2, h, H, F, M, ops.EPI_ACCUM_F32, out=lyr.output.dense.weight.grad, a_mn=True, b_mn=True)
        dhpre = ops.gemm_bf16(dy2, w["w2"], M, F, H, ops.EPI_DGELU_BF16, aux=hpre, b_mn=True)
        ops.colsum_bf16(dhpre, lyr.intermediate.dense.bias.grad)
        ops.gemm_bf16(dhpre, x1, F, H, M, ops.EPI_ACCUM_F32, out=lyr.intermediate.dense.weight.grad, a_mn=True, b_mn=True)
        dx1 = ops.gemm_bf16(dhpre, w["w1"], M, H, F, ops.EPI_NONE_F32, b_mn=True)
        # ---- attention block ------------------------------------------------------------------------
        dz1 = ops.layernorm_bwd(y1, dx1, w["g1"], mean1, rstd1, a.output.LayerNorm.weight.grad, a.output.LayerNorm.bias.grad,
                                dxsum=a.output.dense.bias.grad, bias=w["bo"], resid=x, dres=dz2, drop=d_h1)
        dz1, dy1 = dz1 if isinstance(dz1, tuple) else (dz1, dz1)
        ops.gemm_bf16(dy1, ctx, H, H, M, ops.EPI_ACCUM_F32, out=a.output.dense.weight.grad, a_mn=True, b_mn=True)
        dctx = ops.gemm_bf16(dy1, w["wo"], M, H, H, ops.EPI_BIAS, b_mn=True)
        dqkv = ops.attention_bwd(qkv, ctx, dctx, lse, key_len, R, S, heads, workspace=ws, drop=d_attn)
        ops.colsum_bf16(dqkv, ar.view(a.self.query.bias, (3 * H,), grad=True))
        ops.gemm_bf16(dqkv, x, 3 * H, H, M, ops.EPI_ACCUM_F32, out=ar.view(a.self.query.weight, (3 * H, H), grad=True),
                      a_mn=True, b_mn=True)
        if li > 0:
            dout = ops.gemm_bf16(dqkv, w["wqkv"], M, H, 3 * H, ops.EPI_NONE_F32, b_mn=True)
            dres = dz1
        else:      # the embedding backward takes one fp32 tensor: let this last dgrad add the residual gradient itself
            dout = ops.gemm_bf16(dqkv, w["wqkv"], M, H, 3 * H, ops.EPI_BIAS_RESID_F32, aux=dz1, b_mn=True)
    return dout, dres


def _backward_embed(self, saved, dout):
    c = self.config
    if saved.get("dropout") and c.hidden_dropout_prob > 0:
        ops.dropout_apply(dout, (self._drop_seed, 4 * len(self._compute), c.hidden_dropout_prob))
    e = self.embeddings
    ops.embed_ln_bwd(saved["ids"], e.word_embeddings.weight.data, e.position_embeddings.weight.data,
                     e.token_type_embeddings.weight.data[0], e.LayerNorm.weight.data, c.layer_norm_eps, c.pad_token_id,
                     dout, e.word_embeddings.weight.grad, e.position_embeddings.weight.grad,
                     e.token_type_embeddings.weight.grad[0], e.LayerNorm.weight.grad, e.LayerNorm.bias.grad)




def _backward_eager(self, saved, dout):
    ws = _backward_workspace(self, saved, dout.device)
    dout, _ = _backward_layers(self, saved, dout, None, len(self.encoder.layer) - 1, 0, ws)
    _backward_embed(self, saved, dout)


def _chunk_plan(self, n_chunks=4):
    """Layer chunks of the backward pass (descending) with the arena slice each one finalises: the arena is laid out
    layer 0 .. layer L-1, embeddings (_arena_order), so a chunk of consecutive layers owns one contiguous slice."""
    L = len(self.encoder.layer)
    n = max(1, min(n_chunks, L))
    bounds = [(L * k) // n for k in range(n + 1)]
    ar = self.arena

    def start(li):
        if li >= L:
            return ar.offsets[id(self.embeddings.word_embeddings.weight)]
        return ar.offsets[id(self.encoder.layer[li].attention.self.query.weight)]
    plan = [(bounds[k + 1] - 1, bounds[k], start(bounds[k]), start(bounds[k + 1])) for k in reversed(range(n))]
    return plan, (start(L), ar.numel)


@torch.no_grad()
def _backward_chunked(self, saved, dout, sync):
    """The backward pass in layer chunks; after each chunk `sync(lo, hi)` is called with the slice of the gradient arena
    that chunk has finalised, so the caller can start its all-reduce (async, NCCL's own stream) while the next chunk
    runs.  Used on the last gradient-accumulation micro-step of a multi-GPU run; chunks are captured / replayed as CUDA
    graphs like the one-piece backward (own pool; the tensors carried from chunk to chunk stay referenced)."""
    plan, emb_slice = _chunk_plan(self)
    st = saved.get("graph_state")
    if st is None:
        ws = _backward_workspace(self, saved, dout.device)
        dres = None
        for k, (hi, lo, a, b) in enumerate(plan):
            dout, dres = _backward_layers(self, saved, dout, dres, hi, lo, ws)
            if k == len(plan) - 1:
                _backward_embed(self, saved, dout)
            sync(a, b)
        sync(*emb_slice)
        return
    from . import _lib
    if st.get("bwd_chunks") is None:
        st["dout_c"] = dout.clone()
        st["bwd_ws"] = _backward_workspace(self, saved, dout.device)
        torch.cuda.synchronize()
        pool = torch.cuda.graph_pool_handle()
        chunks, cur, dres = [], st["dout_c"], None
        for k, (hi, lo, a, b) in enumerate(plan):
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(g, pool=pool):
                cur, dres = _backward_layers(self, saved, cur, dres, hi, lo, st["bwd_ws"])
                if k == len(plan) - 1:
                    _backward_embed(self, saved, cur)
            chunks.append((g, _lib.launch_count() - l0, cur, dres))
        st["bwd_chunks"] = chunks
    else:
        st["dout_c"].copy_(dout, non_blocking=True)
    for (g, n, _, _), (hi, lo, a, b) in zip(st["bwd_chunks"], plan):
        g.replay()
        _lib_note_launches(n)
        sync(a, b)
    sync(*emb_slice)


@torch.no_grad()
def _backward(self, saved, dout):
    """dout: gradient w.r.t. the last hidden state, fp32 [R*S, H].  Accumulates into the gradient arena.
    When `saved` came from a replayed forward graph, the backward launches are captured / replayed as a graph too
    (same pool; the activations, the arena and dout's static copy keep their addresses)."""
    sync = getattr(self, "_grad_sync", None)
    if sync is not None:
        return _backward_chunked(self, saved, dout, sync)
    st = saved.get("graph_state")
    if st is None:
        return _backward_eager(self, saved, dout)
    if st["bwd"] is None:
        from . import _lib
        st["dout"] = dout.clone()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(g, pool=st["pool"]):
            _backward_eager(self, saved, st["dout"])
        st["bwd_launches"] = _lib.launch_count() - l0
        st["bwd"] = g
    else:
        st["dout"].copy_(dout, non_blocking=True)
    st["bwd"].replay()
    _lib_note_launches(st["bwd_launches"])


XLMRobertaEncoderB200.ensure_arena = _ensure_arena
XLMRobertaEncoderB200.forward_train = _forward_train
XLMRobertaEncoderB200.backward = _backward
XLMRobertaEncoderB200.sync_compute_weights_arena = _sync_compute_weights_arena
