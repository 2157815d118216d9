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

Inference precision (``encoder.precision`` / ``KBNER_PRECISION``), measured against the fp32 oracle on 24 layers:
  ``"bf16"``        (default, fastest) bf16 operands and bf16 hidden state between kernels     ~9.7e-3 rel-L2
  ``"bf16-res32"``  fp32 residual stream, bf16 operands (unfused LayerNorm)                    ~7.4e-3
  ``"bf16x3"``      fp32 residual stream + every GEMM operand as a bf16 pair hi+lo, one K-concatenated tcgen05 GEMM
                    per projection (3x the tensor work): the mode that meets BASELINE.json's 1e-3 on logits ~2e-4
"""
import collections
import os

import torch

from . import ops


PRECISIONS = ("bf16", "bf16-res32", "bf16x3")


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
        self._gen = 0                 # bumped whenever the compute copies are REPLACED (captured graphs die with them)
        self._ws = {}
        # captured CUDA graphs per (R, S) shape, least-recently-used first.  A training graph pins every saved activation
        # of its shape (~0.9 MB per token for XLM-R-large), so both caches are bounded: a corpus with hundreds of distinct
        # (R, S) keeps the most recent few and replays the rest eagerly until they come back.
        self._graphs = collections.OrderedDict()
        self._tgraphs = collections.OrderedDict()
        self._graph_cap = max(1, int(os.environ.get("KBNER_GRAPH_CACHE", "8")))
        self._tgraph_cap = max(1, int(os.environ.get("KBNER_TRAIN_GRAPH_CACHE", "3")))
        self._use_graphs = os.environ.get("KBNER_GRAPHS", "1") != "0"
        self._fuse_ln = os.environ.get("KBNER_FUSE_LN", "1") != "0"
        # fine-tuning keeps the attention output's rounding residual for the backward's D = rowsum(dO * O) (ops.attention_bwd)
        self._ctx_residual = os.environ.get("KBNER_CTX_RESIDUAL", "1") != "0"
        self.precision = os.environ.get("KBNER_PRECISION", "bf16")
        if self.precision not in PRECISIONS:
            raise ValueError("KBNER_PRECISION must be one of %s" % (PRECISIONS,))

    # ---- checkpointing: only parameters travel; graphs, workspaces, compute copies and the arena are rebuilt ----
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_compute", "arena", "_drop_seed", "_drop_seed_buf", "_ids_hook", "_grad_sync"):
            state[k] = None
        state["_ws"] = {}
        state["_graphs"] = collections.OrderedDict()
        state["_tgraphs"] = collections.OrderedDict()
        state["_compute_static"] = False
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.__dict__.setdefault("_gen", 0)
        self.__dict__.setdefault("precision", "bf16")
        self.__dict__.setdefault("_graph_cap", 8)
        self.__dict__.setdefault("_tgraph_cap", 3)
        self.__dict__.setdefault("_ctx_residual", True)
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
        bad = [(k, tuple(v.shape), tuple(own[k].shape)) for k, v in sd.items() if tuple(v.shape) != tuple(own[k].shape)]
        if bad:
            raise ValueError("encoder parameter shapes differ from the config: %s" % bad[:3])
        self.load_state_dict({k: v.to(torch.float32) for k, v in sd.items()})
        self._compute = None

    def set_precision(self, precision):
        """Switch the inference arithmetic (module docstring); drops the compute copies and every captured graph."""
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % (PRECISIONS,))
        if precision != self.precision:
            self.precision = precision
            self._compute = None
            self._drop_graphs()

    def _drop_graphs(self):
        self._graphs.clear()
        self._tgraphs.clear()
        self._ws = {}
        self._gen += 1

    @torch.no_grad()
    def sync_compute_weights(self):
        """(Re)build the bf16 compute copies from the fp32 masters (after load / optimizer step).  The copies are NEW
        tensors, so every graph that captured the old addresses is dropped here."""
        x3 = self.precision == "bf16x3"

        def w(t):
            t = t.detach().float()
            hi = t.bfloat16()
            if not x3:
                return hi.contiguous()
            lo = (t - hi.float()).bfloat16()
            return torch.cat([hi, hi, lo], 1).contiguous()      # [out, 3*in]: pairs with activation rows [ hi | lo | hi ]
        layers = []
        for lyr in self.encoder.layer:
            a = lyr.attention
            layers.append(dict(
                wqkv=w(torch.cat([a.self.query.weight, a.self.key.weight, a.self.value.weight], 0)),
                bqkv=torch.cat([a.self.query.bias, a.self.key.bias, a.self.value.bias], 0).float().contiguous(),
                wo=w(a.output.dense.weight), bo=a.output.dense.bias.float().contiguous(),
                g1=a.output.LayerNorm.weight.float().contiguous(), b1=a.output.LayerNorm.bias.float().contiguous(),
                w1=w(lyr.intermediate.dense.weight), bi=lyr.intermediate.dense.bias.float().contiguous(),
                w2=w(lyr.output.dense.weight), b2=lyr.output.dense.bias.float().contiguous(),
                g2=lyr.output.LayerNorm.weight.float().contiguous(), bb2=lyr.output.LayerNorm.bias.float().contiguous()))
        self._compute = layers
        self._compute_static = False
        self._drop_graphs()

    def _new_workspace(self, M, dev):
        H, F = self.config.hidden_size, self.config.intermediate_size
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        if self.precision == "bf16":
            ws = dict(x0=e((M, H), bf), x1=e((M, H), bf), qkv=e((M, 3 * H), bf), ctx=e((M, H), bf), y=e((M, H), f32),
                      h=e((M, F), bf))
            if self._fuse_ln and H in (256, 512, 768, 1024):
                ws["ln_ws"] = ops.gemm_ln_workspace(M, H, dev)    # LayerNorm statistics exchange of the fused kernel
            return ws
        k = 3 if self.precision == "bf16x3" else 1
        ws = dict(x=e((M, k * H), bf), r0=e((M, H), f32), r1=e((M, H), f32), qkv=e((M, 3 * H), bf), ctx=e((M, k * H), bf),
                  y=e((M, H), f32), h=e((M, k * F), bf))
        if k == 3:
            ws["yf"] = e((M, F), f32)
        return ws

    def _workspace(self, M, dev):
        """Workspace of the EAGER path: one shape resident.  A captured graph owns a workspace of its own (st["ws"]): these
        buffers go back to the allocator when another shape arrives, and a graph must never write to freed memory."""
        key = (M, str(dev), self.precision)
        ws = self._ws.get(key)
        if ws is None:
            ws = self._new_workspace(M, dev)
            self._ws = {key: ws}
        return ws

    # ---- forward -----------------------------------------------------------------------------------
    @torch.no_grad()
    def _forward_hidden_eager(self, ids, key_len, ws=None):
        c = self.config
        R, S = ids.shape
        M = R * S
        if ws is None:
            ws = self._workspace(M, ids.device)
        if self.precision != "bf16":
            return self._forward_hidden_precise(ids, key_len, ws)
        e = self.embeddings
        x, xn = ws["x0"], ws["x1"]
        ops.embed_ln_fwd(ids, e.word_embeddings.weight, e.position_embeddings.weight,
                         e.token_type_embeddings.weight[0], e.LayerNorm.weight, e.LayerNorm.bias,
                         c.layer_norm_eps, c.pad_token_id, out=x)
        # attention-output and FFN-down projections: bias + residual + LayerNorm fused into the GEMM epilogue, the CTA pairs
        # of a row panel exchanging their LayerNorm partials (csrc/gemm_ln_grid_tcgen05.cu); hidden sizes it is not built for, or
        # KBNER_FUSE_LN=0, take the GEMM (fp32 out) + LayerNorm-with-bias-and-residual pair instead
        fuse = self._fuse_ln and c.hidden_size in (256, 512, 768, 1024)
        y, h, ctx = ws["y"], ws["h"], ws["ctx"]
        for w in self._compute:
            ops.gemm_bf16_tn(x, w["wqkv"], w["bqkv"], epilogue=ops.EPI_BIAS, out=ws["qkv"])
            ops.attention_fwd(ws["qkv"], key_len, R, S, c.num_attention_heads, out=ctx)
            if fuse:
                ops.gemm_ln(ctx, w["wo"], w["bo"], x, w["g1"], w["b1"], c.layer_norm_eps, out=xn, ws=ws["ln_ws"])
            else:
                ops.gemm_bf16_tn(ctx, w["wo"], None, epilogue=ops.EPI_NONE_F32, out=y)
                ops.layernorm_fwd(y, w["g1"], w["b1"], c.layer_norm_eps, out=xn, bias=w["bo"], resid=x)
            ops.gemm_bf16_tn(xn, w["w1"], w["bi"], epilogue=ops.EPI_BIAS_GELU, out=h)
            if fuse:
                ops.gemm_ln(h, w["w2"], w["b2"], xn, w["g2"], w["bb2"], c.layer_norm_eps, out=x, ws=ws["ln_ws"])
            else:
                ops.gemm_bf16_tn(h, w["w2"], None, epilogue=ops.EPI_NONE_F32, out=y)
                ops.layernorm_fwd(y, w["g2"], w["bb2"], c.layer_norm_eps, out=x, bias=w["b2"], resid=xn)
        return x

    @torch.no_grad()
    def _forward_hidden_precise(self, ids, key_len, ws):
        """"bf16-res32" / "bf16x3" (module docstring).  Same kernels for the contractions -- the tcgen05 GEMM over
        K-concatenated operands, the attention kernel writing its output as hi|lo|hi -- plus the fp32-residual LayerNorm
        and the bias + erf-GELU + split pass.  Q/K/V and the probabilities stay single bf16 (2.0e-4 of the 24-layer hidden
        state, scripts/bf16_ablation.py).  Returns the fp32 hidden state [R*S, H]."""
        c = self.config
        R, S = ids.shape
        x3 = self.precision == "bf16x3"
        e = self.embeddings
        x, r_in, r_out = ws["x"], ws["r0"], ws["r1"]
        y, h, ctx, qkv = ws["y"], ws["h"], ws["ctx"], ws["qkv"]
        ops.embed_ln_fwd(ids, e.word_embeddings.weight, e.position_embeddings.weight,
                         e.token_type_embeddings.weight[0], e.LayerNorm.weight, e.LayerNorm.bias,
                         c.layer_norm_eps, c.pad_token_id, out=x, out32=r_in, split=x3)
        for w in self._compute:
            ops.gemm_bf16_tn(x, w["wqkv"], w["bqkv"], epilogue=ops.EPI_BIAS, out=qkv)
            ops.attention_fwd(qkv, key_len, R, S, c.num_attention_heads, out=ctx, split=x3)
            ops.gemm_bf16_tn(ctx, w["wo"], None, epilogue=ops.EPI_NONE_F32, out=y)
            ops.layernorm_fwd_res32(y, w["g1"], w["b1"], c.layer_norm_eps, out=x, out32=r_out, bias=w["bo"], resid=r_in,
                                    split=x3)
            r_in, r_out = r_out, r_in
            if x3:
                ops.gemm_bf16_tn(x, w["w1"], None, epilogue=ops.EPI_NONE_F32, out=ws["yf"])
                ops.bias_gelu_split(ws["yf"], w["bi"], h)
            else:
                ops.gemm_bf16_tn(x, w["w1"], w["bi"], epilogue=ops.EPI_BIAS_GELU, out=h)
            ops.gemm_bf16_tn(h, w["w2"], None, epilogue=ops.EPI_NONE_F32, out=y)
            ops.layernorm_fwd_res32(y, w["g2"], w["bb2"], c.layer_norm_eps, out=x, out32=r_out, bias=w["b2"], resid=r_in,
                                    split=x3)
            r_in, r_out = r_out, r_in
        return r_in

    @torch.no_grad()
    def forward_hidden(self, ids, key_len):
        """ids [R,S] int32 (cuda), key_len [R] int32 -> last hidden state [R*S, H] (bf16; fp32 in the precision modes).
        The returned tensor aliases an internal buffer (valid until the next call).

        The 1 + 5*layers launches of one shape are captured into a CUDA graph on the second call with that shape and
        replayed afterwards (static id / length buffers, a workspace the graph state owns): the forward of a batch is
        one graph launch, which takes ~125 ctypes round trips per batch off the host's critical path.  The cache keeps the
        KBNER_GRAPH_CACHE (8) most recently used shapes; KBNER_GRAPHS=0 disables capture."""
        if self._compute is None:
            if getattr(self, "arena", None) is not None and self.precision == "bf16":
                _sync_compute_weights_arena(self)
            else:
                self.sync_compute_weights()
        if not self._use_graphs:
            return self._forward_hidden_eager(ids, key_len)
        R, S = ids.shape
        key = (R, S, str(ids.device), self._gen, self.precision)
        st = self._graphs.get(key)
        if st is None:                                   # first call with this shape: eager (also warms every kernel up)
            self._graphs[key] = {"graph": None}
            while len(self._graphs) > self._graph_cap:
                self._graphs.popitem(last=False)
            return self._forward_hidden_eager(ids, key_len)
        self._graphs.move_to_end(key)
        if st["graph"] is not None:
            st["ids"].copy_(ids, non_blocking=True)
            st["key_len"].copy_(key_len, non_blocking=True)
            st["graph"].replay()
            _lib_note_launches(st["launches"])
            return st["out"]
        # second call: capture
        from . import _lib
        st["ids"], st["key_len"] = ids.clone(), key_len.clone()
        st["ws"] = self._new_workspace(R * S, ids.device)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(g):
            st["out"] = self._forward_hidden_eager(st["ids"], st["key_len"], st["ws"])
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
    def save_pretrained(self, path, safe_serialization=True):
        """A directory `transformers.XLMRobertaModel.from_pretrained` can read (and from_pretrained below):
        config.json in transformers' layout + model.safetensors (safe_serialization=False: pytorch_model.bin, the only
        format the reference's transformers 3.0.0 knows)."""
        import json
        from .checkpoint_compat import hf_config_dict
        os.makedirs(path, exist_ok=True)
        sd = {k: v.detach().to("cpu").contiguous().clone() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(path, "pytorch_model.bin"))
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(hf_config_dict(self.config), f, indent=1)

    @classmethod
    def from_pretrained(cls, path, **kw):
        """Replaces AutoModel.from_pretrained(model, config=config) for a LOCAL directory (flair/embeddings.py:2951-2953):
        config.json + model.safetensors / pytorch_model.bin as transformers (any version) or this package wrote them, with or
        without a task-model prefix (`roberta.`); pooler / LM-head tensors are ignored, a missing encoder tensor raises."""
        from .checkpoint_compat import config_from_hf_json, normalize_hf_state_dict, read_weight_files
        m = cls(config_from_hf_json(path, **kw))
        m.load_hf_state_dict(normalize_hf_state_dict(read_weight_files(path)))
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
        sizes = [(p.numel() + 7) // 8 * 8 for p in params]           # every view 16-byte aligned, in fp32 and in bf16
        self.numel = sum(sizes)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.shadow = None            # bf16 copy of flat[:shadow.numel()] (the tensor-core operands), see make_shadow
        self.shadow_fresh = False     # set by FusedAdamW.step when its launch has just rewritten the shadow
        self.row_table = None         # row-sparse optimizer passes over one [V,H] table, see enable_row_skipping
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

    def make_shadow(self, n):
        """bf16 shadow of the first n parameters, same offsets as the fp32 arena.  FusedAdamW writes it in the optimizer
        launch; refresh_shadow() is the explicit fp32 -> bf16 pass for every other way the masters can change."""
        n = (int(n) + 7) // 8 * 8
        if self.shadow is None or self.shadow.numel() != n:
            self.shadow = torch.empty(n, dtype=torch.bfloat16, device=self.flat.device)
            self.shadow_fresh = False
        return self.shadow

    def refresh_shadow(self):
        if self.shadow is not None:
            ops.pack_bf16(self.flat[:self.shadow.numel()], self.shadow)
        self.shadow_fresh = False

    def shadow_view(self, p_first, shape):
        off = self.offsets[id(p_first)]
        n = 1
        for s_ in shape:
            n *= s_
        return self.shadow[off:off + n].view(shape)

    def enable_row_skipping(self, table):
        """The optimizer passes (clip norm, AdamW, zero_grad) over `table` ([V,H] parameter of this arena) visit only rows
        marked in `touched` -- rows some sentence has embedded since training began (mark_touched).  Unmarked rows have
        g = m = v = 0 and AdamW (weight decay 0) leaves them unchanged, so the result is bit-identical to the dense passes.
        Valid while every gradient that reaches the table comes with its ids: single GPU, or the sparse row exchange."""
        V, H = table.shape
        lo = self.offsets[id(table)]
        self.row_table = dict(lo=lo, hi=lo + (V * H + 7) // 8 * 8, V=V, H=H,
                              touched=torch.zeros(V, dtype=torch.uint8, device=self.flat.device))
        return self.row_table

    def mark_touched(self, ids):
        if self.row_table is not None:
            ops.mark_rows(ids, self.row_table["touched"])

    def table_view(self, buf):
        rt = self.row_table
        return buf[rt["lo"]:rt["lo"] + rt["V"] * rt["H"]].view(rt["V"], rt["H"])

    def zero_grad(self):
        rt = self.row_table
        if rt is None:
            self.grad.zero_()
            return
        self.grad[:rt["lo"]].zero_()
        ops.zero_rows(self.table_view(self.grad), rt["touched"])
        self.grad[rt["hi"]:].zero_()


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
    """bf16 compute copies as VIEWS of the arena's bf16 shadow (same offsets as the fp32 masters; Q|K|V are one contiguous
    [3H,H] region: no concatenation).  The shadow is a static buffer: CUDA graphs that captured its addresses stay valid.
    FusedAdamW.step rewrites it inside the optimizer launch (and marks it fresh), so the call that follows an optimizer
    step costs nothing; any other change of the masters is picked up by the one fp32 -> bf16 pass here."""
    ar = self.arena
    H = self.config.hidden_size
    if getattr(self, "_compute_static", False) and self._compute is not None and ar.shadow is not None:
        if ar.shadow_fresh:
            ar.shadow_fresh = False
        else:
            ar.refresh_shadow()
        return
    ar.make_shadow(ar.offsets[id(self.embeddings.word_embeddings.weight)])      # the encoder layers (_arena_order)
    ar.refresh_shadow()
    layers = []
    for lyr in self.encoder.layer:
        a = lyr.attention
        layers.append(dict(
            wqkv=ar.shadow_view(a.self.query.weight, (3 * H, H)), bqkv=ar.view(a.self.query.bias, (3 * H,)),
            wo=ar.shadow_view(a.output.dense.weight, tuple(a.output.dense.weight.shape)), bo=a.output.dense.bias.data,
            g1=a.output.LayerNorm.weight.data, b1=a.output.LayerNorm.bias.data,
            w1=ar.shadow_view(lyr.intermediate.dense.weight, tuple(lyr.intermediate.dense.weight.shape)),
            bi=lyr.intermediate.dense.bias.data,
            w2=ar.shadow_view(lyr.output.dense.weight, tuple(lyr.output.dense.weight.shape)), b2=lyr.output.dense.bias.data,
            g2=lyr.output.LayerNorm.weight.data, bb2=lyr.output.LayerNorm.bias.data))
    self._compute = layers
    self._compute_static = True
    self._drop_graphs()


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
        ctx_lo = torch.empty((M, H), dtype=bf, device=dev) if self._ctx_residual else None
        ctx, lse = ops.attention_fwd(qkv, key_len, R, S, c.num_attention_heads, want_lse=True, drop=d_attn, out_lo=ctx_lo)
        y1 = ops.gemm_bf16(ctx, w["wo"], M, H, H, ops.EPI_NONE_F32)
        x1, mean1, rstd1 = ops.layernorm_fwd(y1, w["g1"], w["b1"], c.layer_norm_eps, save_stats=True, bias=w["bo"], resid=x,
                                             drop=d_h1)
        hpre = torch.empty((M, F), dtype=bf, device=dev)
        h = ops.gemm_bf16(x1, w["w1"], M, F, H, ops.EPI_BIAS_GELU, bias=w["bi"], aux_out=hpre)
        y2 = ops.gemm_bf16(h, w["w2"], M, H, F, ops.EPI_NONE_F32)
        xo, mean2, rstd2 = ops.layernorm_fwd(y2, w["g2"], w["bb2"], c.layer_norm_eps, save_stats=True, bias=w["b2"], resid=x1,
                                             drop=d_h2)
        saved["layers"].append((x, qkv, lse, ctx, y1, mean1, rstd1, x1, hpre, h, y2, mean2, rstd2, ctx_lo))
        x = xo
    return x, saved


@torch.no_grad()
def _forward_train(self, ids, key_len):
    """Forward that keeps the activations the backward needs.  Returns (hidden [R*S,H] bf16, saved).
    Like forward_hidden, the launches of one (R, S) shape are captured into a CUDA graph on the second call and
    replayed afterwards; the saved activations then live in the graph's private pool (valid until the next replay)."""
    if self.precision != "bf16":
        raise NotImplementedError("fine-tuning runs in the bf16 arithmetic; precision=%r is an inference mode" % self.precision)
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
    hook = getattr(self, "_ids_hook", None)
    if hook is not None:              # the data-parallel exchange records which embedding rows this step touches
        hook(ids)
    self.arena.mark_touched(ids)      # row-sparse optimizer passes over the word-embedding table (no-op unless enabled)
    if not self._use_graphs:
        return _forward_train_eager(self, ids, key_len)
    R, S = ids.shape
    key = (R, S, str(ids.device), drop_on)
    st = self._tgraphs.get(key)
    if st is None:
        self._tgraphs[key] = {"fwd": None, "bwd": None}
        while len(self._tgraphs) > self._tgraph_cap:          # least recently used shape: its graph pool (saved activations) goes
            self._tgraphs.popitem(last=False)
        return _forward_train_eager(self, ids, key_len)
    self._tgraphs.move_to_end(key)
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


_WGRAD_GROUP = os.environ.get("KBNER_WGRAD_GROUP", "1") != "0"


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
        x, qkv, lse, ctx, y1, mean1, rstd1, x1, hpre, h, y2, mean2, rstd2, ctx_lo = saved["layers"][li]
        # ---- FFN block ------------------------------------------------------------------------------
        dz2 = ops.layernorm_bwd(y2, dout, w["g2"], mean2, rstd2, lyr.output.LayerNorm.weight.grad, lyr.output.LayerNorm.bias.grad,
                                dxsum=lyr.output.dense.bias.grad, bias=w["b2"], resid=x1, dres=dres, drop=d_h2)
        dz2, dy2 = dz2 if isinstance(dz2, tuple) else (dz2, dz2)      # (residual path, through the dropout mask)
        wg = [(dy2, h, lyr.output.dense.weight.grad)]          # the layer's weight gradients: one grouped launch at its end
        dhpre = ops.gemm_bf16(dy2, w["w2"], M, F, H, ops.EPI_DGELU_BF16, aux=hpre, b_mn=True)
        ops.colsum_bf16(dhpre, lyr.intermediate.dense.bias.grad)
        wg.append((dhpre, x1, lyr.intermediate.dense.weight.grad))
        dx1 = ops.gemm_bf16(dhpre, w["w1"], M, H, F, ops.EPI_NONE_F32, b_mn=True)
        # ---- attention block ------------------------------------------------------------------------
        dz1 = ops.layernorm_bwd(y1, dx1, w["g1"], mean1, rstd1, a.output.LayerNorm.weight.grad, a.output.LayerNorm.bias.grad,
                                dxsum=a.output.dense.bias.grad, bias=w["bo"], resid=x, dres=dz2, drop=d_h1)
        dz1, dy1 = dz1 if isinstance(dz1, tuple) else (dz1, dz1)
        wg.append((dy1, ctx, a.output.dense.weight.grad))
        dctx = ops.gemm_bf16(dy1, w["wo"], M, H, H, ops.EPI_BIAS, b_mn=True)
        dqkv = ops.attention_bwd(qkv, ctx, dctx, lse, key_len, R, S, heads, workspace=ws, drop=d_attn, out_lo=ctx_lo)
        ops.colsum_bf16(dqkv, ar.view(a.self.query.bias, (3 * H,), grad=True))
        wg.append((dqkv, x, ar.view(a.self.query.weight, (3 * H, H), grad=True)))
        if _WGRAD_GROUP:
            # dW += dY^T . X for the four projections in ONE stream-K launch over all their tiles (csrc/gemm_group_tcgen05.cu):
            # separately they paid four ramps and four tails for 94 us of work (profiles/r02/wgrad_streamk.json)
            ops.gemm_wgrad_group(wg)
        else:
            for dy_, x_, dw_ in wg:
                ops.gemm_bf16(dy_, x_, dy_.shape[1], x_.shape[1], M, ops.EPI_ACCUM_F32, out=dw_, a_mn=True, b_mn=True)
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
