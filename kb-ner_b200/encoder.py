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
                 type_vocab_size=1, name="xlm-roberta-large", **_unused):
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
    def forward_hidden(self, ids, key_len):
        """ids [R,S] int32 (cuda), key_len [R] int32 -> last hidden state [R*S, H] bf16.
        The returned tensor aliases an internal workspace buffer (valid until the next call)."""
        if self._compute is None:
            self.sync_compute_weights()
        c = self.config
        R, S = ids.shape
        M = R * S
        ws = self._workspace(M, ids.device)
        e = self.embeddings
        x, xn = ws["x0"], ws["x1"]
        ops.embed_ln_fwd(ids, e.word_embeddings.weight, e.position_embeddings.weight,
                         e.token_type_embeddings.weight[0], e.LayerNorm.weight, e.LayerNorm.bias,
                         c.layer_norm_eps, c.pad_token_id, out=x)
        for w in self._compute:
            ops.gemm_bf16_tn(x, w["wqkv"], w["bqkv"], epilogue=ops.EPI_BIAS, out=ws["qkv"])
            ops.attention_fwd(ws["qkv"], key_len, R, S, c.num_attention_heads, out=ws["ctx"])
            ops.gemm_bf16_tn(ws["ctx"], w["wo"], w["bo"], residual=x, epilogue=ops.EPI_BIAS_RESID_F32, out=ws["y"])
            ops.layernorm_fwd(ws["y"], w["g1"], w["b1"], c.layer_norm_eps, out=xn)
            ops.gemm_bf16_tn(xn, w["w1"], w["bi"], epilogue=ops.EPI_BIAS_GELU, out=ws["h"])
            ops.gemm_bf16_tn(ws["h"], w["w2"], w["b2"], residual=xn, epilogue=ops.EPI_BIAS_RESID_F32, out=ws["y"])
            ops.layernorm_fwd(ws["y"], w["g2"], w["bb2"], c.layer_norm_eps, out=x)
        return x

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
