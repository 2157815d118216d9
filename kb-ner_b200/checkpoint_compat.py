"""Loading weights the reference's users already have.

Two sources (SURVEY.md hard part 8):

* a Hugging Face model directory -- what ``AutoModel.from_pretrained`` reads at ``flair/embeddings.py:2951-2953`` and
  ``save_pretrained`` writes at ``flair/trainers/finetune_trainer.py:1297-1298``: ``config.json`` as transformers
  writes it plus ``model.safetensors`` or ``pytorch_model.bin``, with or without a task-model prefix (``roberta.``)
  and with the extras of other heads (``pooler.*``, ``lm_head.*``, ``embeddings.position_ids``);
* a reference-trained ``best-model.pt`` / ``final-model.pt`` (``flair/nn.py:60-108``, ``sequence_tagger_model.py:435-477``,
  ``:1824-1897``): a pickle that embeds the WHOLE ``flair.embeddings.StackedEmbeddings`` object and through it a
  ``transformers==3.0.0`` module tree (Appendix B.7: class and module names are part of the file format).  Neither
  ``flair`` nor transformers 3.0.0 is importable next to this package, so the file is read with an unpickler that maps
  ``flair.data.Dictionary`` to ours and every other ``flair.*`` / ``transformers.*`` / ``tokenizers.*`` /
  ``sentencepiece.*`` class to an attribute bag; tensors are rebuilt by torch itself.  The numbers come from the
  checkpoint's ``state_dict`` (its encoder keys are ``embeddings.list_embedding_0.model.<HF name>`` -- the names this
  package uses too), the architecture from the pickled config object / the tensor shapes.
"""
import io
import json
import os
import pickle

import torch

_PREFIXES = ("roberta.", "xlm_roberta.", "bert.", "model.", "transformer.")
_DROP = ("pooler.", "lm_head.", "classifier.", "cls.", "qa_outputs.")


def normalize_hf_state_dict(sd):
    """HF checkpoint keys -> the encoder's own (= plain XLMRobertaModel) names."""
    out = {}
    for k, v in sd.items():
        for p in _PREFIXES:
            if k.startswith(p):
                k = k[len(p):]
                break
        if k.startswith(_DROP) or k.endswith("position_ids") or k.endswith("token_type_ids"):
            continue
        # pre-2019 checkpoints name the LayerNorm parameters gamma / beta
        if k.endswith("LayerNorm.gamma"):
            k = k[:-5] + "weight"
        elif k.endswith("LayerNorm.beta"):
            k = k[:-4] + "bias"
        out[k] = v
    return out


def read_weight_files(path):
    """All tensors of a model directory: model.safetensors (also sharded, via its index) or pytorch_model.bin."""
    path = str(path)
    st = os.path.join(path, "model.safetensors")
    idx = os.path.join(path, "model.safetensors.index.json")
    binf = os.path.join(path, "pytorch_model.bin")
    if os.path.exists(st) or os.path.exists(idx):
        from safetensors.torch import load_file
        if os.path.exists(st):
            return load_file(st, device="cpu")
        with open(idx) as f:
            shards = sorted(set(json.load(f)["weight_map"].values()))
        sd = {}
        for s in shards:
            sd.update(load_file(os.path.join(path, s), device="cpu"))
        return sd
    if os.path.exists(binf):
        return torch.load(binf, map_location="cpu", weights_only=True)
    raise FileNotFoundError("%s holds neither model.safetensors nor pytorch_model.bin" % path)


def config_from_hf_json(path, **overrides):
    """EncoderConfig from a config.json written by transformers (or by this package)."""
    from .encoder import EncoderConfig
    with open(os.path.join(str(path), "config.json")) as f:
        cfg = json.load(f)
    mt = cfg.get("model_type")
    if mt not in (None, "xlm-roberta", "roberta", "bert"):
        raise NotImplementedError("model_type %r: the encoder kernels implement the post-LN BERT / (XLM-)RoBERTa block" % mt)
    if cfg.get("hidden_act", "gelu") != "gelu":
        raise NotImplementedError("hidden_act %r: the FFN epilogue is erf-GELU" % cfg.get("hidden_act"))
    if cfg.get("position_embedding_type", "absolute") != "absolute":
        raise NotImplementedError("position_embedding_type %r" % cfg.get("position_embedding_type"))
    name = cfg.get("name") or os.path.basename(os.path.normpath(str(path)))
    keep = ("vocab_size", "hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size",
            "max_position_embeddings", "layer_norm_eps", "pad_token_id", "type_vocab_size", "hidden_dropout_prob",
            "attention_probs_dropout_prob")
    kw = {k: cfg[k] for k in keep if k in cfg}
    kw.update(overrides)
    kw.setdefault("name", name)
    return EncoderConfig(**kw)


def hf_config_dict(config):
    """config.json as transformers' XLMRobertaConfig.save_pretrained lays it out, so that
    ``transformers.XLMRobertaModel.from_pretrained(dir)`` (any version) reads a directory this package wrote."""
    c = config
    return {"architectures": ["XLMRobertaModel"], "model_type": "xlm-roberta", "vocab_size": c.vocab_size,
            "hidden_size": c.hidden_size, "num_hidden_layers": c.num_hidden_layers,
            "num_attention_heads": c.num_attention_heads, "intermediate_size": c.intermediate_size,
            "hidden_act": "gelu", "hidden_dropout_prob": c.hidden_dropout_prob,
            "attention_probs_dropout_prob": c.attention_probs_dropout_prob,
            "max_position_embeddings": c.max_position_embeddings, "type_vocab_size": c.type_vocab_size,
            "initializer_range": 0.02, "layer_norm_eps": c.layer_norm_eps, "pad_token_id": c.pad_token_id,
            "bos_token_id": 0, "eos_token_id": 2, "position_embedding_type": "absolute", "output_hidden_states": True,
            "name": c.name}


# ---- reference-trained tagger checkpoints ---------------------------------------------------------------------------------
class RefObject:
    """Attribute bag standing in for a class of the reference's environment that is not importable here."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):     # (dict, slots) form
            state = {**(state[0] or {}), **state[1]}
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state

    def __call__(self, *a, **k):
        return None


_STUB_ROOTS = ("flair", "transformers", "tokenizers", "sentencepiece", "pytorch_transformers", "allennlp", "gensim", "bpemb")
_stub_classes = {}


def _stub(module, name):
    key = (module, name)
    if key not in _stub_classes:
        _stub_classes[key] = type(name, (RefObject,), {"__module__": module, "_ref_class": "%s.%s" % (module, name)})
    return _stub_classes[key]


class _AliasUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split(".")[0]
        if module == "flair.data" and name == "Dictionary":
            from .data import Dictionary
            return Dictionary
        if root in _STUB_ROOTS:
            return _stub(module, name)
        return super().find_class(module, name)


class _AliasPickle:
    """The `pickle_module` torch.load takes: its Unpickler resolves reference classes through the alias table."""
    __name__ = "kbner_b200.checkpoint_compat"
    Unpickler = _AliasUnpickler
    load = staticmethod(lambda f, **kw: _AliasUnpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _AliasUnpickler(io.BytesIO(b), **kw).load())
    dump, dumps, Pickler = pickle.dump, pickle.dumps, pickle.Pickler
    HIGHEST_PROTOCOL, PickleError, UnpicklingError = pickle.HIGHEST_PROTOCOL, pickle.PickleError, pickle.UnpicklingError


def _modules(obj):
    return getattr(obj, "_modules", None) or {}


def _encoder_config_from_checkpoint(ref_twe, sd, prefix):
    """Architecture of the pickled transformers module: its config object when it unpickled, else the tensor shapes."""
    from .encoder import EncoderConfig
    hf = getattr(_modules(ref_twe).get("model"), "config", None)
    g = lambda k, d=None: getattr(hf, k, d) if hf is not None else d
    word = sd[prefix + "embeddings.word_embeddings.weight"]
    pos = sd[prefix + "embeddings.position_embeddings.weight"]
    n_layers = 1 + max(int(k[len(prefix):].split(".")[2]) for k in sd if k.startswith(prefix + "encoder.layer."))
    inter = sd[prefix + "encoder.layer.0.intermediate.dense.weight"].shape[0]
    H = word.shape[1]
    heads = g("num_attention_heads") or H // 64
    return EncoderConfig(vocab_size=word.shape[0], hidden_size=H, num_hidden_layers=n_layers, num_attention_heads=heads,
                         intermediate_size=inter, max_position_embeddings=pos.shape[0],
                         layer_norm_eps=g("layer_norm_eps", 1e-5), pad_token_id=g("pad_token_id", 1),
                         type_vocab_size=sd[prefix + "embeddings.token_type_embeddings.weight"].shape[0],
                         hidden_dropout_prob=g("hidden_dropout_prob", 0.1),
                         attention_probs_dropout_prob=g("attention_probs_dropout_prob", 0.1),
                         name=str(getattr(ref_twe, "name", "xlm-roberta-large")))


def load_reference_checkpoint(model_file, tokenizer=None, device=None, tagger_cls=None):
    """A tagger from a checkpoint written by the REFERENCE's ``Model.save`` (or by this package's).  `tokenizer`: the
    sub-word tokenizer to attach (the reference pickles a tokenizer object whose SentencePiece file lives elsewhere; when
    omitted, transformers.AutoTokenizer is asked for the embedding's name among the local files)."""
    from .embeddings import StackedEmbeddings, TransformerWordEmbeddings
    from .sequence_tagger import FastSequenceTagger
    tagger_cls = tagger_cls or FastSequenceTagger
    state = torch.load(str(model_file), map_location="cpu", pickle_module=_AliasPickle, weights_only=False)
    emb = state["embeddings"]
    if isinstance(emb, StackedEmbeddings):                      # written by this package: nothing to translate
        model = tagger_cls._init_model_with_state_dict(state, testing=True)
    else:
        members = [m for k, m in sorted(_modules(emb).items()) if k.startswith("list_embedding_")]
        if len(members) != 1 or "TransformerWordEmbeddings" not in getattr(members[0], "_ref_class", ""):
            raise NotImplementedError("checkpoint stacks %s: the KB-NER path is one TransformerWordEmbeddings"
                                      % [getattr(m, "_ref_class", type(m).__name__) for m in members])
        ref = members[0]
        sd = state["state_dict"]
        prefix = "embeddings.list_embedding_0.model."
        cfg = _encoder_config_from_checkpoint(ref, sd, prefix)
        layers = ",".join(str(i) for i in getattr(ref, "layer_indexes", [-1]))
        twe = TransformerWordEmbeddings(
            model=cfg.name, layers=layers, pooling_operation=getattr(ref, "pooling_operation", "first"),
            fine_tune=bool(getattr(ref, "fine_tune", False)), allow_long_sentences=getattr(ref, "allow_long_sentences", True),
            maximum_subtoken_length=getattr(ref, "maximum_subtoken_length", 999), use_scalar_mix=getattr(ref, "use_scalar_mix", False),
            sentence_feat=getattr(ref, "sentence_feat", False), tokenizer=tokenizer, config=cfg, device="cpu")
        twe.stride = getattr(ref, "stride", twe.stride)
        twe.max_subtokens_sequence_length = getattr(ref, "max_subtokens_sequence_length", twe.max_subtokens_sequence_length)
        enc_sd = normalize_hf_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)})
        twe.model.load_hf_state_dict(enc_sd)
        own = dict(state)
        own["embeddings"] = StackedEmbeddings([twe])
        own["state_dict"] = {k: v for k, v in sd.items() if not k.startswith("embeddings.")}
        own["use_locked_dropout"] = own.get("use_locked_dropout", 0.0)
        model = _init_from_reference_state(tagger_cls, own)
    model.eval()
    model.to(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    return model


def _init_from_reference_state(tagger_cls, state):
    """_init_model_with_state_dict (:1824-1897) for a state whose head parameters are loaded non-strictly (the encoder
    was filled from the same file above); every option outside the hot path must be off or the constructor refuses."""
    kw = {k: state[k] for k in ("use_mfvi", "use_language_attention", "enhanced_crf", "use_transition_attention",
                                "use_language_vector", "biaf_attention", "token_level_attention", "embedding_selector",
                                "use_rl", "use_gumbel", "use_embedding_masks", "embedding_attention", "multi_view_training",
                                "map_embeddings", "unlabel_entropy_loss") if k in state}
    kw["relearn_embeddings"] = bool(state.get("relearn_embeddings", False))    # an extra Linear (:966): refused by the constructor
    model = tagger_cls(hidden_size=state["hidden_size"], embeddings=state["embeddings"], tag_dictionary=state["tag_dictionary"],
                       tag_type=state["tag_type"], use_crf=state["use_crf"], use_rnn=state["use_rnn"],
                       use_cnn=state.get("use_cnn", False), rnn_layers=state["rnn_layers"], dropout=state.get("use_dropout", 0.0),
                       word_dropout=state.get("use_word_dropout", 0.0), locked_dropout=state.get("use_locked_dropout", 0.0),
                       remove_x=state.get("remove_x", False), sentence_loss=state.get("sentence_level_loss", False),
                       target_languages=state.get("target_languages", 1), config=state.get("config"), testing=True, **kw)
    missing, unexpected = model.load_state_dict(state["state_dict"], strict=False)
    missing = [k for k in missing if not k.startswith("embeddings.")]
    if missing or unexpected:
        raise KeyError("reference checkpoint: head parameters missing %s / unexpected %s" % (missing, unexpected))
    return model
