"""CPU: loading the weights a user of the reference already has (SURVEY hard part 8).

* Hugging Face model directories -- what AutoModel.from_pretrained reads at /root/reference/flair/embeddings.py:2951-2953
  -- written here by the installed transformers' own save_pretrained (safetensors), with and without the task-model
  prefix, and the way back: a directory this package writes loads into transformers.XLMRobertaModel unchanged.
* A tagger checkpoint written by the REFERENCE's own Model.save (flair/nn.py:60-67) from the reference's own classes
  (imported through oracle/ref_shim, only where /root/reference exists) read back by checkpoint_compat's alias unpickler.
* An unknown model name raises instead of silently building a random encoder.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

TINY = dict(vocab_size=120, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
            max_position_embeddings=66, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5)


def _hf_model(seed=0):
    import transformers
    torch.manual_seed(seed)
    cfg = transformers.XLMRobertaConfig(**TINY)
    return transformers.XLMRobertaModel(cfg).eval()


def test_hf_directory_round_trip(tmp_path):
    import transformers
    from kbner_b200.encoder import XLMRobertaEncoderB200
    hf = _hf_model()
    hf.save_pretrained(str(tmp_path / "hf"))                                  # config.json + model.safetensors
    assert (tmp_path / "hf" / "model.safetensors").exists()
    enc = XLMRobertaEncoderB200.from_pretrained(str(tmp_path / "hf"))
    c = enc.config
    assert (c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.intermediate_size, c.vocab_size) == (128, 2, 2, 256, 120)
    hf_sd = hf.state_dict()
    for k, v in enc.state_dict().items():
        assert torch.equal(v, hf_sd[k]), k
    # the way back: what save_finetuned_embedding writes (finetune_trainer.py:1297-1298) loads into transformers
    with torch.no_grad():
        enc.encoder.layer[1].output.dense.weight.mul_(1.5)
    enc.save_pretrained(str(tmp_path / "ours"))
    back = transformers.XLMRobertaModel.from_pretrained(str(tmp_path / "ours"))
    for k, v in enc.state_dict().items():
        assert torch.equal(v, back.state_dict()[k]), k
    # transformers-3.0.0 vintage file format
    enc.save_pretrained(str(tmp_path / "bin"), safe_serialization=False)
    again = XLMRobertaEncoderB200.from_pretrained(str(tmp_path / "bin"))
    assert torch.equal(again.encoder.layer[1].output.dense.weight, enc.encoder.layer[1].output.dense.weight)


def test_task_model_prefix_extras_and_missing_tensors(tmp_path):
    from safetensors.torch import save_file
    from kbner_b200.encoder import XLMRobertaEncoderB200
    hf = _hf_model(1)
    hf.save_pretrained(str(tmp_path / "plain"))
    sd = {"roberta." + k: v.clone() for k, v in hf.state_dict().items()}       # XLMRobertaFor*: 'roberta.' prefix
    sd["lm_head.bias"] = torch.zeros(120)
    sd["roberta.embeddings.position_ids"] = torch.arange(66)[None]
    d = tmp_path / "prefixed"
    d.mkdir()
    save_file(sd, str(d / "model.safetensors"))
    (d / "config.json").write_text((tmp_path / "plain" / "config.json").read_text())
    enc = XLMRobertaEncoderB200.from_pretrained(str(d))
    assert torch.equal(enc.embeddings.word_embeddings.weight, hf.state_dict()["embeddings.word_embeddings.weight"])
    del sd["roberta.encoder.layer.1.attention.self.key.bias"]
    save_file(sd, str(d / "model.safetensors"))
    with pytest.raises(KeyError):
        XLMRobertaEncoderB200.from_pretrained(str(d))


def test_unknown_model_name_raises_instead_of_random_init():
    from kbner_b200.embeddings import SyntheticTokenizer, TransformerWordEmbeddings
    with pytest.raises(FileNotFoundError):
        TransformerWordEmbeddings(model="xlm-roberta-large-not-on-disk", layers="-1", pooling_operation="first",
                                  tokenizer=SyntheticTokenizer(100), device="cpu")


def test_reference_written_checkpoint_loads(tmp_path):
    import ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference is only present in the build container")
    flair = ref_shim.load_flair()
    import flair.embeddings as FE
    import make_golden as G
    from flair.models import FastSequenceTagger as RefTagger
    # the reference's own embedding classes around a transformers module tree; __init__ of TransformerWordEmbeddings needs
    # hub files, so the instance is assembled attribute by attribute (what its __init__ sets, :2929-3023)
    hf = _hf_model(2)
    twe = FE.TransformerWordEmbeddings.__new__(FE.TransformerWordEmbeddings)
    torch.nn.Module.__init__(twe)
    twe.model, twe.tokenizer = hf, None
    twe.name, twe.fine_tune, twe.static_embeddings = "xlm-roberta-tiny", True, False
    twe.layer_indexes, twe.pooling_operation, twe.use_scalar_mix = [-1], "first", False
    twe.allow_long_sentences, twe.max_subtokens_sequence_length, twe.stride = True, 512, 256
    twe.maximum_subtoken_length, twe.sentence_feat = 999, False          # embedding_length / _type are properties (:3879-3890)
    stack = FE.StackedEmbeddings([twe])
    d = G.make_dictionary(flair, 13, with_x=True)
    torch.manual_seed(3)
    ref = RefTagger(hidden_size=256, embeddings=stack, tag_dictionary=d, tag_type="ner", use_crf=True, use_rnn=False,
                    remove_x=True, word_dropout=0.1, locked_dropout=0.0, sentence_loss=True, testing=True)
    path = tmp_path / "best-model.pt"
    ref.save(path)                                                              # flair/nn.py:60-67, the reference's writer
    from kbner_b200.embeddings import SyntheticTokenizer
    from kbner_b200.sequence_tagger import FastSequenceTagger
    ours = FastSequenceTagger.load(path, device="cpu", tokenizer=SyntheticTokenizer(120))
    assert ours.tag_dictionary.get_items() == [i.decode() for i in d.idx2item]
    assert ours.remove_x is True and ours.use_word_dropout == 0.1 and ours.tag_type == "ner"
    assert torch.equal(ours.transitions.data, ref.transitions.data)
    assert torch.equal(ours.linear.weight.data, ref.linear.weight.data) and torch.equal(ours.linear.bias.data, ref.linear.bias.data)
    enc = ours.embeddings.embeddings[0]
    assert enc.name == "xlm-roberta-tiny" and enc.fine_tune is True and enc.model.config.num_hidden_layers == 2
    for k, v in enc.model.state_dict().items():
        assert torch.equal(v, hf.state_dict()[k]), k
    # and a checkpoint this package writes goes through the same entry point
    ours.save(tmp_path / "ours.pt")
    again = FastSequenceTagger.load(tmp_path / "ours.pt", device="cpu")
    assert torch.equal(again.transitions.data, ref.transitions.data)
