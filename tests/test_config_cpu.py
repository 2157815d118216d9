"""YAML -> corpus / tag dictionary / loader through the ConfigParser + Params mirrors (flair/config_parser.py,
flair/utils/params.py) on the fixture derived from the reference's own training YAML (oracle/make_golden_config.py),
and the train.py command line (flag set of /root/reference/train.py:35-64).  No model is built here: that needs the GPU
(tests/test_cli_gpu.py)."""
import json
import os
import shutil

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def write_config(tmp_path, model="xlm-roberta-large", **train_overrides):
    data = tmp_path / "data"
    data.mkdir(exist_ok=True)
    for split in ("train", "dev", "test"):
        shutil.copy(os.path.join(GOLD, "sample_conll.txt"), data / ("%s.txt" % split))
    out = tmp_path / "out"
    text = open(os.path.join(GOLD, "kbner_config.yaml"), encoding="utf-8").read()
    text = text.replace("${DATA}", str(data)).replace("${OUT}", str(out)).replace("${MODEL}", str(model))
    if train_overrides:
        import yaml
        cfg = yaml.safe_load(text)
        cfg["train"].update(train_overrides)
        text = yaml.safe_dump(cfg)
    path = tmp_path / "config.yaml"
    path.write_text(text, encoding="utf-8")
    return str(path), out


def test_config_parser_builds_corpus_and_dictionary(tmp_path):
    from kbner_b200.config_parser import ConfigParser, Params
    from kbner_b200.data import Dictionary
    path, out = write_config(tmp_path)
    params = Params.from_file(path)
    assert "train" in params and params["trainer"] == "ModelFinetuner" and params.get("nope") is None
    cp = ConfigParser(params)
    assert cp.target == cp.tag_type == "ner" and cp.mini_batch_size == 1
    assert cp.corpus_list == ["ColumnCorpus-EN-EnglishDOC"] == cp.corpus.targets and cp.num_corpus == 1
    gold = json.load(open(os.path.join(GOLD, "conll_golden.json"), encoding="utf-8"))
    n = len(gold["bioes"])
    assert (len(cp.corpus.train), len(cp.corpus.dev), len(cp.corpus.test)) == (n, n, n)
    assert [len(x) for x in cp.corpus.train_list] == [n]
    # the reader applied tag_to_bioes + the '# id' comment symbol of the YAML
    assert [t.get_tag("ner").value for t in cp.corpus.train[1].tokens] == gold["bioes"][1]["ner"]
    # dictionary built like Corpus.make_tag_dictionary of the reference (golden from its own code) and pickled
    assert [i.decode() if isinstance(i, bytes) else i for i in cp.tag_dictionary.get_items()] == gold["tag_dictionary"]
    assert (out / "tags.pkl").exists()
    again = ConfigParser(Params.from_file(path))                        # second run LOADS the pickle (:122-124)
    assert again.tag_dictionary.get_items() == Dictionary.load_from_file(out / "tags.pkl").get_items()
    assert str(cp.get_target_path) == str(out / "kbner-fixture") and cp.load_pretrained() is False
    assert cp.tokens[0][0] == "<EOS>"                                    # most frequent token of train + test


def test_config_parser_refuses_what_is_outside_the_path(tmp_path):
    from kbner_b200.config_parser import ConfigParser, Params
    path, _ = write_config(tmp_path)
    params = Params.from_file(path)
    with pytest.raises(NotImplementedError):
        ConfigParser(params, zero_shot=True)
    cp = ConfigParser(params)
    with pytest.raises(NotImplementedError, match="BertEmbeddings"):
        cp.create_embeddings({"BertEmbeddings-0": {"bert_model_or_path": "bert-base-cased"}})
    params["ner"]["Corpus"] = "CONLL_03"
    with pytest.raises(NotImplementedError, match="CONLL_03"):
        ConfigParser(params)


def test_loader_assign_tags(tmp_path):
    from kbner_b200.config_parser import ConfigParser, Params
    from kbner_b200.datasets import ColumnDataLoader
    path, _ = write_config(tmp_path)
    cp = ConfigParser(Params.from_file(path))
    loader = ColumnDataLoader(list(cp.corpus.test), 4, sort_data=False, sentence_level_batch=True)
    loader.assign_tags("ner", cp.tag_dictionary)
    d = cp.tag_dictionary
    for batch in loader:
        T = max(len(s) for s in batch)
        assert tuple(batch.ner_tags.shape) == (len(batch), T)
        for i, s in enumerate(batch):
            want = [d.get_idx_for_item(t.get_tag("ner").value) for t in s.tokens]
            assert s.ner_tags.tolist() == want and batch.ner_tags[i, :len(s)].tolist() == want
            assert batch.ner_tags[i, len(s):].sum() == 0           # padded with 0 = '<unk>'


def test_cli_flag_set_matches_the_reference():
    """Every flag of the reference's parser (train.py:35-64) parses; the ones that leave the hot path are refused."""
    from kbner_b200.train import build_parser, main
    flags = ["config", "test", "zeroshot", "all", "other", "quiet", "nocrf", "parse", "parse_train_and_dev", "keep_order",
             "predict", "debug", "target_dir", "spliter", "recur_parse", "parse_test", "save_embedding", "mst", "test_speed",
             "predict_posterior", "batch_size", "keep_embedding", "remove_x", "v2doc", "eval_train", "num_columns",
             "comment_symbol", "parse_name", "output_dir"]
    ns = build_parser().parse_args(["--config", "c.yaml"])
    assert sorted(vars(ns)) == sorted(flags)
    assert (ns.batch_size, ns.keep_embedding, ns.num_columns, ns.output_dir, ns.spliter) == (-1, -1, 2, "outputs", "\t")
    with pytest.raises(NotImplementedError, match="--zeroshot"):
        main(["--config", "c.yaml", "--zeroshot"])


def test_checkpoint_dictionary_matches_the_reference_key_for_key():
    """tests/golden/state_dict_golden.json is what the REFERENCE's FastSequenceTagger._get_state_dict() writes for the
    KB-NER head (oracle/make_golden_statedict.py).  Ours must carry every key with the same value, so the reference's
    loader (sequence_tagger_model.py:1824-1897) rebuilds the same head from a checkpoint written here."""
    import json
    import os
    from kbner_b200.data import Dictionary
    from kbner_b200.embeddings import StackedEmbeddings, SyntheticTokenizer, TransformerWordEmbeddings
    from kbner_b200.encoder import EncoderConfig
    from kbner_b200.sequence_tagger import FastSequenceTagger
    with open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_golden.json")) as f:
        ref = json.load(f)
    cfg = EncoderConfig(name="t", vocab_size=64, hidden_size=256, num_hidden_layers=1, num_attention_heads=4,
                        intermediate_size=256, max_position_embeddings=40)
    emb = TransformerWordEmbeddings(model="t", layers="-1", pooling_operation="first", fine_tune=True,
                                    tokenizer=SyntheticTokenizer(64), config=cfg, device="cpu")
    d = Dictionary.make_tag_dictionary(["%s-T%d" % ("BIES"[i % 4], i // 4) for i in range(8)], with_x=True)
    assert len(d) == 13
    tagger = FastSequenceTagger(hidden_size=256, embeddings=StackedEmbeddings([emb]), tag_dictionary=d, tag_type="ner",
                                use_crf=True, use_rnn=False, remove_x=True, word_dropout=0.1, sentence_loss=True,
                                testing=True)
    mine = tagger._get_state_dict()
    assert set(ref) <= set(mine), sorted(set(ref) - set(mine))
    for k, v in ref.items():
        if isinstance(v, dict):
            continue                               # state_dict / embeddings / tag_dictionary: objects, checked below
        assert mine[k] == v, (k, mine[k], v)
    head = {n: list(t.shape) for n, t in mine["state_dict"].items() if not n.startswith("embeddings.")}
    want = {n: ([s[0], 256] if n == "linear.weight" else s) for n, s in ref["state_dict"].items()}   # golden used D = 16
    assert head == want
    # and the loader reads its own checkpoint back to the same head
    again = FastSequenceTagger._init_model_with_state_dict(mine, testing=True)
    assert again.remove_x and again.use_word_dropout == 0.1 and again.tagset_size == 13
