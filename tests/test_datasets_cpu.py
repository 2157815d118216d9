"""CoNLL column reader, BIO -> BIOES, span extraction, tag-dictionary construction and batch assembly against the golden
file produced by the REFERENCE's own code (oracle/make_golden_conll.py -> tests/golden/conll_golden.json)."""
import json
import os
import random

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SAMPLE = os.path.join(HERE, "golden", "sample_conll.txt")
FMT = {0: "text", 1: "pos", 2: "upos", 3: "ner"}


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "conll_golden.json"), encoding="utf-8") as f:
        return json.load(f)


def _read(**kw):
    from kbner_b200.datasets import ColumnDataset
    return ColumnDataset(SAMPLE, FMT, comment_symbol="# id", **kw)


@pytest.mark.parametrize("mode", ["bioes", "raw"])
def test_column_reader_matches_reference(golden, mode):
    ds = _read(tag_to_bioes="ner") if mode == "bioes" else _read()
    want = golden[mode]
    assert len(ds) == len(want)
    for s, w in zip(ds, want):
        assert [t.text for t in s.tokens] == w["text"]
        assert [t.get_tag("ner").value for t in s.tokens] == w["ner"]
        assert [t.get_tag("pos").value for t in s.tokens] == w["pos"]
        assert [list(sp) for sp in s.get_spans("ner")] == w["spans"]


def test_context_tokens_become_s_x(golden):
    ds = _read(tag_to_bioes="ner")
    n_ctx = 0
    for s in ds:
        texts = [t.text for t in s.tokens]
        if "<EOS>" in texts:
            k = texts.index("<EOS>")
            assert all(t.get_tag("ner").value == "S-X" for t in s.tokens[k:])
            n_ctx += 1
    assert n_ctx > 5


def test_tag_dictionary_order_matches_reference(golden):
    from kbner_b200.datasets import Corpus
    ds = _read(tag_to_bioes="ner")
    corpus = Corpus(ds.sentences[:15], ds.sentences[15:20], ds.sentences[20:])
    d = corpus.make_tag_dictionary("ner")
    assert d.get_items() == golden["tag_dictionary"]
    assert d.get_idx_for_item("<unk>") == 0 and d.get_items()[-2:] == ["<START>", "<STOP>"]


@pytest.mark.parametrize("mode,kw", [("sentence_level_4", dict(batch_size=4, sentence_level_batch=True)),
                                     ("word_budget_30", dict(batch_size=30)),
                                     ("unsorted_5", dict(batch_size=5, sentence_level_batch=True, sort_data=False))])
def test_batch_assembly_matches_reference(golden, mode, kw):
    from kbner_b200.datasets import ColumnDataLoader
    sents = list(_read(tag_to_bioes="ner"))
    loader = ColumnDataLoader(sents, **kw)
    got = [[sents.index(s) for s in b] for b in loader]
    assert got == golden["loader_" + mode]
    assert all(hasattr(b, "features") for b in loader)        # BatchedData
    # reshuffle permutes batches only; true_reshuffle keeps the multiset of sentences
    random.seed(3)
    loader.reshuffle()
    assert sorted(map(tuple, got)) == sorted(tuple(sents.index(s) for s in b) for b in loader)
    loader.true_reshuffle()
    assert sorted(i for b in got for i in b) == sorted(sents.index(s) for b in loader for s in b)


def test_iob_helpers_edge_cases():
    from kbner_b200.datasets import iob2, iob_iobes
    tags = ["I-PER", "I-PER", "O", "I-LOC", "B-LOC", "I-ORG"]
    assert iob2(tags) and tags == ["B-PER", "I-PER", "O", "B-LOC", "B-LOC", "B-ORG"]
    assert iob_iobes(tags) == ["B-PER", "E-PER", "O", "S-LOC", "S-LOC", "S-ORG"]
    assert not iob2(["O", "X-PER"])
    with pytest.raises(ValueError):
        iob_iobes(["E-PER"])


def test_column_corpus_discovers_files(tmp_path):
    from kbner_b200.datasets import ColumnCorpus
    text = open(SAMPLE, encoding="utf-8").read()
    for name in ("en_train.conll", "en_dev.conll", "en_test.conll"):
        (tmp_path / name).write_text(text, encoding="utf-8")
    c = ColumnCorpus(tmp_path, FMT, tag_to_bioes="ner", comment_symbol="# id")
    assert len(c.train) == len(c.dev) == len(c.test) == 25
    only = tmp_path / "only"
    only.mkdir()
    (only / "train.txt").write_text(text, encoding="utf-8")
    c2 = ColumnCorpus(only, FMT, tag_to_bioes="ner", comment_symbol="# id")      # 10 % splits like the reference
    assert len(c2.test) == round(25 / 10) and len(c2.dev) == round((25 - len(c2.test)) / 10)
    assert len(c2.train) + len(c2.dev) + len(c2.test) == 25
