"""CPU: the host half of TransformerWordEmbeddings (sub-tokenisation bookkeeping, window rows, first-sub-token map --
SURVEY.md 8(a) rows a3 / a6) against golden vectors produced by the REFERENCE's own
`_add_embeddings_to_sentences` + `reconstruct_tokens_from_subtokens` (oracle/make_golden_embed.py drives them with
oracle/fake_tokenizer.py and a stand-in model whose hidden state encodes (row, position, id), so the golden says which
sub-token of which window the reference pools for every token)."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden", "embed_golden.json")


def _load():
    with open(GOLDEN) as f:
        return json.load(f)


def _embeddings(opts, tok_kw):
    from fake_tokenizer import FakeSentencePieceTokenizer
    from kbner_b200.embeddings import TransformerWordEmbeddings
    from kbner_b200.encoder import EncoderConfig
    cfg = EncoderConfig(name="fake", vocab_size=tok_kw["vocab_size"], hidden_size=256, num_hidden_layers=1,
                        num_attention_heads=4, intermediate_size=256, max_position_embeddings=514)
    return TransformerWordEmbeddings(model="fake", layers="-1", pooling_operation="first", fine_tune=False,
                                     tokenizer=FakeSentencePieceTokenizer(**tok_kw), config=cfg, device="cpu",
                                     maximum_subtoken_length=opts.get("maximum_subtoken_length", 999))


@pytest.mark.parametrize("case", _load()["cases"], ids=lambda c: c["name"])
def test_build_batch_matches_reference(case):
    from kbner_b200.data import BatchedData, Sentence
    emb = _embeddings(case["options"], _load()["tokenizer"])
    batch = BatchedData([Sentence(tokens=list(w)) for w in case["sentences"]])
    ids, key_len, row_of, first_idx, lengths, S = emb.build_batch(batch)
    assert lengths == [len(w) for w in case["sentences"]]
    ref_ids = case["input_ids"]
    if ref_ids:
        # the reference builds input_ids [P, S] zero-padded (embeddings.py:3247-3260) for the NON-EMPTY sentences, in order
        assert ids.tolist() == ref_ids
        assert key_len.tolist() == [int(x) for x in case["mask_len"]]
    for b, row in enumerate(case["tokens"]):
        for t, want in enumerate(row):
            fi = int(first_idx[b, t])
            if want is None:                      # zero vector: no sub-token / sentence without any sub-token
                assert fi < 0, (case["name"], b, t, fi)
                continue
            assert fi >= 0, (case["name"], b, t)
            r, p = int(row_of[b]) + fi // S, fi % S
            assert [r, p, int(ids[r, p])] == want, (case["name"], b, t)
    # padding columns beyond a sentence's word count never point anywhere
    T = first_idx.shape[1]
    for b, n in enumerate(lengths):
        assert bool((first_idx[b, n:T] < 0).all())


def test_golden_covers_the_branches():
    names = {c["name"]: c for c in _load()["cases"]}
    assert any(t is None for row in names["dropped_words"]["tokens"] for t in row)           # omitted words
    assert all(t is None for row in names["empty_sentence"]["tokens"] for t in row)          # sentence without sub-tokens
    assert max(names["long_fits"]["mask_len"]) > 500                                          # near-full window
    assert 2 in [i for row in names["with_eos_context"]["input_ids"] for i in row[1:-1]]     # '<EOS>' became </s> inside


def test_build_batch_accepts_the_references_own_sentence_objects():
    """Duck typing at the drop-in boundary: the reference's flair.data.Sentence / Token inside its own
    flair.custom_data_loader.BatchedData (custom_data_loader.py:13-20) go through build_batch / embed's host half unchanged
    and give the same tensors as this package's data classes.  Needs /root/reference (build container only)."""
    import ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference is only present in the build container")
    ref_shim.load_flair()
    from flair.custom_data_loader import BatchedData as RefBatched
    from flair.data import Sentence as RefSentence, Token as RefToken
    from kbner_b200.data import BatchedData, Sentence
    case = next(c for c in _load()["cases"] if c["name"] == "with_eos_context")
    emb = _embeddings(case["options"], _load()["tokenizer"])
    ref_sents = []
    for words in case["sentences"]:
        s = RefSentence()
        for w in words:
            s.add_token(RefToken(w))
        ref_sents.append(s)
    got = emb.build_batch(RefBatched(ref_sents))
    want = emb.build_batch(BatchedData([Sentence(tokens=list(w)) for w in case["sentences"]]))
    for a, b in zip(got[:4], want[:4]):
        assert torch.equal(a, b)
    assert got[4] == want[4] and got[5] == want[5]
    assert hasattr(RefBatched(ref_sents), "features")          # what embed() writes its EncodedBatch into


def test_per_word_piece_cache_agrees_with_the_sentence_level_algorithm():
    """The per-word fast path of subtokenize is only used after it has agreed with the reference's sentence-level matching on
    the first sentences; here: on every golden case (dropped words, <EOS>, truncation to maximum_subtoken_length, windows) the
    two paths give the same ids / n_sub, a verified instance answers from the per-word cache, and a tokenizer whose pieces
    DO depend on the neighbouring word switches the fast path off."""
    from fake_tokenizer import FakeSentencePieceTokenizer
    from kbner_b200.data import Sentence
    g = _load()
    for case in g["cases"]:
        emb = _embeddings(case["options"], g["tokenizer"])
        for w in case["sentences"]:
            words = [emb._eos_text() if x == "<EOS>" else x for x in w]
            assert emb._subtokenize_by_word(words) == emb._subtokenize_words(words), case["name"]
        assert emb._by_word is None
    emb = _embeddings({}, g["tokenizer"])
    emb._verify_left = 2
    s1, s2, s3 = (Sentence(tokens=["alpha", "beta%d" % i, "gamma"]) for i in range(3))
    want = [emb._subtokenize_words([t.text for t in s.tokens]) for s in (s1, s2, s3)]
    assert [emb.subtokenize(s) for s in (s1, s2)] == want[:2] and emb._by_word is True
    calls = []
    real = emb.tokenizer.tokenize
    emb.tokenizer.tokenize = lambda text: (calls.append(text), real(text))[1]
    assert emb.subtokenize(s3) == want[2] and calls == ["beta2"]          # only the never-seen word reached the tokenizer

    class Contextual(FakeSentencePieceTokenizer):                        # merges a word with its left neighbour's last letter
        def tokenize(self, text):
            ws = text.split()
            return [p for i, w in enumerate(ws) for p in super(Contextual, self).tokenize(w if i == 0 else w + ws[i - 1][-1:])]
    emb = _embeddings({}, g["tokenizer"])
    emb.tokenizer = Contextual(**g["tokenizer"])
    emb.subtokenize(Sentence(tokens=["abc", "def", "ghi"]))
    assert emb._by_word is False
