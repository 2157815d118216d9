#!/usr/bin/env python
"""TEST INFRASTRUCTURE: runs the REFERENCE's own TransformerWordEmbeddings._add_embeddings_to_sentences
(/root/reference/flair/embeddings.py:3111-3345 incl. reconstruct_tokens_from_subtokens :3347-3408) on CPU with
oracle/fake_tokenizer.py and a stand-in "transformer" whose last hidden state encodes (window row, position, input id),
so the embedding the reference assigns to every token says exactly WHICH sub-token of WHICH window it pooled.
Writes tests/golden/embed_golden.json: per case the words, the input_ids / mask the reference built and, per token,
[row, position, id] or null for the zero vector.  Run in the build container only:  python oracle/make_golden_embed.py"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from fake_tokenizer import FakeSentencePieceTokenizer  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "embed_golden.json")


def cases():
    rnd = random.Random(11)
    alpha = "abcdefghijklmnopqrstuvwxyzABCDEFGH0123456789-'"

    def word(lo=1, hi=11):
        return "".join(rnd.choice(alpha) for _ in range(rnd.randint(lo, hi)))
    out = []
    # 1: plain sentences of different lengths in one batch (padding of input_ids, several words per window)
    out.append(("plain_batch", [[word() for _ in range(n)] for n in (7, 1, 19, 4)], {}))
    # 2: the KB-NER input form: sentence <EOS> retrieved context (embeddings.py:3139-3163)
    out.append(("with_eos_context", [[word() for _ in range(5)] + ["<EOS>"] + [word() for _ in range(23)],
                                     [word() for _ in range(3)] + ["<EOS>"] + [word() for _ in range(9)] + ["<EOS>"] + [word() for _ in range(4)]], {}))
    # 3: words the tokenizer omits entirely (zero sub-tokens -> zero vector), in the middle and at the end
    out.append(("dropped_words", [[word(), "​", word(), word(), "­​", word()],
                                  [word(), word(), "​"]], {}))
    # 4: mixed case and punctuation-like single characters
    out.append(("case_and_singles", [["Hello", "WORLD", ",", "a", "B", ".", "MiXeD-Case's"], ["X"]], {}))
    # 5: words cut to maximum_subtoken_length sub-tokens (:3183-3197)
    out.append(("max_subtoken_length", [[word(10, 11), word(1, 2), word(12, 14), word(3, 3)]], {"maximum_subtoken_length": 2}))
    # 6: a long sentence that still fits one window (508 sub-tokens + specials <= 512)
    out.append(("long_fits", [[word(3, 3) for _ in range(254)] + ["<EOS>"] + [word(3, 3) for _ in range(253)]], {}))
    # 7: a sentence whose every word is dropped: no sub-token at all -> removed from the batch, zero vectors
    out.append(("empty_sentence", [["​", "­"]], {}))
    # 8-10: more than 510 sub-tokens -> overlapping windows (encode_plus overflow, :3203-3227) stitched by dropping
    # stride // 2 (+ the special token) on each inner edge (:3292-3299): two windows, three windows with multi-piece words
    # and an <EOS> context, and a batch that mixes a windowed sentence with short ones (row bookkeeping)
    out.append(("overflow_two_windows", [[word(3, 3) for _ in range(700)]], {}))
    out.append(("overflow_three_windows", [[word(1, 8) for _ in range(120)] + ["<EOS>"] + [word(1, 8) for _ in range(420)]], {}))
    out.append(("overflow_in_batch", [[word() for _ in range(5)], [word(2, 3) for _ in range(600)], [word() for _ in range(9)]], {}))
    return out


def main():
    import torch
    flair = ref_shim.load_flair()
    import flair.embeddings as FE
    from flair.data import Sentence, Token
    flair.device = torch.device("cpu")

    class FakeModel(torch.nn.Module):
        class config:
            hidden_size = 4

        def forward(self, input_ids, attention_mask=None, inputs_embeds=None):
            self.seen = (input_ids.clone(), attention_mask.clone())
            R, S = input_ids.shape
            h = torch.zeros(R, S, 4)
            h[..., 0] = torch.arange(R, dtype=torch.float32)[:, None] + 1.0          # row + 1 (0 = the zero vector)
            h[..., 1] = torch.arange(S, dtype=torch.float32)[None, :]
            h[..., 2] = input_ids.float()
            h[..., 3] = attention_mask.float()
            return h, h[:, 0], (torch.zeros_like(h), h)

    golden = {"tokenizer": {"vocab_size": 1000, "piece_len": 3}, "cases": []}
    for name, sents, opts in cases():
        emb = object.__new__(FE.TransformerWordEmbeddings)
        torch.nn.Module.__init__(emb)
        emb.tokenizer = FakeSentencePieceTokenizer()
        emb.model = FakeModel()
        emb.name = "fake"
        emb.allow_long_sentences = True
        emb.max_subtokens_sequence_length = 512
        emb.stride = 256
        emb.layer_indexes = [-1]
        emb.pooling_operation = "first"
        emb.use_scalar_mix = False
        emb.fine_tune = False
        emb.static_embeddings = True
        emb.sentence_feat = False
        emb.use_internal_doc = False
        emb.special_tokens = ["<s>", "<s>"]
        emb.begin_offset = 1
        emb.maximum_subtoken_length = opts.get("maximum_subtoken_length", 999)
        ref_sents = []
        for words in sents:
            s = Sentence()
            for w in words:
                s.add_token(Token(w))
            ref_sents.append(s)
        emb._add_embeddings_to_sentences(ref_sents)
        seen = getattr(emb.model, "seen", None)
        case = {"name": name, "sentences": sents, "options": opts,
                "input_ids": seen[0].tolist() if seen else [], "mask_len": seen[1].sum(1).tolist() if seen else [],
                "tokens": []}
        for s in ref_sents:
            row = []
            for tok in s:
                v = tok.get_embedding()
                assert v.numel() == 4
                row.append(None if float(v[0]) == 0.0 else [int(v[0]) - 1, int(v[1]), int(v[2])])
            case["tokens"].append(row)
        golden["cases"].append(case)
    with open(OUT, "w") as f:
        json.dump(golden, f, indent=0)
    print("wrote", OUT, {c["name"]: sum(len(t) for t in c["tokens"]) for c in golden["cases"]})


if __name__ == "__main__":
    main()
