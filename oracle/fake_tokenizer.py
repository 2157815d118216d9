"""TEST INFRASTRUCTURE ONLY -- a deterministic SentencePiece-like tokenizer shared by the golden-vector generator
(oracle/make_golden_embed.py, which drives the REFERENCE's own sub-token bookkeeping with it) and the CPU test that feeds
the same tokenizer to kbner_b200 (tests/test_embed_host_cpu.py).

Behaviour chosen to exercise the branches of /root/reference/flair/embeddings.py:3135-3231 and :3347-3408:
  * words are split into pieces of <= 3 characters, the first piece of a word carries the U+2581 prefix;
  * pieces keep their case (the reference lower-cases when it matches pieces to words);
  * a word made only of characters in DROPPED yields NO piece at all (tokenizers omit such words: zero sub-tokens,
    zero vector);
  * "</s>" / "<s>" stay single pieces with their special ids (the '<EOS>' separator of the KB-NER inputs becomes "</s>");
  * ids are a stable hash into [4, vocab);
  * encode_plus(): [<s>] + ids + [</s>] for inputs that fit.  For longer inputs it restates the DOCUMENTED semantics of
    `truncation=True, stride, return_overflowing_tokens` (transformers tokenization docs: the first max_length - 2 ids are
    kept, "overflowing_tokens" = the removed ids preceded by `stride` ids of overlap, IN ORDER), which is what
    transformers >= 3.1 implements.  The 3.0.0 source the reference pins is not available offline, so the reference's
    window + stitching code (:3203-3227, :3292-3299) is pinned GIVEN this overflow rule, not 3.0.0's own implementation.
"""
from typing import List

DROPPED = "​­"          # zero-width space, soft hyphen


class FakeSentencePieceTokenizer:
    bos_token, eos_token, pad_token, unk_token = "<s>", "</s>", "<pad>", "<unk>"
    cls_token, sep_token = "<s>", "</s>"
    _bos_token, _eos_token, _sep_token, _cls_token = "<s>", "</s>", "</s>", "<s>"
    bos_token_id, pad_token_id, eos_token_id, unk_token_id = 0, 1, 2, 3

    def __init__(self, vocab_size=1000, piece_len=3, model_max_length=512):
        self.vocab_size, self.piece_len, self.model_max_length = vocab_size, piece_len, model_max_length

    def tokenize(self, text: str) -> List[str]:
        out = []
        for w in text.split():
            if w in (self.eos_token, self.bos_token):
                out.append(w)
                continue
            w = "".join(ch for ch in w if ch not in DROPPED)
            for i in range(0, len(w), self.piece_len):
                out.append(("▁" if i == 0 else "") + w[i:i + self.piece_len])
        return out

    def convert_tokens_to_ids(self, tokens: List[str]) -> List[int]:
        ids = []
        for t in tokens:
            if t == self.eos_token:
                ids.append(self.eos_token_id)
            elif t == self.bos_token:
                ids.append(self.bos_token_id)
            else:
                h = 2166136261
                for ch in t.encode("utf-8"):
                    h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
                ids.append(4 + h % (self.vocab_size - 4))
        return ids

    def encode_plus(self, ids, max_length=None, stride=0, return_overflowing_tokens=False, truncation=True, **_kw):
        ids = list(ids)
        if max_length is None or len(ids) + 2 <= max_length:
            return {"input_ids": [self.bos_token_id] + ids + [self.eos_token_id]}
        if not truncation:
            raise ValueError("input longer than max_length and truncation is off")
        n_remove = len(ids) + 2 - max_length
        kept = ids[:len(ids) - n_remove]
        out = {"input_ids": [self.bos_token_id] + kept + [self.eos_token_id]}
        if return_overflowing_tokens:
            out["overflowing_tokens"] = ids[max(0, len(kept) - stride):]       # `stride` ids of overlap, then the removed ids
        return out
