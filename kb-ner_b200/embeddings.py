"""``TransformerWordEmbeddings`` / ``StackedEmbeddings`` on the B200 encoder.

Mirror of ``/root/reference/flair/embeddings.py``: ``TransformerWordEmbeddings`` (:2906-3416, same constructor
keywords and attribute surface), ``Embeddings.embed`` (:75-101), ``assign_batch_features`` (:108-124) and
``StackedEmbeddings`` (:155-211).  What changes is where the work happens:

* sub-tokenisation and the word -> first-sub-token map are host logic (cached per sentence -- the reference
  re-tokenises every word on every call, :3103-3109);
* the encoder runs on the sm_100a kernels (``encoder.XLMRobertaEncoderB200``), once per batch, producing only
  the last hidden state;
* word pooling is *not* a Python loop over tokens with a ``.cpu()`` round trip (:3288-3345, :122): the batch
  keeps an ``EncodedBatch`` (device hidden state + index tensors) and the tagger fuses gather + word dropout
  + projection into one kernel.  Per-token ``Token._embeddings`` vectors are materialised only on request
  (``materialize_token_embeddings``), which is what ``embeddings_storage_mode='none'`` configs never need.

Target configs use ``layers='-1'``, ``pooling_operation='first'`` (config/*.yaml embeddings block); other
values raise instead of silently computing something else.
"""
import itertools
import operator
import re
from typing import List, Optional

import numpy as np
import torch

from .data import BatchedData
from .encoder import EncoderConfig, XLMRobertaEncoderB200


_chain = itertools.chain.from_iterable
_first, _second = operator.itemgetter(0), operator.itemgetter(1)

class SyntheticTokenizer:
    """Deterministic stand-in for the SentencePiece tokenizer (no tokenizer files exist offline).

    SentencePiece-style surface: words are split into pieces of <= ``piece_len`` characters, the first piece
    of every word carries the U+2581 prefix; ids are a stable hash into [3, vocab).  Exposes exactly the
    members the reference touches (embeddings.py:2951-2966, :3139-3181)."""

    FILE = "synthetic_tokenizer.json"
    bos_token, eos_token, pad_token, unk_token = "<s>", "</s>", "<pad>", "<unk>"
    _bos_token, _eos_token, _sep_token, _cls_token = "<s>", "</s>", "</s>", "<s>"
    bos_token_id, pad_token_id, eos_token_id, unk_token_id = 0, 1, 2, 3

    def __init__(self, vocab_size=250002, piece_len=4, model_max_length=512):
        self.vocab_size = vocab_size
        self.piece_len = piece_len
        self.model_max_length = model_max_length

    def save_pretrained(self, path):
        """Written next to the encoder by `save_finetuned_embedding` (finetune_trainer.py:1297); a directory holding this
        file loads back as a SyntheticTokenizer."""
        import json
        import os
        os.makedirs(str(path), exist_ok=True)
        with open(os.path.join(str(path), self.FILE), "w") as f:
            json.dump({"vocab_size": self.vocab_size, "piece_len": self.piece_len,
                       "model_max_length": self.model_max_length}, f)

    @classmethod
    def from_pretrained(cls, path):
        import json
        import os
        with open(os.path.join(str(path), cls.FILE)) as f:
            return cls(**json.load(f))

    def tokenize(self, text: str) -> List[str]:
        out = []
        for w in text.split():
            if w in (self.eos_token, self.bos_token):
                out.append(w)
                continue
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


def _strip_markup(piece: str) -> str:
    """Sub-token -> surface text: the four substitutions of the reference (embeddings.py:3093-3101), applied one after the
    other like its re.sub chain, as plain string tests (130 k regex calls per 32 x 510-word batch were 100 ms of host time)."""
    if piece[:1] == "Ġ":
        piece = piece[1:]
    if piece[:2] == "##":
        piece = piece[2:]
    if piece[:1] == "▁":
        piece = piece[1:]
    if piece[-4:] == "</w>":
        piece = piece[:-4]
    return piece


class EncodedBatch:
    """Device-side result of embedding one batch: what the fused pooling kernel consumes."""

    def __init__(self, hidden, S, row_of, first_idx, lengths, key_len, ids, saved=None, encoder=None):
        self.saved = saved            # activations kept for the backward pass (fine-tuning only)
        self.encoder = encoder
        self.hidden = hidden          # [R*S, H] bf16
        self.S = S
        self.row_of = row_of          # [B] int32: window row that starts each sentence
        self.first_idx = first_idx    # [B,T] int32: sub-token index (row-relative, may span rows) or -1
        self.lengths = lengths        # python list of word counts
        self.key_len = key_len
        self.ids = ids


class TransformerWordEmbeddings(torch.nn.Module):
    def __init__(self, model="xlm-roberta-large", layers: str = "-1", pooling_operation: str = "first",
                 batch_size: int = 1, use_scalar_mix: bool = False, fine_tune: bool = False,
                 allow_long_sentences: bool = True, stride: int = -1, maximum_window: bool = False,
                 document_extraction: bool = False, embedding_name: Optional[str] = None, doc_batch_size: int = 32,
                 maximum_subtoken_length: int = 999, v2_doc: bool = False, ext_doc: bool = False,
                 sentence_feat: bool = False, use_internal_doc: bool = False, tokenizer=None, config=None,
                 device=None, **kwargs):
        super().__init__()
        if pooling_operation != "first" or [int(x) for x in str(layers).split(",")] != [-1]:
            raise NotImplementedError("kbner_b200 implements the KB-NER configuration: layers='-1', "
                                      "pooling_operation='first' (got layers=%r pooling=%r)" % (layers, pooling_operation))
        if document_extraction or v2_doc or ext_doc or use_internal_doc or sentence_feat or use_scalar_mix:
            raise NotImplementedError("document-context / sentence_feat / scalar-mix variants are outside the hot path")
        self.device_ = torch.device(device if device is not None else "cuda")
        # ---- encoder + tokenizer ---------------------------------------------------------------
        if isinstance(model, XLMRobertaEncoderB200):
            self.model = model
            name = model.config.name
        else:
            name = str(model)
            if config is None:
                import os
                if os.path.isdir(name) and os.path.exists(os.path.join(name, "config.json")):
                    self.model = XLMRobertaEncoderB200.from_pretrained(name)
                    if tokenizer is None and os.path.exists(os.path.join(name, SyntheticTokenizer.FILE)):
                        tokenizer = SyntheticTokenizer.from_pretrained(name)
                else:
                    # the reference would download `name` from the hub here (AutoModel.from_pretrained, :2951-2953); there is
                    # nothing to load it from, and a silently random-initialised encoder is exactly the "computing something
                    # else" this package refuses: random init only happens when the caller passes an explicit `config`
                    raise FileNotFoundError(
                        "no local weights for %r: pass a directory written by save_pretrained() (config.json + "
                        "model.safetensors / pytorch_model.bin, HF or kbner_b200), or config=EncoderConfig(...) for a "
                        "randomly initialised encoder" % name)
            if not hasattr(self, "model"):
                with torch.device(self.device_):
                    self.model = XLMRobertaEncoderB200(config)
        self.model.to(self.device_)
        if tokenizer is None:
            try:
                from transformers import AutoTokenizer
                tokenizer = AutoTokenizer.from_pretrained(name, local_files_only=True, **kwargs)
            except Exception as e:       # no tokenizer files offline: say so, do not guess silently
                raise RuntimeError("no tokenizer files for %r are available locally (%s); pass tokenizer=... "
                                   "(e.g. kbner_b200.embeddings.SyntheticTokenizer for synthetic data)" % (name, e))
        self.tokenizer = tokenizer
        # ---- the attribute surface the reference's callers read -------------------
        self.allow_long_sentences = allow_long_sentences
        mml = min(getattr(tokenizer, "model_max_length", 512) or 512, 512)
        self.max_subtokens_sequence_length = mml
        self.stride = mml // 2 if allow_long_sentences else 0
        if allow_long_sentences and stride != -1:
            if not maximum_window:
                self.max_subtokens_sequence_length = stride * 2
            self.stride = stride
        self.name = str(name) if embedding_name is None else embedding_name
        self.layer_indexes = [-1]
        self.pooling_operation = pooling_operation
        self.use_scalar_mix = use_scalar_mix
        self.fine_tune = fine_tune
        self.static_embeddings = not fine_tune
        self.batch_size = batch_size
        self.sentence_feat = sentence_feat
        self.use_internal_doc = use_internal_doc
        self.document_extraction = document_extraction
        self.v2_doc, self.ext_doc = v2_doc, ext_doc
        self.doc_batch_size = doc_batch_size
        self.begin_offset = 1
        self.maximum_subtoken_length = maximum_subtoken_length
        self.embedding_type = "word-level"
        self._tok_cache = {}
        self._plan_cache = {}
        self._word_cache = {}
        self._piece_cache = {}
        self._piece1 = {}             # word -> its single sub-token id (words of exactly one piece; subset of _piece_cache)
        import os
        self._by_word = False if os.environ.get("KBNER_WORD_CACHE", "1") == "0" else None     # None: still being verified
        self._verify_left = 64
        self._stage = None
        self.model.eval()

    @property
    def embedding_length(self) -> int:
        return len(self.layer_indexes) * self.model.config.hidden_size

    def train(self, mode=True):
        # not fine-tuning ("feature-based"): never in training mode (embeddings.py:3410-3416)
        if self.fine_tune:
            super().train(mode)
        return self

    # ---- host logic: sub-tokenisation -----------------------------------------------------------
    def _eos_text(self):
        eos = getattr(self.tokenizer, "_eos_token", None) or getattr(self.tokenizer, "_sep_token", None)
        return getattr(eos, "content", eos)

    def _word_text(self, text: str) -> str:
        """_get_processed_token_text (:3103-3109): the word re-tokenised on its own, markup stripped, lower-cased.
        Memoised per word string (the reference recomputes it for every word of every sentence on every call)."""
        hit = self._word_cache.get(text)
        if hit is None:
            hit = "".join(_strip_markup(p) for p in self.tokenizer.tokenize(text)).lower()
            if len(self._word_cache) < 1000000:
                self._word_cache[text] = hit
        return hit

    def subtokenize(self, sentence):
        """-> (sub-token ids without specials, n_sub per word).  '<EOS>' words become the tokenizer's EOS
        (embeddings.py:3139-3163); words are matched to sub-tokens by reconstructing their surface text
        (:3347-3408); words longer than maximum_subtoken_length are cut (:3183-3197).

        Two levels of caching.  Per sentence (the reference re-tokenises every sentence on every call).  Per WORD: for
        tokenizers that segment whitespace-separated words independently (SentencePiece with split_by_whitespace, i.e.
        XLM-R; WordPiece) the sentence's pieces are the concatenation of its words' pieces, so a never-seen sentence costs
        a dictionary lookup per word instead of a tokenizer call + the matching loop (55 -> 6 ms per 32 x 510-word batch).
        That property is CHECKED, not assumed: the first `_verify_left` sentences go through both paths and must agree
        exactly; one disagreement switches the per-word path off for good (KBNER_WORD_CACHE=0 never uses it)."""
        words = [t.text for t in sentence.tokens]
        key = tuple(words)
        hit = self._tok_cache.get(key)
        if hit is not None:
            return hit
        if "<EOS>" in words:
            eos = self._eos_text()
            if eos:
                words = [eos if w == "<EOS>" else w for w in words]
        state = self._by_word
        if state is True:
            out = self._subtokenize_by_word(words)
        else:
            out = self._subtokenize_words(words)
            if state is None:
                if self._subtokenize_by_word(words) != out:
                    self._by_word = False
                else:
                    self._verify_left -= 1
                    if self._verify_left <= 0:
                        self._by_word = True
        if len(self._tok_cache) < 200000:
            self._tok_cache[key] = out
        return out

    def _subtokenize_by_word(self, words):
        """Concatenation of the words' cached pieces.  The loops run inside the interpreter's C code (dict.get through map,
        chain, list): an explicit Python loop with `ids += ...; n_sub.append(...)` was the largest single term of a cold
        batch (1.6 us per word, 26 of 42 profiled ms per 32 x 510-word batch)."""
        flat = self._piece1
        ids = list(map(flat.get, words))              # words known to be ONE piece: a flat word -> id table
        if None not in ids:
            return ids, [1] * len(ids)
        cache = self._piece_cache
        ents = list(map(cache.get, words))
        if None in ents:
            for i, e in enumerate(ents):
                if e is None:
                    w = words[i]
                    e = cache.get(w)
                    if e is None:
                        wi, wn = self._subtokenize_words([w])
                        e = (wi, wn[0])
                        if len(cache) < 2000000:
                            cache[w] = e
                            if wn[0] == 1 and len(wi) == 1:
                                flat[w] = wi[0]
                    ents[i] = e
        return list(_chain(map(_first, ents))), list(map(_second, ents))

    def _subtokenize_words(self, words):
        """The reference's algorithm on a list of word strings: tokenise the joined text, then walk the pieces and the
        words' own surface forms in step (:3347-3408)."""
        pieces = self.tokenizer.tokenize(" ".join(words))
        n_sub, wi, acc, cnt = [], 0, "", 0
        targets = [self._word_text(w) for w in words]
        for pc in pieces:
            surface = _strip_markup(pc).lower()
            # tokenizers may drop a word entirely: it gets 0 sub-tokens (zero vector later)
            while wi < len(words) and cnt == 0 and not targets[wi].startswith(surface):
                n_sub.append(0)
                wi += 1
            if wi >= len(words):
                break
            acc += surface
            cnt += 1
            if acc == targets[wi]:
                n_sub.append(cnt)
                wi, acc, cnt = wi + 1, "", 0
        while len(n_sub) < len(words):
            n_sub.append(cnt)
            cnt = 0
        if any(n > self.maximum_subtoken_length for n in n_sub):
            kept, i = [], 0
            for n in n_sub:
                kept += pieces[i:i + min(n, self.maximum_subtoken_length)]
                i += n
            pieces = kept
            n_sub = [min(n, self.maximum_subtoken_length) for n in n_sub]
        return self.tokenizer.convert_tokens_to_ids(pieces), n_sub

    def windows(self, n_ids: int):
        """Window starts for a sentence of n_ids sub-tokens: [<s>] ids[start:start+W-2] [</s>] with overlap
        `stride` (encode_plus(max_length, stride, return_overflowing_tokens), embeddings.py:3203-3227)."""
        cap = self.max_subtokens_sequence_length - 2
        starts = [0]
        if self.allow_long_sentences:
            while starts[-1] + cap < n_ids:
                starts.append(starts[-1] + cap - self.stride)
        return starts, cap

    def _sentence_plan(self, sentence):
        """Cached per sentence: window rows (np.int32 arrays incl. <s> </s>) and, per word, the first-sub-token
        position as (window, index-in-window) or -1."""
        key = tuple(t.text for t in sentence.tokens)
        hit = self._plan_cache.get(key)
        if hit is not None:
            return hit
        tok = self.tokenizer
        bos, eos = getattr(tok, "bos_token_id", 0), getattr(tok, "eos_token_id", 2)
        ids, n_sub = self.subtokenize(sentence)
        starts, cap = self.windows(len(ids))
        if not self.allow_long_sentences:
            ids = ids[:cap]
        n_sub_a = np.asarray(n_sub, dtype=np.int64)
        if len(starts) == 1:
            # one window (the common case): every word sits in window 0 at 1 + its first sub-token's index
            row = np.empty(len(ids) + 2, dtype=np.int32)
            row[0], row[-1] = bos, eos
            row[1:-1] = ids
            rows = [row]
            inwin = np.cumsum(n_sub_a)
            inwin += 1 - n_sub_a                      # 1 + exclusive prefix sum
            valid = (n_sub_a > 0) & (inwin <= len(ids))
            plan = (rows, valid.astype(np.int64) - 1, inwin)      # window 0 where valid, -1 otherwise
        else:
            rows = [np.asarray([bos] + ids[st:st + cap] + [eos], dtype=np.int32) for st in starts]
            # stitched index of sub-token g: window w owns [starts[w]+half, starts[w+1]+half) after dropping stride//2
            # (+1 special) on each inner edge (:3292-3299)
            half = self.stride // 2
            g = np.concatenate([[0], np.cumsum(n_sub_a)[:-1]]) if len(n_sub) else np.zeros(0, np.int64)
            starts_a = np.asarray(starts, dtype=np.int64)
            w = np.searchsorted(starts_a[1:] + half, g, side="right")
            inwin = 1 + g - starts_a[w]
            valid = (n_sub_a > 0) & (g < len(ids))
            plan = (rows, np.where(valid, w, -1).astype(np.int64), inwin.astype(np.int64))
        if len(self._plan_cache) < 200000:
            self._plan_cache[key] = plan
        return plan

    def build_batch(self, sentences):
        """Host side of _add_embeddings_to_sentences (:3135-3260): ids / key_len / first-sub-token map.
        The reference pads input_ids with 0 (:3247-3251); so do we (pads are masked as keys)."""
        plans = [self._sentence_plan(s) for s in sentences]
        lengths = [len(s.tokens) for s in sentences]
        B, T = len(sentences), max(lengths)
        R = sum(len(p[0]) for p in plans)
        S = max(len(r) for p in plans for r in p[0])
        ids = np.zeros((R, S), dtype=np.int32)
        key_len = np.zeros((R,), dtype=np.int32)
        row_of = np.zeros((B,), dtype=np.int32)
        first_idx = np.full((B, T), -1, dtype=np.int32)
        r = 0
        for b, (rows, w, inwin) in enumerate(plans):
            row_of[b] = r
            for row in rows:
                ids[r, :len(row)] = row
                key_len[r] = len(row)
                r += 1
            fi = np.where(w >= 0, w * S + inwin, -1)          # rows of one sentence are consecutive
            first_idx[b, :len(fi)] = fi
        return (torch.from_numpy(ids), torch.from_numpy(key_len), torch.from_numpy(row_of),
                torch.from_numpy(first_idx), lengths, S)

    # ---- device side ----------------------------------------------------------------------------------
    def _staging(self, n):
        """Rotating pinned int32 staging buffers (one H2D copy per batch; a buffer is reused only after the copy
        that last read it has completed)."""
        if not hasattr(self, "_stage") or self._stage is None:
            self._stage, self._stage_i = [], 0
        if len(self._stage) < 4:
            buf = torch.empty(max(n, 1 << 16), dtype=torch.int32)
            if torch.cuda.is_available():
                buf = buf.pin_memory()
            self._stage.append([buf, None])
            return self._stage[-1]
        slot = self._stage[self._stage_i % 4]
        self._stage_i += 1
        if slot[1] is not None:
            slot[1].synchronize()
        if slot[0].numel() < n:
            slot[0] = torch.empty(n, dtype=torch.int32).pin_memory()
        return slot

    def encode(self, sentences) -> EncodedBatch:
        ids, key_len, row_of, first_idx, lengths, S = self.build_batch(sentences)
        dev = self.device_
        len_t = torch.tensor(lengths, dtype=torch.int32)
        parts = (ids, key_len, row_of, first_idx, len_t)
        sizes = [t.numel() for t in parts]
        n = sum(sizes)
        slot = self._staging(n)
        host = slot[0][:n]
        off = 0
        for t, k in zip(parts, sizes):
            host[off:off + k] = t.reshape(-1)
            off += k
        packed = host.to(dev, non_blocking=True)
        if dev.type == "cuda":
            ev = torch.cuda.Event()
            ev.record()
            slot[1] = ev
        o0, o1, o2 = sizes[0], sizes[0] + sizes[1], sizes[0] + sizes[1] + sizes[2]
        ids_d = packed[:o0].view(ids.shape)
        key_d, row_d = packed[o0:o1], packed[o1:o2]
        o3 = o2 + sizes[3]
        first_d = packed[o2:o3].view(first_idx.shape)
        len_d = packed[o3:]
        if self.fine_tune and self.training:
            # gradients are enabled iff (fine_tune and self.training), embeddings.py:3280
            hidden, saved = self.model.forward_train(ids_d, key_d)
            eb = EncodedBatch(hidden, S, row_d, first_d, lengths, key_d, ids_d, saved=saved, encoder=self.model)
        else:
            hidden = self.model.forward_hidden(ids_d, key_d)
            eb = EncodedBatch(hidden, S, row_d, first_d, lengths, key_d, ids_d)
        eb.lengths_d = len_d
        return eb

    def embed(self, sentences):
        """Embeddings.embed (:75-101): afterwards ``sentences.features[self.name]`` holds the batch.  Here the
        feature is an EncodedBatch (device resident) instead of a padded [B,T,D] CPU tensor."""
        if not isinstance(sentences, (list, BatchedData)):
            sentences = [sentences]
        if not hasattr(sentences, "features"):
            sentences = BatchedData(sentences)
        # The reference short-circuits here when static embeddings are already stored on the batch (:3030-3037).  Its cache
        # holds real tensors; an EncodedBatch aliases the encoder's workspace / graph output (valid until the next forward),
        # so a cached one may describe ANOTHER batch by now: always re-encode (11 ms for 32 x 512 -- the reference's cache
        # exists because its encoder costs seconds).
        sentences.features[self.name] = self.encode(sentences)
        return sentences

    def materialize_token_embeddings(self, sentences):
        """Optional: write per-token vectors into Token._embeddings (what :3343 does for every token)."""
        enc = sentences.features[self.name]
        H = self.embedding_length
        hid = enc.hidden.float()
        for b, s in enumerate(sentences):
            base = int(enc.row_of[b]) * enc.S
            for t, tok in enumerate(s.tokens):
                fi = int(enc.first_idx[b, t])
                tok.set_embedding(self.name, hid[base + fi].clone() if fi >= 0 else torch.zeros(H, device=hid.device))

    def __getstate__(self):
        state = self.__dict__.copy()
        state["tokenizer"] = None if not isinstance(self.tokenizer, SyntheticTokenizer) else self.tokenizer
        state["_tok_cache"] = {}
        state["_plan_cache"] = {}
        state["_word_cache"] = {}
        state["_piece_cache"] = {}
        state["_piece1"] = {}
        state["_by_word"], state["_verify_left"] = None, 64
        state["_stage"] = None
        return state

    def __setstate__(self, d):
        self.__dict__ = d
        self.__dict__.setdefault("_word_cache", {})
        self.__dict__.setdefault("_piece_cache", {})
        self.__dict__.setdefault("_piece1", {})
        self.__dict__.setdefault("_by_word", None)
        self.__dict__.setdefault("_verify_left", 64)
        if self.tokenizer is None:
            from transformers import AutoTokenizer
            self.tokenizer = AutoTokenizer.from_pretrained(self.name.split("/")[-1])

    def extra_repr(self):
        return "model=%s" % self.name


class StackedEmbeddings(torch.nn.Module):
    """flair/embeddings.py:155-211: the tagger always receives a stack, here of exactly one member."""

    def __init__(self, embeddings: List[TransformerWordEmbeddings]):
        super().__init__()
        if len(embeddings) != 1:
            raise NotImplementedError("the KB-NER configs stack exactly one TransformerWordEmbeddings")
        self.embeddings = embeddings
        for i, e in enumerate(embeddings):
            self.add_module("list_embedding_%d" % i, e)
        self.name = "Stack"
        self.static_embeddings = all(e.static_embeddings for e in embeddings)
        self.embedding_type = embeddings[0].embedding_type
        self.embedding_length = sum(e.embedding_length for e in embeddings)

    def embed(self, sentences, static_embeddings: bool = True, embedding_mask=None):
        for e in self.embeddings:
            sentences = e.embed(sentences)
        return sentences

    def forward(self, sentences):
        return self.embed(sentences)
