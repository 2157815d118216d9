"""Minimal data model of the hot path's callers: Dictionary / Label / Token / Sentence / BatchedData.

Mirrors the parts of ``/root/reference/flair/data.py`` (Dictionary :21-101, Label :104-142, Token :164-262,
Sentence :340-) and ``flair/custom_data_loader.py`` (BatchedData :13-20) that the hot path reads, with the
same method names and on-disk formats (the tag-dictionary pickles in resources/taggers/*.pkl load unchanged).
When the reference's own ``flair.data`` objects are passed in instead, the hot path only relies on the
attributes used here (duck typing).
"""
import pickle
from typing import Dict, List, Optional


class Dictionary:
    """String <-> id map with byte-string keys (flair/data.py:21-101)."""

    def __init__(self, add_unk=True):
        self.item2idx: Dict[bytes, int] = {}
        self.idx2item: List[bytes] = []
        self.multi_label = False
        if add_unk:
            self.add_item("<unk>")

    def add_item(self, item: str) -> int:
        b = item.encode("utf-8")
        if b not in self.item2idx:
            self.idx2item.append(b)
            self.item2idx[b] = len(self.idx2item) - 1
        return self.item2idx[b]

    def get_idx_for_item(self, item: str) -> int:
        return self.item2idx.get(item.encode("utf-8"), 0)     # unknown -> 0 (<unk>)

    def get_items(self) -> List[str]:
        return [i.decode("utf-8") for i in self.idx2item]

    def get_item_for_index(self, idx) -> str:
        return self.idx2item[int(idx)].decode("utf-8")

    def __len__(self):
        return len(self.idx2item)

    def save(self, path):
        with open(path, "wb") as f:
            pickle.dump({"idx2item": self.idx2item, "item2idx": self.item2idx}, f)

    @classmethod
    def load_from_file(cls, path):
        d = cls(add_unk=False)
        with open(path, "rb") as f:
            m = pickle.load(f, encoding="latin1")
        d.idx2item, d.item2idx = m["idx2item"], m["item2idx"]
        return d

    @classmethod
    def make_tag_dictionary(cls, tags, with_x=True):
        """<unk>, O, tags..., [S-X], <START>, <STOP>  -- the order Corpus.make_tag_dictionary produces
        (flair/data.py:1083-1104)."""
        d = cls(add_unk=True)
        d.add_item("O")
        for t in tags:
            d.add_item(t)
        if with_x:
            d.add_item("S-X")
        d.add_item("<START>")
        d.add_item("<STOP>")
        return d


class Label:
    __slots__ = ("value", "score")

    def __init__(self, value: Optional[str], score: float = 1.0):
        self.value = value if value else ""
        # score clamped to [0, 1] (flair/data.py:118-125)
        self.score = score if 0.0 <= score <= 1.0 else (1.0 if score > 1.0 else 0.0)

    def to_dict(self):
        return {"value": self.value, "confidence": self.score}

    def __repr__(self):
        return "%s (%.4f)" % (self.value, self.score)

    def __eq__(self, other):
        return isinstance(other, Label) and self.value == other.value and self.score == other.score


class LabelSeq:
    """Read-only sequence of Labels for one sentence, backed by the decoded tag indices / confidences of the batch.
    Label objects are created when an element is accessed (a batch of 32 x 510 tokens is 16k Python objects -- 8 ms of
    pure interpreter time per batch if built eagerly, which was more than half of the GPU time of the whole batch)."""
    __slots__ = ("_names", "_tags", "_conf")

    def __init__(self, names, tags, conf):
        self._names, self._tags, self._conf = names, tags, conf

    def __len__(self):
        return len(self._tags)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [Label(self._names[a], c) for a, c in zip(self._tags[i], self._conf[i])]
        return Label(self._names[self._tags[i]], self._conf[i])

    def __iter__(self):
        names = self._names
        for a, c in zip(self._tags, self._conf):
            yield Label(names[a], c)

    def tag_indices(self):
        return self._tags

    def __eq__(self, other):
        return list(self) == list(other)

    def __repr__(self):
        return repr(list(self))


class Token:
    def __init__(self, text: str, idx: Optional[int] = None):
        self.text = text
        self.idx = idx
        self.tags: Dict[str, Label] = {}
        self._embeddings = {}

    def add_tag(self, tag_type: str, tag_value: str, confidence=1.0):
        self.tags[tag_type] = Label(tag_value, confidence)

    def add_tag_label(self, tag_type: str, label: Label):
        self.tags[tag_type] = label

    def get_tag(self, tag_type: str) -> Label:
        return self.tags.get(tag_type, Label(""))

    def set_embedding(self, name, vector):
        self._embeddings[name] = vector

    def clear_embeddings(self, names=None):
        if names is None:
            self._embeddings = {}
        else:
            for n in names:
                self._embeddings.pop(n, None)

    def __repr__(self):
        return "Token: %s %s" % (self.idx, self.text)


class Sentence:
    """Whitespace-tokenised sentence (the CoNLL reader of the reference always supplies tokens)."""

    def __init__(self, text: Optional[str] = None, tokens: Optional[List[str]] = None):
        self.tokens: List[Token] = []
        words = tokens if tokens is not None else (text.split() if text else [])
        for w in words:
            self.add_token(w)

    def add_token(self, token):
        if isinstance(token, str):
            token = Token(token)
        token.idx = len(self.tokens) + 1
        self.tokens.append(token)

    def to_tokenized_string(self) -> str:
        return " ".join(t.text for t in self.tokens)

    def clear_embeddings(self, names=None):
        for t in self.tokens:
            t.clear_embeddings(names)

    def get_spans(self, tag_type: str, min_score=-1):
        """(type, start, end_exclusive, text) of every span, with the reference's rules (flair/data.py:455-532), which
        matter for MALFORMED predicted sequences: anything that is not B-/I-/O-/E-/S- is a single-token tag; only B- and
        S- (or a type change right after an S-) open a new span -- an E- does not close one and an I- of another type
        does not split one; a span's type is the weighted majority of its tags (weight 1.1 for the tag that opened it,
        first-seen wins ties); spans whose mean tag confidence is not above min_score are dropped."""
        spans, cur, weights = [], [], {}
        prev_prefix, prev_type = "O-", ""

        def emit():
            nonlocal cur, weights
            if cur:
                scores = [self.tokens[i].get_tag(tag_type).score for i in cur]
                if sum(scores) / len(scores) > min_score:
                    best = max(weights.values())
                    typ = next(t for t, w in weights.items() if w == best)      # dicts keep insertion order
                    spans.append((typ, cur[0], cur[-1] + 1, " ".join(self.tokens[i].text for i in cur)))
            cur, weights = [], {}

        for i, tok in enumerate(self.tokens):
            v = tok.get_tag(tag_type).value
            if v == "" or v == "O":
                v = "O-"
            if v[0:2] not in ("B-", "I-", "O-", "E-", "S-"):
                v = "S-" + v
            prefix, typ = v[0:2], v[2:]
            in_span = prefix != "O-"
            opens = prefix in ("B-", "S-") or (prev_prefix == "S-" and prev_type != typ and in_span)
            if opens or not in_span:
                emit()
            if in_span:
                cur.append(i)
                weights[typ] = weights.get(typ, 0.0) + (1.1 if opens else 1.0)
            prev_prefix, prev_type = prefix, typ
        emit()
        return spans

    def __getitem__(self, i):
        return self.tokens[i]

    def __iter__(self):
        return iter(self.tokens)

    def __len__(self):
        return len(self.tokens)


class BatchedData(list):
    """list of sentences + per-batch feature cache (flair/custom_data_loader.py:13-20)."""

    def __init__(self, items):
        super().__init__(items)
        self.features = {}
        self.img_features = {}
        self.teacher_features = {}
        self.sentence_features = {}
