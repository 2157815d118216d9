"""The data formats either side of the hot path: the CoNLL column reader, the BIO -> BIOES
conversion and the batch assembler.

Mirrors, for the options the KB-NER configs use
(``config/xlmr-large-...ner20.yaml:46-55``: ``column_format {0: text, 1: pos, 2: upos, 3: ner}``, ``comment_symbol '# id'``,
``tag_to_bioes: ner``):

* ``ColumnDataset`` / ``ColumnCorpus``  -- ``/root/reference/flair/datasets.py:852-1004`` / ``:21-128``
* ``iob2`` / ``iob_iobes`` / ``convert_tag_scheme`` -- ``flair/data.py:1122-1164`` / ``:630-645``
* ``Corpus.make_tag_dictionary`` -- ``flair/data.py:1083-1104``
* ``ColumnDataLoader`` -- ``flair/custom_data_loader.py:25-149`` (sort by word count, chunk by sentence count or word budget,
  ``reshuffle`` / ``true_reshuffle``)

The KB-augmented files carry ``sentence <EOS> retrieved context`` per sentence with the context tokens tagged ``B-X``
(``kb/context_process.py:219-223,424-426``); after BIOES conversion those become ``S-X``, which the tagger's remove-X logic
keys on.  Parity is pinned by ``tests/golden/conll_golden.json`` (the reference's own reader run on ``tests/golden/sample_conll.txt``).
"""
import random
import re
from pathlib import Path
from typing import Dict, List, Optional, Union

from .data import BatchedData, Dictionary, Sentence, Token

_WS = re.compile(r"\s+")


def iob2(tags: List[str]) -> bool:
    """IOB1 -> IOB2 in place; False when a tag is not O / B-x / I-x (the list is then left partially converted, as in the
    reference).  flair/data.py:1122-1142."""
    for i, tag in enumerate(tags):
        if tag == "O":
            continue
        split = tag.split("-")
        if len(split) != 2 or split[0] not in ("I", "B"):
            return False
        if split[0] == "B":
            continue
        if i == 0 or tags[i - 1] == "O":
            tags[i] = "B" + tag[1:]
        elif tags[i - 1][1:] == tag[1:]:
            continue
        else:
            tags[i] = "B" + tag[1:]
    return True


def iob_iobes(tags: List[str]) -> List[str]:
    """IOB2 -> IOBES (flair/data.py:1145-1164).  str.replace semantics kept: 'B-' / 'I-' are replaced wherever they occur."""
    out = []
    for i, tag in enumerate(tags):
        if tag == "O":
            out.append(tag)
        elif tag.split("-")[0] == "B":
            if i + 1 != len(tags) and tags[i + 1].split("-")[0] == "I":
                out.append(tag)
            else:
                out.append(tag.replace("B-", "S-"))
        elif tag.split("-")[0] == "I":
            if i + 1 < len(tags) and tags[i + 1].split("-")[0] == "I":
                out.append(tag)
            else:
                out.append(tag.replace("I-", "E-"))
        else:
            raise ValueError("Invalid IOB format: %r" % tag)
    return out


def convert_tag_scheme(sentence: Sentence, tag_type: str = "ner", target_scheme: str = "iob"):
    """Sentence.convert_tag_scheme (flair/data.py:630-645)."""
    tags = [tok.get_tag(tag_type).value for tok in sentence.tokens]
    if target_scheme == "iob":
        iob2(tags)
    if target_scheme == "iobes":
        iob2(tags)
        tags = iob_iobes(tags)
    for tok, tag in zip(sentence.tokens, tags):
        tok.add_tag(tag_type, tag)


class ColumnDataset:
    """In-memory CoNLL column file.  A line starting with `comment_symbol` is skipped; a whitespace-only line ends a
    sentence; other lines are split on runs of whitespace, column `text` is the token, the other mapped columns become tags
    (a line with fewer fields simply gets fewer tags -- note that the trailing newline yields one empty extra field, so a
    line with exactly one field gets an EMPTY tag for column 1, as in the reference)."""

    def __init__(self, path_to_column_file: Union[str, Path], column_name_map: Dict[int, str], tag_to_bioes: Optional[str] = None,
                 comment_symbol: Optional[str] = None, in_memory: bool = True):
        path = Path(path_to_column_file)
        if not path.exists():
            raise FileNotFoundError(str(path))
        if not in_memory:
            raise NotImplementedError("only in_memory=True (the reference's default and what the configs use)")
        self.path_to_column_file = path
        self.column_name_map = {int(k): v for k, v in column_name_map.items()}
        self.tag_to_bioes = tag_to_bioes
        self.comment_symbol = comment_symbol
        self.in_memory = True
        self.text_column = 0
        for col, name in self.column_name_map.items():
            if name == "text":
                self.text_column = col
        try:
            with open(str(path), encoding="utf-8") as f:
                f.read(10)
            encoding = "utf-8"
        except UnicodeDecodeError:
            encoding = "latin1"
        self.sentences: List[Sentence] = []
        cur = Sentence(tokens=[])
        with open(str(path), encoding=encoding) as f:
            for line in f:
                if comment_symbol is not None and line.startswith(comment_symbol):
                    continue
                if line.isspace():
                    if len(cur) > 0:
                        self._finish(cur)
                    cur = Sentence(tokens=[])
                    continue
                fields = _WS.split(line)
                tok = Token(fields[self.text_column])
                for col, name in self.column_name_map.items():
                    if len(fields) > col and col != self.text_column:
                        tok.add_tag(name, fields[col])
                cur.add_token(tok)
        if len(cur) > 0:
            self._finish(cur)
        self.total_sentence_count = len(self.sentences)

    def _finish(self, sentence):
        if self.tag_to_bioes is not None:
            convert_tag_scheme(sentence, tag_type=self.tag_to_bioes, target_scheme="iobes")
        self.sentences.append(sentence)

    def is_in_memory(self):
        return True

    def __len__(self):
        return self.total_sentence_count

    def __getitem__(self, i):
        return self.sentences[i]

    def __iter__(self):
        return iter(self.sentences)


class Corpus:
    """train / dev / test with the helpers the trainer and the config parser use (flair/data.py:1007-1104)."""

    def __init__(self, train, dev, test, name: str = "corpus"):
        self.train, self.dev, self.test, self.name = train, dev, test, name

    def get_all_sentences(self):
        return list(self.train) + list(self.dev) + list(self.test)

    def make_tag_dictionary(self, tag_type: str) -> Dictionary:
        """'<unk>' (index 0), 'O', then every tag value in order of first appearance over train + dev + test, '<START>', '<STOP>'
        (flair/data.py:1083-1104)."""
        d = Dictionary(add_unk=True)
        d.add_item("O")
        for s in self.get_all_sentences():
            for tok in s.tokens:
                d.add_item(tok.get_tag(tag_type).value)
        d.add_item("<START>")
        d.add_item("<STOP>")
        return d


class ColumnCorpus(Corpus):
    """flair/datasets.py:21-128: explicit file names or discovery by name ('train'; 'dev' / 'testa' -> dev; 'testb' -> test,
    else any 'test'); a missing test / dev split is a random 10 % of train (torch.utils.data.random_split, i.e. torch's
    global generator, like the reference)."""

    def __init__(self, data_folder: Union[str, Path], column_format: Dict[int, str], train_file=None, test_file=None,
                 dev_file=None, tag_to_bioes=None, comment_symbol: Optional[str] = None, in_memory: bool = True):
        folder = Path(data_folder)
        train_file = folder / train_file if train_file is not None else None
        test_file = folder / test_file if test_file is not None else None
        dev_file = folder / dev_file if dev_file is not None else None
        if train_file is None:
            for f in folder.iterdir():
                n = f.name
                if n.endswith(".gz") or n.endswith(".swp") or n.endswith(".pkl"):
                    continue
                if "train" in n:
                    train_file = f
                if "dev" in n:
                    dev_file = f
                if "testa" in n:
                    dev_file = f
                if "testb" in n:
                    test_file = f
            if test_file is None:
                for f in folder.iterdir():
                    if f.name.endswith(".gz"):
                        continue
                    if "test" in f.name:
                        test_file = f
        if train_file is None:
            raise FileNotFoundError("no train file under %s" % folder)
        read = lambda p: ColumnDataset(p, column_format, tag_to_bioes, comment_symbol=comment_symbol, in_memory=in_memory)
        train = read(train_file)
        if test_file is not None:
            test = read(test_file)
        else:
            train, test = self._split(train)
        if dev_file is not None:
            dev = read(dev_file)
        else:
            train, dev = self._split(train)
        super().__init__(train, dev, test, name=folder.name)

    @staticmethod
    def _split(train):
        from torch.utils.data import random_split
        n = len(train)
        k = round(n / 10)
        a, b = random_split(train, [n - k, k])
        return a, b


class ListCorpus(Corpus):
    """Several corpora side by side (flair/list_data.py:2-19): `train_list / dev_list / test_list` keep the per-corpus
    splits (final_test reports per corpus, `targets` names them), `train / dev / test` are their concatenations.
    Also accepts flat lists of sentences (one anonymous corpus)."""

    def __init__(self, train, dev=None, test=None, name: str = "listcorpus", targets=None):
        def as_lists(x):
            x = list(x) if x is not None else []
            if x and hasattr(x[0], "tokens"):          # a flat list of sentences
                return [x]
            return [list(d) for d in x]
        self.train_list, self.dev_list, self.test_list = as_lists(train), as_lists(dev), as_lists(test)
        cat = lambda ls: [s for d in ls for s in d]
        super().__init__(cat(self.train_list), cat(self.dev_list), cat(self.test_list), name=name)
        self.targets = list(targets) if targets is not None else ["corpus-%d" % i for i in range(len(self.train_list))]

    def get_train_full_tokenset(self, max_tokens: int = -1, min_freq: int = -1, attr: str = "text"):
        """[tokens of train+test sorted by frequency, character set] (flair/data.py:1018-1050 as the config parser calls it;
        only the FastWord / character embeddings, which are outside the hot path, consume it)."""
        from collections import Counter
        cnt = Counter(getattr(t, attr, t.text) for s in self.train + self.test for t in s.tokens)
        toks = [w for w, c in cnt.most_common() if c > min_freq]
        if max_tokens > 0:
            toks = toks[:max_tokens]
        return [toks, sorted({ch for w in toks for ch in w})]


class ColumnDataLoader:
    """Batch assembler (flair/custom_data_loader.py:25-149): sentences stably sorted by WORD count (use_bert=False is what
    the trainer passes for TransformerWordEmbeddings), then chunked either by sentence count (`sentence_level_batch`, the
    KB-NER setting) or by a word budget; batches are `BatchedData`.  `reshuffle` permutes the batches, `true_reshuffle`
    re-chunks first; both use Python's global `random` like the reference (seed it on every rank for data-parallel runs).
    A sentence longer than the word budget starts a new batch; the reference would first emit the EMPTY current batch in
    that case -- empty batches are not emitted here."""

    def __init__(self, data, batch_size: int, shuffle: bool = False, args=None, grouped_data: bool = False, use_bert: bool = False,
                 tokenizer=None, sort_data: bool = True, sentence_level_batch: bool = False, model=None):
        if grouped_data or use_bert:
            raise NotImplementedError("grouped_data / use_bert batching is not used by the KB-NER configs")
        if sentence_level_batch and batch_size > 500:
            raise ValueError("batch size too large for sentence-level batching (the reference asserts here)")
        self.batch_size = batch_size
        self.shuffled = shuffle
        self.sentence_level_batch = sentence_level_batch
        self.sort_data = sort_data
        self.model = model
        data = list(data)
        self.num_examples = len(data)
        self.data = self.chunk_batches(data, sort_data=sort_data)

    def chunk_batches(self, data, sort_data: bool = True):
        if sort_data:
            data = sorted(data, key=len)
        res, cur, curlen = [], [], 0
        for x in data:
            if self.sentence_level_batch:
                full = len(cur) >= self.batch_size
            else:
                full = len(x) + curlen > self.batch_size
            if full:
                if cur:
                    res.append(BatchedData(cur))
                cur, curlen = [], 0
            cur.append(x)
            curlen += len(x)
        if curlen > 0:
            res.append(BatchedData(cur))
        return res

    def assign_tags(self, tag_type: str, tag_dictionary, teacher_input=None, grouped_data: bool = False):
        """Gold tag indices once per sentence (flair/custom_data_loader.py:199-382, the part the CRF loss reads): every
        sentence gets `<tag_type>_tags` (int64, CPU), every batch a zero-padded `[B, T]` tensor of the same name
        (0 = '<unk>' is the padding value, Appendix B.13)."""
        import torch
        if grouped_data:
            raise NotImplementedError("grouped_data batches are not used by the KB-NER configs")
        batches = [teacher_input] if teacher_input is not None else self.data
        for batch in batches:
            rows = []
            for s in batch:
                idx = torch.tensor([tag_dictionary.get_idx_for_item(t.get_tag(tag_type).value) for t in s.tokens],
                                   dtype=torch.int64)
                setattr(s, tag_type + "_tags", idx)
                rows.append(idx)
            T = max((r.numel() for r in rows), default=0)
            padded = torch.zeros((len(rows), T), dtype=torch.int64)
            for i, r in enumerate(rows):
                padded[i, :r.numel()] = r
            setattr(batch, tag_type + "_tags", padded)
        return self

    def reshuffle(self):
        random.shuffle(self.data)

    def true_reshuffle(self):
        self.data = self.chunk_batches([s for b in self.data for s in b], sort_data=self.sort_data)
        random.shuffle(self.data)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        if not isinstance(i, int):
            raise TypeError
        if i < 0 or i >= len(self.data):
            raise IndexError
        return self.data[i]

    def __iter__(self):
        return iter(self.data)
