#!/usr/bin/env python
"""TEST INFRASTRUCTURE: runs the REFERENCE's own CoNLL reader, tag dictionary builder and batch assembler
(/root/reference/flair/datasets.py ColumnDataset, flair/data.py make_tag_dictionary, flair/custom_data_loader.py
ColumnDataLoader) on tests/golden/sample_conll.txt and writes tests/golden/conll_golden.json.
Run in the build container only (the reference tree is not on the GPU box):  python oracle/make_golden_conll.py"""
import json
import os
import random
import sys
from pathlib import Path

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

SAMPLE = os.path.join(ROOT, "tests", "golden", "sample_conll.txt")
OUT = os.path.join(ROOT, "tests", "golden", "conll_golden.json")


def write_sample():
    rnd = random.Random(7)
    types = ["PER", "LOC", "CORP", "GRP", "PROD", "CW"]
    lines = []
    for s in range(23):
        lines.append("# id %08x\tdomain=en" % rnd.getrandbits(32))
        n = rnd.randint(1, 14)
        i = 0
        while i < n:
            if rnd.random() < 0.35:
                t = rnd.choice(types)
                ln = rnd.randint(1, 4)
                first = "I" if rnd.random() < 0.2 else "B"          # IOB1-style openings exercise iob2()
                for k in range(ln):
                    lines.append("w%d_%d _ _ %s-%s" % (s, i, first if k == 0 else "I", t))
                    i += 1
            else:
                sep = "\t" if rnd.random() < 0.3 else " "
                lines.append(sep.join(["w%d_%d" % (s, i), "_", "_", "O"]))
                i += 1
        if rnd.random() < 0.6:                                       # KB context: <EOS> then B-X tokens
            lines.append("<EOS> _ _ B-X")
            for k in range(rnd.randint(1, 12)):
                lines.append("ctx%d_%d _ _ B-X" % (s, k))
        lines.append("" if rnd.random() < 0.7 else "   ")
        if rnd.random() < 0.2:
            lines.append("")                                         # doubled blank line
    lines.append("-DOCSTART- -X- -X- O")
    lines.append("")
    lines.append("last _ _ B-LOC")
    lines.append("one _ _ I-LOC")                                    # no trailing blank line / newline at EOF
    with open(SAMPLE, "w", encoding="utf-8") as f:
        f.write("\n".join(lines))


def main():
    write_sample()
    ref_shim.load_flair()
    import flair.datasets as D
    from flair.data import Corpus
    from flair.custom_data_loader import ColumnDataLoader
    fmt = {0: "text", 1: "pos", 2: "upos", 3: "ner"}
    out = {}
    for name, kw in (("bioes", dict(tag_to_bioes="ner", comment_symbol="# id")), ("raw", dict(comment_symbol="# id"))):
        ds = D.ColumnDataset(Path(SAMPLE), fmt, **kw)
        out[name] = [{"text": [t.text for t in s.tokens], "ner": [t.get_tag("ner").value for t in s.tokens],
                      "pos": [t.get_tag("pos").value for t in s.tokens],
                      "spans": [[sp.tag, sp.tokens[0].idx - 1, sp.tokens[-1].idx, sp.text] for sp in s.get_spans("ner")]}
                     for s in ds.sentences]
    ds = D.ColumnDataset(Path(SAMPLE), fmt, tag_to_bioes="ner", comment_symbol="# id")
    corpus = Corpus(ds.sentences[:15], ds.sentences[15:20], ds.sentences[20:])
    d = corpus.make_tag_dictionary("ner")
    out["tag_dictionary"] = [x.decode("utf-8") for x in d.idx2item]
    sents = list(ds.sentences)
    for mode, kw in (("sentence_level_4", dict(batch_size=4, sentence_level_batch=True)),
                     ("word_budget_30", dict(batch_size=30)), ("unsorted_5", dict(batch_size=5, sentence_level_batch=True, sort_data=False))):
        loader = ColumnDataLoader(sents, **kw)
        out["loader_" + mode] = [[sents.index(s) for s in batch] for batch in loader.data]
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(out, f, indent=0)
    print("wrote", OUT, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
