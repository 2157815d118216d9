"""Span-level evaluation bookkeeping with the reference's observable behaviour.

Mirrors what SequenceTagger.evaluate and the trainers read from /root/reference/flair/training_utils.py:
`Result` (:15-23: main_score / log_header / log_line / detailed_results / macro_score), `Metric` (:26-190: per-class
tp / fp / fn / tn counters; precision, recall, F1 and accuracy each rounded to 4 places BEFORE they are combined, so the
F1 is the F1 of the rounded P and R; macro F1 = plain mean of the rounded per-class F1, not rounded again), and
`EvaluationMetric` (:240-248).  The known-answer values of the reference's own tests (tests/test_utils.py:7-95) are
checked in tests/test_training_utils_cpu.py.
"""
from collections import Counter
from enum import Enum
from typing import Iterable, List, Optional

__all__ = ["Result", "Metric", "EvaluationMetric", "store_embeddings", "span_counts"]


class Result:
    def __init__(self, main_score: float, log_header: str, log_line: str, detailed_results: str,
                 macro_score: Optional[float] = None, counts=None):
        self.main_score = main_score
        self.log_header = log_header
        self.log_line = log_line
        self.detailed_results = detailed_results
        self.macro_score = macro_score
        self.counts = dict(counts or {})       # tp / fp / fn totals: what a multi-GPU run all-reduces

    # dict-style access kept for callers written against the first version of evaluate()
    def __getitem__(self, key):
        if key in self.counts:
            return self.counts[key]
        return getattr(self, key)

    def get(self, key, default=None):
        try:
            return self[key]
        except AttributeError:
            return default

    def __repr__(self):
        return "Result(main_score=%s, log_line=%r)" % (self.main_score, self.log_line)


def _ratio(num: int, den: int) -> float:
    return round(num / den, 4) if den > 0 else 0.0


class Metric:
    """Four counters per class; `None` as class name means "over all classes"."""

    _KINDS = ("tp", "fp", "fn", "tn")

    def __init__(self, name: str):
        self.name = name
        self._c = {k: Counter() for k in self._KINDS}

    def _add(self, kind, class_name, n=1):
        self._c[kind][class_name] += n

    def add_tp(self, class_name): self._add("tp", class_name)
    def add_fp(self, class_name): self._add("fp", class_name)
    def add_fn(self, class_name): self._add("fn", class_name)
    def add_tn(self, class_name): self._add("tn", class_name)

    def _get(self, kind, class_name=None) -> int:
        if class_name is None:
            return sum(self._c[kind][c] for c in self.get_classes())
        return self._c[kind][class_name]

    def get_tp(self, class_name=None): return self._get("tp", class_name)
    def get_fp(self, class_name=None): return self._get("fp", class_name)
    def get_fn(self, class_name=None): return self._get("fn", class_name)
    def get_tn(self, class_name=None): return self._get("tn", class_name)

    def get_classes(self) -> List[str]:
        seen = set()
        for k in self._KINDS:
            seen.update(c for c in self._c[k] if c is not None)
        return sorted(seen)

    def precision(self, class_name=None) -> float:
        tp = self.get_tp(class_name)
        return _ratio(tp, tp + self.get_fp(class_name))

    def recall(self, class_name=None) -> float:
        tp = self.get_tp(class_name)
        return _ratio(tp, tp + self.get_fn(class_name))

    def f_score(self, class_name=None) -> float:
        p, r = self.precision(class_name), self.recall(class_name)
        return round(2 * (p * r) / (p + r), 4) if p + r > 0 else 0.0

    def accuracy(self, class_name=None) -> float:
        tp = self.get_tp(class_name)
        return _ratio(tp, tp + self.get_fp(class_name) + self.get_fn(class_name))

    def micro_avg_f_score(self) -> float:
        return self.f_score(None)

    def macro_avg_f_score(self) -> float:
        scores = [self.f_score(c) for c in self.get_classes()]
        return sum(scores) / len(scores) if scores else 0.0

    def micro_avg_accuracy(self) -> float:
        return self.accuracy(None)

    def macro_avg_accuracy(self) -> float:
        acc = [self.accuracy(c) for c in self.get_classes()]
        return round(sum(acc) / len(acc), 4) if acc else 0.0

    def merge(self, other: "Metric") -> "Metric":
        for k in self._KINDS:
            self._c[k].update(other._c[k])
        return self

    def to_tsv(self) -> str:
        return "{}\t{}\t{}\t{}".format(self.precision(), self.recall(), self.accuracy(), self.micro_avg_f_score())

    @staticmethod
    def tsv_header(prefix=None) -> str:
        cols = ("PRECISION", "RECALL", "ACCURACY", "F-SCORE")
        return "\t".join(("%s_%s" % (prefix, c)) if prefix else c for c in cols)

    @staticmethod
    def to_empty_tsv() -> str:
        return "\t_\t_\t_\t_"

    def _line(self, label, c) -> str:
        return ("{0:<10}\ttp: {1} - fp: {2} - fn: {3} - tn: {4} - precision: {5:.4f} - recall: {6:.4f} - "
                "accuracy: {7:.4f} - f1-score: {8:.4f}").format(label, self.get_tp(c), self.get_fp(c), self.get_fn(c),
                                                                 self.get_tn(c), self.precision(c), self.recall(c),
                                                                 self.accuracy(c), self.f_score(c))

    def __str__(self):
        return "\n".join([self._line(self.name, None)] + [self._line(c, c) for c in self.get_classes()])

    # ---- what SequenceTagger.evaluate prints (sequence_tagger_model.py:2708-2729) -----------------------------
    def detailed_results(self) -> str:
        out = ("\nMICRO_AVG: acc %s - f1-score %s\nMACRO_AVG: acc %s - f1-score %s"
               % (self.micro_avg_accuracy(), self.micro_avg_f_score(), self.macro_avg_accuracy(), self.macro_avg_f_score()))
        for c in self.get_classes():
            out += ("\n{0:<10} tp: {1} - fp: {2} - fn: {3} - tn: {4} - precision: {5:.4f} - recall: {6:.4f} - "
                    "accuracy: {7:.4f} - f1-score: {8:.4f}").format(c, self.get_tp(c), self.get_fp(c), self.get_fn(c),
                                                                     self.get_tn(c), self.precision(c), self.recall(c),
                                                                     self.accuracy(c), self.f_score(c))
        return out

    def to_result(self) -> Result:
        return Result(main_score=self.micro_avg_f_score(),
                      log_line="%s\t%s\t%s" % (self.precision(), self.recall(), self.micro_avg_f_score()),
                      log_header="PRECISION\tRECALL\tF1", detailed_results=self.detailed_results(),
                      macro_score=self.macro_avg_f_score(),
                      counts={"tp": self.get_tp(), "fp": self.get_fp(), "fn": self.get_fn()})

    # ---- multi-GPU: counters as a flat vector in a fixed class order ------------------------------------------
    def to_vector(self, classes: Iterable[str]) -> List[int]:
        return [self._c[k][c] for c in classes for k in self._KINDS]

    @classmethod
    def from_vector(cls, name: str, classes: Iterable[str], vec: Iterable[int]) -> "Metric":
        m, it = cls(name), iter(vec)
        for c in classes:
            for k in cls._KINDS:
                n = int(next(it))
                if n:
                    m._add(k, c, n)
        return m


class EvaluationMetric(Enum):
    MICRO_ACCURACY = "micro-average accuracy"
    MICRO_F1_SCORE = "micro-average f1-score"
    MACRO_ACCURACY = "macro-average accuracy"
    MACRO_F1_SCORE = "macro-average f1-score"
    MEAN_SQUARED_ERROR = "mean squared error"


def span_counts(metric: Metric, gold_spans, predicted_spans, gold_is_x=None, remove_x: bool = False) -> None:
    """One sentence of SequenceTagger.evaluate's matching (sequence_tagger_model.py:2644-2686).

    A span is (type, start, end, text) with `start:end` a token range.  With remove_x, predicted spans that touch a
    token whose GOLD tag is S-X are dropped and gold spans of type X are dropped.  A predicted span is a true positive
    when an identical gold span exists, else a false positive; a gold span without an identical prediction is a false
    negative, with one a "true negative" (the reference's naming)."""
    pred = list(predicted_spans)
    gold = list(gold_spans)
    if remove_x:
        pred = [s for s in pred if not any(gold_is_x[s[1]:s[2]])]
        gold = [s for s in gold if s[0] != "X"]
    gold_set, pred_set = set(gold), set(pred)
    for s in pred:
        (metric.add_tp if s in gold_set else metric.add_fp)(s[0])
    for s in gold:
        (metric.add_tn if s in pred_set else metric.add_fn)(s[0])


class frozen_gc:
    """Context manager around the batch loops (evaluate / predict / a training epoch): everything alive at entry -- the corpus'
    Sentence / Token objects, the tokenisation caches, the model -- is moved to the collector's permanent generation
    (gc.freeze), so the collections that the per-batch Label / list / tuple churn triggers scan only what the loop itself
    created.  Measured on the cold-sentence speed test: 94 -> 28 ms of host time per 32 x 510-word batch (a full collection
    over a heap holding a corpus walks millions of objects).  Restored on exit."""

    def __enter__(self):
        import gc
        self._gc = gc
        gc.freeze()
        return self

    def __exit__(self, *exc):
        self._gc.unfreeze()
        return False


def store_embeddings(sentences, storage_mode: str) -> None:
    """store_embeddings (:331-358): with 'none' every per-token embedding is dropped after the batch."""
    if storage_mode == "none":
        for s in sentences:
            s.clear_embeddings()
