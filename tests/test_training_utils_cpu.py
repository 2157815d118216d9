"""Metric / Result / get_spans against the known-answer values of the reference's OWN tests
(/root/reference/tests/test_utils.py:7-95 and tests/test_data.py:468-574): same inputs, same expected numbers."""
from kbner_b200.data import Sentence
from kbner_b200.training_utils import Metric, Result, span_counts


def test_metric_get_classes():                       # test_utils.py:7-18
    m = Metric("Test")
    m.add_fn("class-1"); m.add_fn("class-3"); m.add_tn("class-1"); m.add_tp("class-2")
    assert m.get_classes() == ["class-1", "class-2", "class-3"]


def test_metric_with_classes():                      # test_utils.py:43-95
    m = Metric("Test")
    for c in ("class-1", "class-2", "class-4"):
        m.add_tp(c); m.add_tn(c); m.add_tn(c); m.add_fp(c)
    for _ in range(10):
        m.add_tp("class-3")
    for _ in range(90):
        m.add_fp("class-3")
    assert [m.precision(c) for c in m.get_classes()] == [0.5, 0.5, 0.1, 0.5]
    assert [m.recall(c) for c in m.get_classes()] == [1, 1, 1, 1]
    assert m.accuracy() == m.micro_avg_accuracy() and m.f_score() == m.micro_avg_f_score()
    assert [m.f_score(c) for c in m.get_classes()] == [0.6667, 0.6667, 0.1818, 0.6667]
    assert [m.accuracy(c) for c in m.get_classes()] == [0.5, 0.5, 0.1, 0.5]
    assert m.micro_avg_f_score() == 0.2184
    assert m.macro_avg_f_score() == 0.5454749999999999
    assert m.micro_avg_accuracy() == 0.1226 and m.macro_avg_accuracy() == 0.4
    assert m.precision() == 0.1226 and m.recall() == 1
    r = m.to_result()
    assert isinstance(r, Result) and r.main_score == 0.2184 and r.macro_score == 0.5454749999999999
    assert r.log_header == "PRECISION\tRECALL\tF1" and r.log_line == "0.1226\t1.0\t0.2184"
    assert r["tp"] == 13 and r["fp"] == 93 and r["fn"] == 0 and r["main_score"] == 0.2184
    assert "class-3    tp: 10 - fp: 90 - fn: 0 - tn: 0 - precision: 0.1000" in r.detailed_results
    # the flat-vector round trip the multi-GPU evaluation all-reduces
    classes = m.get_classes()
    m2 = Metric.from_vector("x", classes, [2 * v for v in m.to_vector(classes)])
    assert m2.get_tp() == 26 and m2.micro_avg_f_score() == m.micro_avg_f_score()


def _tag(sentence, tags, scores=None):
    for i, t in tags.items():
        sentence[i].add_tag("ner", t, 1.0 if scores is None else scores[i])


def test_spans_reference_cases():                    # test_data.py:468-574
    s = Sentence("Zalando Research is located in Berlin .")
    _tag(s, {0: "B-ORG", 1: "E-ORG", 5: "S-LOC"})                                    # bioes
    assert [(t, x) for t, _, _, x in s.get_spans("ner")] == [("ORG", "Zalando Research"), ("LOC", "Berlin")]
    _tag(s, {0: "B-ORG", 1: "I-ORG", 5: "B-LOC"})                                    # bio
    assert [(t, x) for t, _, _, x in s.get_spans("ner")] == [("ORG", "Zalando Research"), ("LOC", "Berlin")]
    _tag(s, {0: "I-ORG", 1: "E-ORG", 5: "I-LOC"})                                    # broken
    assert [(t, x) for t, _, _, x in s.get_spans("ner")] == [("ORG", "Zalando Research"), ("LOC", "Berlin")]
    _tag(s, {0: "I-ORG", 1: "E-ORG", 2: "aux", 3: "verb", 4: "preposition", 5: "I-LOC"})    # all tags
    sp = s.get_spans("ner")
    assert len(sp) == 5 and sp[0][0::3] == ("ORG", "Zalando Research") and sp[4][0::3] == ("LOC", "Berlin")
    _tag(s, {0: "I-ORG", 1: "S-LOC", 2: "aux", 3: "B-relation", 4: "E-preposition", 5: "S-LOC"})   # all weird tags
    sp = s.get_spans("ner")
    assert len(sp) == 5
    assert sp[0][0::3] == ("ORG", "Zalando") and sp[1][0::3] == ("LOC", "Research")
    assert sp[3][0::3] == ("relation", "located in")
    s = Sentence("A woman was charged on Friday with terrorist offences after three Irish Republican Army mortar "
                 "bombs were found in a Belfast house , police said . ")
    _tag(s, {11: "S-MISC", 12: "B-MISC", 13: "E-MISC"})
    assert [x for _, _, _, x in s.get_spans("ner")] == ["Irish", "Republican Army"]
    s = Sentence("Zalando Research is located in Berlin .")                          # confidences
    _tag(s, {0: "B-ORG", 1: "E-ORG", 5: "S-LOC"}, {0: 1.0, 1: 0.9, 5: 0.5})
    assert len(s.get_spans("ner", min_score=0.0)) == 2
    assert [x for _, _, _, x in s.get_spans("ner", min_score=0.6)] == ["Zalando Research"]
    assert s.get_spans("ner", min_score=0.99) == []


def test_span_counts_remove_x():
    """sequence_tagger_model.py:2644-2686: predicted spans touching a gold S-X token and gold spans of type X do not count."""
    s = Sentence("Marie Curie won <EOS> Curie was Polish")
    gold = {0: "B-PER", 1: "E-PER", 2: "O", 3: "S-X", 4: "S-X", 5: "S-X", 6: "S-X"}
    pred = {0: "B-PER", 1: "E-PER", 2: "S-LOC", 3: "O", 4: "S-PER", 5: "O", 6: "O"}
    for i in gold:
        s[i].add_tag("ner", gold[i])
        s[i].add_tag("predicted", pred[i])
    gold_x = [gold[i] == "S-X" for i in range(7)]
    m = Metric("e")
    span_counts(m, s.get_spans("ner"), s.get_spans("predicted"), gold_x, remove_x=True)
    assert (m.get_tp(), m.get_fp(), m.get_fn()) == (1, 1, 0) and m.get_classes() == ["LOC", "PER"]
    m = Metric("e")
    span_counts(m, s.get_spans("ner"), s.get_spans("predicted"), None, remove_x=False)
    # without the filter the context prediction is a false positive and the X spans are unmatched gold spans
    assert m.get_tp() == 1 and m.get_fp() == 2 and m.get_fn("X") == 4
