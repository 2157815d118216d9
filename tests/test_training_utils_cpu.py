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


def test_linear_schedule_matches_transformers():
    """finetune_trainer.py:679-688 builds transformers' get_linear_schedule_with_warmup(optimizer, 0, t_total) and steps it
    after every optimizer step; FusedAdamW.scheduler_step must produce the same learning-rate sequence for both groups."""
    import pytest
    import torch
    from transformers import get_linear_schedule_with_warmup
    from kbner_b200.encoder import ParamArena
    from kbner_b200.optim import FusedAdamW
    t_total, base, rate = 7, 5e-6, 10000.0
    p = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))]
    ref_opt = torch.optim.SGD([{"params": [p[0]], "lr": base * rate}, {"params": [p[1]], "lr": base}], lr=base)
    sched = get_linear_schedule_with_warmup(ref_opt, num_warmup_steps=0, num_training_steps=t_total)
    q = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))]
    opt = FusedAdamW([{"arena": ParamArena([q[0]]), "lr": base * rate}, {"arena": ParamArena([q[1]]), "lr": base}])
    opt.set_linear_schedule(t_total)
    for _ in range(t_total + 2):                  # two steps past the end: the factor clamps at zero
        want = [g["lr"] for g in ref_opt.param_groups]
        got = [g["lr"] for g in opt.groups]
        assert got == pytest.approx(want, rel=1e-12, abs=1e-18)
        ref_opt.step()
        sched.step()
        opt.steps += 1                            # what FusedAdamW.step() does before its kernels (CUDA only)
        opt.scheduler_step()


def test_parameter_groups_follow_the_reference_name_rule():
    """finetune_trainer.py:552-553 groups parameters by NAME: 'embedding' in the name, or linear.weight / linear.bias ->
    learning_rate; everything else (the CRF transitions) -> learning_rate * lr_rate.  Applied to OUR tagger's
    named_parameters() that rule must give exactly the arenas build_reference_optimizer creates, with every parameter in
    exactly one of them."""
    import torch
    from kbner_b200.data import Dictionary
    from kbner_b200.embeddings import StackedEmbeddings, SyntheticTokenizer, TransformerWordEmbeddings
    from kbner_b200.encoder import EncoderConfig
    from kbner_b200.optim import build_reference_optimizer
    from kbner_b200.sequence_tagger import FastSequenceTagger
    cfg = EncoderConfig(name="t", vocab_size=64, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                        intermediate_size=256, max_position_embeddings=40)
    emb = TransformerWordEmbeddings(model="t", layers="-1", pooling_operation="first", fine_tune=True,
                                    tokenizer=SyntheticTokenizer(64), config=cfg, device="cpu")
    d = Dictionary.make_tag_dictionary(["B-A", "I-A", "E-A", "S-A"], with_x=True)
    tagger = FastSequenceTagger(hidden_size=256, embeddings=StackedEmbeddings([emb]), tag_dictionary=d, tag_type="ner",
                                use_crf=True, use_rnn=False, word_dropout=0.0, locked_dropout=0.0)
    named = list(tagger.named_parameters())
    finetune = {n for n, _ in named if "embedding" in n or n == "linear.weight" or n == "linear.bias"}      # :552
    other = {n for n, _ in named if "embedding" not in n and n != "linear.weight" and n != "linear.bias"}   # :553
    assert other == {"transitions"} and finetune | other == {n for n, _ in named}
    lr, rate = 5e-6, 10000.0
    opt = build_reference_optimizer(tagger, lr=lr, lr_rate=rate)
    by_ptr = {}
    for g in opt.groups:
        ar = g["arena"]
        lo, hi = ar.flat.data_ptr(), ar.flat.data_ptr() + ar.flat.numel() * 4
        by_ptr[(lo, hi)] = g["lr"]
    for n, p in tagger.named_parameters():                   # after the arenas took the parameters over
        owners = [v for (lo, hi), v in by_ptr.items() if lo <= p.data_ptr() < hi]
        assert len(owners) == 1, n
        assert owners[0] == (lr * rate if n in other else lr), n
        assert p.grad is not None and p.grad.shape == p.shape        # .grad is a view into the arena's gradient buffer
