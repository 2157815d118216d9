"""ModelFinetuner (finetune_trainer.py:379-1348 step semantics) on a small synthetic corpus: loss goes down, the
checkpoints round-trip through save / load with identical predictions, final_test accepts train.py's keywords."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_model_finetuner_trains_saves_and_reloads(tmp_path):
    from test_api_gpu import SMALL, _models, _sentences
    from kbner_b200.data import BatchedData
    from kbner_b200.trainer import ListCorpus, ModelFinetuner
    tagger, emb, params, ocfg = _models(SMALL, 13, seed=11)
    tagger.use_word_dropout = 0.0
    d = tagger.tag_dictionary
    rng = np.random.RandomState(4)

    def tagged(n, seed):
        sents = _sentences(n, 4, 24, seed=seed)
        for s in sents:
            # a learnable rule: the tag depends on the first letter of the word
            for tok in s.tokens:
                tok.add_tag("ner", "S-T0" if tok.text[0] in "abcde" else "O")
        return sents

    assert "S-T0" in d.item2idx or b"S-T0" in d.item2idx
    corpus = ListCorpus(tagged(24, 1), tagged(8, 2), tagged(8, 3))
    trainer = ModelFinetuner(tagger, corpus=corpus)
    before = float(tagger.forward_loss(BatchedData(corpus.dev)).detach())
    out = trainer.train(tmp_path, learning_rate=3e-4, lr_rate=100.0, mini_batch_size=4, max_epochs=3,
                        gradient_accumulation_steps=2, log_every=2)
    assert len(out["history"]) == 3 and all("dev_f1" in h for h in out["history"])
    tagger.eval()
    emb.fine_tune = True
    after = float(tagger.forward_loss(BatchedData(corpus.dev)).detach())
    print("dev loss before/after:", before, after, out["history"])
    assert after < before * 0.8
    assert os.path.exists(tmp_path / "final-model.pt")
    # reload: identical decode on the test split
    with torch.no_grad():
        f0 = tagger.forward(BatchedData(corpus.test))
        t0, _ = tagger._decode_batch(f0)
    loaded = type(tagger).load(tmp_path / "final-model.pt")
    with torch.no_grad():
        f1 = loaded.forward(BatchedData(corpus.test))
        t1, _ = loaded._decode_batch(f1)
    assert torch.equal(t0, t1)
    score = ModelFinetuner(loaded, corpus=corpus).final_test(tmp_path, eval_mini_batch_size=4, overall_test=True,
                                                              quiet_mode=True, nocrf=False, predict_posterior=False,
                                                              keep_embedding=-1, sort_data=False, eval_train=False)
    assert 0.0 <= score <= 1.0
