"""`python -m kbner_b200.train` end to end on the fixture YAML (the reference's own key set) with a tiny random encoder:
fine-tune -> best/final checkpoints + saved fine-tuned encoder -> --test -> --test_speed -> --parse of a CoNLL directory
(prediction file "text gold pred score") -> --save_embedding.  This is BASELINE configs[0] (`train.py --test` plumbing)
on the B200 path instead of the CPU."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny_encoder_dir(tmp_path):
    from kbner_b200.embeddings import SyntheticTokenizer
    from kbner_b200.encoder import EncoderConfig, XLMRobertaEncoderB200
    torch.manual_seed(5)
    cfg = EncoderConfig(vocab_size=1000, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                        max_position_embeddings=514, name="tiny-xlmr")
    enc = XLMRobertaEncoderB200(cfg)
    d = tmp_path / "tiny-xlmr"
    enc.save_pretrained(str(d))
    SyntheticTokenizer(cfg.vocab_size).save_pretrained(str(d))
    return d


def test_train_py_pipeline(tmp_path, capsys):
    from test_config_cpu import write_config
    from kbner_b200.train import main
    model_dir = _tiny_encoder_dir(tmp_path)
    cfg_path, out = write_config(tmp_path, model=model_dir, max_epochs=2, learning_rate=2e-4, lr_rate=100,
                                 mini_batch_size=2, gradient_accumulation_steps=2, train_with_dev=False)
    hist = main(["--config", cfg_path])
    assert len(hist["history"]) == 2 and "dev_macro_f1" in hist["history"][-1]
    base = out / "kbner-fixture"
    assert (base / "final-model.pt").exists() and (base / "best-model.pt").exists()
    # save_finetuned_embedding: true -> encoder + tokenizer next to the checkpoints (finetune_trainer.py:1290-1298)
    assert (base / "tiny-xlmr" / "config.json").exists() and (base / "tiny-xlmr" / "synthetic_tokenizer.json").exists()
    # --test: final_test on the test split (train.py passes eval_train, Appendix B.11)
    score = main(["--config", cfg_path, "--test", "--batch_size", "4", "--eval_train"])
    assert 0.0 <= score <= 1.0
    lines = open(base / "test.tsv", encoding="utf-8").read().split("\n")
    assert len(lines[0].split(" ")) == 4                                # text gold pred score (:2626-2643)
    # --test_speed: forward + Viterbi only
    r = main(["--config", cfg_path, "--test_speed"])
    assert r["sentences"] == 25 and r["sentences_per_sec"] > 0
    # --parse of a directory of CoNLL files, 4 columns, input order kept
    os.chdir(tmp_path)
    res = main(["--config", cfg_path, "--parse", "--target_dir", str(tmp_path / "data"), "--keep_order", "--num_columns", "4",
                "--comment_symbol", "# id", "--parse_name", "fixture", "--output_dir", str(tmp_path / "pred")])
    pred_file = tmp_path / "pred" / "train.kbner-fixture.fixture..conllu"
    assert pred_file.exists() and 0.0 <= res.main_score <= 1.0 and "MICRO_AVG" in res.detailed_results
    first = open(pred_file, encoding="utf-8").readline().split(" ")
    assert first[0] == "w0_0" and first[1] == "O"                       # keep_order: first sentence of the file first
    # --save_embedding on the loaded checkpoint
    main(["--config", cfg_path, "--save_embedding"])
    assert (base / "tiny-xlmr" / "model.safetensors").exists()       # HF layout: transformers.XLMRobertaModel.from_pretrained reads it
