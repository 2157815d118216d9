"""Command line of the reference's train.py (/root/reference/train.py:35-64 flags, :88-174 flow), on the B200 path.

    python -m kbner_b200.train --config config/<name>.yaml                 # fine-tune   (train.py:416)
    python -m kbner_b200.train --config ... --test [--batch_size N]        # final_test  (train.py:160-174)
    python -m kbner_b200.train --config ... --test_speed                   # sentences/s (train.py:148-158)
    python -m kbner_b200.train --config ... --parse --target_dir DIR --keep_order [--num_columns 4] [--parse_name X]
    python -m kbner_b200.train --config ... --save_embedding               # dump the fine-tuned encoder (train.py:259-266)

Multi-GPU: launch under torchrun (one process per GPU); the trainer shards sentence batches and all-reduces gradients.
Flags of the reference that select code outside the hot path are accepted by the parser (so existing scripts do not break
on argument parsing) and refused with a message when set.
"""
import argparse
import logging
import os
import sys
from pathlib import Path

log = logging.getLogger("kbner_b200")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser("train.py")
    p.add_argument("--config", help="configuration YAML file.")
    p.add_argument("--test", action="store_true", help="Whether testing the pretrained model.")
    p.add_argument("--zeroshot", action="store_true")
    p.add_argument("--all", action="store_true")
    p.add_argument("--other", action="store_true")
    p.add_argument("--quiet", action="store_true", help="print results only")
    p.add_argument("--nocrf", action="store_true")
    p.add_argument("--parse", action="store_true", help="parse files")
    p.add_argument("--parse_train_and_dev", action="store_true")
    p.add_argument("--keep_order", action="store_true", help="keep the parse order for the prediction")
    p.add_argument("--predict", action="store_true")
    p.add_argument("--debug", action="store_true")
    p.add_argument("--target_dir", default="", help="file dir to parse")
    p.add_argument("--spliter", default="\t")
    p.add_argument("--recur_parse", action="store_true")
    p.add_argument("--parse_test", action="store_true", help="parse the test set")
    p.add_argument("--save_embedding", action="store_true", help="save the pretrained embeddings")
    p.add_argument("--mst", action="store_true")
    p.add_argument("--test_speed", action="store_true", help="test the running speed")
    p.add_argument("--predict_posterior", action="store_true")
    p.add_argument("--batch_size", default=-1, help="manually setting the mini batch size for testing")
    p.add_argument("--keep_embedding", default=-1)
    p.add_argument("--remove_x", action="store_true", help="forcing the remove_x to be activated")
    p.add_argument("--v2doc", action="store_true")
    p.add_argument("--eval_train", action="store_true")
    p.add_argument("--num_columns", type=int, default=2, help="for prediction")
    p.add_argument("--comment_symbol", type=str, default=None)
    p.add_argument("--parse_name", default="", help="for naming the output file")
    p.add_argument("--output_dir", default="outputs", help="for naming the output dir")
    return p


_REFUSED = ("zeroshot", "all", "other", "nocrf", "predict", "mst", "predict_posterior", "v2doc", "recur_parse")


def count_parameters(model) -> int:
    return sum(int(p.numel()) for _, p in model.named_parameters())


def _loader(sentences, batch_size, student, trainer_cfg, keep_order):
    from .datasets import ColumnDataLoader
    loader = ColumnDataLoader(list(sentences), batch_size, use_bert=student.use_bert, model=student,
                              sort_data=not keep_order, sentence_level_batch=trainer_cfg.get("sentence_level_batch", True))
    loader.assign_tags(student.tag_type, student.tag_dictionary)
    return loader


def _report(result):
    if result is not None and hasattr(result, "main_score"):
        print("Current accuracy: " + str(result.main_score * 100))
        print(result.detailed_results)


def main(argv=None):
    args = build_parser().parse_args(argv)
    bad = [f for f in _REFUSED if getattr(args, f)] + (["keep_embedding"] if int(args.keep_embedding) >= 0 else [])
    if bad:
        raise NotImplementedError("flags outside the KB-NER hot path: %s" % ", ".join("--" + b for b in bad))
    if args.quiet:
        log.disabled = True
    import torch
    from . import trainer as trainers
    from .config_parser import ConfigParser, Params
    from .datasets import ColumnCorpus

    if "RANK" in os.environ and not torch.distributed.is_initialized():      # torchrun: one process per GPU
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        torch.distributed.init_process_group("nccl")

    config = ConfigParser(Params.from_file(args.config), all=args.all, zero_shot=args.zeroshot, other_shot=args.other,
                          predict=args.predict, save_embedding=args.save_embedding)
    if args.save_embedding:
        # no corpus / tag dictionary is read in this mode (config_parser.py:69-75): the tagger comes from its checkpoint
        from .sequence_tagger import FastSequenceTagger
        base_path = Path(config.config["target_dir"]) / config.config["model_name"]
        ckpt = base_path / "best-model.pt" if (base_path / "best-model.pt").exists() else base_path / "final-model.pt"
        if not ckpt.exists():
            raise FileNotFoundError(str(base_path) + " not exist!")
        student = FastSequenceTagger.load(ckpt)
        for e in student.embeddings.embeddings:
            e.fine_tune = True                      # final_test / load leave it False; the dump is of the tuned encoder
        trainers.ModelFinetuner(student, corpus=None).save_finetuned_embeddings(base_path)
        return None
    student = config.create_student(nocrf=args.nocrf)
    log.info("Model Size: %d", count_parameters(student))
    corpus = config.corpus

    cfg = config.config
    trainer_name = cfg["trainer"] if "trainer" in cfg else ("ModelFinetuner" if "ModelFinetuner" in cfg else "ModelDistiller")
    trainer_cls = getattr(trainers, trainer_name, None)
    if trainer_cls is None:
        raise NotImplementedError("trainer %r is outside the hot path (KB-NER's configs use ModelFinetuner)" % trainer_name)
    trainer_cfg = dict(cfg.get(trainer_name) or {})
    trainer_cfg.setdefault("distill_mode", False)
    trainer = trainer_cls(student, None, corpus, config=cfg, **trainer_cfg, is_test=args.test)

    train_config = dict(cfg["train"])
    train_config["base_path"] = config.get_target_path
    eval_mini_batch_size = int(args.batch_size) if int(args.batch_size) > 0 else int(cfg["train"]["mini_batch_size"])

    if args.test_speed:
        student.eval()
        print(count_parameters(student))
        loader = _loader(trainer.corpus.test, 32, student, {"sentence_level_batch": True}, keep_order=True)
        result, _ = student.evaluate(loader, embeddings_storage_mode="none", speed_test=True)
        print(result["sentences_per_sec"])
        return result
    if args.test:
        student.eval()
        trainer.embeddings_storage_mode = "cpu"
        return trainer.final_test(config.get_target_path, eval_mini_batch_size=eval_mini_batch_size, overall_test=True,
                                  quiet_mode=args.quiet, nocrf=args.nocrf, predict_posterior=args.predict_posterior,
                                  sort_data=not args.keep_order, eval_train=args.eval_train)
    if args.parse:
        print("Batch Size:", eval_mini_batch_size)
        base_path = Path(cfg["target_dir"]) / cfg["model_name"]
        if (base_path / "best-model.pt").exists():
            print("Loading pretraining best model")
            student = student.load(base_path / "best-model.pt")
        elif (base_path / "final-model.pt").exists():
            print("Loading pretraining final model")
            student = student.load(base_path / "final-model.pt")
        else:
            raise FileNotFoundError(str(base_path) + " not exist!")
        if args.remove_x:
            student.remove_x = True
            student.tag_dictionary.add_item("S-X")
        if not hasattr(student, "use_bert"):
            student.use_bert = False
        results = {}
        if args.parse_train_and_dev:
            os.makedirs("system_pred", exist_ok=True)
            print("Current Model: ", cfg["model_name"])
            for split, lists in (("dev", corpus.dev_list), ("train", corpus.train_list), ("test", corpus.test_list)):
                print("Current Set: ", split)
                for name, sub in zip(corpus.targets, lists):
                    if len(sub) == 0:
                        continue
                    print("Current Lang: ", name)
                    r, _ = student.evaluate(_loader(sub, eval_mini_batch_size, student, trainer_cfg, args.keep_order),
                                            embeddings_storage_mode="none",
                                            out_path=Path("system_pred/%s.%s.conllu" % (split, cfg["model_name"])))
                    _report(r)
                    results[(split, name)] = r
            return results
        if args.target_dir != "":
            fmt = {0: "text", 1: "upos", 2: "xpos", 3: "ner"} if args.num_columns == 4 else {0: "text", 1: "ner"}
            parsed = ColumnCorpus(args.target_dir, column_format=fmt, tag_to_bioes="ner", comment_symbol=args.comment_symbol)
            if args.parse_test:
                data, out = parsed.test, Path("system_pred/test.%s.%s.conllu" % (cfg["model_name"], args.parse_name))
            else:
                data = parsed.train
                out = Path("%s/train.%s.%s..conllu" % (args.output_dir, cfg["model_name"], args.parse_name))
        else:
            data, out = corpus.train, Path("outputs/train.%s.%s.conllu" % (cfg["model_name"], corpus.targets[0]))
        out.parent.mkdir(parents=True, exist_ok=True)
        r, _ = student.evaluate(_loader(data, eval_mini_batch_size, student, trainer_cfg, args.keep_order), out_path=out,
                                embeddings_storage_mode="none", prediction_mode=True)
        _report(r)
        return r
    return trainer.train(**train_config)


if __name__ == "__main__":
    main(sys.argv[1:])
