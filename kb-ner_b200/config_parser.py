"""YAML -> corpus / tag dictionary / embeddings / tagger, for the KB-NER configurations.

Mirror of /root/reference/flair/config_parser.py (ConfigParser :27-135 constructor, create_embeddings :145-188,
create_model :189-243, create_student :245-249, load_pretrained :298-302, get_target :304-309, get_corpus :311-358,
get_target_path :601-603) and of flair/utils/params.py (Params.from_file :97-109), reduced to what the shipped
`config/*.yaml` of KB-NER use: `ColumnCorpus-*` corpora, `TransformerWordEmbeddings-*` embeddings, a
`FastSequenceTagger` / `SequenceTagger` model.  Classes are resolved BY NAME from the YAML keys exactly like the
reference does (`getattr(module, key.split('-')[0])`), so a YAML that names something outside the hot path fails with a
message that says so instead of silently building something else.
"""
import copy
import logging
from pathlib import Path
from typing import Dict, List

import yaml

from . import datasets as datasets
from . import embeddings as Embeddings
from .data import Dictionary
from .datasets import ListCorpus

log = logging.getLogger("kbner_b200")

__all__ = ["ConfigParser", "Params"]


def _merge(dst: dict, src: dict) -> dict:
    """Recursive dict merge (flair/algorithms/dict_merge.py as used by Params.from_file for comma-separated files)."""
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


class Params:
    """Thin mapping wrapper with the members train.py / ConfigParser touch (`[]`, `in`, `.get`, `.params`)."""

    def __init__(self, params: dict):
        self.params = params

    @classmethod
    def from_file(cls, params_file_list: str) -> "Params":
        merged: dict = {}
        for name in str(params_file_list).split(","):
            with open(name, encoding="utf-8") as f:
                if name.endswith(".yaml") or name.endswith(".yml"):
                    _merge(merged, yaml.safe_load(f) or {})
                elif name.endswith(".json"):
                    import json
                    merged = json.load(f)
                else:
                    raise NotImplementedError("config files are .yaml or .json (got %r)" % name)
        return cls(merged)

    def __getitem__(self, k): return self.params[k]
    def __setitem__(self, k, v): self.params[k] = v
    def __contains__(self, k): return k in self.params
    def __iter__(self): return iter(self.params)
    def get(self, k, default=None): return self.params.get(k, default)
    def keys(self): return self.params.keys()
    def items(self): return self.params.items()
    def duplicate(self): return Params(copy.deepcopy(self.params))
    def __repr__(self): return "Params(%r)" % (self.params,)


def _resolve(module, key: str, what: str):
    name = key.split("-")[0]
    cls = getattr(module, name, None)
    if cls is None:
        raise NotImplementedError("%s %r (YAML key %r) is outside the hot path kbner_b200 implements; available: %s"
                                  % (what, name, key, sorted(n for n in dir(module) if n[:1].isupper())))
    return cls


class ConfigParser:
    def __init__(self, config, all: bool = False, zero_shot: bool = False, other_shot: bool = False,
                 predict: bool = False, save_embedding: bool = False):
        if all or zero_shot or other_shot or predict:
            raise NotImplementedError("--all / --zeroshot / --other / --predict select the reference's built-in multi-corpus "
                                      "benchmarks (PANX, UD, CoNLL-03 ...), none of which KB-NER's configs use")
        self.config = config
        self.mini_batch_size = self.config["train"]["mini_batch_size"]
        self.target: str = self.get_target
        self.tag_type = self.target
        if save_embedding:                                  # (:69-75) no data needed to dump the fine-tuned encoder
            self.corpus, self.tokens, self.tag_dictionary, self.num_corpus = None, None, {}, None
            return
        self.corpus: ListCorpus = self.get_corpus
        train_cfg = self.config["train"]
        self.tokens = self.corpus.get_train_full_tokenset(-1, min_freq=train_cfg.get("min_freq", -1))
        if train_cfg.get("use_unlabeled_data", False):
            raise NotImplementedError("use_unlabeled_data (semi-supervised KD) is outside the hot path")
        self.corpus_list: List[str] = self.config[self.target]["Corpus"].split(":")
        # keep the tag dictionary consistent between runs (:122-129): load it when the pickle exists, else build + save
        dict_path = self.config[self.target].get("tag_dictionary")
        if dict_path and Path(dict_path).exists():
            self.tag_dictionary = Dictionary.load_from_file(dict_path)
        else:
            self.tag_dictionary = self.corpus.make_tag_dictionary(tag_type=self.target)
            if dict_path:
                Path(dict_path).parent.mkdir(parents=True, exist_ok=True)
                self.tag_dictionary.save(dict_path)
        log.info(self.tag_dictionary.item2idx)
        self.num_corpus = len(self.corpus.targets)

    # ---- corpus ---------------------------------------------------------------------------------------------------
    @property
    def get_target(self) -> str:
        targets = self.config.get("targets").split(":")
        if len(targets) > 1:
            log.info("Warning! Not support multitask now!")
        return targets[0]

    @property
    def get_corpus(self) -> ListCorpus:
        lists: Dict[str, list] = {"train": [], "dev": [], "test": []}
        names = self.config[self.target]["Corpus"].split(":")
        for corpus in names:
            if "ColumnCorpus" not in corpus:
                raise NotImplementedError("corpus %r: only ColumnCorpus-* entries (CoNLL column files with the <EOS> + "
                                          "context convention) are on the KB-NER path" % corpus)
            cls = _resolve(datasets, corpus, "corpus class")
            current = cls(**self.config[self.target][corpus])
            lists["train"].append(current.train)
            lists["dev"].append(current.dev)
            lists["test"].append(current.test)
        return ListCorpus(**lists, targets=names)

    @property
    def get_target_path(self) -> Path:
        return Path(self.config["target_dir"]) / self.config["model_name"]

    # ---- model ----------------------------------------------------------------------------------------------------
    def create_embeddings(self, embeddings: dict):
        """-> (StackedEmbeddings, word_map, char_map, lemma_map, postag_map); the maps belong to embedding classes that are
        outside the hot path and are always None here."""
        built = []
        for key, kw in embeddings.items():
            cls = _resolve(Embeddings, key, "embedding class")
            built.append(cls(**kw) if isinstance(kw, dict) else cls())
        return Embeddings.StackedEmbeddings(embeddings=built), None, None, None, None

    def create_model(self, config=None, pretrained: bool = False, is_student: bool = False, crf: bool = True):
        from . import sequence_tagger as models
        if config is None:
            config = self.config
        if config.get("is_toy", False):
            pretrained = False
        embeddings, word_map, char_map, lemma_map, postag_map = self.create_embeddings(config["embeddings"])
        classname = list(config["model"].keys())[0]
        kwargs = copy.deepcopy(config["model"][classname])
        if not crf:
            kwargs["use_crf"] = False
        kwargs.update(embeddings=embeddings, tag_type=self.target, tag_dictionary=self.tag_dictionary)
        if not pretrained:
            kwargs["target_languages"] = self.num_corpus
        tagger = _resolve(models, classname, "model class")(**kwargs, config=config)
        tagger.word_map, tagger.char_map, tagger.lemma_map, tagger.postag_map = word_map, char_map, lemma_map, postag_map
        if pretrained:
            base_path = Path(config["target_dir"]) / config["model_name"]
            if (base_path / "best-model.pt").exists():
                log.info("Loading pretraining best model")
                tagger = tagger.load(base_path / "best-model.pt")
            elif (base_path / "final-model.pt").exists():
                log.info("Loading pretraining final model")
                tagger = tagger.load(base_path / "final-model.pt")
            else:
                raise FileNotFoundError(str(base_path) + " not exist!")
        tagger.use_bert = any("bert" in key.lower() for key in config["embeddings"])      # (:227-231)
        return tagger

    def create_student(self, nocrf: bool = False):
        return self.create_model(self.config, pretrained=self.load_pretrained(self.config), is_student=True, crf=not nocrf)

    def create_teachers(self, is_professor: bool = False):
        raise NotImplementedError("teacher models belong to knowledge distillation (distill_mode), outside the hot path")

    create_teachers_list = create_teachers

    def load_pretrained(self, config=None) -> bool:
        try:
            return bool(self.config["load_pretrained"])
        except KeyError:
            return False
