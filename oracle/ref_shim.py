"""TEST INFRASTRUCTURE ONLY -- loader for the *reference's own* Python code.

Imports ``flair`` from ``/root/reference`` (read-only, only present in the
build container, never on the GPU box) after installing stub modules for the
reference's missing third-party dependencies.  It is
used by ``oracle/make_golden.py`` to generate the committed golden vectors in
``tests/golden/`` and by the CPU tests that pin the restated oracle
(``oracle/crf_oracle.py`` / ``oracle/crf_oracle.c``) against the reference.

Nothing in the product path may import this file.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KBNER_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = (
    "segtok", "gensim", "bpemb", "pytorch_transformers", "h5py", "matplotlib",
    "IPython", "hyperopt", "allennlp", "pyhocon", "boto3", "botocore", "spacy",
    "stog", "mock", "conllu", "overrides", "nltk", "wikipedia2vec", "tensorboardX",
    "lmdb", "elasticsearch", "sklearn_crfsuite", "pytorch_pretrained_bert", "apex",
)


class _StubModule(types.ModuleType):
    """A module whose every attribute is a fresh subclassable dummy class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__call__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "flair"))


def install():
    """Make ``import flair`` resolve to the reference tree. Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch
    import transformers

    # stubs only for modules that are really missing
    really_missing = []
    for root in _STUB_ROOTS:
        try:
            if importlib.util.find_spec(root) is None:
                really_missing.append(root)
        except (ImportError, ValueError):
            really_missing.append(root)
    global _STUB_ROOTS_ACTIVE
    finder = _StubFinder()
    finder_roots = tuple(really_missing)

    def find_spec(fullname, path=None, target=None, _roots=finder_roots):
        if fullname.split(".")[0] in _roots:
            return importlib.machinery.ModuleSpec(fullname, finder, is_package=True)
        return None
    finder.find_spec = find_spec
    sys.meta_path.append(finder)

    # transformers 5.x dropped AdamW (reference: flair/trainers/finetune_trainer.py:8-11)
    # (transformers 5.x re-creates its lazy module object the first time some
    #  tokenizer classes are imported, so trigger those imports first and patch
    #  the module object that finally sits in sys.modules)
    from transformers import (XLNetTokenizer, T5Tokenizer, GPT2Tokenizer, AutoTokenizer,  # noqa: F401
                              AutoConfig, AutoModel, XLNetModel, BertTokenizer, BertModel,
                              XLMRobertaModel, XLMRobertaTokenizer,
                              get_linear_schedule_with_warmup)
    if not hasattr(sys.modules["transformers"], "AdamW"):
        sys.modules["transformers"].AdamW = torch.optim.AdamW
    # hard-coded .cuda() on the hot path (sequence_tagger_model.py:1028,2555-2563)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_flair():
    install()
    import flair  # noqa: F401  (the reference's package)
    import flair.models  # noqa: F401
    return flair
