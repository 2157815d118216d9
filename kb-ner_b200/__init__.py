"""kbner_b200 -- B200-native KB-NER token-classification hot path (XLM-R encoder + CRF).

Hand-written sm_100a CUDA kernels behind a C ABI (include/kbner_b200.h), bound with ctypes,
exposed through mirrors of the reference's flair classes.  See DESIGN.md.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401

from . import data, datasets, encoder, embeddings, sequence_tagger  # noqa: F401,E402
from .data import BatchedData, Dictionary, Label, Sentence, Token  # noqa: F401,E402
from .embeddings import StackedEmbeddings, SyntheticTokenizer, TransformerWordEmbeddings  # noqa: F401,E402
from .sequence_tagger import FastSequenceTagger, SequenceTagger  # noqa: F401,E402

__all__ = ["_lib", "ops", "data", "datasets", "encoder", "embeddings", "sequence_tagger", "TransformerWordEmbeddings",
           "StackedEmbeddings", "SyntheticTokenizer", "SequenceTagger", "FastSequenceTagger", "Sentence", "Token",
           "Label", "Dictionary", "BatchedData"]
