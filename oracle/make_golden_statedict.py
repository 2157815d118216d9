#!/usr/bin/env python
"""TEST INFRASTRUCTURE: the checkpoint dictionary the REFERENCE's FastSequenceTagger writes
(/root/reference/flair/models/sequence_tagger_model.py:435-477 `_get_state_dict`) for the KB-NER head configuration
(use_crf, no RNN, remove_x, sentence_loss) -> tests/golden/state_dict_golden.json: every key with its value (values that
are not plain JSON -- state_dict, embeddings, tag_dictionary -- are recorded as the list of state_dict entries / a type
name).  Run in the build container only:  python oracle/make_golden_statedict.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "state_dict_golden.json")


def main():
    flair = ref_shim.load_flair()
    import make_golden as G
    d = G.make_dictionary(flair, 13, with_x=True)
    tagger = G.build_tagger(flair, d, remove_x=True, seed=0)
    st = tagger._get_state_dict()
    out = {}
    for k, v in st.items():
        if k == "state_dict":
            out[k] = {n: list(t.shape) for n, t in v.items()}
        elif isinstance(v, (bool, int, float, str)) or v is None:
            out[k] = v
        else:
            out[k] = {"__type__": type(v).__name__}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, len(out), "keys")


if __name__ == "__main__":
    main()
