"""TEST INFRASTRUCTURE ONLY -- golden-vector generator (run in the build container).

Executes the *reference's own* CRF / Viterbi / loss code
(``/root/reference/flair/models/sequence_tagger_model.py``:
``_viterbi_decode`` :1248-1327, ``_forward_alg`` :1329-1394,
``_score_sentence`` :2544-2591, ``_calculate_loss`` :2426-2539,
``_obtain_labels`` :1157-1246) on seeded synthetic inputs and writes the
inputs + outputs to ``tests/golden/crf_golden.npz``.

The committed ``.npz`` travels to the GPU box; ``/root/reference`` does not.

    python oracle/make_golden.py            # regenerates tests/golden/*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def make_dictionary(flair, L, with_x=True):
    """<unk>, O, real tags..., [S-X], <START>, <STOP>  (data.py:1083-1104 order)."""
    d = flair.data.Dictionary(add_unk=True)
    d.add_item("O")
    n_real = L - 4 - (1 if with_x else 0)
    for i in range(n_real):
        d.add_item("%s-T%d" % ("BIES"[i % 4], i // 4))
    if with_x:
        d.add_item("S-X")
    d.add_item("<START>")
    d.add_item("<STOP>")
    assert len(d) == L, (len(d), L)
    return d


class _FakeEmbeddings(torch.nn.Module):
    """Only what SequenceTagger.__init__ reads (sequence_tagger_model.py:165-388)."""

    def __init__(self, dim):
        super().__init__()
        self.embedding_length = dim
        self.embeddings = [self]
        self.name = "fake"


class _Tok:
    pass


class _Sent:
    def __init__(self, n, tags):
        self.tokens = [_Tok() for _ in range(n)]
        self.ner_tags = tags

    def __len__(self):
        return len(self.tokens)


def build_tagger(flair, tag_dictionary, remove_x, seed):
    from flair.models import FastSequenceTagger
    torch.manual_seed(seed)
    tagger = FastSequenceTagger(
        hidden_size=256, embeddings=_FakeEmbeddings(16), tag_dictionary=tag_dictionary,
        tag_type="ner", use_crf=True, use_rnn=False, remove_x=remove_x,
        word_dropout=0.1, sentence_loss=True, testing=True)
    return tagger


def run_case(flair, name, B, T, L, lens, remove_x, seed, scale=3.0, real_dict=None):
    d = real_dict if real_dict is not None else make_dictionary(flair, L, with_x=True)
    L = len(d)
    tagger = build_tagger(flair, d, remove_x, seed)
    start, stop = d.get_idx_for_item("<START>"), d.get_idx_for_item("<STOP>")
    x_idx = d.get_idx_for_item("S-X")
    g = torch.Generator().manual_seed(seed + 1)
    emis = (torch.randn(B, T, L, generator=g) * scale).float()
    # gold tags: uniform over non-special indices (never <unk>, S-X, START, STOP)
    legal = [i for i in range(L) if i not in (0, x_idx, start, stop)]
    tags = torch.tensor(legal)[torch.randint(0, len(legal), (B, T), generator=g)]
    lens_t = torch.tensor(lens, dtype=torch.long)
    mask = (torch.arange(T)[None, :] < lens_t[:, None]).float()
    # padded gold positions are 0 = <unk>  (custom_data_loader.py:362-365)
    tags = tags * mask.long()
    keep = mask.clone()
    if remove_x:
        # sentence || <EOS> || context : the first n_sent words are the sentence,
        # the rest carry S-X  (kb/context_process.py:219-223,424-426)
        for b in range(B):
            n_sent = int(torch.randint(1, max(2, lens[b] // 2 + 1), (1,), generator=g)) if lens[b] > 0 else 0
            n_sent = min(n_sent, lens[b])
            tags[b, n_sent:lens[b]] = x_idx
            keep[b, n_sent:] = 0
    sentences = [_Sent(lens[b], tags[b].clone()) for b in range(B)]

    emis_req = emis.clone().requires_grad_(True)
    tagger.transitions.grad = None
    tagger.mask = mask.clone()                        # what forward() leaves behind (:1028)
    # --- loss through the reference's own _calculate_loss --------------------
    loss = tagger._calculate_loss(emis_req, sentences, mask.clone())
    loss.backward()
    d_emis = emis_req.grad.detach().clone()
    d_trans = tagger.transitions.grad.detach().clone()
    ref_keep = tagger.mask.detach().clone()           # overwritten when remove_x (:2448-2453)
    assert torch.equal(ref_keep, keep)

    # --- per-sentence logZ / gold on the compacted rows ----------------------
    with torch.no_grad():
        klen = keep.sum(-1).long()
        Tm = max(int(klen.max()), 1)
        cf = torch.zeros(B, Tm, L)
        ct = torch.zeros(B, Tm, dtype=torch.long)
        cm = torch.zeros(B, Tm)
        for b in range(B):
            sel = keep[b].bool()
            cf[b, :klen[b]] = emis[b][sel]
            ct[b, :klen[b]] = tags[b][sel]
            cm[b, :klen[b]] = 1
        logz = tagger._forward_alg(cf, klen)
        gold = tagger._score_sentence(cf, ct, klen, mask=cm)

        # --- decode through the reference's own _obtain_labels ---------------
        labels, _ = tagger._obtain_labels(emis, sentences)
        vit = torch.full((B, T), -1, dtype=torch.int32)
        conf = torch.zeros(B, T)
        for b in range(B):
            for t, lab in enumerate(labels[b]):
                vit[b, t] = d.get_idx_for_item(lab.value)
                conf[b, t] = lab.score
    out = {
        "L": np.int32(L), "start": np.int32(start), "stop": np.int32(stop), "x_idx": np.int32(x_idx),
        "remove_x": np.int32(remove_x),
        "emis": emis.numpy(), "trans": tagger.transitions.detach().numpy().copy(),
        "lens": np.asarray(lens, np.int32), "tags": tags.numpy().astype(np.int32),
        "keep": keep.numpy().astype(np.uint8),
        "logz": logz.numpy(), "gold": gold.numpy(), "loss": np.float32(loss.item()),
        "d_emis": d_emis.numpy(), "d_trans": d_trans.numpy(),
        "viterbi": vit.numpy(), "conf": conf.numpy(),
    }
    return {"%s/%s" % (name, k): v for k, v in out.items()}


def main():
    flair = ref_shim.load_flair()
    real = flair.data.Dictionary.load_from_file(
        os.path.join(ref_shim.REFERENCE_ROOT, "resources/taggers/EN-English_x.pkl"))
    cases = {}
    rng = np.random.RandomState(7)
    specs = [
        # name,          B,  T,   L,  lens,                           remove_x
        ("t1_l13",       1,  1,   13, [1],                            0),
        ("b3_t7_l9",     3,  7,   9,  [7, 3, 1],                      0),
        ("b4_t33_l13",   4,  33,  13, [33, 20, 5, 1],                 0),
        ("b4_t40_l13_x", 4,  40,  13, [40, 31, 12, 2],                1),
        ("b2_t128_l29",  2,  128, 29, [128, 77],                      0),
        ("b3_t96_l29_x", 3,  96,  29, [96, 50, 9],                    1),
        ("b2_t512_l13",  2,  512, 13, [512, 300],                     0),
        ("b2_t512_l29x", 2,  512, 29, [512, 411],                     1),
        ("b6_t64_l32",   6,  64,  32, list(rng.randint(1, 65, 6)),    0),
    ]
    names = []
    for i, (name, B, T, L, lens, rx) in enumerate(specs):
        lens = [int(x) for x in lens]
        rd = real if (L == 29) else None
        cases.update(run_case(flair, name, B, T, L, lens, bool(rx), seed=100 + i, real_dict=rd))
        names.append(name)
        print("case", name, "loss", float(cases[name + "/loss"]))
    cases["names"] = np.array(names)
    out = os.path.join(ROOT, "tests", "golden", "crf_golden.npz")
    np.savez_compressed(out, **cases)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
