"""CPU: the restated oracle (oracle/crf_oracle.c) against the golden vectors produced by the
reference's own code (oracle/make_golden.py -> tests/golden/crf_golden.npz)."""
import numpy as np
import pytest

import crf_oracle as O


def _case(g, n):
    return {k: g["%s/%s" % (n, k)] for k in ("emis", "trans", "keep", "tags", "lens", "start", "stop", "x_idx",
                                             "viterbi", "conf", "logz", "gold", "loss", "d_emis", "d_trans")}


def test_golden_has_cases(golden):
    assert len(golden["names"]) >= 8


def test_viterbi_bit_exact(golden):
    for n in golden["names"]:
        c = _case(golden, n)
        pos, klen = O.compact(c["keep"])
        tags, conf = O.viterbi(c["emis"], c["trans"], klen, slen=c["lens"], pos=pos, start=int(c["start"]),
                               stop=int(c["stop"]), x_idx=int(c["x_idx"]))
        assert np.array_equal(tags, c["viterbi"]), n
        np.testing.assert_allclose(conf, c["conf"], rtol=2e-6, atol=1e-7, err_msg=str(n))


def test_logz_gold_loss(golden):
    for n in golden["names"]:
        c = _case(golden, n)
        pos, klen = O.compact(c["keep"])
        logz, gold = O.crf_nll(c["emis"], c["tags"], c["trans"], klen, pos=pos, start=int(c["start"]),
                               stop=int(c["stop"]))
        np.testing.assert_allclose(logz, c["logz"], rtol=1e-6, err_msg=str(n))
        np.testing.assert_allclose(gold, c["gold"], rtol=1e-5, atol=1e-4, err_msg=str(n))
        loss = O.crf_loss(c["emis"], c["tags"], c["trans"], c["keep"], start=int(c["start"]), stop=int(c["stop"]))
        np.testing.assert_allclose(loss, c["loss"], rtol=1e-5, err_msg=str(n))


def test_backward_matches_reference_autograd(golden):
    for n in golden["names"]:
        c = _case(golden, n)
        pos, klen = O.compact(c["keep"])
        B = c["emis"].shape[0]
        w = np.full(B, 1.0 / B, np.float32)
        de, dt = O.crf_nll_bwd(c["emis"], c["tags"], c["trans"], klen, w, pos=pos, start=int(c["start"]),
                               stop=int(c["stop"]))
        np.testing.assert_allclose(de, c["d_emis"], atol=1e-4, err_msg=str(n))
        scale = max(1.0, float(np.abs(c["d_trans"]).max()))
        assert np.abs(dt - c["d_trans"]).max() / scale < 1e-4, n


def test_compact_edge_cases():
    keep = np.array([[0, 0, 0, 0], [1, 1, 1, 1], [0, 1, 0, 1]], np.uint8)
    pos, klen = O.compact(keep)
    assert klen.tolist() == [0, 4, 2]
    assert pos[1].tolist() == [0, 1, 2, 3] and pos[2].tolist()[:2] == [1, 3] and pos[0].tolist() == [-1] * 4


def test_viterbi_empty_and_single():
    rng = np.random.RandomState(0)
    L = 13
    trans = rng.randn(L, L).astype(np.float32)
    trans[L - 2, :] = -1e12
    trans[:, L - 1] = -1e12
    emis = rng.randn(2, 3, L).astype(np.float32)
    tags, conf = O.viterbi(emis, trans, np.array([0, 1], np.int32), slen=np.array([2, 3], np.int32), x_idx=10)
    assert tags[0].tolist() == [10, 10, -1]
    assert tags[1, 1:].tolist() == [10, 10] and 0 <= tags[1, 0] < L - 2
