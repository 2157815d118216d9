"""Host-only entry points of the C ABI that can be exercised without a GPU: argument validation happens before any CUDA
call, and the workspace size of the fused GEMM + LayerNorm kernel is plain arithmetic."""
import ctypes

import pytest


@pytest.fixture(scope="module")
def lib():
    import kbner_b200
    from kbner_b200 import _lib
    return _lib.load()


def test_gemm_ln_workspace_bytes(lib):
    # 16 control bytes + one 16-byte {mean, tag, M2, tag} slot per (256-row panel, 256-column tile, column half, row)
    assert lib.kbner_gemm_ln_workspace_bytes(16384, 1024) == 16 + 64 * 4 * 2 * 256 * 16
    assert lib.kbner_gemm_ln_workspace_bytes(300, 256) == 16 + 2 * 1 * 2 * 256 * 16
    assert lib.kbner_gemm_ln_workspace_bytes(4096, 1000) == 0          # N must be a multiple of 256
    assert lib.kbner_gemm_ln_workspace_bytes(0, 1024) == 0


def test_wgrad_group_rejects_bad_arguments_before_touching_the_device(lib):
    vp, ci = ctypes.c_void_p * 1, ctypes.c_int * 1
    one = vp(16)
    assert lib.kbner_gemm_wgrad_group(0, one, one, one, ci(8), ci(8), ci(8), ci(8), 64, None) != 0
    assert b"1..4 problems" in lib.kbner_last_error()
    assert lib.kbner_gemm_wgrad_group(5, one, one, one, ci(8), ci(8), ci(8), ci(8), 64, None) != 0
    assert lib.kbner_gemm_wgrad_group(1, one, one, one, ci(8), ci(8), ci(8), ci(8), 0, None) != 0
    assert lib.kbner_gemm_wgrad_group(1, one, one, one, ci(12), ci(8), ci(12), ci(8), 64, None) != 0   # extents: multiples of 8
    assert b"multiples of 8" in lib.kbner_last_error()


def test_fused_layernorm_rejects_unsupported_hidden_sizes(lib):
    rc = lib.kbner_gemm_bias_resid_layernorm_ws(16, 16, None, None, 16, 16, 1e-5, 16, 128, 1000, 64, 64, 64, 16, 1 << 20, None)
    assert rc != 0 and b"256, 512, 768, 1024" in lib.kbner_last_error()


def test_bench_strict_sentences_are_one_window_each():
    """The e2e.strict leg of bench.py promises the headline's shape on never-seen sentences: 510 one-piece words = one
    512-sub-token window per sentence.  (Five-character words were two pieces of the stand-in tokenizer each: four
    overflow windows per sentence, 128 x 512 rows per batch -- the leg measured 4x the headline's work until round 2.)"""
    import torch
    import bench
    from kbner_b200.data import BatchedData
    from kbner_b200.embeddings import SyntheticTokenizer, TransformerWordEmbeddings
    from kbner_b200.encoder import EncoderConfig
    cfg = EncoderConfig.xlmr_base()
    cfg.num_hidden_layers = 1
    emb = TransformerWordEmbeddings(model=cfg.name, layers="-1", pooling_operation="first", fine_tune=False,
                                    tokenizer=SyntheticTokenizer(cfg.vocab_size), config=cfg, device=torch.device("cpu"))
    for seed in (1, 2):
        ids, key_len, row_of, first_idx, lengths, S = emb.build_batch(BatchedData(bench.fresh_sentences(4, seed)))
        assert tuple(ids.shape) == (4, bench.S_LEN) and S == bench.S_LEN
        assert key_len.tolist() == [bench.S_LEN] * 4 and row_of.tolist() == [0, 1, 2, 3]
        assert lengths == [bench.S_LEN - 2] * 4 and int(first_idx.min()) >= 1
    a = [t.text for t in bench.fresh_sentences(1, 3)[0].tokens]
    b = [t.text for t in bench.fresh_sentences(1, 4)[0].tokens]
    assert a != b                                            # different seeds: sentences no sentence-level cache has seen
