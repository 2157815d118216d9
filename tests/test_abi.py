"""CPU: the C-ABI library loads and exports every symbol include/kbner_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "kbner_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kbner_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ("kbner_crf_viterbi", "kbner_crf_nll_fwd", "kbner_crf_nll_bwd", "kbner_gemm_bf16_tn",
                 "kbner_attention_fwd", "kbner_layernorm_fwd", "kbner_embed_ln_fwd", "kbner_gather_tagproj_fwd"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import kbner_b200
    lib = ctypes.CDLL(kbner_b200._lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert set(_declared()) == set(kbner_b200._lib.SIGNATURES), "ctypes table out of sync with the header"
    assert kbner_b200._lib.load().kbner_abi_version() == 2


def test_no_silent_fallback_when_library_missing(monkeypatch):
    import kbner_b200
    monkeypatch.setattr(kbner_b200._lib, "_lib", None)
    monkeypatch.setattr(kbner_b200._lib, "LIB_PATH", "/nonexistent/libkbner_b200.so")
    with pytest.raises(kbner_b200._lib.KbnerError):
        kbner_b200._lib.load()


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kb-ner_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "crf_oracle" not in src and "ref_shim" not in src and "encoder_oracle" not in src, f
