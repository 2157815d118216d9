"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/crf_oracle.c.

CPU restatement of the reference's CRF path (see crf_oracle.c for the
reference file:line of every function).  Pinned against
tests/golden/crf_golden.npz in tests/test_oracle_golden.py.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkbner_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "crf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


_f, _i, _u8 = ctypes.c_float, ctypes.c_int32, ctypes.c_uint8


def compact(keep):
    keep = np.ascontiguousarray(keep, np.uint8)
    B, T = keep.shape
    pos = np.empty((B, T), np.int32)
    klen = np.empty((B,), np.int32)
    lib().kbner_oracle_compact(_p(keep, _u8), B, T, _p(pos, _i), _p(klen, _i))
    return pos, klen


def viterbi(emis, trans, klen, slen=None, pos=None, start=None, stop=None, x_idx=0):
    emis = np.ascontiguousarray(emis, np.float32)
    trans = np.ascontiguousarray(trans, np.float32)
    B, T, L = emis.shape
    klen = np.ascontiguousarray(klen, np.int32)
    slen = klen if slen is None else np.ascontiguousarray(slen, np.int32)
    pos = None if pos is None else np.ascontiguousarray(pos, np.int32)
    start = L - 2 if start is None else start
    stop = L - 1 if stop is None else stop
    tags = np.empty((B, T), np.int32)
    conf = np.empty((B, T), np.float32)
    lib().kbner_oracle_viterbi(_p(emis, _f), _p(pos, _i), _p(klen, _i), _p(slen, _i), _p(trans, _f),
                               B, T, L, int(start), int(stop), int(x_idx), _p(tags, _i), _p(conf, _f))
    return tags, conf


def crf_nll(emis, tags, trans, klen, pos=None, start=None, stop=None, want_alpha=False):
    emis = np.ascontiguousarray(emis, np.float32)
    trans = np.ascontiguousarray(trans, np.float32)
    tags = np.ascontiguousarray(tags, np.int32)
    B, T, L = emis.shape
    klen = np.ascontiguousarray(klen, np.int32)
    pos = None if pos is None else np.ascontiguousarray(pos, np.int32)
    start = L - 2 if start is None else start
    stop = L - 1 if stop is None else stop
    logz = np.empty((B,), np.float32)
    gold = np.empty((B,), np.float32)
    alpha = np.zeros((B, T, L), np.float32) if want_alpha else None
    lib().kbner_oracle_crf_nll(_p(emis, _f), _p(tags, _i), _p(pos, _i), _p(klen, _i), _p(trans, _f),
                               B, T, L, int(start), int(stop), _p(logz, _f), _p(gold, _f), _p(alpha, _f))
    return (logz, gold, alpha) if want_alpha else (logz, gold)


def crf_nll_bwd(emis, tags, trans, klen, w, pos=None, start=None, stop=None):
    emis = np.ascontiguousarray(emis, np.float32)
    trans = np.ascontiguousarray(trans, np.float32)
    tags = np.ascontiguousarray(tags, np.int32)
    w = np.ascontiguousarray(w, np.float32)
    B, T, L = emis.shape
    klen = np.ascontiguousarray(klen, np.int32)
    pos = None if pos is None else np.ascontiguousarray(pos, np.int32)
    start = L - 2 if start is None else start
    stop = L - 1 if stop is None else stop
    d_emis = np.empty((B, T, L), np.float32)
    d_trans = np.empty((L, L), np.float32)
    lib().kbner_oracle_crf_nll_bwd(_p(emis, _f), _p(tags, _i), _p(pos, _i), _p(klen, _i), _p(trans, _f),
                                   _p(w, _f), B, T, L, int(start), int(stop), _p(d_emis, _f), _p(d_trans, _f))
    return d_emis, d_trans


def crf_loss(emis, tags, trans, keep, start=None, stop=None):
    """FastSequenceTagger._calculate_loss (use_crf, remove_x): mean_b(logZ - gold), :2490-2506."""
    pos, klen = compact(keep)
    logz, gold = crf_nll(emis, tags, trans, klen, pos=pos, start=start, stop=stop)
    return np.float32(np.mean((logz - gold).astype(np.float32)))
