"""Tensor-level wrappers over the C ABI (include/kbner_b200.h).

torch is used only for device memory and the current CUDA stream; every function here ends in
exactly one call into libkbner_b200.so.  No function has a PyTorch / CPU fallback.
"""
import os

import torch

from . import _lib

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESID_F32, EPI_NONE_F32 = 0, 1, 2, 3


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name, ndim=None):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.KbnerError("%s must be a CUDA tensor (kbner_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise _lib.KbnerError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _lib.KbnerError("%s must be contiguous" % name)
    if ndim is not None and t.dim() != ndim:
        raise _lib.KbnerError("%s must have %d dims, got %d" % (name, ndim, t.dim()))


# ---- CRF ---------------------------------------------------------------------------------------
def crf_compact(keep):
    """keep [B,T] uint8 -> (pos [B,T] int32, klen [B] int32)."""
    _chk(keep, torch.uint8, "keep", 2)
    B, T = keep.shape
    pos = torch.empty((B, T), dtype=torch.int32, device=keep.device)
    klen = torch.empty((B,), dtype=torch.int32, device=keep.device)
    _lib.check(_lib.load().kbner_crf_compact(_ptr(keep), B, T, _ptr(pos), _ptr(klen), _stream()), "crf_compact")
    return pos, klen


def crf_viterbi(emis, trans, klen, slen, start_idx, stop_idx, x_idx=0, pos=None):
    """Returns (tags [B,T] int32, conf [B,T] float32)."""
    _chk(emis, torch.float32, "emis", 3)
    _chk(trans, torch.float32, "trans", 2)
    _chk(klen, torch.int32, "klen", 1)
    _chk(slen, torch.int32, "slen", 1)
    _chk(pos, torch.int32, "pos", 2)
    B, T, L = emis.shape
    tags = torch.empty((B, T), dtype=torch.int32, device=emis.device)
    conf = torch.empty((B, T), dtype=torch.float32, device=emis.device)
    _lib.check(_lib.load().kbner_crf_viterbi(_ptr(emis), _ptr(pos), _ptr(klen), _ptr(slen), _ptr(trans), B, T, L,
                                             int(start_idx), int(stop_idx), int(x_idx), _ptr(tags), _ptr(conf),
                                             _stream()), "crf_viterbi")
    return tags, conf


def crf_nll_fwd(emis, tags, trans, klen, start_idx, stop_idx, pos=None, want_alpha=False):
    """Returns (logz [B], gold [B], alpha) ; alpha = (ahat [B,T,L] f32, scale [B,T] f64) or None."""
    _chk(emis, torch.float32, "emis", 3)
    _chk(tags, torch.int32, "tags", 2)
    _chk(trans, torch.float32, "trans", 2)
    _chk(klen, torch.int32, "klen", 1)
    _chk(pos, torch.int32, "pos", 2)
    B, T, L = emis.shape
    logz = torch.empty((B,), dtype=torch.float32, device=emis.device)
    gold = torch.empty((B,), dtype=torch.float32, device=emis.device)
    ahat = torch.empty((B, T, L), dtype=torch.float32, device=emis.device) if want_alpha else None
    scale = torch.empty((B, T), dtype=torch.float64, device=emis.device) if want_alpha else None
    _lib.check(_lib.load().kbner_crf_nll_fwd(_ptr(emis), _ptr(tags), _ptr(pos), _ptr(klen), _ptr(trans), B, T, L,
                                             int(start_idx), int(stop_idx), _ptr(logz), _ptr(gold), _ptr(ahat),
                                             _ptr(scale), _stream()), "crf_nll_fwd")
    return logz, gold, ((ahat, scale) if want_alpha else None)


def crf_nll_bwd(emis, tags, trans, klen, alpha, w, start_idx, stop_idx, pos=None):
    """Returns (d_emis [B,T,L], d_trans [L,L]) of sum_b w[b]*(logZ_b - gold_b); alpha from crf_nll_fwd."""
    ahat, scale = alpha
    _chk(ahat, torch.float32, "alpha", 3)
    _chk(scale, torch.float64, "alpha_scale", 2)
    _chk(w, torch.float32, "w", 1)
    B, T, L = emis.shape
    d_emis = torch.empty_like(emis)
    d_trans = torch.zeros((L, L), dtype=torch.float32, device=emis.device)
    _lib.check(_lib.load().kbner_crf_nll_bwd(_ptr(emis), _ptr(tags), _ptr(pos), _ptr(klen), _ptr(trans), _ptr(ahat),
                                             _ptr(scale), _ptr(w), B, T, L, int(start_idx), int(stop_idx),
                                             _ptr(d_emis), _ptr(d_trans), _stream()), "crf_nll_bwd")
    return d_emis, d_trans


# ---- encoder -----------------------------------------------------------------------------------
def embed_ln_fwd(ids, word_emb, pos_emb, type_emb, gamma, beta, eps, pad_id, out=None, out32=None, split=False):
    """out [R*S,H] bf16 (split: [R*S,3H] rows hi|lo|hi); out32 (optional) receives the same values in fp32."""
    _chk(ids, torch.int32, "ids", 2)
    _chk(out32, torch.float32, "out32", 2)
    for n, t in (("word_emb", word_emb), ("pos_emb", pos_emb), ("type_emb", type_emb), ("gamma", gamma),
                 ("beta", beta)):
        _chk(t, torch.float32, n)
    R, S = ids.shape
    V, H = word_emb.shape
    P = pos_emb.shape[0]
    width = 3 * H if split else H
    if out is None:
        out = torch.empty((R * S, width), dtype=torch.bfloat16, device=ids.device)
    else:
        _chk(out, torch.bfloat16, "out", 2)
        if tuple(out.shape) != (R * S, width):
            raise _lib.KbnerError("embed_ln_fwd: out must be [%d, %d]" % (R * S, width))
    _lib.check(_lib.load().kbner_embed_ln_fwd_ex(_ptr(ids), _ptr(word_emb), _ptr(pos_emb), _ptr(type_emb), _ptr(gamma),
                                                 _ptr(beta), float(eps), int(pad_id), R, S, H, V, P, _ptr(out),
                                                 _ptr(out32), int(bool(split)), _stream()), "embed_ln_fwd")
    return out


def layernorm_fwd_res32(x, gamma, beta, eps, out, out32=None, bias=None, resid=None, split=False):
    """Precision modes: out32 = LayerNorm(x + bias + resid) (fp32 residual stream), out = bf16 / split bf16 copy."""
    _chk(x, torch.float32, "x", 2)
    _chk(gamma, torch.float32, "gamma", 1)
    _chk(beta, torch.float32, "beta", 1)
    _chk(bias, torch.float32, "bias", 1)
    _chk(resid, torch.float32, "resid", 2)
    _chk(out32, torch.float32, "out32", 2)
    _chk(out, torch.bfloat16, "out", 2)
    M, H = x.shape
    if tuple(out.shape) != (M, 3 * H if split else H):
        raise _lib.KbnerError("layernorm_fwd_res32: out must be [%d, %d]" % (M, 3 * H if split else H))
    _lib.check(_lib.load().kbner_add_layernorm_fwd_res32(_ptr(x), _ptr(bias), _ptr(resid), _ptr(gamma), _ptr(beta), float(eps),
                                                         M, H, _ptr(out32), _ptr(out), int(bool(split)), _stream()),
               "layernorm_fwd_res32")
    return out


def bias_gelu_split(x, bias, out3):
    """out3 [M,3F] bf16 = hi|lo|hi of gelu_erf(x + bias), x fp32 [M,F]."""
    _chk(x, torch.float32, "x", 2)
    _chk(bias, torch.float32, "bias", 1)
    _chk(out3, torch.bfloat16, "out3", 2)
    M, F = x.shape
    if tuple(out3.shape) != (M, 3 * F):
        raise _lib.KbnerError("bias_gelu_split: out3 must be [%d, %d]" % (M, 3 * F))
    _lib.check(_lib.load().kbner_bias_gelu_split(_ptr(x), _ptr(bias), M, F, _ptr(out3), _stream()), "bias_gelu_split")
    return out3


def _drop_args(drop):
    """drop = None | (seed int32[2] cuda tensor, site int, p float) -> (seed ptr, site, p) of the C ABI."""
    if drop is None:
        return None, 0, 0.0
    seed, site, p = drop
    _chk(seed, torch.int32, "drop_seed", 1)
    if seed.numel() != 2:
        raise _lib.KbnerError("drop_seed must hold 2 words")
    return _ptr(seed), int(site), float(p)


def layernorm_fwd(x, gamma, beta, eps, out=None, save_stats=False, bias=None, resid=None, drop=None):
    """y = LayerNorm(dropout(x + bias) + resid) * gamma + beta (bias / resid / drop optional)."""
    _chk(x, torch.float32, "x", 2)
    _chk(gamma, torch.float32, "gamma", 1)
    _chk(beta, torch.float32, "beta", 1)
    _chk(bias, torch.float32, "bias", 1)
    _chk(resid, torch.bfloat16, "resid", 2)
    M, H = x.shape
    if out is None:
        out = torch.empty((M, H), dtype=torch.bfloat16, device=x.device)
    mean = rstd = None
    if save_stats:
        mean = torch.empty((M,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
    sp, site, p = _drop_args(drop)
    _lib.check(_lib.load().kbner_add_layernorm_fwd(_ptr(x), _ptr(bias), _ptr(resid), _ptr(gamma), _ptr(beta), float(eps), M, H,
                                                   _ptr(out), _ptr(mean), _ptr(rstd), sp, site, p, _stream()),
               "layernorm_fwd")
    return (out, mean, rstd) if save_stats else out


def dropout_apply(x, drop):
    """In-place dropout of a [M,H] bf16 / fp32 tensor with the counter-hash mask of site drop[1]."""
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.KbnerError("dropout_apply: bf16 or fp32 expected")
    _chk(x, x.dtype, "x", 2)
    sp, site, p = _drop_args(drop)
    _lib.check(_lib.load().kbner_dropout_apply(_ptr(x), 1 if x.dtype == torch.float32 else 0, x.shape[0], x.shape[1], sp, site, p,
                                               _stream()), "dropout_apply")
    return x


def gather_tagproj_fwd(hidden, row_of, first_idx, W, bias, S, drop_keep=None):
    f32 = hidden.dtype == torch.float32          # precision modes hand the last LayerNorm's fp32 output over
    _chk(hidden, torch.float32 if f32 else torch.bfloat16, "hidden", 2)
    _chk(row_of, torch.int32, "row_of", 1)
    _chk(first_idx, torch.int32, "first_idx", 2)
    _chk(W, torch.float32, "W", 2)
    _chk(bias, torch.float32, "bias", 1)
    _chk(drop_keep, torch.uint8, "drop_keep", 1)
    B, T = first_idx.shape
    L, H = W.shape
    logits = torch.empty((B, T, L), dtype=torch.float32, device=hidden.device)
    fn = _lib.load().kbner_gather_tagproj_fwd_f32 if f32 else _lib.load().kbner_gather_tagproj_fwd
    _lib.check(fn(_ptr(hidden), _ptr(row_of), _ptr(first_idx), _ptr(drop_keep),
                  _ptr(W), _ptr(bias), B, T, int(S), H, L, _ptr(logits), _stream()), "gather_tagproj_fwd")
    return logits


def gemm_bf16_tn(A, B, bias=None, residual=None, epilogue=EPI_BIAS, out=None):
    """C[M,N] = epilogue(A[M,K] @ B[N,K]^T).  bf16 out for EPI_BIAS / EPI_BIAS_GELU, fp32 otherwise."""
    _chk(A, torch.bfloat16, "A", 2)
    _chk(B, torch.bfloat16, "B", 2)
    _chk(bias, torch.float32, "bias", 1)
    _chk(residual, torch.bfloat16, "residual", 2)
    M, K = A.shape
    N, K2 = B.shape
    if K != K2:
        raise _lib.KbnerError("gemm: K mismatch %d vs %d" % (K, K2))
    odt = torch.bfloat16 if epilogue in (EPI_BIAS, EPI_BIAS_GELU) else torch.float32
    if out is None:
        out = torch.empty((M, N), dtype=odt, device=A.device)
    else:
        _chk(out, odt, "out", 2)
    _lib.check(_lib.load().kbner_gemm_bf16_tn(_ptr(A), _ptr(B), _ptr(bias), _ptr(residual), _ptr(out), M, N, K,
                                              K, K, N, int(epilogue), _stream()), "gemm_bf16_tn")
    return out


_GEMM_LN_IMPL = None      # tests: "grid" | "cluster" forces one kernel; None: the library picks (one round of tiles -> grid)


def gemm_ln_workspace(M, N, device):
    """Zeroed workspace of the fused GEMM + LayerNorm kernel for an [M, N] output (statistics slots tagged with the
    workspace's own launch epoch: one allocation serves every launch of that shape, also inside a replayed CUDA graph)."""
    n = int(_lib.load().kbner_gemm_ln_workspace_bytes(int(M), int(N)))
    if n <= 0:
        raise _lib.KbnerError("gemm_ln_workspace: unsupported shape M=%d N=%d" % (M, N))
    return torch.zeros(n, dtype=torch.uint8, device=device)


def gemm_ln(A, W, bias, resid, gamma, beta, eps, out=None, ws=None):
    """out[M,N] (bf16) = LayerNorm(A[M,K] @ W[N,K]^T + bias + resid) * gamma + beta, one kernel (N in 256..1024 step 256).
    ws: gemm_ln_workspace(M, N) of the caller (a fresh one is allocated when omitted)."""
    _chk(A, torch.bfloat16, "A", 2)
    _chk(W, torch.bfloat16, "W", 2)
    _chk(bias, torch.float32, "bias", 1)
    _chk(resid, torch.bfloat16, "resid", 2)
    _chk(gamma, torch.float32, "gamma", 1)
    _chk(beta, torch.float32, "beta", 1)
    M, K = A.shape
    N, K2 = W.shape
    if K != K2:
        raise _lib.KbnerError("gemm_ln: K mismatch %d vs %d" % (K, K2))
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=A.device)
    else:
        _chk(out, torch.bfloat16, "out", 2)
    if resid is not None and resid.data_ptr() == out.data_ptr():
        raise _lib.KbnerError("gemm_ln: out must not alias resid (other CTAs still read the residual rows)")
    if _GEMM_LN_IMPL == "cluster":
        _lib.check(_lib.load().kbner_gemm_bias_resid_layernorm(_ptr(A), _ptr(W), _ptr(bias), _ptr(resid), _ptr(gamma), _ptr(beta),
                                                               float(eps), _ptr(out), M, N, K, K, K, _stream()), "gemm_ln")
        return out
    if ws is None:
        ws = gemm_ln_workspace(M, N, A.device)
    fn = _lib.load().kbner_gemm_ln_grid if _GEMM_LN_IMPL == "grid" else _lib.load().kbner_gemm_bias_resid_layernorm_ws
    _lib.check(fn(_ptr(A), _ptr(W), _ptr(bias), _ptr(resid), _ptr(gamma), _ptr(beta), float(eps), _ptr(out), M, N, K, K, K,
                  _ptr(ws), ws.numel(), _stream()), "gemm_ln")
    return out


def attention_fwd(qkv, key_len, R, S, heads, out=None, want_lse=False, drop=None, split=False, out_lo=None):
    """out [R*S,H] bf16; split=True: out is the [R*S,3H] operand of the attention-output GEMM, rows hi|lo|hi;
    out_lo (optional [R*S,H] bf16, not with split): receives the rounding residual bf16(o - out)."""
    _chk(qkv, torch.bfloat16, "qkv", 2)
    _chk(key_len, torch.int32, "key_len", 1)
    H = heads * 64
    if qkv.shape != (R * S, 3 * H):
        raise _lib.KbnerError("attention: qkv must be [R*S, 3*H] = [%d, %d], got %s" % (R * S, 3 * H, tuple(qkv.shape)))
    width = 3 * H if split else H
    if out is None:
        out = torch.empty((R * S, width), dtype=torch.bfloat16, device=qkv.device)
    else:
        _chk(out, torch.bfloat16, "out", 2)
        if tuple(out.shape) != (R * S, width):
            raise _lib.KbnerError("attention: out must be [%d, %d]" % (R * S, width))
    lse = torch.empty((R, heads, S), dtype=torch.float32, device=qkv.device) if want_lse else None
    sp, site, p = _drop_args(drop)
    _chk(out_lo, torch.bfloat16, "out_lo", 2)
    if split:
        if out_lo is not None:
            raise _lib.KbnerError("attention: split already writes the residual")
        lo_p, hi2_p = out.data_ptr() + 2 * H, out.data_ptr() + 4 * H
    else:
        if out_lo is not None and tuple(out_lo.shape) != (R * S, H):
            raise _lib.KbnerError("attention: out_lo must be [%d, %d]" % (R * S, H))
        lo_p, hi2_p = _ptr(out_lo), 0
    _lib.check(_lib.load().kbner_attention_fwd_ex(_ptr(qkv), _ptr(key_len), R, S, heads, _ptr(out), width, lo_p, hi2_p,
                                                  _ptr(lse), sp, site, p, _stream()), "attention_fwd")
    return (out, lse) if want_lse else out


# ---- fine-tuning step -----------------------------------------------------------------------------
def layernorm_bwd(x, dout, gamma, mean, rstd, dgamma, dbeta, out=None, dxsum=None, bias=None, resid=None, dres=None,
                  drop=None, out_masked=None):
    """Backward of layernorm_fwd.  Returns dx (bf16 [M,H], grad w.r.t. the LayerNorm input z = dropout(x+bias)+resid); with
    dropout returns (dx, dx_masked) where dx_masked is the gradient of x (through the mask).  dgamma / dbeta (fp32 [H]) are
    accumulated into; dxsum (optional) += colsum(dx_masked or dx) = the bias gradient; dres (bf16) is added to dout."""
    _chk(x, torch.float32, "x", 2)
    _chk(dout, torch.float32, "dout", 2)
    for n, t in (("gamma", gamma), ("mean", mean), ("rstd", rstd), ("dgamma", dgamma), ("dbeta", dbeta)):
        _chk(t, torch.float32, n, 1)
    _chk(bias, torch.float32, "bias", 1)
    _chk(resid, torch.bfloat16, "resid", 2)
    _chk(dres, torch.bfloat16, "dres", 2)
    M, H = x.shape
    if out is None:
        out = torch.empty((M, H), dtype=torch.bfloat16, device=x.device)
    _chk(dxsum, torch.float32, "dxsum", 1)
    sp, site, p = _drop_args(drop)
    if sp is not None and p > 0.0 and out_masked is None:
        out_masked = torch.empty((M, H), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().kbner_add_layernorm_bwd(_ptr(x), _ptr(bias), _ptr(resid), _ptr(dout), _ptr(dres), _ptr(gamma),
                                                   _ptr(mean), _ptr(rstd), M, H, _ptr(out), _ptr(out_masked), _ptr(dgamma),
                                                   _ptr(dbeta), _ptr(dxsum), sp, site, p, _stream()), "layernorm_bwd")
    return (out, out_masked) if out_masked is not None else out


def colsum_bf16(dY, db):
    _chk(dY, torch.bfloat16, "dY", 2)
    _chk(db, torch.float32, "db", 1)
    M, N = dY.shape
    _lib.check(_lib.load().kbner_colsum_bf16(_ptr(dY), M, N, _ptr(db), _stream()), "colsum_bf16")
    return db


def embed_ln_bwd(ids, word_emb, pos_emb, type_emb, gamma, eps, pad_id, dout, d_word, d_pos, d_type, dgamma, dbeta):
    _chk(ids, torch.int32, "ids", 2)
    _chk(dout, torch.float32, "dout", 2)
    for n, t in (("d_word", d_word), ("d_pos", d_pos), ("d_type", d_type), ("dgamma", dgamma), ("dbeta", dbeta)):
        _chk(t, torch.float32, n)
    R, S = ids.shape
    H = word_emb.shape[1]
    _lib.check(_lib.load().kbner_embed_ln_bwd(_ptr(ids), _ptr(word_emb), _ptr(pos_emb), _ptr(type_emb), _ptr(gamma),
                                              float(eps), int(pad_id), R, S, H, _ptr(dout), _ptr(d_word), _ptr(d_pos),
                                              _ptr(d_type), _ptr(dgamma), _ptr(dbeta), _stream()), "embed_ln_bwd")


def gather_tagproj_bwd(hidden, row_of, first_idx, W, dlogits, S, d_hidden, dW, db, drop_keep=None):
    _chk(hidden, torch.bfloat16, "hidden", 2)
    _chk(dlogits, torch.float32, "dlogits", 3)
    _chk(d_hidden, torch.float32, "d_hidden", 2)
    _chk(dW, torch.float32, "dW", 2)
    _chk(db, torch.float32, "db", 1)
    B, T = first_idx.shape
    L, H = W.shape
    _lib.check(_lib.load().kbner_gather_tagproj_bwd(_ptr(hidden), _ptr(row_of), _ptr(first_idx), _ptr(drop_keep), _ptr(W),
                                                    _ptr(dlogits), B, T, int(S), H, L, _ptr(d_hidden), _ptr(dW),
                                                    _ptr(db), _stream()), "gather_tagproj_bwd")


def sumsq_f32(g, out):
    _chk(g, torch.float32, "g", 1)
    _chk(out, torch.float32, "out", 1)
    _lib.check(_lib.load().kbner_sumsq_f32(_ptr(g), g.numel(), _ptr(out), _stream()), "sumsq_f32")
    return out


def clip_coef(sumsq, pre_scale, max_norm, coef):
    _lib.check(_lib.load().kbner_clip_coef(_ptr(sumsq), float(pre_scale), float(max_norm), _ptr(coef), _stream()),
               "clip_coef")
    return coef


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, gscale_dev=None, gscale_host=1.0, shadow=None):
    """Fused AdamW over a flat arena.  g: fp32, or the bf16 buffer an all-reduce of packed gradients left behind.
    shadow (optional, bf16, numel <= p.numel()): receives the updated leading parameters as bf16 (compute copies)."""
    for n, t in (("p", p), ("m", m), ("v", v)):
        _chk(t, torch.float32, n, 1)
    gb = g.dtype == torch.bfloat16
    _chk(g, torch.bfloat16 if gb else torch.float32, "g", 1)
    _chk(shadow, torch.bfloat16, "shadow", 1)
    if g.numel() != p.numel():
        raise _lib.KbnerError("adamw_step: gradient has %d elements, parameters %d" % (g.numel(), p.numel()))
    _lib.check(_lib.load().kbner_adamw_step_ex(_ptr(p), None if gb else _ptr(g), _ptr(g) if gb else None, _ptr(m), _ptr(v),
                                               p.numel(), float(lr), float(beta1), float(beta2), float(eps),
                                               float(weight_decay), int(step), _ptr(gscale_dev), float(gscale_host),
                                               _ptr(shadow), 0 if shadow is None else shadow.numel(), _stream()),
               "adamw_step")


def pack_bf16(src, dst, scale=1.0):
    """dst (bf16) = src (fp32) * scale, flat buffers with a multiple of 4 elements."""
    _chk(src, torch.float32, "src", 1)
    _chk(dst, torch.bfloat16, "dst", 1)
    if src.numel() != dst.numel():
        raise _lib.KbnerError("pack_bf16: size mismatch")
    _lib.check(_lib.load().kbner_pack_bf16(_ptr(src), _ptr(dst), src.numel(), float(scale), _stream()), "pack_bf16")
    return dst


def mark_rows(ids, touched):
    """touched[id] = 1 for every id in ids (int32, any shape; ids outside [0, V) are ignored)."""
    _chk(ids, torch.int32, "ids")
    _chk(touched, torch.uint8, "touched", 1)
    _lib.check(_lib.load().kbner_mark_rows(_ptr(ids), ids.numel(), touched.numel(), _ptr(touched), _stream()), "mark_rows")


def adamw_rows(p, g, m, v, touched, lr, beta1, beta2, eps, weight_decay, step, gscale_dev=None, gscale_host=1.0):
    """AdamW over the marked rows of a [V,H] table (all fp32, same arithmetic as adamw_step)."""
    for n, t in (("p", p), ("g", g), ("m", m), ("v", v)):
        _chk(t, torch.float32, n, 2)
    _chk(touched, torch.uint8, "touched", 1)
    V, H = p.shape
    _lib.check(_lib.load().kbner_adamw_rows(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(touched), V, H, float(lr), float(beta1),
                                            float(beta2), float(eps), float(weight_decay), int(step), _ptr(gscale_dev),
                                            float(gscale_host), _stream()), "adamw_rows")


def sumsq_rows(g, touched, out, partials):
    """out[0] += sum of squares over the marked rows of g [V,H] fp32, fixed summation order."""
    _chk(g, torch.float32, "g", 2)
    _chk(touched, torch.uint8, "touched", 1)
    _chk(partials, torch.float32, "partials", 1)
    _lib.check(_lib.load().kbner_sumsq_rows_det(_ptr(g), _ptr(touched), g.shape[0], g.shape[1], _ptr(partials), partials.numel(),
                                                _ptr(out), _stream()), "sumsq_rows")
    return out


def zero_rows(g, touched):
    _chk(g, torch.float32, "g", 2)
    _chk(touched, torch.uint8, "touched", 1)
    _lib.check(_lib.load().kbner_zero_rows(_ptr(g), _ptr(touched), g.shape[0], g.shape[1], _stream()), "zero_rows")


def rows_gather_bf16(src, ids, rows, zero_src=False):
    """rows[i] = bf16(src[ids[i]]) (zeros where ids[i] < 0); zero_src clears the gathered source rows.  src fp32 [V,H]."""
    _chk(src, torch.float32, "src", 2)
    _chk(ids, torch.int32, "ids", 1)
    _chk(rows, torch.bfloat16, "rows", 2)
    V, H = src.shape
    if tuple(rows.shape) != (ids.numel(), H):
        raise _lib.KbnerError("rows_gather: rows must be [%d, %d]" % (ids.numel(), H))
    _lib.check(_lib.load().kbner_rows_gather_bf16(_ptr(src), _ptr(ids), ids.numel(), V, H, _ptr(rows), int(bool(zero_src)),
                                                  _stream()), "rows_gather_bf16")
    return rows


def rows_scatter_add_bf16(rows, ids, dst):
    """dst[ids[i]] += rows[i] for ids[i] >= 0; ids unique within the call.  dst fp32 [V,H]."""
    _chk(dst, torch.float32, "dst", 2)
    _chk(ids, torch.int32, "ids", 1)
    _chk(rows, torch.bfloat16, "rows", 2)
    V, H = dst.shape
    if tuple(rows.shape) != (ids.numel(), H):
        raise _lib.KbnerError("rows_scatter_add: rows must be [%d, %d]" % (ids.numel(), H))
    _lib.check(_lib.load().kbner_rows_scatter_add_bf16(_ptr(rows), _ptr(ids), ids.numel(), V, H, _ptr(dst), _stream()),
               "rows_scatter_add_bf16")
    return dst


def sumsq(g, out, partials=None):
    """out[0] += sum g^2 over a flat fp32 or bf16 buffer.  With `partials` (fp32 scratch, >= 32 slots) the summation order
    is fixed: bit-identical across runs and across data-parallel ranks (what the clip coefficient needs)."""
    if partials is not None:
        _chk(g, g.dtype, "g", 1)
        if g.dtype not in (torch.bfloat16, torch.float32):
            raise _lib.KbnerError("sumsq: fp32 or bf16 expected")
        _chk(partials, torch.float32, "partials", 1)
        _chk(out, torch.float32, "out", 1)
        _lib.check(_lib.load().kbner_sumsq_det(_ptr(g), g.numel(), 1 if g.dtype == torch.bfloat16 else 0, _ptr(partials),
                                               partials.numel(), _ptr(out), _stream()), "sumsq_det")
        return out
    if g.dtype == torch.bfloat16:
        _chk(g, torch.bfloat16, "g", 1)
        _chk(out, torch.float32, "out", 1)
        _lib.check(_lib.load().kbner_sumsq_bf16(_ptr(g), g.numel(), _ptr(out), _stream()), "sumsq_bf16")
        return out
    return sumsq_f32(g, out)


EPI_DGELU_BF16, EPI_ACCUM_F32 = 4, 5


def gemm_bf16(A, B, M, N, K, epilogue, bias=None, aux=None, aux_out=None, out=None, a_mn=False, b_mn=False):
    """General tensor-core GEMM  C[M,N] = epilogue(sum_k A(m,k) B(n,k)).
    a_mn / b_mn: the operand is stored [K][M|N] (MN-major) instead of [M|N][K].  See include/kbner_b200.h."""
    _chk(A, torch.bfloat16, "A", 2)
    _chk(B, torch.bfloat16, "B", 2)
    _chk(bias, torch.float32, "bias", 1)
    _chk(aux, torch.bfloat16, "aux", 2)
    _chk(aux_out, torch.bfloat16, "aux_out", 2)
    exp_a = (K, M) if a_mn else (M, K)
    exp_b = (K, N) if b_mn else (N, K)
    if tuple(A.shape) != exp_a or tuple(B.shape) != exp_b:
        raise _lib.KbnerError("gemm: operand shapes %s / %s do not match M=%d N=%d K=%d (a_mn=%s b_mn=%s)" %
                              (tuple(A.shape), tuple(B.shape), M, N, K, a_mn, b_mn))
    odt = torch.bfloat16 if epilogue in (EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU_BF16) else torch.float32
    if out is None:
        if epilogue == EPI_ACCUM_F32:
            raise _lib.KbnerError("gemm: EPI_ACCUM_F32 accumulates into `out`; pass it")
        out = torch.empty((M, N), dtype=odt, device=A.device)
    else:
        _chk(out, odt, "out", 2)
    _lib.check(_lib.load().kbner_gemm_bf16(_ptr(A), _ptr(B), _ptr(bias), _ptr(aux), _ptr(aux_out), _ptr(out), M, N, K,
                                           A.shape[1], B.shape[1], N, int(bool(a_mn)), int(bool(b_mn)), int(epilogue),
                                           _stream()), "gemm_bf16")
    return out


def gemm_wgrad_group(problems):
    """One launch for up to four weight gradients over the same token rows: problems = [(dY [tokens, out] bf16,
    X [tokens, in] bf16, dW [out, in] fp32 -- accumulated into), ...]."""
    import ctypes
    n = len(problems)
    if not 1 <= n <= 4:
        raise _lib.KbnerError("gemm_wgrad_group: 1..4 problems, got %d" % n)
    tokens = problems[0][0].shape[0]
    for dy, x, dw in problems:
        _chk(dy, torch.bfloat16, "dY", 2)
        _chk(x, torch.bfloat16, "X", 2)
        _chk(dw, torch.float32, "dW", 2)
        if dy.shape[0] != tokens or x.shape[0] != tokens or tuple(dw.shape) != (dy.shape[1], x.shape[1]):
            raise _lib.KbnerError("gemm_wgrad_group: shapes dY %s X %s dW %s" % (tuple(dy.shape), tuple(x.shape), tuple(dw.shape)))
    vp, ci = ctypes.c_void_p * n, ctypes.c_int * n
    _lib.check(_lib.load().kbner_gemm_wgrad_group(
        n, vp(*[p[0].data_ptr() for p in problems]), vp(*[p[1].data_ptr() for p in problems]),
        vp(*[p[2].data_ptr() for p in problems]), ci(*[p[0].shape[1] for p in problems]), ci(*[p[1].shape[1] for p in problems]),
        ci(*[p[0].stride(0) for p in problems]), ci(*[p[1].stride(0) for p in problems]), int(tokens), _stream()),
        "gemm_wgrad_group")


def attention_bwd(qkv, out, d_out, lse, key_len, R, S, heads, dqkv=None, workspace=None, drop=None, out_lo=None):
    """dQ | dK | dV ([R*S, 3H] bf16) of attention_fwd.  workspace = (d_scratch [R,heads,S] f32, dq_acc [R*S,H] f32).
    out_lo: the forward's rounding residual (attention_fwd(out_lo=...)); D = rowsum(dO * (out + out_lo)) when given."""
    _chk(out_lo, torch.bfloat16, "out_lo", 2)
    _chk(qkv, torch.bfloat16, "qkv", 2)
    _chk(out, torch.bfloat16, "out", 2)
    _chk(d_out, torch.bfloat16, "d_out", 2)
    _chk(lse, torch.float32, "lse", 3)
    _chk(key_len, torch.int32, "key_len", 1)
    H = heads * 64
    if dqkv is None:
        dqkv = torch.empty((R * S, 3 * H), dtype=torch.bfloat16, device=qkv.device)
    if workspace is None:
        workspace = (torch.empty((R, heads, S), dtype=torch.float32, device=qkv.device),
                     torch.empty((R * S, H), dtype=torch.float32, device=qkv.device))
    sp, site, p = _drop_args(drop)
    _lib.check(_lib.load().kbner_attention_bwd_ex(_ptr(qkv), _ptr(out), _ptr(out_lo), _ptr(d_out), _ptr(lse), _ptr(key_len), R, S,
                                                  heads, _ptr(workspace[0]), _ptr(workspace[1]), _ptr(dqkv), sp, site, p,
                                                  _stream()), "attention_bwd")
    return dqkv
