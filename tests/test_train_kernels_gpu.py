"""GPU parity of the fine-tuning kernels (through the C ABI) against fp32 torch restatements of the same ops:
operand-layout variants of the tensor-core GEMM (dgrad / wgrad read W, dY, X in place), GELU-derivative and
accumulate epilogues, LayerNorm / embedding / tag-projection backward, bias gradients, gradient norm, AdamW."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import kbner_b200
    from kbner_b200 import ops as o
    kbner_b200._lib.check(kbner_b200._lib.load().kbner_device_check(0), "device_check")
    return o


def _rand(g, *shape, scale=0.5):
    return (torch.randn(*shape, device="cuda", generator=g) * scale)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 1024, 1024), (1000, 264, 200), (4096, 1024, 4096)])
def test_gemm_dgrad_layout(ops, M, N, K):
    """dX[M,N] = dY[M,K] . W[K,N]  -- W (torch Linear weight [out=K, in=N]) is read MN-major, no transpose."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    dy = _rand(g, M, K).bfloat16()
    w = _rand(g, K, N).bfloat16()
    out = ops.gemm_bf16(dy, w, M, N, K, ops.EPI_NONE_F32, b_mn=True)
    ref = dy.float() @ w.float()
    assert bool(((out - ref).abs() <= 1e-5 * ref.abs() + 2e-3 * math.sqrt(K / 1024.0)).all())


@pytest.mark.parametrize("T,N1,N2", [(64, 256, 256), (512, 1024, 1024), (1000, 264, 520), (4096, 4096, 1024)])
def test_gemm_wgrad_layout_accumulate(ops, T, N1, N2):
    """dW[N1,N2] += dY[T,N1]^T . X[T,N2]  -- both operands read MN-major in place, fp32 accumulate epilogue."""
    g = torch.Generator(device="cuda").manual_seed(T + N1 + N2)
    dy = _rand(g, T, N1).bfloat16()
    x = _rand(g, T, N2).bfloat16()
    dw0 = _rand(g, N1, N2)
    dw = dw0.clone()
    ops.gemm_bf16(dy, x, N1, N2, T, ops.EPI_ACCUM_F32, out=dw, a_mn=True, b_mn=True)
    ref = dw0 + dy.float().t() @ x.float()
    assert bool(((dw - ref).abs() <= 1e-5 * ref.abs() + 2e-3 * math.sqrt(T / 1024.0)).all())


def test_gemm_gelu_saves_preactivation_and_dgelu(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 512, 1024, 256
    x = _rand(g, M, K).bfloat16()
    w = _rand(g, N, K).bfloat16()
    b = _rand(g, N)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h = ops.gemm_bf16(x, w, M, N, K, ops.EPI_BIAS_GELU, bias=b, aux_out=pre)
    ref_pre = x.float() @ w.float().t() + b
    ref_h = torch.nn.functional.gelu(ref_pre)
    assert bool(((pre.float() - ref_pre).abs() <= 2.0 ** -8 * ref_pre.abs() + 1e-2).all())
    assert bool(((h.float() - ref_h).abs() <= 2.0 ** -8 * ref_h.abs() + 1e-2).all())
    # dgrad through the GELU: d_pre = (d_h . W2) * gelu'(pre)
    K2 = 256
    dy = _rand(g, M, K2).bfloat16()
    w2 = _rand(g, K2, N).bfloat16()                      # Linear(N -> K2) weight [out=K2, in=N]
    dpre = ops.gemm_bf16(dy, w2, M, N, K2, ops.EPI_DGELU_BF16, aux=pre, b_mn=True)
    p = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(p).backward(dy.float() @ w2.float())
    assert bool(((dpre.float() - p.grad).abs() <= 2.0 ** -7 * p.grad.abs() + 2e-2).all())


@pytest.mark.parametrize("H", [256, 1024])
def test_layernorm_bwd(ops, H):
    g = torch.Generator(device="cuda").manual_seed(H)
    M = 777
    x = (_rand(g, M, H, scale=2.0) + 0.3).requires_grad_(True)
    gamma = (torch.rand(H, device="cuda", generator=g) + 0.5).requires_grad_(True)
    beta = _rand(g, H, scale=0.1).requires_grad_(True)
    dout = _rand(g, M, H, scale=1.0)
    y, mean, rstd = ops.layernorm_fwd(x.detach(), gamma.detach(), beta.detach(), 1e-5, save_stats=True)
    torch.nn.functional.layer_norm(x, (H,), gamma, beta, 1e-5).backward(dout)
    dgamma = torch.zeros(H, device="cuda")
    dbeta = torch.zeros(H, device="cuda")
    dxsum = torch.zeros(H, device="cuda")
    dx = ops.layernorm_bwd(x.detach(), dout, gamma.detach(), mean, rstd, dgamma, dbeta, dxsum=dxsum)
    assert bool(((dx.float() - x.grad).abs() <= 2.0 ** -8 * x.grad.abs() + 1e-4).all())
    torch.testing.assert_close(dxsum, x.grad.sum(0), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dgamma, gamma.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dbeta, beta.grad, rtol=1e-4, atol=1e-3)


def test_colsum(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    dy = _rand(g, 1500, 3072).bfloat16()
    db = torch.ones(3072, device="cuda")
    ops.colsum_bf16(dy, db)
    torch.testing.assert_close(db, 1.0 + dy.float().sum(0), rtol=1e-4, atol=1e-3)


def test_embed_ln_bwd(ops):
    g = torch.Generator(device="cuda").manual_seed(2)
    R, S, H, V, P, pad = 3, 70, 256, 500, 80, 1
    ids = torch.randint(3, V, (R, S), device="cuda", generator=g, dtype=torch.int32)
    ids[:, 0] = 0
    ids[1, 40:] = 0
    ids[2, 5] = pad
    word = _rand(g, V, H, scale=0.05).requires_grad_(True)
    posw = _rand(g, P, H, scale=0.05).requires_grad_(True)
    typ = _rand(g, 1, H, scale=0.05).requires_grad_(True)
    gamma = (torch.rand(H, device="cuda", generator=g) + 0.5).requires_grad_(True)
    beta = _rand(g, H, scale=0.1).requires_grad_(True)
    dout = _rand(g, R * S, H, scale=1.0)
    mask = (ids != pad).int()
    position = (torch.cumsum(mask, 1) * mask + pad).long()
    x = word[ids.long()] + typ[0][None, None, :] + posw[position]
    torch.nn.functional.layer_norm(x, (H,), gamma, beta, 1e-5).reshape(R * S, H).backward(dout)
    d_word, d_pos = torch.zeros_like(word), torch.zeros_like(posw)
    d_type, dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    ops.embed_ln_bwd(ids, word.detach(), posw.detach(), typ.detach()[0].contiguous(), gamma.detach(), 1e-5, pad, dout,
                     d_word, d_pos, d_type, dg, db)
    torch.testing.assert_close(d_word, word.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(d_pos, posw.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(d_type, typ.grad[0], rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dg, gamma.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db, beta.grad, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("L,H", [(13, 1024), (29, 256)])
def test_gather_tagproj_bwd(ops, L, H):
    g = torch.Generator(device="cuda").manual_seed(L)
    R, S, B, T = 3, 64, 4, 30
    hidden = _rand(g, R * S, H, scale=1.0).bfloat16()
    row_of = torch.tensor([0, 1, 1, 2], dtype=torch.int32, device="cuda")
    first = torch.stack([torch.randperm(S - 2, device="cuda", generator=g)[:T] + 1 for _ in range(B)]).to(torch.int32)
    first[1] += 0
    first[0, 3] = -1
    first[3, 20:] = -1
    first[2] = (first[2] % 30) + 31            # sentences 1 and 2 share a window row: keep their rows disjoint
    first[1] = (first[1] % 30) + 1
    # make indices within a sentence unique
    for b in range(B):
        vals = first[b][first[b] >= 0]
        uniq = torch.unique(vals)
        first[b] = -1
        first[b, :uniq.numel()] = uniq
    W = _rand(g, L, H, scale=0.05).requires_grad_(True)
    bias = _rand(g, L).requires_grad_(True)
    keep = (torch.rand(T, device="cuda", generator=g) > 0.2).to(torch.uint8)
    dlog = _rand(g, B, T, L, scale=1.0)
    hf = hidden.float().requires_grad_(True)
    rows = row_of.long()[:, None] * S + first.long().clamp(min=0)
    live = (first >= 0).float() * keep.float()[None, :]
    ((hf[rows] * live[..., None]) @ W.t() + bias).backward(dlog)
    d_hidden = torch.zeros(R * S, H, device="cuda")
    dW, db = torch.zeros(L, H, device="cuda"), torch.zeros(L, device="cuda")
    ops.gather_tagproj_bwd(hidden, row_of, first, W.detach(), dlog, S, d_hidden, dW, db, drop_keep=keep)
    torch.testing.assert_close(d_hidden, hf.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dW, W.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(db, bias.grad, rtol=1e-4, atol=1e-3)


def test_adamw_sumsq_clip(ops):
    """transformers-3.0.0 AdamW semantics (correct_bias, eps outside sqrt, decoupled decay) + clip_grad_norm_."""
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 1_000_003
    p0 = _rand(g, n + 1)[:n].contiguous() if False else _rand(g, n)
    p = p0.clone()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    rp, rm, rv = p0.double().clone(), torch.zeros(n, device="cuda", dtype=torch.float64), torch.zeros(n, device="cuda", dtype=torch.float64)
    lr, b1, b2, eps, wd, max_norm, accum = 5e-3, 0.9, 0.999, 1e-6, 0.01, 5.0, 4
    for step in range(1, 4):
        grad = _rand(g, n, scale=3.0)
        ss = torch.zeros(1, device="cuda")
        ops.sumsq_f32(grad, ss)
        torch.testing.assert_close(ss[0], grad.double().pow(2).sum().float(), rtol=1e-4, atol=0)
        coef = torch.empty(1, device="cuda")
        ops.clip_coef(ss, 1.0 / accum, max_norm, coef)
        gn = grad.double() / accum
        ref_coef = min(1.0, max_norm / (float(gn.norm()) + 1e-6))
        assert abs(float(coef) - ref_coef) < 1e-5
        ops.adamw_step(p, grad, m, v, lr, b1, b2, eps, wd, step, gscale_dev=coef, gscale_host=1.0 / accum)
        gd = gn * ref_coef
        rm = b1 * rm + (1 - b1) * gd
        rv = b2 * rv + (1 - b2) * gd * gd
        rp = rp - lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step) * rm / (rv.sqrt() + eps)
        rp = rp - lr * wd * rp
        torch.testing.assert_close(p.double(), rp, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("R,S,heads,lens", [(2, 128, 2, [128, 60]), (2, 512, 16, [512, 300]), (3, 200, 4, [200, 129, 1]),
                                            (1, 384, 4, [257])])
def test_attention_bwd(ops, R, S, heads, lens):
    """dQ, dK, dV against autograd through an fp32 softmax attention on the same bf16 inputs.  Only query rows inside
    the window carry upstream gradient (padding rows never reach the loss), as in the model."""
    g = torch.Generator(device="cuda").manual_seed(R * 7 + S + heads)
    H = heads * 64
    qkv = _rand(g, R * S, 3 * H, scale=1.0).bfloat16()
    key_len = torch.tensor(lens, dtype=torch.int32, device="cuda")
    out, lse = ops.attention_fwd(qkv, key_len, R, S, heads, want_lse=True)
    d_out = _rand(g, R * S, H, scale=1.0)
    valid_q = (torch.arange(S, device="cuda")[None, :] < key_len[:, None]).reshape(R * S, 1)
    d_out = (d_out * valid_q).bfloat16()
    dqkv = ops.attention_bwd(qkv, out, d_out, lse, key_len, R, S, heads)
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(R, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) * 0.125
    kmask = torch.arange(S, device="cuda")[None, :] < key_len[:, None]
    sc = sc.masked_fill(~kmask[:, None, None, :], float("-inf"))
    o = (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(R * S, H)
    o.backward(d_out.float())
    ref = x.grad
    # rows beyond the window have no key/value role and no query gradient: compare rows inside the window
    rows = valid_q.squeeze(1)
    diff = (dqkv.float() - ref)[rows].abs()
    scale = ref[rows].abs().max().item()
    assert diff.max().item() < 3e-2 * max(scale, 1.0), (diff.max().item(), scale)
    assert (diff.mean() / ref[rows].abs().mean()).item() < 1e-2
    # key rows outside the window receive exactly zero dK / dV
    if (~rows).any():
        assert float(dqkv.float()[~rows][:, H:].abs().max()) == 0.0


def _seed_tensor(a, b):
    return torch.tensor([a, b], dtype=torch.int32, device="cuda")


@pytest.mark.parametrize("H,p", [(256, 0.0), (1024, 0.1), (1024, 0.3)])
def test_add_layernorm_fwd_bwd_with_dropout(ops, H, p):
    """y = LN(dropout(x + bias) + resid): forward and every gradient against torch autograd with the host-rebuilt mask."""
    from test_api_gpu import host_dropout_mask
    g = torch.Generator(device="cuda").manual_seed(H + int(p * 100))
    M, site = 555, 7
    x = _rand(g, M, H, scale=1.5).requires_grad_(True)
    bias = _rand(g, H, scale=0.3).requires_grad_(True)
    resid = _rand(g, M, H, scale=1.0).bfloat16()
    rf = resid.float().requires_grad_(True)
    gamma = (torch.rand(H, device="cuda", generator=g) + 0.5).requires_grad_(True)
    beta = _rand(g, H, scale=0.1).requires_grad_(True)
    dout = _rand(g, M, H, scale=1.0)
    dres = _rand(g, M, H, scale=0.5).bfloat16()
    seed = _seed_tensor(1234567, 89)
    drop = (seed, site, p) if p > 0 else None
    mask = (torch.from_numpy(host_dropout_mask(np.array([1234567, 89], np.uint32), site, p, n_elems=M * H)).view(M, H).cuda()
            if p > 0 else torch.ones(M, H, device="cuda"))
    y, mean, rstd = ops.layernorm_fwd(x.detach(), gamma.detach(), beta.detach(), 1e-5, save_stats=True, bias=bias.detach(),
                                      resid=resid, drop=drop)
    z = (x + bias) * mask + rf
    ref = torch.nn.functional.layer_norm(z, (H,), gamma, beta, 1e-5)
    assert bool(((y.float() - ref).abs() <= 2.0 ** -8 * ref.abs() + 1e-5).all())
    ref.backward(dout + dres.float())
    dgamma, dbeta, dxsum = (torch.zeros(H, device="cuda") for _ in range(3))
    res = ops.layernorm_bwd(x.detach(), dout, gamma.detach(), mean, rstd, dgamma, dbeta, dxsum=dxsum, bias=bias.detach(),
                            resid=resid, dres=dres, drop=drop)
    dz, dxm = res if isinstance(res, tuple) else (res, res)
    assert (p > 0) == isinstance(res, tuple)
    assert bool(((dz.float() - rf.grad).abs() <= 2.0 ** -8 * rf.grad.abs() + 1e-4).all())        # residual path: unmasked
    assert bool(((dxm.float() - x.grad).abs() <= 2.0 ** -8 * x.grad.abs() + 1e-4).all())         # through the mask
    torch.testing.assert_close(dxsum, bias.grad, rtol=1e-4, atol=2e-3)
    torch.testing.assert_close(dgamma, gamma.grad, rtol=1e-4, atol=2e-3)
    torch.testing.assert_close(dbeta, beta.grad, rtol=1e-4, atol=2e-3)


def test_dropout_apply_bf16_and_f32(ops):
    from test_api_gpu import host_dropout_mask
    g = torch.Generator(device="cuda").manual_seed(5)
    M, H, site, p = 333, 1024, 96, 0.1
    mask = torch.from_numpy(host_dropout_mask(np.array([42, 4242], np.uint32), site, p, n_elems=M * H)).view(M, H).cuda()
    seed = _seed_tensor(42, 4242)
    xf = _rand(g, M, H, scale=1.0)
    xb = xf.bfloat16()
    want_f, want_b = xf * mask, (xb.float() * mask).bfloat16()
    ops.dropout_apply(xf, (seed, site, p))
    ops.dropout_apply(xb, (seed, site, p))
    torch.testing.assert_close(xf, want_f, rtol=1e-6, atol=0)
    assert torch.equal(xb, want_b)


@pytest.mark.parametrize("R,S,heads,lens", [(2, 512, 16, [512, 300]), (3, 200, 4, [200, 129, 1])])
def test_attention_fwd_bwd_with_dropout(ops, R, S, heads, lens):
    """Attention-probability dropout: O and dQ / dK / dV against autograd through softmax -> mask -> P.V with the
    host-rebuilt mask (same tolerances as the dropout-free test)."""
    from test_api_gpu import host_dropout_mask
    g = torch.Generator(device="cuda").manual_seed(R * 11 + S + heads)
    H, p, site = heads * 64, 0.1, 40
    qkv = _rand(g, R * S, 3 * H, scale=1.0).bfloat16()
    key_len = torch.tensor(lens, dtype=torch.int32, device="cuda")
    seed = _seed_tensor(2024, 925)
    drop = (seed, site, p)
    out, lse = ops.attention_fwd(qkv, key_len, R, S, heads, want_lse=True, drop=drop)
    out0, lse0 = ops.attention_fwd(qkv, key_len, R, S, heads, want_lse=True)
    torch.testing.assert_close(lse, lse0, rtol=0, atol=0)            # the softmax statistics ignore the mask
    mask = torch.from_numpy(host_dropout_mask(np.array([2024, 925], np.uint32), site, p, attn=(R, heads, S))).cuda()
    d_out = _rand(g, R * S, H, scale=1.0)
    valid_q = (torch.arange(S, device="cuda")[None, :] < key_len[:, None]).reshape(R * S, 1)
    d_out = (d_out * valid_q).bfloat16()
    dqkv = ops.attention_bwd(qkv, out, d_out, lse, key_len, R, S, heads, drop=drop)
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(R, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) * 0.125
    kmask = torch.arange(S, device="cuda")[None, :] < key_len[:, None]
    sc = sc.masked_fill(~kmask[:, None, None, :], float("-inf"))
    o = ((torch.softmax(sc, -1) * mask) @ v).permute(0, 2, 1, 3).reshape(R * S, H)
    rows = valid_q.squeeze(1)
    do = (out.float() - o.detach())[rows].abs()
    assert do.max().item() < 3e-2 and do.mean().item() < 3e-3, (do.max().item(), do.mean().item())
    assert (out.float() - out0.float())[rows].abs().max().item() > 0.05          # the mask really was applied
    o.backward(d_out.float())
    ref = x.grad
    diff = (dqkv.float() - ref)[rows].abs()
    scale = ref[rows].abs().max().item()
    assert diff.max().item() < 3e-2 * max(scale, 1.0), (diff.max().item(), scale)
    assert (diff.mean() / ref[rows].abs().mean()).item() < 1e-2


@pytest.mark.parametrize("tokens", [4096, 1000])
def test_gemm_wgrad_group_matches_the_separate_launches(ops, tokens):
    """The four weight gradients of a layer in one stream-K launch (csrc/gemm_group_tcgen05.cu) against fp32 matmuls and
    against four kbner_gemm_bf16 launches with the reduce-add epilogue; dW is accumulated INTO (gradient accumulation over
    micro-batches), so it starts non-zero.  1000 tokens: a K extent that is not a multiple of the 64-token k-block."""
    g = torch.Generator(device="cuda").manual_seed(tokens)
    shapes = [(1024, 4096), (4096, 1024), (1024, 1024), (3072, 1024)]       # (out, in): FFN-down, FFN-up, attention-out, QKV
    probs, refs, seps = [], [], []
    for n_out, n_in in shapes:
        dy = (torch.randn(tokens, n_out, device="cuda", generator=g) * 0.1).bfloat16()
        x = (torch.randn(tokens, n_in, device="cuda", generator=g) * 0.5).bfloat16()
        dw0 = torch.randn(n_out, n_in, device="cuda", generator=g)
        probs.append((dy, x, dw0.clone()))
        refs.append(dw0 + dy.float().t() @ x.float())
        sep = dw0.clone()
        ops.gemm_bf16(dy, x, n_out, n_in, tokens, ops.EPI_ACCUM_F32, out=sep, a_mn=True, b_mn=True)
        seps.append(sep)
    ops.gemm_wgrad_group(probs)
    for (dy, x, dw), ref, sep in zip(probs, refs, seps):
        tol = 1e-5 * ref.abs() + 2e-3 * math.sqrt(tokens / 1024.0)
        assert bool(((dw - ref).abs() <= tol).all()), float((dw - ref).abs().max())
        assert bool(((dw - sep).abs() <= 1e-5 * sep.abs() + 1e-4).all()), float((dw - sep).abs().max())
    # a group of one, and of two problems with ragged extents
    dy = (torch.randn(300, 200, device="cuda", generator=g) * 0.1).bfloat16()
    x = (torch.randn(300, 72, device="cuda", generator=g) * 0.5).bfloat16()
    dw = torch.zeros(200, 72, device="cuda")
    dw2 = torch.zeros(200, 72, device="cuda")
    ops.gemm_wgrad_group([(dy, x, dw)])
    ops.gemm_wgrad_group([(dy, x, dw2), (x, dy, torch.zeros(72, 200, device="cuda"))])
    ref = dy.float().t() @ x.float()
    assert bool(((dw - ref).abs() <= 1e-5 * ref.abs() + 1e-3).all()) and bool(((dw2 - ref).abs() <= 1e-5 * ref.abs() + 1e-3).all())
