"""CPU, gloo, world_size 2: the N>1 path -- sentence sharding (no data-path collective), metric-count reduction,
and the gradient all-reduce semantics of fine-tuning (mean over ranks == global batch)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kbner_b200.distributed import GradExchange, allreduce_counts, shard_indices
    from kbner_b200.encoder import EncoderConfig, ParamArena, XLMRobertaEncoderB200, _chunk_plan
    n = 11
    mine = shard_indices(n, rank, world)
    allidx = [None] * world
    dist.all_gather_object(allidx, mine)
    counts = allreduce_counts([len(set(mine)), rank + 1, 7])
    # the host logic of the exchange is exercised with a torch stand-in for the pack kernel (the product binds the CUDA one)
    pack = lambda src, dst, scale=1.0: dst.copy_(src * scale)
    # gradient averaging through the PRODUCT's exchange class: each rank holds the gradient of its own shard's mean loss in
    # a flat arena; sum over ranks / world must equal the global-batch gradient (loss is a per-batch mean,
    # sequence_tagger_model.py:2506)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 3))
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 3))
    ref.load_state_dict(model.state_dict())
    arena = ParamArena(model.parameters())
    x = torch.randn(4 * world, 8)
    y = torch.randn(4 * world, 3)
    ((ref(x) - y) ** 2).mean().backward()
    ref_flat = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    errs = {}
    for payload in ("fp32", "bf16"):
        arena.zero_grad()
        ex = GradExchange(None, [arena], payload=payload, overlap=False, pack=pack)
        loss = ((model(x[rank * 4:(rank + 1) * 4]) - y[rank * 4:(rank + 1) * 4]) ** 2).mean()
        ex.backward(loss, boundary=True)
        (g,) = ex.reduce()
        got = torch.cat([(g[arena.offsets[id(p)]:arena.offsets[id(p)] + p.numel()]).float() for p in model.parameters()]) / world
        errs[payload] = float((got - ref_flat).abs().max() / ref_flat.abs().max())
        errs[payload + "_dtype"] = str(g.dtype)
        errs[payload + "_bytes"] = ex.bytes_per_step
        if payload == "fp32":
            norm = float(torch.sqrt((g ** 2).sum()))           # the clip norm is taken post-reduce: same on both ranks
    # overlapped exchange: a stand-in for the encoder's chunked backward reports the arena slices the real one finalises
    # (encoder._chunk_plan); every rank must end with the SUM over ranks in every arena, the hook must be gone
    # afterwards, and the plan must tile the arena exactly
    cfg = EncoderConfig(name="t", vocab_size=50, hidden_size=256, num_hidden_layers=5, num_attention_heads=4,
                        intermediate_size=256, max_position_embeddings=32)
    enc = XLMRobertaEncoderB200(cfg)
    ar = enc.ensure_arena()
    plan, emb_slice = _chunk_plan(enc)
    tiles = sorted([(a, b) for _, _, a, b in plan] + [emb_slice])
    tiled = tiles[0][0] == 0 and tiles[-1][1] == ar.numel and all(tiles[i][1] == tiles[i + 1][0] for i in range(len(tiles) - 1))
    descending = all(plan[i][1] == plan[i + 1][0] + 1 for i in range(len(plan) - 1)) and plan[-1][1] == 0
    head = ParamArena([torch.nn.Parameter(torch.zeros(7))])
    total = float(sum(range(1, world + 1)))
    overlap_ok = tiled and descending
    for payload in ("fp32", "bf16"):
        ar.grad.fill_(float(rank + 1))
        head.grad.fill_(float(rank + 1))
        ex = GradExchange(enc, [ar, head], payload=payload, overlap=True, pack=pack)
        hook_seen = []

        class FakeLoss:                               # what encoder._backward_chunked does after each chunk
            def backward(self_inner):
                hook_seen.append(enc._grad_sync is not None)
                for _, _, a, b in plan:
                    enc._grad_sync(a, b)
                enc._grad_sync(*emb_slice)
        ex.backward(FakeLoss(), boundary=True)
        g_enc, g_head = ex.reduce()
        overlap_ok = (overlap_ok and hook_seen == [True] and enc._grad_sync is None and bool((g_enc.float() == total).all())
                      and bool((g_head.float() == total).all())
                      and ex.bytes_per_step == (ar.numel + head.numel) * (2 if payload == "bf16" else 4))
    # sparse exchange of an embedding table's gradient: touched rows + ids are all-gathered and added back in rank order;
    # the dense rest of the arena goes through the all-reduce; the optimizer receives (lo, hi, buffer) segments
    def rows_gather(src, ids, rows, zero_src):
        ok = ids >= 0
        rows.zero_()
        rows[ok] = src[ids[ok].long()].bfloat16()
        if zero_src:
            src[ids[ok].long()] = 0.0

    def rows_scatter_add(rows, ids, dst):
        ok = ids >= 0
        dst[ids[ok].long()] += rows[ok].float()
    V, Hd = 40, 8
    table = torch.nn.Parameter(torch.zeros(V, Hd))
    other = torch.nn.Parameter(torch.zeros(24))
    sar = ParamArena([other, table, torch.nn.Parameter(torch.zeros(8))])
    g = torch.Generator().manual_seed(100 + rank)
    my_ids = torch.randint(0, V, (6,), generator=g, dtype=torch.int32)
    sar.grad.zero_()
    dense_table = torch.zeros(V, Hd)
    for t in my_ids.tolist():                                  # what embed_ln_bwd does: scatter-add per token
        dense_table[t] += float(rank + 1) * 0.5
    tl = sar.offsets[id(table)]
    sar.grad[tl:tl + V * Hd] = dense_table.reshape(-1)
    sar.grad[:24] = float(rank + 1)
    sar.grad[tl + V * Hd:] = float(10 * (rank + 1))
    def mark_rows(ids, touched):
        touched[ids[ids >= 0].long()] = 1
    ex = GradExchange(None, [sar], payload="bf16", overlap=False,
                      kernels=dict(pack=pack, rows_gather=rows_gather, rows_scatter_add=rows_scatter_add, mark_rows=mark_rows))
    sparse_on = ex.enable_sparse_rows(sar, table, cap_tokens=8)
    ex.note_ids(my_ids[:4].view(2, 2))
    ex.note_ids(my_ids[4:].view(1, 2))
    (segs,) = ex.reduce()
    all_tables = [torch.zeros(V, Hd) for _ in range(world)]
    for r in range(world):
        gg = torch.Generator().manual_seed(100 + r)
        for t in torch.randint(0, V, (6,), generator=gg, dtype=torch.int32).tolist():
            all_tables[r][t] += float(r + 1) * 0.5
    want_table = sum(all_tables)
    got = torch.zeros(sar.numel)
    for lo, hi, buf in segs:
        got[lo:hi] = buf.float()
    sparse_ok = (sparse_on and isinstance(segs, list) and len(segs) == 3 and segs[1][2].dtype == torch.float32
                 and segs[0][2].dtype == torch.bfloat16
                 and torch.allclose(got[tl:tl + V * Hd].view(V, Hd), want_table)
                 and bool((got[:24] == total).all()) and bool((got[tl + V * Hd:] == 10 * total).all())
                 and ex._sparse["n"] == 0 and bool((ex._sparse["ids"] == -1).all())
                 # the exchange switched the arena's row skipping on and marked the rows ANY rank touched
                 and sar.row_table is not None
                 and bool((sar.row_table["touched"].bool() == (want_table.abs().sum(1) > 0)).all()))
    q.put((rank, allidx, counts, errs, norm, overlap_ok and sparse_ok))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    for rank, allidx, counts, errs, norm, overlap_ok in res:
        assert sorted(set(sum(allidx, []))) == list(range(11))          # every sentence covered
        assert len(allidx[0]) == len(allidx[1]) == 6                    # same number of steps on every rank
        assert counts[1] == 3 and counts[2] == 14
        assert errs["fp32"] < 1e-6 and errs["fp32_dtype"] == "torch.float32"   # mean of rank grads == global-batch grad
        assert errs["bf16"] < 1e-2 and errs["bf16_dtype"] == "torch.bfloat16"  # bf16 payload: one rounding per addend
        assert errs["bf16_bytes"] * 2 == errs["fp32_bytes"]
        assert overlap_ok                                               # chunked all-reduce: arena tiled, sums right, hook removed
    assert abs(res[0][4] - res[1][4]) < 1e-7                            # identical clip norm on both ranks


def test_shard_indices_edge_cases():
    from kbner_b200.distributed import shard_indices
    assert shard_indices(0, 0, 4) == []
    assert shard_indices(3, 3, 4) == [3 % 3]                             # fewer items than ranks: padded by wrapping
    assert shard_indices(8, 1, 4, pad=False) == [1, 5]
