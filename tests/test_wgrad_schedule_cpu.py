"""Work assignment of the grouped weight-gradient launch, restated from csrc/gemm_group_tcgen05.cu (group_next_item and the
host set-up in kbner_gemm_wgrad_group): every (problem, tile, k-block) is computed exactly once, in both modes, and the
default mode keeps the tiles of a round in lockstep (whole tiles, all of K, `lanes` clusters)."""
import collections

import pytest

BM = BN = 256
BK = 64


def schedule(n_out, n_in, tokens, pairs, stream):
    num_kb = (tokens + BK - 1) // BK
    unit_begin, num_n, units = [], [], 0
    for o, i in zip(n_out, n_in):
        nn = (i + BN - 1) // BN
        num_n.append(nn)
        unit_begin.append(units)
        units += ((o + BM - 1) // BM) * nn * num_kb
    unit_begin.append(units)
    count = len(n_out)
    items = []                                            # (cluster, p, m_blk, n_blk, kb0, kb1)
    if stream:
        U = (units + pairs - 1) // pairs
        U += U & 1
        U = max(U, 4)
        clusters = (units + U - 1) // U
        for c in range(clusters):
            begin, end, cursor = c * U, min(c * U + U, units), 0
            while begin + cursor < end:
                pos = begin + cursor
                p = 0
                while p + 1 < count and pos >= unit_begin[p + 1]:
                    p += 1
                local = pos - unit_begin[p]
                tile = local // num_kb
                kb0 = local - tile * num_kb
                kb1 = min(num_kb, kb0 + (end - pos))
                items.append((c, p, tile // num_n[p], tile % num_n[p], kb0, kb1))
                cursor += kb1 - kb0
    else:
        tiles = units // num_kb
        rounds = (tiles + pairs - 1) // pairs
        lanes = (tiles + rounds - 1) // rounds
        clusters = lanes
        for c in range(clusters):
            cursor = 0
            while True:
                w = c + cursor * lanes
                pos = w * num_kb
                if pos >= units:
                    break
                p = 0
                while p + 1 < count and pos >= unit_begin[p + 1]:
                    p += 1
                tile = (pos - unit_begin[p]) // num_kb
                items.append((c, p, tile // num_n[p], tile % num_n[p], 0, num_kb))
                cursor += 1
    return items, clusters, num_kb


LAYER = ([1024, 4096, 1024, 3072], [4096, 1024, 1024, 1024])       # FFN-down, FFN-up, attention-out, QKV (out, in)


@pytest.mark.parametrize("stream", [False, True])
@pytest.mark.parametrize("shapes,tokens,pairs", [(LAYER, 4096, 74), (LAYER, 1000, 74), (LAYER, 4096, 66),
                                                 (([200, 72], [72, 200]), 300, 74), (([256], [256]), 64, 74)])
def test_every_k_block_of_every_tile_exactly_once(shapes, tokens, pairs, stream):
    n_out, n_in = shapes
    items, clusters, num_kb = schedule(n_out, n_in, tokens, pairs, stream)
    assert 1 <= clusters <= pairs
    seen = collections.Counter()
    for c, p, mb, nb, kb0, kb1 in items:
        assert 0 <= kb0 < kb1 <= num_kb and mb * BM < n_out[p] and nb * BN < n_in[p]
        for kb in range(kb0, kb1):
            seen[(p, mb, nb, kb)] += 1
    want = sum(((o + BM - 1) // BM) * ((i + BN - 1) // BN) * num_kb for o, i in zip(n_out, n_in))
    assert len(seen) == want and set(seen.values()) == {1}


def test_default_mode_is_three_lockstep_rounds_for_a_layer():
    items, clusters, num_kb = schedule(*LAYER, 4096, 74, stream=False)
    assert clusters == 64 and len(items) == 192 and num_kb == 64
    per_cluster = collections.defaultdict(list)
    for c, p, mb, nb, kb0, kb1 in items:
        per_cluster[c].append(p)
        assert (kb0, kb1) == (0, 64)                         # whole tiles: one reduce-add per output element
    # round r of every cluster belongs to the same problem group: FFN-down, FFN-up, then attention-out / QKV
    assert all(v[0] == 0 and v[1] == 1 and v[2] in (2, 3) for v in per_cluster.values())
    assert sum(1 for v in per_cluster.values() if v[2] == 2) == 16
