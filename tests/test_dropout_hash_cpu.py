"""Integer identities the attention backward's dropout path relies on (csrc/attention_bwd_tcgen05.cu), restated in numpy
uint32: the incremental hash input equals the reference form drop_bits(key, pair) of csrc/common.cuh for the column each
lane of a pair hashes, and the shifted 32-bit compare equals the 16-bit keep rule."""
import numpy as np

C = np.uint32(0x9E3779B1)


def fmix32(x):
    x = x.astype(np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x85EBCA6B)
    x ^= x >> np.uint32(13)
    x *= np.uint32(0xC2B2AE35)
    x ^= x >> np.uint32(16)
    return x


def test_incremental_hash_input_matches_drop_bits():
    rng = np.random.default_rng(0)
    with np.errstate(over="ignore"):
        for _ in range(200):
            dkey = np.uint32(rng.integers(0, 2 ** 32))
            dwin = np.uint32(rng.integers(0, 4096) * 512)
            i = np.uint32(rng.integers(0, 4))
            key_row = np.uint32(rng.integers(0, 512))
            kpair, odd = key_row >> np.uint32(1), key_row & np.uint32(1)
            col = np.arange(0, 128, 2, dtype=np.uint32)                       # even query column of each pair
            hbase = ((dwin + i * np.uint32(128) + odd) * np.uint32(256) + kpair) * C + dkey
            mine = fmix32(hbase + col * (np.uint32(256) * C))
            ref = fmix32(((dwin + i * np.uint32(128) + col + odd) * np.uint32(256) + kpair) * C + dkey)   # drop_bits(dkey, pair)
            assert np.array_equal(mine, ref)


def test_shifted_compare_is_the_16_bit_keep_rule():
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2 ** 32, size=200000, dtype=np.uint64).astype(np.uint32)
    for thresh in (0, 1, 6554, 19661, 65535):
        thr = np.uint32(thresh)
        dthr = np.uint32(thresh << 16)
        keep_hi = (bits >> np.uint32(16)) >= thr                            # drop_keep_hi
        keep_lo = (bits & np.uint32(0xFFFF)) >= thr                         # drop_keep_lo
        with np.errstate(over="ignore"):
            assert np.array_equal((bits << np.uint32(0)) >= dthr, keep_hi)   # odd key: dshift = 0
            assert np.array_equal((bits << np.uint32(16)) >= dthr, keep_lo)  # even key: dshift = 16
