"""CPU checks of the arithmetic the walk kernel relies on (no GPU): the bucket sampling index
selects exactly the edge of the flat inverse-CDF rule for EVERY possible draw, the bytewise
compare / byte-permute tricks of the device step are what the format restatement says, the
split 85-bit product equals floor(k53 * S / 2^53), and fp32 division gives the reference's
float64-division-then-cast weights."""
import numpy as np
import pytest

from oracle import oracle as O


def _graph(rng, degs, wmax, n_dst=4000):
    src = np.concatenate([np.full(d, v) for v, d in enumerate(degs)]).astype(np.int64)
    dst = rng.integers(0, n_dst, size=src.size).astype(np.int64)
    w = (0.5 * rng.integers(1, wmax + 1, size=src.size)).astype(np.float32)
    perm = rng.permutation(src.size)
    ei = np.stack([src[perm], dst[perm]])
    return O.csr_build(ei, w[perm], max(len(degs), n_dst), 1)


@pytest.mark.parametrize("slots", [8, 6])
@pytest.mark.parametrize("wmax", [1, 10, 60, 300])
def test_bucket_index_picks_the_flat_rule_edge_for_every_t(wmax, slots):
    """slots = 8: PB200_LEAF_BUCKET (24-bit ids); slots = 6: PB200_LEAF_BUCKET32 (32-bit ids, ids up to 2^31)."""
    rng = np.random.Generator(np.random.PCG64(wmax))
    degs = [0, 1, 2, 5, 6, 7, 8, 9, 10, 17, 64, 65, 300, 1000]
    row_ptr, col, cum = _graph(rng, degs, wmax)
    if slots == 6:
        col = (col.astype(np.int64) * 500000 + 7).astype(np.int32)      # ids up to 2e9: beyond 24 bits
        assert col.max() > (1 << 30)
    built = O.walk_bucket_index(row_ptr, col, cum, slots)
    assert built is not None
    meta, leaf = built
    assert meta[:, 3].max() <= 7 and leaf[:, :8].max() <= 128
    for v, d in enumerate(degs):
        a = int(row_ptr[v])
        S = int(cum[a + d - 1]) if d else 0
        assert meta[v, 1] == d and meta[v, 2] == S
        if d == 0:
            assert O.walk_bucket_pick(meta, leaf, v, 123, slots) == -1
            continue
        # every t in [0, S): drive the pick with a k53 that lands exactly on t
        ts = np.arange(S) if S <= 6000 else np.unique(np.concatenate(
            [np.arange(3000), rng.integers(0, S, 3000), np.arange(S - 3000, S)]))
        want = col[a + np.searchsorted(cum[a:a + d], ts.astype(np.uint32), side="right")]
        for t, wnt in zip(ts.tolist(), want.tolist()):
            k53 = -((-t << 53) // S)                      # smallest k with floor(k S / 2^53) == t
            assert (k53 * S) >> 53 == t and k53 < (1 << 53)
            assert O.walk_bucket_pick(meta, leaf, v, k53, slots) == wnt


def test_bucket_index_refuses_zero_weight_edges():
    ei = np.array([[0, 0, 0], [1, 2, 3]], np.int64)
    w = np.array([1.0, 0.0, 2.0], np.float32)
    row_ptr, col, cum = O.csr_build(ei, w, 4, 1)
    assert O.walk_bucket_index(row_ptr, col, cum) is None


def _byte_perm(x, y, s):
    b = [(x >> (8 * i)) & 255 for i in range(4)] + [(y >> (8 * i)) & 255 for i in range(4)]
    return sum(b[(s >> (4 * i)) & 7] << (8 * i) for i in range(4))


def test_device_step_bit_tricks_match_the_format():
    """bucket_step (csrc/walk_topt.cu): count of rel <= tr via (0x80 + tr - rel) bit 7 per byte, and
    the 24-bit id assembled from the three byte planes with four byte permutes."""
    rng = np.random.Generator(np.random.PCG64(5))
    for _ in range(2000):
        n = int(rng.integers(1, 9))
        rel = np.sort(rng.integers(1, 129, n))
        rel = np.concatenate([rel, np.full(8 - n, 128)]).astype(np.int64)
        tr = int(rng.integers(0, 128))
        w0 = sum(int(rel[i]) << (8 * i) for i in range(4))
        w1 = sum(int(rel[4 + i]) << (8 * i) for i in range(4))
        t4 = (tr * 0x01010101 + 0x80808080) & 0xFFFFFFFF
        c = bin(((t4 - w0) & 0xFFFFFFFF) & 0x80808080).count("1") + bin(((t4 - w1) & 0xFFFFFFFF) & 0x80808080).count("1")
        assert c == int(np.sum(rel <= tr))
        if c > 7:
            continue
        ids = rng.integers(0, 1 << 24, 8)
        planes = [[sum(((int(ids[4 * h + i]) >> (8 * p)) & 255) << (8 * i) for i in range(4)) for h in range(2)]
                  for p in range(3)]
        r = [_byte_perm(planes[p][0], planes[p][1], c) for p in range(3)]
        got = _byte_perm(_byte_perm(r[0], r[1], 0x0040), r[2], 0x0410) & 0xFFFFFF
        assert got == int(ids[c])


def test_device_step_bit_tricks_six_slot_form():
    """bucket_pick<true>: the count masks bytes 6..7 of the second word (they hold 128), the id is word 2 + c."""
    rng = np.random.Generator(np.random.PCG64(15))
    for _ in range(2000):
        n = int(rng.integers(1, 7))
        rel = np.sort(rng.integers(1, 129, n))
        rel = np.concatenate([rel, np.full(8 - n, 128)]).astype(np.int64)
        tr = int(rng.integers(0, 128))
        w0 = sum(int(rel[i]) << (8 * i) for i in range(4))
        w1 = sum(int(rel[4 + i]) << (8 * i) for i in range(4))
        t4 = (tr * 0x01010101 + 0x80808080) & 0xFFFFFFFF
        c = bin(((t4 - w0) & 0xFFFFFFFF) & 0x80808080).count("1") + bin(((t4 - w1) & 0xFFFFFFFF) & 0x00008080).count("1")
        assert c == int(np.sum(rel[:6] <= tr))
        if c > 5:
            continue
        ids = rng.integers(0, 1 << 31, 6).tolist()
        odd = c & 1
        lo, mid, hi = (ids[1] if odd else ids[0]), (ids[3] if odd else ids[2]), (ids[5] if odd else ids[4])
        assert (hi if c >= 4 else (mid if c >= 2 else lo)) == ids[c]


def test_split_product_equals_floor_k53_total_over_2_53():
    rng = np.random.Generator(np.random.PCG64(6))
    a = rng.integers(0, 1 << 32, 20000, dtype=np.uint64)
    b = rng.integers(0, 1 << 32, 20000, dtype=np.uint64)
    S = np.concatenate([rng.integers(1, 1 << 32, 19990, dtype=np.uint64),
                        np.array([1, 2, 3, (1 << 32) - 1, (1 << 32) - 2, 1 << 31, 255, 256, 65535, 65536], np.uint64)])
    for ai, bi, si in zip(a.tolist(), b.tolist(), S.tolist()):
        k53 = ((ai >> 5) << 26) | (bi >> 6)
        kh, kl = ai >> 11, (((ai >> 5) << 26) | (bi >> 6)) & 0xFFFFFFFF
        assert (kh << 32) | kl == k53
        t = (kh * si + ((kl * si) >> 32)) >> 21
        assert t == (k53 * si) >> 53 and t < (1 << 32)


def test_fp32_division_equals_float64_division_then_cast():
    """weights = count / sum(kept counts) in Python floats, cast by torch.tensor(list) to fp32
    (model/pinsage.py:140).  For count <= total <= 255 the correctly rounded fp32 quotient is the
    same number, so the kernel divides in fp32."""
    c = np.arange(0, 256, dtype=np.int64)[:, None]
    t = np.arange(1, 256, dtype=np.int64)[None, :]
    via64 = (c.astype(np.float64) / t.astype(np.float64)).astype(np.float32)
    via32 = c.astype(np.float32) / t.astype(np.float32)
    mask = c <= t
    np.testing.assert_array_equal(via64[mask], via32[mask])
