"""Host-side restatement of the group kernel's layouts (csrc/qmk_device2.cuh) and the properties they must have.
Pure index arithmetic, no GPU: the K permutation + chunk swizzle of a packed weight row is a bijection and hands lane q
of warp w the contiguous activation run it reads; the attention split policy covers every position exactly once."""

import itertools

import pytest


def pack_row_word_order(k_words: int, nk: int):
    """source word index (32-bit = 2 bf16) stored at packed word position p of a row whose K block has k_words words.
    pack_chunk2: chunk ch -> slice w = ch / (2 nk), m = ch % (2 nk); the chunk holds source words w*8nk + 2nk*q + m, q = 0..3."""
    chunks = k_words // 4
    order = []
    for ch in range(chunks):
        w, m = divmod(ch, 2 * nk)
        order += [w * 8 * nk + 2 * nk * q + m for q in range(4)]
    return order


@pytest.mark.parametrize("k_elems,nk", [(1024, 8), (256, 2), (384, 3)])
def test_k_permutation_is_a_bijection_and_matches_the_b_fragments(k_elems, nk):
    k_words = k_elems // 2
    order = pack_row_word_order(k_words, nk)
    assert sorted(order) == list(range(k_words))                       # every source word exactly once
    # tensor-core step j of warp w uses chunks 2j (k-half 0) and 2j+1 (k-half 1) of the warp's slice; lane q = lane % 4 gets
    # packed word q of each chunk as its A columns (2q, 2q+1) / (2q+8, 2q+9) and must find the matching B values at
    # activation elements  slice*16nk + q*4nk + 4j + {0,1}  and  + {2,3}  (mma_tiles: contiguous run of 4 nk elements)
    for w, j, q in itertools.product(range(8), range(nk), range(4)):
        for half in (0, 1):
            ch = w * 2 * nk + 2 * j + half
            src_word = order[ch * 4 + q]
            want_elem = w * 16 * nk + q * 4 * nk + 4 * j + 2 * half
            assert src_word * 2 == want_elem


@pytest.mark.parametrize("row_bytes", [2048, 512, 768])
def test_chunk_swizzle_is_conflict_free_for_ldmatrix(row_bytes):
    """chunk t of tile row r is stored at chunk t ^ (r & 7): a bijection inside the row, and the eight rows of an 8x8
    ldmatrix tile (same logical chunk) land in eight different 16-byte bank groups"""
    chunks = row_bytes // 16
    for r in range(16):
        assert sorted(t ^ (r & 7) for t in range(chunks)) == list(range(chunks))
    for t in range(chunks):
        for r0 in (0, 8):
            groups = {((r * row_bytes) // 16 + (t ^ (r & 7))) % 8 for r in range(r0, r0 + 8)}
            assert len(groups) == 8


def attn_item2(position: int, j: int, att_round=40, solo_rounds=2, s_max=16):
    n = position + 1
    s0 = 1 if n <= att_round * solo_rounds else min(s_max, -(-n // att_round))
    c = -(-n // s0)
    s = -(-n // c)
    if s == 1:
        return s, 0, n, True
    p0 = j * c
    return s, p0, min(p0 + c, n), j < s


@pytest.mark.parametrize("position", [0, 15, 16, 39, 40, 79, 80, 81, 159, 639, 640, 1000, 2047, 8191])
def test_attention_split_covers_every_position_once(position):
    n = position + 1
    s = attn_item2(position, 0)[0]
    if s == 1:
        assert n <= 80 and attn_item2(position, 7)[1:3] == (0, n)    # every CTA of the group covers the whole context
        return
    covered = []
    for j in range(16):
        _, p0, p1, has = attn_item2(position, j)
        if has:
            assert p1 > p0
            covered += list(range(p0, p1))
        assert has == (j < s)
    assert covered == list(range(n))
    assert s <= 16


def batched_split_chunk(position: int, z: int, nz: int):
    """kb_qkv_attention_split (csrc/qmk_batched.cu): chunk z of a context of position + 1 rows split over nz CTAs"""
    n = position + 1
    c = ((-(-n // nz)) + 7) & ~7
    p_lo = z * c
    p_hi = min(p_lo + c, n)
    owner = p_lo <= position < p_lo + c
    return p_lo, p_hi, owner


@pytest.mark.parametrize("nz", [2, 4])
@pytest.mark.parametrize("position", [0, 5, 7, 8, 255, 256, 257, 1023, 1500, 2046, 2047])
def test_batched_attention_split_covers_every_position_once(position, nz):
    """every position belongs to exactly one chunk, exactly one chunk appends the new row, trailing chunks may be empty"""
    covered, owners = [], 0
    for z in range(nz):
        p_lo, p_hi, owner = batched_split_chunk(position, z, nz)
        covered += list(range(p_lo, max(p_lo, p_hi)))
        owners += owner
        if owner:
            assert p_lo <= position < p_hi
    assert covered == list(range(position + 1)) and owners == 1


def test_group_row_assignment_partitions_every_matrix():
    """QKV / gate-up rows and the O / down slabs of the 8 x 16 CTAs tile every weight matrix exactly once"""
    q, k, v, gu, o, d = set(), set(), set(), set(), set(), set()
    for g, j in itertools.product(range(8), range(16)):
        for r in range(16):
            q.add(2 * g * 128 + 16 * j + r)
        for r in range(8):
            k.add(g * 128 + 8 * j + r)
            v.add(g * 128 + 8 * j + r)
        for r in range(24):
            gu.add(384 * g + 24 * j + r)
        for r in range(64):
            o.update((64 * j + r, c) for c in range(256 * g, 256 * g + 256, 64))   # sampled columns of the K block
            d.update((64 * j + r, c) for c in range(384 * g, 384 * g + 384, 96))
    assert q == set(range(2048)) and k == set(range(1024)) and v == set(range(1024)) and gu == set(range(3072))
    assert o == {(r, c) for r in range(1024) for c in range(0, 2048, 64)}
    assert d == {(r, c) for r in range(1024) for c in range(0, 3072, 96)}
