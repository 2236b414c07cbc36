"""CPU: the oracle's multimodal-RoPE row selection and rotation against vectors generated from HF transformers' own
functions (tests/golden/make_golden_mrope.py -> tests/golden/mrope.npz)."""

import os

import numpy as np
import torch

from conftest import GOLDEN, bf16_from_bits
from oracle.tts_oracle import _rope, mrope_axis_map, mrope_rows
from qwen_megakernel.synthetic import rope_tables


def test_axis_maps():
    ch = mrope_axis_map((24, 20, 20), False)
    assert ch == [0] * 24 + [1] * 20 + [2] * 20
    il = mrope_axis_map((24, 20, 20), True)
    assert il[:6] == [0, 1, 2, 0, 1, 2] and il[57:] == [0, 1, 2, 0, 0, 0, 0] and (il.count(0), il.count(1), il.count(2)) == (24, 20, 20)


def test_chunked_mrope_matches_hf_bit_exactly():
    g = np.load(os.path.join(GOLDEN, "mrope.npz"))
    cos_t, sin_t = rope_tables(256)
    axis = mrope_axis_map(tuple(int(v) for v in g["section"]), False)
    q, k = bf16_from_bits(g["q_bits"]), bf16_from_bits(g["k_bits"])          # [1, 2, n, 128], [1, 1, n, 128]
    q_ref, k_ref = bf16_from_bits(g["q_chunked_bits"]), bf16_from_bits(g["k_chunked_bits"])
    for i, pos in enumerate(g["positions"]):
        c, s = mrope_rows(cos_t, sin_t, tuple(int(v) for v in pos), axis)
        for h in range(2):
            assert torch.equal(_rope(q[0, h, i], c, s), q_ref[0, h, i])
        assert torch.equal(_rope(k[0, 0, i], c, s), k_ref[0, 0, i])
    # equal positions on all axes = the standard table row
    c, s = mrope_rows(cos_t, sin_t, (5, 5, 5), axis)
    assert torch.equal(c, cos_t[5]) and torch.equal(s, sin_t[5])


def test_interleaved_mrope_rows_match_hf():
    g = np.load(os.path.join(GOLDEN, "mrope.npz"))
    cos_t, sin_t = rope_tables(256)
    axis = mrope_axis_map(tuple(int(v) for v in g["section"]), True)
    cos_ref, sin_ref = bf16_from_bits(g["cos_interleaved_bits"]), bf16_from_bits(g["sin_interleaved_bits"])
    for i, pos in enumerate(g["positions"]):
        c, s = mrope_rows(cos_t, sin_t, tuple(int(v) for v in pos), axis)
        assert torch.equal(c, cos_ref[i]) and torch.equal(s, sin_ref[i])


def test_engine_axis_bits_equal_the_oracle_map():
    """The C layer packs the same map as 2 bits per frequency (qmk_model_set_mrope); restated here from the header's rule."""
    for inter in (False, True):
        axis = mrope_axis_map((24, 20, 20), inter)
        words = [0, 0]
        for i in range(64):
            if inter:
                a = 1 if (i % 3 == 1 and i < 60) else 2 if (i % 3 == 2 and i < 60) else 0
            else:
                a = 0 if i < 24 else 1 if i < 44 else 2
            words[i >> 5] |= a << (2 * (i & 31))
        assert [(words[i >> 5] >> (2 * (i & 31))) & 3 for i in range(64)] == axis
