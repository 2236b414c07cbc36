"""Host-side restatement of the group kernel's L2 accumulator words (csrc/qmk_device2.cuh: acc_word / acc_done / acc_value)
and the properties the exchange protocol relies on.  Pure integer arithmetic, no GPU.

A word is {arrival count : 8 | signed fixed-point sum (2^-24) : 56}; a producer adds (1 << 56) + round(v * 2^24); totals are
cumulative (never zeroed), a consumer subtracts the total it read last time; a use is complete when the count advanced by 8."""

import random

MASK = (1 << 64) - 1
NGRP, SHIFT = 8, 24


def acc_word(v: float) -> int:
    return ((1 << 56) + int(round(v * (1 << SHIFT)))) & MASK


def acc_done(now: int, prev: int) -> bool:
    return ((((now - prev) & MASK) + (1 << 55)) & MASK) >> 56 == NGRP


def acc_value(now: int, prev: int) -> float:
    d = ((now - prev) - NGRP * (1 << 56)) & MASK
    if d >= 1 << 63:
        d -= 1 << 64
    return d / float(1 << SHIFT)


def test_sum_is_order_independent_and_exact_in_fixed_point():
    rng = random.Random(1)
    for _ in range(200):
        parts = [rng.uniform(-300.0, 300.0) * rng.choice([1.0, 1e-3, 1e-6]) for _ in range(NGRP)]
        prev = rng.getrandbits(64)
        totals = set()
        for _ in range(5):
            rng.shuffle(parts)
            now = prev
            for k, v in enumerate(parts):
                assert not acc_done(now, prev), "complete before the eighth arrival"
                now = (now + acc_word(v)) & MASK
            assert acc_done(now, prev)
            totals.add(now)
        assert len(totals) == 1
        want = sum(int(round(v * (1 << SHIFT))) for v in parts) / float(1 << SHIFT)
        assert acc_value(totals.pop(), prev) == want


def test_counts_and_sums_wrap_without_reset():
    """thousands of uses of one word: the 8-bit count wraps every 32 uses, the 56-bit sum wraps too; differences stay exact"""
    rng = random.Random(2)
    total = 0
    for use in range(5000):
        prev = total
        parts = [rng.uniform(-5e4, 5e4) for _ in range(NGRP)]
        for i, v in enumerate(parts):
            total = (total + acc_word(v)) & MASK
            assert acc_done(total, prev) == (i == NGRP - 1)
        want = sum(int(round(v * (1 << SHIFT))) for v in parts) / float(1 << SHIFT)
        assert acc_value(total, prev) == want


def test_negative_partial_sums_do_not_disturb_the_count():
    prev = 0x0123456789ABCDEF
    now = prev
    for i in range(NGRP):
        now = (now + acc_word(-1e-7 * (i + 1))) & MASK      # borrows ripple into the count byte of the raw word
        assert acc_done(now, prev) == (i == NGRP - 1)
    assert acc_value(now, prev) < 0
