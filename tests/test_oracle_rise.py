"""The builder-defined RISE mask generator (csrc/common.h: rise_mask_key / rise_keep, restated in oracle/loops.py): Bernoulli
statistics, independence across masks and seeds, and known-answer bits that pin the hash constants."""
import numpy as np

from oracle import loops


def test_keep_probability_and_independence():
    for p in (0.1, 0.5, 0.75):
        m = loops.rise_keep_mask(0, 0, 1025, 400, p)
        assert m.shape == (1025, 400) and abs(m.mean() - p) < 4 * np.sqrt(p * (1 - p) / m.size) + 1e-4
    a = loops.rise_keep_mask(0, 1, 1025, 400, 0.5)
    b = loops.rise_keep_mask(0, 2, 1025, 400, 0.5)
    c = loops.rise_keep_mask(1, 1, 1025, 400, 0.5)
    for x, y in ((a, b), (a, c)):
        assert abs((x == y).mean() - 0.5) < 0.005                      # different mask index / seed: uncorrelated bits
    assert abs((a[:, 1:] == a[:, :-1]).mean() - 0.5) < 0.005            # neighbouring frames uncorrelated
    assert abs((a[1:, :] == a[:-1, :]).mean() - 0.5) < 0.005            # neighbouring bins uncorrelated
    assert loops.rise_keep_mask(0, 0, 1025, 10, 1.0).all() and not loops.rise_keep_mask(0, 0, 1025, 10, 0.0).any()


def test_known_answer_bits():
    m = loops.rise_keep_mask(0, 3, 1025, 50, 0.5)
    # pinned from the C expression of the hash evaluated independently (see the self-check in the commit that added RISE)
    def lb(x):
        x &= 0xFFFFFFFF; x ^= x >> 16; x = (x * 0x7FEB352D) & 0xFFFFFFFF; x ^= x >> 15; x = (x * 0x846CA68B) & 0xFFFFFFFF; x ^= x >> 16
        return x
    key = lb((0 * 0x9E3779B9 + 3 * 0x85EBCA6B + 0x165667B1) & 0xFFFFFFFF)
    for f, t in ((0, 0), (17, 3), (512, 25), (1024, 49)):
        u = lb(key ^ (((t * 1025 + f) * 0xC2B2AE35) & 0xFFFFFFFF))
        assert bool(m[f, t]) == (u < 2 ** 31)
