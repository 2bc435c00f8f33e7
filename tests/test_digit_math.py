"""The arithmetic behind the digit-sliced mixture kernels (multiclust_b200/csrc/mc_digit.cuh),
restated with Python integers: what the IMMA accumulators hold, how the eight 8-bit digits
recombine, and the error bound DESIGN.md 4.3 states.  No GPU: this pins the claim, the GPU
tests pin the kernels."""
from fractions import Fraction

import numpy as np


def _digits(x):
    return [(x >> (8 * d)) & 0xff for d in range(8)]


def test_e_pass_fixed_point_is_exact_above_a_quarter():
    """|log p| >= 0.25 has an exact image in 2^-54 fixed point, so sum_l c * log p comes out
    as the correctly rounded sum; smaller values are rounded at 2^-55 absolute"""
    rng = np.random.default_rng(1)
    p = np.concatenate([rng.random(400) * 0.75 + 1e-9, [1e-8, 1e-300, 0.77, 0.7788007830714049]])
    lp = -np.log(p)
    c = rng.integers(0, 5, size=p.size)
    X = [int(round(float(v) * 2.0 ** 54)) if v * 2.0 ** 54 < 2 ** 63
         else int(Fraction(float(v)) * 2 ** 54) for v in lp]
    for v, x in zip(lp, X):
        if v >= 0.25:
            assert Fraction(x, 2 ** 54) == Fraction(float(v))       # exact image
        else:
            assert abs(Fraction(x, 2 ** 54) - Fraction(float(v))) <= Fraction(1, 2 ** 55)
        assert 0 <= x < 2 ** 64
    # what the tensor path accumulates: one 32-bit integer per digit
    acc = [sum(int(ci) * _digits(x)[d] for ci, x in zip(c, X)) for d in range(8)]
    assert max(acc) < 2 ** 31
    total = sum(a << (8 * d) for d, a in enumerate(acc))
    assert total == sum(int(ci) * x for ci, x in zip(c, X))         # the integer sum is exact
    # recombination as the kernel does it: pairs of digits exactly, then three FP64 adds
    pairs = [float(acc[2 * t] + 256 * acc[2 * t + 1]) * 2.0 ** (16 * t) for t in range(4)]
    got = -((pairs[0] + pairs[1]) + (pairs[2] + pairs[3])) * 2.0 ** -54
    want = -sum(Fraction(int(ci)) * Fraction(float(v)) for ci, v in zip(c, lp))
    small = sum(int(ci) for ci, v in zip(c, lp) if v < 0.25)
    assert abs(Fraction(got) - want) <= 3 * Fraction(np.spacing(abs(float(want)))) \
        + small * Fraction(1, 2 ** 55)


def test_m_pass_column_scaling_keeps_relative_precision():
    """v 2^(64 - e_k) with max v < 2^e_k: a nearly empty class loses nothing against its
    own largest posterior (the absolute 2^-63 grid would: DESIGN.md 4.3)"""
    rng = np.random.default_rng(2)
    for scale in (1.0, 1e-9, 1e-200):
        v = rng.random(300) * scale
        c = rng.integers(0, 3, size=v.size)
        m, e = np.frexp(v.max())
        assert v.max() < 2.0 ** int(e)
        X = [int(Fraction(float(x)) * Fraction(2) ** (64 - int(e)) + Fraction(1, 2)) for x in v]
        assert max(X) < 2 ** 64
        acc = [sum(int(ci) * _digits(x)[d] for ci, x in zip(c, X)) for d in range(8)]
        assert max(acc) < 2 ** 31
        total = sum(a << (8 * d) for d, a in enumerate(acc))
        got = Fraction(total) / Fraction(2) ** (64 - int(e))
        want = sum(Fraction(int(ci)) * Fraction(float(x)) for ci, x in zip(c, v))
        # every term is rounded at 2^(e - 65): relative to the sum ~ I 2^-64
        assert abs(got - want) <= int(c.sum()) * Fraction(2) ** (int(e) - 65)
        assert abs(got - want) <= Fraction(1, 10 ** 15) * want


def test_accumulator_bound_of_the_planner():
    """a chunk is at most floor((2^31 - 1) / (255 P)) elements long: the largest digit times
    the largest count per element cannot overflow a signed 32-bit accumulator"""
    for P in (1, 2, 4, 15):
        n = (2 ** 31 - 1) // (255 * P)
        assert n * 255 * P <= 2 ** 31 - 1 < (n + 1) * 255 * P
