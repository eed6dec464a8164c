"""Property tests of the oracle's statistics: hypothesis-generated inputs, CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle

vals = st.floats(min_value=-50, max_value=50, allow_nan=False, allow_infinity=False, width=64)
vec = lambda lo, hi: st.lists(vals, min_size=lo, max_size=hi).map(lambda v: np.array(v, dtype=np.float64))


@settings(max_examples=150, deadline=None)
@given(vec(1, 40), vec(1, 40))
def test_ks_numerator_is_symmetric_and_bounded(a, b):
    rc1, d1, p1, e1 = oracle.ks(a, b)
    rc2, d2, p2, e2 = oracle.ks(b, a)
    assert rc1 == rc2 == 0 and d1 == d2 and e1 == e2
    assert 0 <= d1 <= a.size * b.size and 0.0 <= p1 <= 1.0
    assert abs(p1 - p2) <= 1e-12
    # invariance under a strictly increasing map of both samples
    rc3, d3, p3, e3 = oracle.ks(3.0 * a + 1.0, 3.0 * b + 1.0)
    if np.unique(np.concatenate([3.0 * a + 1.0, 3.0 * b + 1.0])).size == np.unique(np.concatenate([a, b])).size:
        assert d3 == d1


@settings(max_examples=150, deadline=None)
@given(vec(2, 40), vec(2, 40))
def test_welch_is_antisymmetric(a, b):
    rc1, t1, df1, p1 = oracle.welch(a, b)
    rc2, t2, df2, p2 = oracle.welch(b, a)
    assert rc1 == rc2
    # two all-zero groups give 0/0 = NaN in t.test itself (stderr 0 is not < 0); stderr^4 can underflow to df = NaN
    if rc1 == 0 and not (np.isnan(t1) or np.isnan(df1)):
        assert t1 == -t2 and df1 == df2 and p1 == p2 and 0.0 <= p1 <= 1.0


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(vals, vals), min_size=3, max_size=60))
def test_pearson_is_affine_invariant_and_bounded(xy):
    x = np.array([p[0] for p in xy]); y = np.array([p[1] for p in xy])
    rc, r, p = oracle.pearson(x, y)
    if rc == 0:
        assert -1.0 <= r <= 1.0 and 0.0 <= p <= 1.0
        rc2, r2, p2 = oracle.pearson(2.0 * x - 3.0, -0.5 * y + 1.0)
        if rc2 == 0 and np.ptp(x) > 1e-6 and np.ptp(y) > 1e-6:
            assert abs(r2 + r) < 1e-7


@settings(max_examples=100, deadline=None)
@given(vec(3, 120))
def test_order_is_a_stable_descending_permutation(x):
    perm = oracle.order_desc(x)
    assert sorted(perm.tolist()) == list(range(x.size))
    xs = x[perm]
    assert np.all(xs[:-1] >= xs[1:])
    ties = xs[:-1] == xs[1:]
    assert np.all(perm[:-1][ties] < perm[1:][ties])            # ties keep the original sample order
    m = x.size // 3
    groups = [set(perm[:m]), set(perm[m:2 * m]), set(perm[2 * m:3 * m])]
    assert sum(len(g) for g in groups) == 3 * m and not (groups[0] & groups[1]) and not (groups[1] & groups[2])


@settings(max_examples=100, deadline=None)
@given(vec(2, 60), vec(2, 60))
def test_guard_matches_sd_of_recycled_differences(a, b):
    L = max(a.size, b.size)
    d = np.array([a[i % a.size] - b[i % b.size] for i in range(L)])
    assert oracle.guard(a, b) == (0 if np.all(d == d[0]) else 1)


def test_sixty_comparator_network_of_the_rank_kernel_sorts_every_zero_one_input():
    """The 16 words a thread holds after the bucket placement are ordered by the 60-comparator network written out in
    epx_warpsort.cuh (sort16_thread). Zero-one principle: a comparator network that sorts all 2^16 zero-one inputs
    sorts everything. The comparators are read from the shipped header, not from a copy."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "epimethex_b200", "csrc",
                            "epx_warpsort.cuh")).read()
    body = src[src.index("static void sort16_thread"):src.index("#undef EPX_CE")]
    ces = [(int(a), int(b)) for a, b in re.findall(r"EPX_CE\((\d+), (\d+)\)", body)]
    assert len(ces) == 60 and all(0 <= a < b < 16 for a, b in ces)
    x = ((np.arange(1 << 16)[:, None] >> np.arange(16)[None, :]) & 1).astype(np.int8)
    for a, b in ces:
        lo, hi = np.minimum(x[:, a], x[:, b]), np.maximum(x[:, a], x[:, b])
        x[:, a], x[:, b] = lo, hi
    assert np.all(x[:, :-1] <= x[:, 1:])


def test_slot_number_in_the_long_row_word_names_the_slot_of_arrival():
    """epx_ctabucket.cuh keeps the low six bits of a value's slot of arrival in its word; after the short network a
    value sits fewer than 32 slots away, and p0 = k - 32 + ((lid - (k - 32)) mod 64) must give the slot back, also at the
    start of the row where k - 32 is negative (the kernel computes in unsigned arithmetic)."""
    for k in list(range(0, 70)) + [8999, 16383]:
        for delta in range(-31, 32):
            p0 = k + delta
            if p0 < 0:
                continue
            lid = p0 & 63
            got = (k - 32) + (((lid - ((k - 32) & 0xffffffff)) & 0xffffffff) & 63)
            assert got == p0, (k, delta, got)
