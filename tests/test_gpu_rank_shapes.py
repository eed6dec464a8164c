"""Row shapes chosen against the bucket placement of the rank kernels (epx_warpsort.cuh order_bucket, epx_ctabucket.cuh):
every shape must either be settled by the short network or be caught by its check and take the full network / the
merge sort -- the results may not depend on which. Compared with the oracle through the C ABI (medians, KS numerators,
tie handling and the sample order of the expression rows all depend on the order rows). Run with -m gpu on a B200."""
import numpy as np
import pytest

from epimethex_b200 import _native, synth
from parity import assert_parity, run_gpu, run_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sess():
    s = _native.Session(0)
    yield s
    s.close()


def _shape_rows(y, rng, t_mode=False):
    """Overwrite the rows of y, cycling through the shapes; returns the shape name of every row.
    t_mode: the Welch statistics of rows whose values agree to 1e-9 (shapes 3 and 12) are differences of means that
    agree to ten digits: R's long double accumulations (and the oracle's) decide their last digits and no fp64 sum
    follows them to 1e-8, so those two shapes are left to the KS runs, where only the order matters."""
    n = y.shape[1]
    names = []
    for p in range(y.shape[0]):
        k = p % 14
        if t_mode and k in (3, 12):
            k = 4
        u = rng.random(n)
        if k == 0:
            row, name = np.full(n, 0.25), "constant"
        elif k == 1:
            row, name = np.where(u < 0.5, 0.125, 0.875), "two values"
        elif k == 2:
            row, name = 0.2 + 0.6 * u, "uniform"
            row[rng.integers(n)] = 1.0e6                       # one far outlier: everything else shares a bucket
        elif k == 3:
            row, name = 0.9 + 1e-13 * rng.integers(0, 2000, n), "narrow range (low words decide)"
        elif k == 4:
            row, name = u - 0.5, "negative and positive"
        elif k == 5:
            row, name = np.round(u, 1), "eleven levels"       # tie runs of ~ n / 11 in the middle of the range
        elif k == 6:
            row, name = 10.0 ** (-300.0 * u), "300 decades"
        elif k == 7:
            row, name = np.clip(0.05 + 0.08 * rng.standard_normal(n), 0.002, 0.998), "pile at the minimum"
        elif k == 8:
            row, name = np.clip(0.95 + 0.08 * rng.standard_normal(n), 0.002, 0.998), "pile at the maximum"
        elif k == 9:
            row, name = np.where(u < 0.3, 0.0, np.where(u < 0.6, 1.0, u)), "piles at both ends"
        elif k == 10:
            row, name = np.sort(u), "ascending already"
        elif k == 11:
            row, name = np.sort(u)[::-1].copy(), "descending"
        elif k == 12:
            row, name = 0.5 + 1e-9 * rng.standard_normal(n), "values closer than the key resolution"
            row[0] = 0.0; row[1] = 1.0
        else:
            row, name = np.where(u < 0.9, 0.5, u), "pile in the middle"
        y[p] = row
        names.append(name)
    return names


@pytest.mark.parametrize("te,tm", [(True, False), (False, True)])
@pytest.mark.parametrize("n", [257, 300, 473, 512, 513, 514, 700, 1000, 1024, 1025, 1500, 4097])
def test_row_shapes(sess, n, te, tm):
    G, P = 24, 140
    rng = np.random.default_rng(1000 + n)
    ann = synth.annotation(P, G)
    x = synth.expression(G, n, ties=True)
    # expression rows: zero piles and constants too (descending order, R/EpimethexAnalysis.R:117-122)
    x[1] = np.where(rng.random(n) < 0.4, 0.0, x[1])
    x[2] = np.round(x[2], 0)
    x[3] = 7.0
    x[4] = -x[4]
    y = synth.methylation(P, n, ann, threads=2)
    _shape_rows(y, rng, t_mode=tm)
    idx = synth.probe_index(ann, 0, G)
    got = run_gpu(sess, x, y, idx, te=te, tm=tm)
    assert_parity(got, run_oracle(x, y, idx, te=te, tm=tm), f"shapes n={n} te={te} tm={tm}")


def test_row_shapes_at_9000_and_16384_samples(sess):
    for n in (9000, 16384):
        G, P = 6, 42
        rng = np.random.default_rng(n)
        ann = synth.annotation(P, G)
        x = synth.expression(G, n, ties=True)
        y = synth.methylation(P, n, ann, threads=2)
        _shape_rows(y, rng)
        idx = synth.probe_index(ann, 0, G)
        got = run_gpu(sess, x, y, idx)
        assert_parity(got, run_oracle(x, y, idx), f"shapes n={n}")


def test_full_network_counter_reports_the_rows_that_failed_the_check():
    """The counter of epx_timing is a performance figure only. Rows whose values all but two lie inside ONE bucket, in no
    order, cannot be settled by the short network and must show up in it (tie runs, however long, do not: equal values
    arrive in sample order, which is their final order)."""
    n, G, P = 473, 12, 60
    rng = np.random.default_rng(7)
    ann = synth.annotation(P, G)
    x = synth.expression(G, n, ties=False)
    y = 0.5 + 1e-4 * rng.random((P, n))
    y[:, 0] = 0.0
    y[:, 1] = 1.0
    idx = synth.probe_index(ann, 0, G)
    with _native.Session(0) as fresh:
        got = run_gpu(fresh, x, y, idx)
        ties = run_gpu(fresh, x, np.round(rng.random((P, n)), 1), idx)
    assert_parity(got, run_oracle(x, y, idx), "one bucket")
    assert got.timing["sort_full_rows"] > 0
    assert ties.timing["sort_full_rows"] == got.timing["sort_full_rows"]     # eleven levels: nothing added
