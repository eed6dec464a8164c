"""The oracle (oracle/epx_oracle.c) against the committed golden fixture and against scipy/mpmath as an
independent second opinion on every statistic. CPU only."""
import json
import os

import mpmath as mp
import numpy as np
import pytest
from scipy import stats

import oracle
from epimethex_b200 import analysis

HERE = os.path.dirname(os.path.abspath(__file__))


def test_readme_fixture_matches_golden(readme_frames):
    gold = json.load(open(os.path.join(HERE, "golden", "readme_fixture.json")))
    dfe, dfa, dfm = readme_frames
    prep = analysis.prepare(dfe, dfa, dfm, 1, 3)
    r = oracle.analyze(prep.expr, prep.meth, prep.index.row_ptr, prep.index.probe_idx,
                       data_linear=True, test_expr_t=True, test_meth_t=False)
    assert r.counts == (2, 2, 2)
    assert (r.pair_na == 0).all() and (r.gene_na == 0).all()
    samples = list(dfe.columns[1:])
    for j in range(prep.index.n_pairs):
        g = prep.index.pair_gene[j]
        name = f"{prep.index.cg[j]}_{prep.index.position[j]}_{prep.gene_names[g]}"
        want = np.array(gold["rows"][name])
        got = np.concatenate([r.pair[j], r.gene[g, :6]])
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14, err_msg=name)
        assert [samples[i] for i in r.perm[g]] == gold["sample_order"][name]
        assert prep.index.island[j] == gold["islands"][name]
    # exact KS for m = n = 2: D numerators in {2, 4}
    assert set(np.unique(r.ks_dnum)) <= {2, 4}


def test_fc_quirk_vs_means(readme_frames):
    """the reference's gene fold changes come from q99/q66/q33, not from the group means."""
    dfe, dfa, dfm = readme_frames
    prep = analysis.prepare(dfe, dfa, dfm, 1, 3)
    q = oracle.analyze(prep.expr, prep.meth, prep.index.row_ptr, prep.index.probe_idx)
    m = oracle.analyze(prep.expr, prep.meth, prep.index.row_ptr, prep.index.probe_idx, fc_from_means=True)
    np.testing.assert_allclose(q.gene[0, :3], [0.836435, 1.0842450196547, 1.84578972236385], rtol=1e-12)
    np.testing.assert_allclose(m.gene[0, :3], [0.6219, 2.08485, 4.68740], rtol=2e-5)


@pytest.mark.parametrize("n", [2, 3, 10, 157, 158, 333])
def test_welch_matches_scipy(n):
    rng = np.random.default_rng(n)
    a, b = rng.normal(0.5, 0.1, n), rng.normal(0.52, 0.3, n + (n % 2))
    rc, t, df, p = oracle.welch(a, b)
    s = stats.ttest_ind(a, b, equal_var=False)
    assert rc == 0
    np.testing.assert_allclose([t, p], [s.statistic, s.pvalue], rtol=1e-11)
    np.testing.assert_allclose(df, s.df, rtol=1e-12)


def test_welch_errors_are_reported():
    assert oracle.welch([1.0], [1.0, 2.0])[0] == 1          # not enough observations
    assert oracle.welch([3.0, 3.0, 3.0], [3.0, 3.0])[0] == 2  # data are essentially constant


def test_pt_matches_mpmath():
    mp.mp.dps = 40
    for df in (1, 2.5, 10, 157.3, 471, 2999.5, 8998):
        for t in (1e-6, 0.3, 1, 2.5, 6, 15, 40):
            x = mp.mpf(df) / (mp.mpf(df) + mp.mpf(t) ** 2)
            want = mp.betainc(mp.mpf(df) / 2, mp.mpf(1) / 2, 0, x, regularized=True) / 2
            got = oracle.pt(-t, df)
            assert abs(got - want) <= 1e-13 * want + 1e-300, (t, df)
            assert abs(oracle.pt(t, df) - (1 - want)) < 1e-15


@pytest.mark.parametrize("m,n", [(2, 2), (5, 5), (20, 20), (99, 99), (33, 34), (7, 50)])
def test_ks_exact_matches_scipy(m, n):
    rng = np.random.default_rng(m * 100 + n)
    a, b = rng.normal(0, 1, m), rng.normal(0.4, 1, n)
    rc, dnum, p, exact = oracle.ks(a, b)
    s = stats.ks_2samp(a, b, method="exact")
    assert rc == 0 and exact
    assert dnum == round(s.statistic * m * n)
    np.testing.assert_allclose(p, s.pvalue, rtol=1e-10, atol=1e-15)


def test_ks_asymptotic_is_r_pkstwo_not_the_true_limit():
    """for x < 1 R sums only k = 1 of the alternative series; replicated, not 'fixed'."""
    for x in (0.3, 0.7, 0.9, 0.99):
        want = np.sqrt(2 * np.pi) / x * np.exp(-np.pi ** 2 / (8 * x * x))
        np.testing.assert_allclose(oracle.pkstwo(x), want, rtol=1e-14)
    assert abs((1 - oracle.pkstwo(0.99)) - stats.kstwobign.sf(0.99)) > 1e-5   # R differs from the true limit here
    for x in (1.0, 1.3, 2.0, 3.5):
        np.testing.assert_allclose(1 - oracle.pkstwo(x), stats.kstwobign.sf(x), rtol=1e-6, atol=1e-12)
    # large groups -> asymptotic; D numerator still exact
    rng = np.random.default_rng(3)
    a, b = rng.normal(0, 1, 158), rng.normal(0.2, 1, 158)
    rc, dnum, p, exact = oracle.ks(a, b)
    s = stats.ks_2samp(a, b, method="asymp")
    assert not exact and dnum == round(s.statistic * 158 * 158)
    en = np.sqrt(158 * 158 / 316.0)
    np.testing.assert_allclose(p, min(1.0, max(0.0, 1 - oracle.pkstwo(en * s.statistic))), rtol=1e-13)


def test_ks_ties_force_asymptotic_and_use_run_ends():
    a = np.array([0.1, 0.2, 0.2, 0.5])
    b = np.array([0.2, 0.3, 0.5, 0.9])
    rc, dnum, p, exact = oracle.ks(a, b)
    assert not exact
    # ECDF difference at the distinct values only: after 0.1: 1/4; after 0.2: 3/4-1/4; after 0.3: 3/4-2/4 ...
    assert dnum == 8  # max |cA*4 - cB*4| = |3*4 - 1*4|
    np.testing.assert_allclose(dnum / 16, stats.ks_2samp(a, b).statistic)


def test_guard_semantics():
    assert oracle.guard([1.0, 2.0, 3.0], [0.5, 1.5, 2.5]) == 0     # constant difference -> sd == 0 -> NA
    assert oracle.guard([1.0, 2.0, 3.0], [0.5, 1.5, 2.6]) == 1
    assert oracle.guard([1.0], [2.0]) == -1                       # sd of one value is NA
    assert oracle.guard([1.0, 2.0, 3.0], [1.0, 2.0]) == 1          # mapply recycles: 0, 0, 2
    assert oracle.guard([1.0, 1.0, 1.0], [1.0, 1.0]) == 0


def test_order_is_stable_descending_with_signed_zero_ties():
    x = np.array([0.0, 3.0, -0.0, 3.0, 1.0, 0.0])
    assert oracle.order_desc(x).tolist() == [1, 3, 4, 0, 2, 5]


@pytest.mark.parametrize("n", [6, 7, 100, 473])
def test_quantile7_and_median(n):
    rng = np.random.default_rng(n)
    x = np.sort(rng.normal(size=n))
    for p in (0.0, 0.33, 1 / 3, 0.66, 2 / 3, 0.99, 1.0):
        np.testing.assert_allclose(oracle.quantile7(x, p), np.quantile(x, p), rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(oracle.median(rng.permutation(x)), np.median(x), rtol=1e-15)


def test_binvar_counts():
    """SURVEY 8a-4 [calc]: sizes of the expression-side thirds when the reference gene has no ties."""
    for n, want in ((6, (2, 2, 2)), (471, (157, 157, 157)), (474, (158, 158, 158)), (999, (333, 333, 333)),
                    (9000, (3000, 3000, 3000)), (473, (158, 157, 158)), (1000, (333, 333, 334))):
        x = np.sort(np.random.default_rng(n).normal(size=n))[::-1]
        assert oracle.binvar_counts(x) == want, n
    with pytest.raises(ValueError):
        oracle.binvar_counts(np.array([1.0, 0, 0, 0, 0, 0, 0]))


def test_fold_changes():
    assert oracle.linear_fc(4.0, 2.0) == 2.0 and oracle.linear_fc(2.0, 4.0) == -2.0
    assert oracle.linear_fc(1.5, -0.5) == 2.0 and oracle.linear_fc(-0.5, 1.5) == -2.0
    assert oracle.linear_fc(0.0, 2.0) == -2.0          # sign(0) = 0 -> "different sign" -> max - min
    assert oracle.linear_fc(2.0, 2.0) == -1.0          # equality takes the -fc branch
    assert oracle.log_fc(3.0, 1.0) == 4.0 and oracle.log_fc(1.0, 3.0) == -4.0 and oracle.log_fc(1.0, 1.0) == -1.0


def test_pearson_matches_scipy_and_na():
    rng = np.random.default_rng(9)
    x = rng.normal(size=473); y = 0.3 * x + rng.normal(size=473)
    rc, r, p = oracle.pearson(x, y)
    s = stats.pearsonr(x, y)
    np.testing.assert_allclose([r, p], [s.statistic, s.pvalue], rtol=1e-10)
    assert oracle.pearson(x, np.full(473, 0.25))[0] == 1  # zero sd -> NA


def _naive_pair(expr_row, meth_row, test_meth_t):
    """independent numpy/scipy restatement of one pair (Loop A) for cross-checking epo_analyze"""
    n = expr_row.size; m = n // 3
    perm = np.argsort(-expr_row, kind="stable")
    y = meth_row[perm]
    up, mid, down = y[:m], y[m:2 * m], y[2 * m:3 * m]
    out = [np.median(down), np.median(mid), np.median(up)]
    out += [out[2] - out[1], out[2] - out[0], out[1] - out[0]]
    for a, b in ((up, mid), (up, down), (mid, down)):
        if np.std(a - b) == 0:
            out.append(np.nan)
        elif test_meth_t:
            out.append(stats.ttest_ind(a, b, equal_var=False).pvalue)
        else:
            ties = np.unique(np.concatenate([a, b])).size < 2 * m
            if m * m < 10000 and not ties:
                out.append(stats.ks_2samp(a, b, method="exact").pvalue)
            else:
                d = stats.ks_2samp(a, b, method="asymp").statistic
                out.append(min(1.0, max(0.0, 1 - oracle.pkstwo(np.sqrt(m / 2) * d))))
    r, p = stats.pearsonr(expr_row[perm], y)
    return out + [r, p]


@pytest.mark.parametrize("n,tm", [(30, False), (30, True), (31, False), (474, False), (473, True)])
def test_analyze_pairs_against_naive(n, tm):
    from epimethex_b200 import synth
    ann = synth.annotation(120, 12)
    x = synth.expression(12, n)
    y = synth.methylation(120, n, ann, decimals=3 if n == 31 else None, threads=1)
    idx = synth.probe_index(ann, 0, 12)
    r = oracle.analyze(x, y, idx.row_ptr, idx.probe_idx, test_meth_t=tm)
    for j in range(0, idx.n_pairs, 3):
        want = _naive_pair(x[idx.pair_gene[j]], y[idx.probe_idx[j]], tm)
        np.testing.assert_allclose(r.pair[j], want, rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=str(j))


def test_missing_probe_rows_are_all_na():
    from epimethex_b200 import synth
    ann = synth.annotation(80, 6)
    x = synth.expression(6, 12); y = synth.methylation(80, 12, ann, threads=1)
    has = np.ones(80, bool); has[::3] = False
    idx = synth.probe_index(ann, 0, 6, probe_has_data=has)
    r = oracle.analyze(x, y, idx.row_ptr, idx.probe_idx)
    miss = idx.probe_idx < 0
    assert miss.any() and (r.pair_na[miss] == 0x7FF).all() and np.isnan(r.pair[miss]).all()
    assert (r.pair_na[~miss] & 0x600 == 0).all()
