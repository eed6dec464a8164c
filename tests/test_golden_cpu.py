"""The oracle (oracle/epx_oracle.c) against the golden tables written by the INDEPENDENT second pipeline
(tests/golden/rpipeline.py: numpy / big integers / mpmath, from the R sources, no shared code): exact and asymptotic
KS, ties, the n mod 3 extension, both tests on both sides, fold-change variants, pooled CG_of_a_genes rows.  CPU only.
This is the pin of the oracle: neither side is R output (parity unpinned), but a misreading of R now has to be made
twice to pass."""
import numpy as np
import pytest

import goldens
import oracle


@pytest.mark.parametrize("name", goldens.CASES)
def test_oracle_matches_the_independent_pipeline(name):
    g, x, y, row_ptr, probe_idx = goldens.load(name)
    for tm in (False, True):
        r = oracle.analyze(x, y, row_ptr, probe_idx, test_expr_t=True, test_meth_t=tm, threads=4)
        goldens.check_pairs(g, r, tm, "oracle")
    r = oracle.analyze(x, y, row_ptr, probe_idx, test_expr_t=False, test_meth_t=False, threads=4)
    goldens.check_genes(g, r, False, "oracle")
    r = oracle.analyze(x, y, row_ptr, probe_idx, fc_from_means=True, threads=4)
    goldens._close(r.gene[:, :3], g["gene_fc_means_linear"], "fold change of the group means", **goldens.TOL["oracle"])
    r = oracle.analyze(x, y, row_ptr, probe_idx, data_linear=False, threads=4)
    goldens._close(r.gene[:, :3], g["gene_fc_quantiles_log"], "log fold change of the quantile rows", **goldens.TOL["oracle"])


@pytest.mark.parametrize("name", goldens.CASES)
def test_oracle_pooled_rows_match_the_independent_pipeline(name):
    g, x, y, row_ptr, probe_idx = goldens.load(name)
    gp, gl = goldens.gene_groups(row_ptr, probe_idx)
    for tm in (False, True):
        r = oracle.analyze_groups(x, y, row_ptr, probe_idx, gp, gl, test_meth_t=tm, threads=4)
        goldens.check_groups(g, r, tm, "oracle")


def test_goldens_cover_what_they_claim():
    """exact KS p-values present at n = 99 without ties, absent with ties; asymptotic-only at n = 474; unlabelled
    samples at n = 473; NA rows (probes without data) and guard hits somewhere"""
    g99, *_ = goldens.load("n99_exact")
    g474, *_ = goldens.load("n474_asymptotic")
    g473, *_ = goldens.load("n473_extension")
    gt, *_ = goldens.load("n99_ties3")
    assert g99["params"]["n"] // 3 == 33 and (g99["pair_dnum_ks"] >= 0).any()
    assert (g474["pair_dnum_ks"].max() <= 158 * 158) and g473["params"]["n"] % 3 == 2
    assert tuple(g473["counts"]) != tuple(g474["counts"])
    for g in (g99, g474, g473, gt):
        t = g["pair_ks"]
        assert np.isnan(t).all(axis=1).any(), "no all-NA row (annotated probe without data)"
        assert (np.isnan(t[:, 6:9]).any(axis=1) & ~np.isnan(t[:, 0])).any(), "no guard hit"
        assert (np.isnan(t[:, 6:9]).sum(axis=1) == 1).any(), "no pair with exactly one guarded comparison"
        assert np.isnan(g["gene_t"][-1, 3:6]).all(), "the constant gene's tests are not NA"
    # ties change the p-value for the same numerator (exact -> asymptotic): the two n = 99 tables must differ
    assert not np.allclose(np.nan_to_num(g99["pair_ks"][:, 6:9]), np.nan_to_num(gt["pair_ks"][:, 6:9]))
