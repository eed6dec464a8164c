"""Shared checker: libepx (CUDA, through the C ABI) against the CPU oracle on the same inputs.
Bar (BASELINE.json north_star): group assignments, ranks, probe ordering and KS D numerators bit-exact;
correlations, fold changes, t statistics and p-values within 1e-10 absolute or 1e-8 relative."""
import numpy as np

import oracle
from epimethex_b200 import _native

RTOL, ATOL = 1e-8, 1e-10


def run_gpu(sess, expr, meth, idx, *, te=True, tm=False, linear=True, fc_means=False, upload=True):
    if upload:
        sess.set_methylation(meth)
    opts = _native.make_options(linear, te, tm, fc_from_means=fc_means)
    return sess.analyze(expr, idx.row_ptr, idx.probe_idx, opts)


def run_oracle(expr, meth, idx, *, te=True, tm=False, linear=True, fc_means=False, threads=8):
    return oracle.analyze(expr, meth, idx.row_ptr, idx.probe_idx, data_linear=linear, test_expr_t=te,
                          test_meth_t=tm, fc_from_means=fc_means, threads=threads)


def assert_parity(got, ref, label=""):
    assert got.counts == ref.counts, f"{label}: expression-side group sizes {got.counts} != {ref.counts}"
    assert got.ref_gene == ref.ref_gene, f"{label}: reference gene"
    assert np.array_equal(got.perm.astype(np.int32), ref.perm), f"{label}: sample order (ranks) differ"
    assert np.array_equal(got.ks_dnum, ref.ks_dnum), f"{label}: pair KS numerators differ"
    assert np.array_equal(got.gene_ks_dnum, ref.gene_ks_dnum), f"{label}: gene KS numerators differ"
    assert np.array_equal(got.pair_na, ref.pair_na), f"{label}: pair NA masks differ"
    assert np.array_equal(got.gene_na, ref.gene_na), f"{label}: gene NA masks differ"
    for name, a, b in (("pair", got.pair, ref.pair), ("pair_stat", got.pair_stat, ref.pair_stat),
                       ("gene", got.gene, ref.gene)):
        ok = np.isclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True)
        if not ok.all():
            bad = np.argwhere(~ok)[:8]
            detail = [(tuple(i), a[tuple(i)], b[tuple(i)]) for i in bad]
            raise AssertionError(f"{label}: {name} mismatch ({(~ok).sum()} cells): {detail}")
