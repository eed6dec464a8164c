"""GPU parity of the pooled probe-group analyses (CG_of_a_genes / CG_by_position / CG_by_islands rows,
R/EpimethexAnalysis.R:476-620) against the oracle, through the C ABI. Run with -m gpu on a B200."""
import numpy as np
import pytest

import oracle
from epimethex_b200 import _native, synth
from epimethex_b200.annotation import build_groups
from parity import ATOL, RTOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sess():
    s = _native.Session(0)
    yield s
    s.close()


def _synth(n_genes, n, n_probes, *, decimals=None, missing_every=0, ties=True):
    ann = synth.annotation(n_probes, n_genes)
    x = synth.expression(n_genes, n, ties=ties)
    y = synth.methylation(n_probes, n, ann, decimals=decimals, threads=4)
    has = None
    if missing_every:
        has = np.ones(n_probes, bool); has[::missing_every] = False
    return x, y, synth.probe_index(ann, 0, n_genes, probe_has_data=has)


def assert_group_parity(got, ref, label=""):
    assert np.array_equal(got.group_na, ref.group_na), f"{label}: NA masks differ at {np.nonzero(got.group_na != ref.group_na)[0][:8]}"
    assert np.array_equal(got.group_dnum, ref.group_dnum), \
        f"{label}: KS numerators differ at {np.argwhere(got.group_dnum != ref.group_dnum)[:8].tolist()}"
    # medians and beta differences are order statistics of the inputs: bit-exact
    assert np.array_equal(got.group[:, :6], ref.group[:, :6], equal_nan=True), f"{label}: medians differ"
    for name, a, b in (("group", got.group, ref.group), ("group_stat", got.group_stat, ref.group_stat)):
        ok = np.isclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True)
        if not ok.all():
            bad = np.argwhere(~ok)[:8]
            raise AssertionError(f"{label}: {name} mismatch ({(~ok).sum()} cells): "
                                 f"{[(tuple(i), a[tuple(i)], b[tuple(i)]) for i in bad]}")


def _both(sess, x, y, idx, gi, tm):
    sess.set_methylation(y)
    sess.analyze(x, idx.row_ptr, idx.probe_idx, _native.make_options(True, True, tm), full=False)
    got = sess.analyze_groups(gi.group_ptr, gi.group_pairs)
    ref = oracle.analyze_groups(x, y, idx.row_ptr, idx.probe_idx, gi.group_ptr, gi.group_pairs, test_meth_t=tm, threads=8)
    return got, ref


@pytest.mark.parametrize("tm", [False, True])
@pytest.mark.parametrize("n", [6, 30, 99, 473, 474, 1000])
@pytest.mark.parametrize("by", ["gene", "position", "island"])
def test_groups_against_oracle(sess, by, n, tm):
    G, P = (150, 3000) if n <= 99 else (60, 2500)
    x, y, idx = _synth(G, n, P, missing_every=9)
    gi = build_groups(idx, by)
    assert gi.n_groups > 0
    got, ref = _both(sess, x, y, idx, gi, tm)
    assert_group_parity(got, ref, f"by={by} n={n} tm={tm}")


def test_single_probe_groups_equal_the_pair_table(sess):
    """A group of one probe is Loop A's column for that pair (R/EpimethexAnalysis.R:496-503 keeps it unstacked)."""
    x, y, idx = _synth(80, 474, 2000)
    sess.set_methylation(y)
    res = sess.analyze(x, idx.row_ptr, idx.probe_idx, _native.make_options(True, True, False))
    sel = np.nonzero(idx.probe_idx >= 0)[0].astype(np.int32)
    got = sess.analyze_groups(np.arange(sel.size + 1, dtype=np.int64), sel)
    assert np.array_equal(got.group_dnum, res.ks_dnum[sel].astype(np.int64))
    assert np.array_equal(got.group_na, res.pair_na[sel])
    assert np.array_equal(got.group[:, :6], res.pair[sel, :6])
    np.testing.assert_allclose(got.group, res.pair[sel], rtol=RTOL, atol=ATOL)


def test_large_gene_needs_global_merge_passes(sess):
    """One gene with 300 probes of 474 samples: 18 sorted blocks, 5 merge passes, 142,200 stacked values."""
    n, k = 474, 300
    rng = np.random.default_rng(5)
    x = rng.normal(6, 2, (3, n))
    y = np.round(rng.random((k + 40, n)), 3)      # 3 decimals: plenty of ties across probes
    row_ptr = np.array([0, k, k + 25, k + 40], np.int64)
    probe_idx = np.arange(k + 40, dtype=np.int32)
    sess.set_methylation(y)
    for tm in (False, True):
        sess.analyze(x, row_ptr, probe_idx, _native.make_options(True, True, tm), full=False)
        gp, gl = row_ptr, np.arange(k + 40, dtype=np.int32)
        got = sess.analyze_groups(gp, gl)
        ref = oracle.analyze_groups(x, y, row_ptr, probe_idx, gp, gl, test_meth_t=tm, threads=4)
        assert_group_parity(got, ref, f"large tm={tm}")


def test_long_rows(sess):
    """n > 1024 (CTA-level sort path of the main kernels) and n > 8192 (one probe per sorted block)."""
    for n, k in ((1500, 12), (9000, 5)):
        rng = np.random.default_rng(n)
        x = rng.normal(6, 2, (2, n))
        y = rng.random((2 * k, n))
        row_ptr = np.array([0, k, 2 * k], np.int64)
        probe_idx = np.arange(2 * k, dtype=np.int32)
        sess.set_methylation(y)
        sess.analyze(x, row_ptr, probe_idx, _native.make_options(True, True, False), full=False)
        got = sess.analyze_groups(row_ptr, probe_idx)
        ref = oracle.analyze_groups(x, y, row_ptr, probe_idx, row_ptr, probe_idx, threads=2)
        assert_group_parity(got, ref, f"n={n}")


def test_degenerate_groups(sess):
    """empty group -> NA row; constant probes -> guard NA and Pearson NA; identical probes -> ties everywhere."""
    n = 30
    rng = np.random.default_rng(2)
    x = rng.normal(5, 1, (2, n))
    y = rng.random((6, n))
    y[1] = 0.3                       # constant probe
    y[3] = y[2]                      # duplicate probe
    row_ptr = np.array([0, 3, 6], np.int64)
    probe_idx = np.array([0, 1, -1, 2, 3, 1], np.int32)
    gp = np.array([0, 0, 2, 3, 5, 6], np.int64)
    gl = np.array([0, 1, 1, 3, 4, 5], np.int32)   # {}, {0,1}, {1}, {3,4}, {5}
    sess.set_methylation(y)
    for tm in (False, True):
        sess.analyze(x, row_ptr, probe_idx, _native.make_options(True, True, tm), full=False)
        got = sess.analyze_groups(gp, gl)
        ref = oracle.analyze_groups(x, y, row_ptr, probe_idx, gp, gl, test_meth_t=tm)
        assert_group_parity(got, ref, f"degenerate tm={tm}")
        assert got.group_na[0] == (1 << 11) - 1 and np.isnan(got.group[0]).all()


def test_bad_group_lists_are_rejected(sess):
    x, y, idx = _synth(20, 30, 400, missing_every=5)
    sess.set_methylation(y)
    sess.analyze(x, idx.row_ptr, idx.probe_idx, _native.make_options(), full=False)
    nodata = int(np.nonzero(idx.probe_idx < 0)[0][0])
    with pytest.raises(_native.EpxError) as e:
        sess.analyze_groups(np.array([0, 1], np.int64), np.array([nodata], np.int32))
    assert e.value.code == -1
    g0 = int(idx.row_ptr[1]) - 1
    with pytest.raises(_native.EpxError):
        sess.analyze_groups(np.array([0, 2], np.int64), np.array([g0, int(idx.row_ptr[-1]) - 1], np.int32))


def test_public_api_writes_the_pooled_tables(tmp_path):
    """epimethex.analysis writes four files (R/EpimethexAnalysis.R:470, :545, :577, :615)."""
    from epimethex_b200 import epimethex_analysis
    dfe, dfa, dfm, _ = synth.frames(40, 30, 600)
    out = epimethex_analysis(dfe, dfa, dfm, 1, 40, 2, True, True, False, output_dir=str(tmp_path), return_all=True)
    d = tmp_path / "1-40"
    assert set(out) >= {"individually", "genes", "position"}
    assert (d / "CG_data_individually.csv").exists() and (d / "CG_of_a_genes.csv").exists() and (d / "CG_by_position.csv").exists()
    genes = out["genes"]
    assert genes.shape[1] == 17 and all(name.startswith("CG~G") for name in genes.index)
    head = (d / "CG_of_a_genes.csv").read_text().splitlines()[0]
    assert head.startswith('"","medianDown","medianMedium","medianUP"') and head.count(",") == 17
    pos = out["position"]
    assert all(name.split("_")[0] in synth.GROUPS for name in pos.index)
    # a gene pooled over all its probes with a single data-bearing probe repeats that probe's row
    ind = out["individually"]
    single = [g for g in genes.index if sum(i.endswith("_" + g[3:]) for i in ind.dropna(subset=["medianUP"]).index) == 1]
    for g in single[:5]:
        row = ind.dropna(subset=["medianUP"]).filter(like="_" + g[3:], axis=0).iloc[0]
        np.testing.assert_allclose(genes.loc[g].to_numpy(float), row.to_numpy()[:17].astype(float), rtol=1e-8, atol=1e-10)
