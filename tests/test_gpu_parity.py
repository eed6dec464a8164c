"""GPU parity tests proper: the CUDA path through the C ABI against the oracle. Run with -m gpu on a B200."""
import json
import os

import numpy as np
import pytest

from epimethex_b200 import _native, analysis, synth
from parity import assert_parity, run_gpu, run_oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sess():
    s = _native.Session(0)
    yield s
    s.close()


def _synth(n_genes, n, n_probes, *, decimals=None, ties=True, missing_every=0):
    ann = synth.annotation(n_probes, n_genes)
    x = synth.expression(n_genes, n, ties=ties)
    y = synth.methylation(n_probes, n, ann, decimals=decimals, threads=4)
    has = None
    if missing_every:
        has = np.ones(n_probes, bool); has[::missing_every] = False
    return x, y, synth.probe_index(ann, 0, n_genes, probe_has_data=has)


def test_device_is_b200():
    info = _native.device_info(0)
    assert "B200" in info["name"] and info["sm_count"] == 148, info


def test_readme_fixture_against_golden_and_oracle(sess, readme_frames):
    """BASELINE config 1: epimethex.analysis(Expr, Annot, Meth, 1, 3, 2, TRUE, TRUE, FALSE)"""
    dfe, dfa, dfm = readme_frames
    prep = analysis.prepare(dfe, dfa, dfm, 1, 3)
    got = run_gpu(sess, prep.expr, prep.meth, prep.index)
    assert_parity(got, run_oracle(prep.expr, prep.meth, prep.index), "readme")
    gold = json.load(open(os.path.join(HERE, "golden", "readme_fixture.json")))
    df = analysis.assemble_table(prep, got)
    for name, want in gold["rows"].items():
        np.testing.assert_allclose(df.loc[name].to_numpy()[:17].astype(float), want, rtol=1e-8, atol=1e-10)
        assert df.loc[name, "island"] == gold["islands"][name]


def test_public_api_writes_the_reference_csv(tmp_path, readme_frames):
    from epimethex_b200 import epimethex_analysis
    dfe, dfa, dfm = readme_frames
    df = epimethex_analysis(dfe, dfa, dfm, 1, 3, 2, True, True, False, output_dir=str(tmp_path))
    text = (tmp_path / "1-3" / "CG_data_individually.csv").read_text().splitlines()
    assert len(text) == 4 and text[0].count(",") == 18
    assert text[1].startswith('"cg11663302_Body_ARHGEF10L","0.9841","0.9825","0.9872"')
    assert df.shape == (3, 18)


@pytest.mark.parametrize("te,tm", [(True, False), (True, True), (False, False), (False, True)])
@pytest.mark.parametrize("n", [6, 30, 99, 474, 999])
def test_modes_and_sizes_divisible_by_3(sess, n, te, tm):
    G, P = (300, 4000) if n <= 99 else (120, 2500)
    x, y, idx = _synth(G, n, P)
    got = run_gpu(sess, x, y, idx, te=te, tm=tm)
    assert_parity(got, run_oracle(x, y, idx, te=te, tm=tm), f"n={n} te={te} tm={tm}")


@pytest.mark.parametrize("n", [7, 31, 473, 1000])
def test_extension_for_n_not_divisible_by_3(sess, n):
    """the reference stops here; extension = the n mod 3 lowest-expression samples are unlabelled."""
    x, y, idx = _synth(100, n, 2000)
    for tm in (False, True):
        got = run_gpu(sess, x, y, idx, tm=tm)
        assert_parity(got, run_oracle(x, y, idx, tm=tm), f"n={n} tm={tm}")
    with pytest.raises(_native.EpxError) as e:
        sess.run(_native.make_options(strict_n3=True))
    assert e.value.code == -6


@pytest.mark.parametrize("n,decimals", [(30, 2), (99, 2), (474, 3), (999, 4)])
def test_ties_in_methylation(sess, n, decimals):
    """Xena-style rounded beta values: KS evaluated at tie-run ends; exact p only without pooled ties."""
    x, y, idx = _synth(150, n, 3000, decimals=decimals)
    got = run_gpu(sess, x, y, idx)
    ref = run_oracle(x, y, idx)
    assert_parity(got, ref, f"ties n={n}")


def test_missing_probes_constant_rows_and_guard(sess):
    x, y, idx = _synth(80, 30, 1500, missing_every=4)
    y = y.copy()
    rows = np.unique(idx.probe_idx[idx.probe_idx >= 0])
    y[rows[0]] = 0.25                                   # constant row: guard NA x3, Pearson NA
    y[rows[1]] = np.tile([0.1, 0.2, 0.3], 10)           # differences between rank-paired groups may be constant
    g = idx.pair_gene[np.nonzero(idx.probe_idx == rows[2])[0][0]]
    perm = np.argsort(-x[g], kind="stable")
    v = np.empty(30); v[perm] = np.concatenate([np.linspace(0.5, 0.6, 10), np.linspace(0.4, 0.5, 10), np.linspace(0.1, 0.9, 10)])
    y[rows[2]] = v                                      # UP - Medium is exactly constant for gene g: guard hits one test
    for tm in (False, True):
        got = run_gpu(sess, x, y, idx, tm=tm)
        ref = run_oracle(x, y, idx, tm=tm)
        assert_parity(got, ref, f"edge tm={tm}")
        assert (ref.pair_na == 0x7FF).any() and ((ref.pair_na & 0x1C0) != 0).any()


@pytest.mark.parametrize("n,decimals", [(1500, None), (2001, 3), (4500, 2)])
def test_long_rows_use_the_cta_per_pair_kernel(sess, n, decimals):
    """n > 1024 (towards the 9,000-sample pan-cancer shape): CTA-per-row sort + CTA-per-pair statistics"""
    x, y, idx = _synth(40, n, 600, decimals=decimals)
    for te, tm in ((True, False), (False, True)):
        got = run_gpu(sess, x, y, idx, te=te, tm=tm)
        assert_parity(got, run_oracle(x, y, idx, te=te, tm=tm), f"n={n} te={te} tm={tm}")


def test_log_fold_change_and_means_variant(sess):
    x, y, idx = _synth(200, 60, 1000, ties=False)
    for linear, fcm in ((False, False), (True, True), (False, True)):
        got = run_gpu(sess, x, y, idx, linear=linear, fc_means=fcm)
        assert_parity(got, run_oracle(x, y, idx, linear=linear, fc_means=fcm), f"linear={linear} means={fcm}")


def test_session_reuse_across_gene_ranges(sess):
    """README: users call per gene range; the uploaded methylation matrix stays on the device."""
    ann = synth.annotation(3000, 200)
    y = synth.methylation(3000, 99, ann, threads=4)
    sess.set_methylation(y)
    for g0, g1 in ((0, 70), (70, 200), (0, 200)):
        x = synth.expression(g1 - g0, 99, g0=g0)
        idx = synth.probe_index(ann, g0, g1)
        got = run_gpu(sess, x, y, idx, upload=False)
        assert_parity(got, run_oracle(x, y, idx), f"range {g0}-{g1}")


def test_sample_major_upload_matches_row_major(sess):
    x, y, idx = _synth(60, 45, 700)
    a = run_gpu(sess, x, y, idx)
    sess.set_methylation_columns([np.ascontiguousarray(y[:, j]) for j in range(y.shape[1])])
    b = run_gpu(sess, x, y, idx, upload=False)
    assert np.array_equal(a.pair, b.pair, equal_nan=True) and np.array_equal(a.ks_dnum, b.ks_dnum)


def test_breaks_not_unique_is_reported(sess):
    """bin.var stops with "'breaks' are not unique" when the reference gene is mostly tied (reference behaviour)."""
    x = np.zeros((3, 9)); x[:, 0] = [1, 2, 3]
    y = np.random.default_rng(0).random((4, 9))
    sess.set_methylation(y)
    with pytest.raises(_native.EpxError) as e:
        sess.analyze(x, np.array([0, 1, 2, 4]), np.array([0, 1, 2, 3], np.int32))
    assert e.value.code == -5


def test_full_size_properties_skcm_slice(sess):
    """BASELINE-size rows (473 samples), a 1,500-gene slice: size-independent properties instead of the oracle:
    (i) relabelling symmetry: reversing every expression row swaps UP and Down (no ties): medians swap, KS
    numerators of UPvsMID / MIDvsDOWN swap, Pearson r flips sign; (ii) affine invariance of r."""
    ann = synth.annotation(40000, 1500)
    y = synth.methylation(40000, 474, ann, threads=8)
    idx = synth.probe_index(ann, 0, 1500)
    x = synth.expression(1500, 474, ties=False)
    sess.set_methylation(y)
    a = run_gpu(sess, x, y, idx, upload=False)
    b = run_gpu(sess, -x, y, idx, upload=False)
    ok = a.pair_na == 0
    np.testing.assert_array_equal(a.pair[ok][:, [0, 1, 2]], b.pair[ok][:, [2, 1, 0]])
    np.testing.assert_array_equal(a.ks_dnum[ok][:, [0, 1, 2]], b.ks_dnum[ok][:, [2, 1, 0]])
    np.testing.assert_allclose(a.pair[ok][:, 9], -b.pair[ok][:, 9], rtol=1e-9, atol=1e-12)
    c = run_gpu(sess, 3.0 * x + 1.0, y, idx, upload=False)
    np.testing.assert_allclose(a.pair[ok][:, 9], c.pair[ok][:, 9], rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(a.ks_dnum, c.ks_dnum)
    # and a sample of genes against the oracle
    sel = np.arange(0, 1500, 25)
    ptr = np.zeros(sel.size + 1, np.int64); pi = []
    for k, g in enumerate(sel):
        pi.append(idx.probe_idx[idx.row_ptr[g]:idx.row_ptr[g + 1]]); ptr[k + 1] = ptr[k] + pi[-1].size
    sub = type(idx)(ptr, np.concatenate(pi), None, None, None, None, None)
    ref = run_oracle(x[sel], y, sub)
    got = run_gpu(sess, x[sel], y, sub, upload=False)
    assert_parity(got, ref, "skcm slice sample")


@pytest.mark.parametrize("n", [3, 4, 5, 8])
def test_tiny_sample_counts(sess, n):
    """m = 1 or 2: the guard is NA (sd of one difference) -> KS still runs, Welch is NA; exact KS with m*m < 10000"""
    x, y, idx = _synth(30, n, 200, ties=False)
    for te, tm in ((True, False), (False, True), (False, False)):
        got = run_gpu(sess, x, y, idx, te=te, tm=tm)
        assert_parity(got, run_oracle(x, y, idx, te=te, tm=tm), f"n={n} te={te} tm={tm}")


def test_degenerate_index_shapes(sess):
    """one gene; genes without any pair; a probe listed twice in one gene; no pairs at all"""
    rng = np.random.default_rng(11)
    y = rng.random((12, 21))
    x = rng.normal(5, 2, (4, 21))
    x[2] = 7.0                                            # constant gene: sd 0 -> Pearson NA, tests NA
    Idx = type(synth.probe_index(synth.annotation(10, 2), 0, 2))
    def mk(row_ptr, pi):
        return Idx(np.asarray(row_ptr, np.int64), np.asarray(pi, np.int32), None, None, None, None, None)
    for name, xx, idx in (("one gene", x[:1], mk([0, 3], [0, 5, 5])),
                          ("empty genes", x, mk([0, 0, 4, 6, 6], [1, 2, 3, -1, 7, 8])),
                          ("no pairs", x, mk([0, 0, 0, 0, 0], []))):
        got = run_gpu(sess, xx, y, idx)
        assert_parity(got, run_oracle(xx, y, idx), name)


def test_heavily_tied_expression(sess):
    """RNA-seq style zeros: the Medium/Down boundary falls inside a run of ties (stable order by sample index);
    expression-side KS with pooled ties"""
    ann = synth.annotation(3000, 200)
    y = synth.methylation(3000, 60, ann, threads=4)
    idx = synth.probe_index(ann, 0, 200)
    x = synth.expression(200, 60)
    x[x < 4.0] = 0.0                                       # most genes now have long zero runs
    x[:, :3] += np.arange(3)[None, :] * 1e-3 + 9.0         # keep every gene non-constant and one gene tie-free enough
    x[0] = np.linspace(1, 60, 60)                          # a reference gene without ties (bin.var needs unique breaks)
    for te in (True, False):
        got = run_gpu(sess, x, y, idx, te=te)
        assert_parity(got, run_oracle(x, y, idx, te=te), f"tied expression te={te}")


@pytest.mark.parametrize("n", [30, 474, 1000, 1500])
def test_values_closer_than_the_sort_keys_coarse_resolution(sess, n):
    """The warp sort orders rows by a 22/23-bit coarse key packed with the sample index and repairs neighbours that share a
    coarse key from the full doubles. Rows whose values differ only in the last mantissa bits make every neighbour pair
    collide: long runs of distinct values with equal coarse keys, plus exact ties mixed in. (The rows keep an ordinary
    spread: a row that varies ONLY in its last bits makes the reference's own long-double Pearson noisy at 1e-8.)"""
    rng = np.random.default_rng(n)
    G, P = 24, 160
    ulp = 2.0 ** -52
    x = rng.normal(7.0, 2.0, (G, n))                                    # expression: ordinary values ...
    half = n // 2
    x[:, half:2 * half] = x[:, :half] * (1 + ulp * rng.integers(-2, 3, (G, half)))   # ... each with a twin 0-2 ulps away
    y = rng.random((P, n))
    y[:40, half:2 * half] = y[:40, :half] * (1 + ulp * rng.integers(-2, 3, (40, half)))      # twins 0-2 ulps apart
    y[40:80] = np.round(y[40:80], 2) * (1 + ulp * rng.integers(0, 3, (40, n)))               # ~100 levels x 3 near-values: long runs
    y[80:120] = np.where(y[80:120] < 0.5, 0.25, 0.75) * (1 + ulp * rng.integers(0, 4, (40, n)))  # two clusters of 4 near-values
    row_ptr = np.arange(0, G * 20 + 1, 20, dtype=np.int64)
    probe_idx = rng.integers(0, P, G * 20).astype(np.int32)

    class Idx:
        pass
    idx = Idx(); idx.row_ptr, idx.probe_idx = row_ptr, probe_idx
    for te, tm in ((True, False), (False, True)):
        got = run_gpu(sess, x, y, idx, te=te, tm=tm)
        assert_parity(got, run_oracle(x, y, idx, te=te, tm=tm), f"near-ties n={n} te={te} tm={tm}")


@pytest.mark.parametrize("tm", [False, True])
def test_repeated_and_chunked_runs_are_identical(sess, tm):
    """The rank and pair kernels draw their rows / work items from ticket counters that the last warp re-arms:
    a second run on the same session, and a run cut into chunks of work items, must give the same tables."""
    x, y, idx = _synth(400, 473, 6000, missing_every=11)
    first = run_gpu(sess, x, y, idx, tm=tm)
    assert_parity(first, run_oracle(x, y, idx, tm=tm), f"first run tm={tm}")
    opts = _native.make_options(True, True, tm)
    for n_chunks in (1, 3, 7):
        sess.run_prepare(opts)
        for c in range(n_chunks):
            sess.run_pairs(c, n_chunks)
        again = sess.fetch()
        for name in ("pair", "pair_stat", "gene"):
            assert np.array_equal(getattr(first, name), getattr(again, name), equal_nan=True), (name, n_chunks)
        for name in ("ks_dnum", "pair_na", "gene_na", "perm"):
            assert np.array_equal(getattr(first, name), getattr(again, name)), (name, n_chunks)
    covered = 0
    for c in range(7):
        b, cnt = sess.pairs_chunk(c, 7)
        assert b == covered
        covered += cnt
    assert covered == idx.probe_idx.size
