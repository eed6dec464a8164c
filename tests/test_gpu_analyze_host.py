"""epx_analyze_host: the methylation matrix stays in host memory and only the rows the index names are uploaded
(the methylation filter of R/EpimethexAnalysis.R:278-282 folded into the upload). Against the oracle, against the
two-call path (set_methylation + analyze: the tables must be identical bit for bit, the kernels see the same rows), and
the state rules of the partial matrix. Run with -m gpu on a B200."""
from types import SimpleNamespace

import numpy as np
import pytest

from epimethex_b200 import _native, synth
from parity import assert_parity, run_gpu, run_oracle

pytestmark = pytest.mark.gpu

TABLES = ("pair", "pair_stat", "gene")
INTS = ("ks_dnum", "gene_ks_dnum", "pair_na", "gene_na", "perm")


def _synth(n_genes, n, n_probes, first=0, last=None, missing_every=0):
    ann = synth.annotation(n_probes, n_genes)
    x = synth.expression(n_genes, n, ties=True)
    y = synth.methylation(n_probes, n, ann, decimals=None, threads=4)
    has = None
    if missing_every:
        has = np.ones(n_probes, bool); has[::missing_every] = False
    last = n_genes if last is None else last
    return x[first:last], y, synth.probe_index(ann, first, last, probe_has_data=has)


def _pinned(a):
    p = _native.pinned_empty(a.shape)
    p[...] = a
    return p


def _same(a, b, label):
    assert a.counts == b.counts and a.ref_gene == b.ref_gene, label
    for k in TABLES:
        assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True), f"{label}: {k} differs"
    for k in INTS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), f"{label}: {k} differs"


@pytest.mark.parametrize("n", [6, 99, 473, 1000, 1500])
@pytest.mark.parametrize("tm", [False, True])
def test_page_locked_matrix_against_oracle_and_two_call_path(n, tm):
    """a gene range that names a fraction of the rows; probes without data (-1) in the index"""
    G, P = (200, 3000) if n <= 473 else (60, 1200)
    x, y, idx = _synth(G, n, P, first=G // 4, last=G // 2, missing_every=7)
    named = np.unique(idx.probe_idx[idx.probe_idx >= 0])
    assert 0 < named.size < P // 2                     # most rows stay behind
    opts = _native.make_options(True, True, tm)
    yp = _pinned(y)
    with _native.Session(0) as s:
        got = s.analyze_host(yp, x, idx.row_ptr, idx.probe_idx, opts)
        assert got.timing["n_unique_probes"] == named.size
    assert_parity(got, run_oracle(x, y, idx, tm=tm), f"analyze_host n={n} tm={tm}")
    with _native.Session(0) as s:
        _same(got, run_gpu(s, x, y, idx, tm=tm), f"n={n} tm={tm}")


def test_rows_no_gene_names_are_never_read():
    """NaN in a row outside the index: the reference drops the row before any arithmetic (na.omit + subset,
    R/EpimethexAnalysis.R:278-282); the two-call path, which uploads every row, refuses the matrix"""
    x, y, idx = _synth(120, 99, 2000, first=0, last=40)
    named = np.zeros(y.shape[0], bool)
    named[idx.probe_idx[idx.probe_idx >= 0]] = True
    free = np.nonzero(~named)[0]
    want = run_oracle(x, y, idx)
    y2 = y.copy()
    y2[free[::3], 5] = np.nan
    y2[free[1::3], 0] = np.inf
    yp = _pinned(y2)
    with _native.Session(0) as s:
        got = s.analyze_host(yp, x, idx.row_ptr, idx.probe_idx)
        assert_parity(got, want, "NaN outside the index")
        with pytest.raises(_native.EpxError) as e:
            s.set_methylation(y2)
        assert e.value.code == -1
    # ... and a NaN in a row the index names is refused here too, by probe and sample
    bad_row = int(np.nonzero(named)[0][17])
    yp[bad_row, 42] = np.nan
    with _native.Session(0) as s:
        with pytest.raises(_native.EpxError) as e:
            s.analyze_host(yp, x, idx.row_ptr, idx.probe_idx)
        assert e.value.code == -1 and f"probe {bad_row} " in str(e.value) and "sample 42" in str(e.value)
        with pytest.raises(_native.EpxError) as e:      # no usable matrix is left behind
            s.run()
        assert e.value.code == -4


def test_row_stride_and_unaligned_rows():
    """rows of 473 doubles start at every multiple of 8 bytes; a padded leading dimension; a matrix that starts in the
    middle of a page-locked block"""
    x, y, idx = _synth(80, 473, 1500, first=10, last=50)
    want = run_oracle(x, y, idx)
    with _native.Session(0) as s:
        for ld, shift in ((473, 0), (473, 5), (480, 0), (501, 3)):
            block = _native.pinned_empty((y.shape[0] * ld + shift + 8,))
            block[...] = -7.0                                   # a wrong value wherever the kernel must not read
            view = block[shift:shift + y.shape[0] * ld].reshape(y.shape[0], ld)[:, :473]
            view[...] = y
            got = s.analyze_host(view, x, idx.row_ptr, idx.probe_idx)
            assert_parity(got, want, f"ld={ld} shift={shift}")


def test_pageable_matrix_takes_the_staged_upload():
    x, y, idx = _synth(100, 99, 1500, first=0, last=60, missing_every=5)
    with _native.Session(0) as s:
        got = s.analyze_host(y, x, idx.row_ptr, idx.probe_idx)      # plain numpy memory
        assert_parity(got, run_oracle(x, y, idx), "pageable")
        # the whole matrix is resident after the fallback: another index is fine
        x2, _, idx2 = _synth(100, 99, 1500, first=60, last=100, missing_every=5)
        got2 = s.analyze(x2, idx2.row_ptr, idx2.probe_idx)
        assert_parity(got2, run_oracle(x2, y, idx2), "pageable, next range")


def test_partial_matrix_state_rules():
    x, y, idx = _synth(150, 99, 2500, first=0, last=50)
    x2, _, idx2 = _synth(150, 99, 2500, first=50, last=150)
    yp = _pinned(y)
    with _native.Session(0) as s:
        a = s.analyze_host(yp, x, idx.row_ptr, idx.probe_idx)
        # same index again, other options: the rows are there
        s.run(_native.make_options(True, True, True))
        b = s.fetch()
        assert_parity(b, run_oracle(x, y, idx, tm=True), "re-run on the partial matrix")
        s.set_index(idx.row_ptr, idx.probe_idx)                     # the same index is accepted
        for call in (lambda: s.set_index(idx2.row_ptr, idx2.probe_idx),
                     lambda: s.analyze(x2, idx2.row_ptr, idx2.probe_idx)):
            with pytest.raises(_native.EpxError) as e:
                call()
            assert e.value.code == -4 and "epx_analyze_host" in str(e.value)
        # the intended use: one call per gene range; then a changed matrix; then back to a resident matrix
        c = s.analyze_host(yp, x2, idx2.row_ptr, idx2.probe_idx)
        assert_parity(c, run_oracle(x2, y, idx2), "next range")
        yp[...] = np.round(y, 2)
        d = s.analyze_host(yp, x, idx.row_ptr, idx.probe_idx)
        assert_parity(d, run_oracle(x, np.round(y, 2), idx), "changed matrix, first index again")
        s.set_methylation(y)
        _same(a, s.analyze(x, idx.row_ptr, idx.probe_idx), "resident matrix again")
        assert_parity(s.analyze(x2, idx2.row_ptr, idx2.probe_idx), run_oracle(x2, y, idx2), "resident, next range")


def test_pooled_groups_after_analyze_host():
    import oracle
    from epimethex_b200.annotation import build_groups
    x, y, idx = _synth(90, 99, 1500, first=20, last=70)
    gi = build_groups(idx, "gene")
    yp = _pinned(y)
    with _native.Session(0) as s:
        s.analyze_host(yp, x, idx.row_ptr, idx.probe_idx, full=False)
        gg = s.analyze_groups(gi.group_ptr, gi.group_pairs)
    gr = oracle.analyze_groups(x, y, idx.row_ptr, idx.probe_idx, gi.group_ptr, gi.group_pairs)
    assert np.array_equal(gg.group_dnum, gr.group_dnum) and np.array_equal(gg.group_na, gr.group_na)
    assert np.isclose(gg.group, gr.group, rtol=1e-8, atol=1e-10, equal_nan=True).all()


def test_empty_index_and_bad_arguments():
    x, y, idx = _synth(30, 30, 400)
    yp = _pinned(y)
    with _native.Session(0) as s:
        rp0 = np.zeros(x.shape[0] + 1, np.int64)
        got = s.analyze_host(yp, x, rp0, np.zeros(0, np.int32))      # no pairs at all: gene table only
        assert got.pair.shape[0] == 0
        assert_parity(got, run_oracle(x, y, SimpleNamespace(row_ptr=rp0, probe_idx=np.zeros(0, np.int32))), "empty")
        bad = idx.probe_idx.copy(); bad[3] = y.shape[0]
        with pytest.raises(_native.EpxError) as e:
            s.analyze_host(yp, x, idx.row_ptr, bad)
        assert e.value.code == -1
        with pytest.raises(ValueError):
            s.analyze_host(yp[:, :20], x, idx.row_ptr, idx.probe_idx)   # sample counts differ
        assert_parity(s.analyze_host(yp, x, idx.row_ptr, idx.probe_idx), run_oracle(x, y, idx), "after the errors")


def test_several_devices_one_process_each_fetches_the_rows_of_its_gene_range():
    """epx_multi_analyze_host. On a one-GPU box the sessions share device 0 -- the gene cuts, the per-device indices and
    the row fetch are the same code."""
    from epimethex_b200.annotation import build_groups
    nd = min(_native.device_count(), 4)
    devices = list(range(nd)) if nd >= 2 else [0, 0, 0]
    x, y, idx = _synth(300, 99, 5000, missing_every=37)
    x2, _, idx2 = _synth(300, 99, 5000, first=0, last=150)
    assert (idx.probe_idx < 0).any()
    yp = _pinned(y)
    gi = build_groups(idx, "gene")
    with _native.Session(0) as s:
        s.set_methylation(y)
        one = s.analyze(x, idx.row_ptr, idx.probe_idx)
        one_groups = s.analyze_groups(gi.group_ptr, gi.group_pairs)
        one2 = s.analyze(x2, idx2.row_ptr, idx2.probe_idx)
    with _native.MultiSession(devices) as ms:
        for _ in range(2):      # twice: the sessions and their allocations are reused
            _same(ms.analyze_host(yp, x, idx.row_ptr, idx.probe_idx), one, "multi analyze_host")
        got = ms.analyze_groups(gi.group_ptr, gi.group_pairs)
        assert np.array_equal(got.group, one_groups.group, equal_nan=True) and np.array_equal(got.group_dnum, one_groups.group_dnum)
        # every device holds the rows of ITS range of THIS index only: the resident-matrix call refuses another index
        with pytest.raises(_native.EpxError, match="EPX_ESTATE"):
            ms.analyze(x2, idx2.row_ptr, idx2.probe_idx)
        _same(ms.analyze_host(yp, x2, idx2.row_ptr, idx2.probe_idx), one2, "multi analyze_host, next range")
        # pageable matrix: replicas through the staged upload, after which any index works
        _same(ms.analyze_host(y, x, idx.row_ptr, idx.probe_idx), one, "multi analyze_host, pageable")
        _same(ms.analyze(x2, idx2.row_ptr, idx2.probe_idx), one2, "replicas resident")
        # a NaN in a named row is refused by the device that fetches it
        yp[int(idx.probe_idx[idx.probe_idx >= 0][-1]), 3] = np.nan
        with pytest.raises(_native.EpxError, match="EPX_EINVAL"):
            ms.analyze_host(yp, x, idx.row_ptr, idx.probe_idx)
