"""The .Call shim driven on the GPU with fake SEXPs (tests/_rstub/): C_epx_upload -> C_epx_analyze -> C_epx_groups ->
C_epx_release on the README fixture and on a synthetic slice, against the oracle; PROTECT balance under a collector
that runs at every allocation, NA_real_ (payload 1954) vs plain NaN, 1-based perm, integer columns, an interrupt in
the middle of a run, errors of libepx surfacing as R errors.  Run with -m gpu on a B200.
Reference boundary replaced: NAMESPACE:3, R/EpimethexAnalysis.R:70-72."""
import numpy as np
import pytest

import oracle
import rshim
from epimethex_b200 import analysis, synth
from epimethex_b200.annotation import build_groups

pytestmark = pytest.mark.gpu


def _clean(rep):
    assert rep.protect_depth == 0, "PROTECT/UNPROTECT imbalance"
    assert rep.ralloc_live == 0, "R_alloc memory outlived the call"


def _compare(out, ref, n):
    assert np.array_equal(out["perm"], ref.perm + 1), "perm must be 1-based"
    assert tuple(out["counts"][:3]) == ref.counts and out["counts"][3] == ref.ref_gene + 1
    dn = ref.ks_dnum.copy()
    want_dn = np.where(dn < 0, rshim.NA_INT, dn)
    assert np.array_equal(out["dnum"], want_dn)
    # cells the oracle flags NA must be R's NA_real_; every other NaN must be a plain NaN
    flagged = ((ref.pair_na[:, None] >> np.arange(11)[None, :]) & 1) == 1
    assert np.array_equal(out["is_na"] == 1, flagged), "NA_real_ pattern differs from the NA mask"
    assert np.isclose(out["pair"], ref.pair, rtol=1e-8, atol=1e-10, equal_nan=True).all()
    gflag = ((ref.gene_na[:, None] >> np.arange(15)[None, :]) & 1) == 1
    gene = np.where(gflag, np.nan, out["gene"])
    assert np.isclose(gene, np.where(gflag, np.nan, ref.gene), rtol=1e-8, atol=1e-10, equal_nan=True).all()


def test_readme_fixture_through_the_call_boundary(readme_frames):
    dfe, dfa, dfm = readme_frames
    prep = analysis.prepare(dfe, dfa, dfm, 1, 3)
    cols = [np.ascontiguousarray(prep.meth[:, j]) for j in range(prep.meth.shape[1])]   # the data.frame's columns
    h, rep = rshim.upload(cols)
    assert rep.rc == 0, rep.message
    _clean(rep)
    assert rep.live_finalizers == 1
    out, rep = rshim.analyze(h, prep.expr, prep.index.row_ptr, prep.index.probe_idx, opts=(1, 1, 0, 0, 0))
    assert rep.rc == 0, rep.message
    _clean(rep)
    ref = oracle.analyze(prep.expr, prep.meth, prep.index.row_ptr, prep.index.probe_idx)
    _compare(out, ref, prep.expr.shape[1])
    rep = rshim.release(h)
    assert rep.rc == 0 and rep.live_finalizers == 0
    out, rep = rshim.analyze(h, prep.expr, prep.index.row_ptr, prep.index.probe_idx)
    assert rep.rc == 1 and "released" in rep.message
    rshim.finish()


@pytest.mark.parametrize("tm", [0, 1])
def test_synthetic_slice_with_missing_probes_groups_and_integer_columns(tm):
    n, G, P = 473, 150, 2500
    ann = synth.annotation(P, G)
    x = synth.expression(G, n)
    y = synth.methylation(P, n, ann, decimals=3, threads=4)
    has = np.ones(P, bool); has[::9] = False
    idx = synth.probe_index(ann, 0, G, probe_has_data=has)
    used = np.unique(idx.probe_idx[idx.probe_idx >= 0])
    y[used[0]] = 0.25                          # constant probe: guard NA x 3, Pearson NA -> NA_real_ cells
    y[:, 5] = np.rint(y[:, 5])                 # a 0/1 column: read.csv would hand it over as integer
    cols = [np.ascontiguousarray(y[:, j]) for j in range(n)]
    as_int = [0] * n; as_int[5] = 1
    h, rep = rshim.upload(cols, as_int=as_int)
    assert rep.rc == 0, rep.message
    _clean(rep)
    out, rep = rshim.analyze(h, x, idx.row_ptr, idx.probe_idx, opts=(1, 1, tm, 0, 0))
    assert rep.rc == 0, rep.message
    _clean(rep)
    ref = oracle.analyze(x, y, idx.row_ptr, idx.probe_idx, test_meth_t=bool(tm), threads=8)
    _compare(out, ref, n)
    assert (out["is_na"].sum(axis=1) == 11).any() and (out["is_na"].sum(axis=1) > 0).sum() > (out["is_na"].sum(axis=1) == 11).sum()
    # pooled rows (CG_of_a_genes) through C_epx_groups
    gi = build_groups(idx, "gene")
    grp, rep = rshim.groups(h, gi.group_ptr, gi.group_pairs)
    assert rep.rc == 0, rep.message
    _clean(rep)
    gref = oracle.analyze_groups(x, y, idx.row_ptr, idx.probe_idx, gi.group_ptr, gi.group_pairs, test_meth_t=bool(tm), threads=8)
    gflag = ((gref.group_na[:, None] >> np.arange(11)[None, :]) & 1) == 1
    assert np.isclose(np.where(gflag, np.nan, grp), np.where(gflag, np.nan, gref.group), rtol=1e-8, atol=1e-10, equal_nan=True).all()
    # an interrupt at the first poll (before anything is launched): the call unwinds, the session survives
    _, rep = rshim.analyze(h, x, idx.row_ptr, idx.probe_idx, interrupt_at=1)
    assert rep.rc == 2 and rep.protect_depth == 0 and rep.ralloc_live == 0
    out2, rep = rshim.analyze(h, x, idx.row_ptr, idx.probe_idx, opts=(1, 1, tm, 0, 0))
    assert rep.rc == 0 and np.array_equal(out2["pair"], out["pair"], equal_nan=True)
    # malformed arguments and libepx errors arrive as R errors, nothing leaks
    for kwargs, text in ((dict(opts=(1, 1, 0, 0)), "opts must hold 5"), (dict(opts=(1, 1, 0, 0, 1)), "differing number of rows")):
        _, rep = rshim.analyze(h, x, idx.row_ptr, idx.probe_idx, **kwargs)
        assert rep.rc == 1 and text in rep.message, rep.message
        _clean(rep)
    _, rep = rshim.analyze(h, x, idx.row_ptr[:-1], idx.probe_idx)
    assert rep.rc == 1 and "rowPtr has" in rep.message
    _, rep = rshim.analyze(h, x, idx.row_ptr, idx.probe_idx[:-1])
    assert rep.rc == 1 and "probeIdx has" in rep.message
    bad = idx.probe_idx.copy(); bad[3] = P + 7
    _, rep = rshim.analyze(h, x, idx.row_ptr, bad)
    assert rep.rc == 1 and "n_probes" in rep.message
    xx = x.copy(); xx[3, 3] = np.nan
    _, rep = rshim.analyze(h, xx, idx.row_ptr, idx.probe_idx)
    assert rep.rc == 1 and "NaN" in rep.message
    rep = rshim.finish()                       # end of the R session: the finalizer releases the device copy
    assert rep.live_finalizers == 0


def test_interrupt_between_slices_of_the_pair_kernel():
    """more than 200,000 pairs: the shim cuts the pair kernel into 8 slices and polls R_CheckUserInterrupt between them"""
    n, G, P = 30, 4000, 230000
    ann = synth.annotation(P, G)
    x = synth.expression(G, n)
    y = synth.methylation(P, n, ann, threads=8)
    idx = synth.probe_index(ann, 0, G)
    assert idx.n_pairs > 200000
    h, rep = rshim.upload([np.ascontiguousarray(y[:, j]) for j in range(n)])
    assert rep.rc == 0, rep.message
    _, rep = rshim.analyze(h, x, idx.row_ptr, idx.probe_idx, interrupt_at=4)   # poll 1 before, polls 2.. between slices
    assert rep.rc == 2 and rep.protect_depth == 0 and rep.ralloc_live == 0
    out, rep = rshim.analyze(h, x, idx.row_ptr, idx.probe_idx)
    assert rep.rc == 0, rep.message
    ref = oracle.analyze(x, y, idx.row_ptr, idx.probe_idx, threads=8)
    _compare(out, ref, n)
    rshim.finish()


def test_device_info_names_the_gpu():
    import ctypes as C
    name = C.create_string_buffer(256)
    rep = rshim.Report()
    assert rshim.lib().rshim_device_info(name, 256, C.byref(rep)) == 0 and b"B200" in name.value
