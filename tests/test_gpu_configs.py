"""Parity on BASELINE.json's own configurations: the CUDA path through the C ABI against the oracle at the sizes the
benchmark is quoted on (configs[1], configs[2] at its defined n = 474 AND at n = 473, configs[3] EPIC at n = 999 and
1,000 with KS on both sides, configs[4]'s 9,000-sample rows). Only the methylation rows some pair names are generated
(probe-sliced copy, as distributed.slice_probes does for a rank), so the host matrices stay small enough for any box.
Reference lines: Loop A R/EpimethexAnalysis.R:351-453, Analysis R/Functions.R:85-132."""
import os
from dataclasses import replace

import numpy as np
import pytest

from epimethex_b200 import _native, distributed, synth
from parity import assert_parity, run_gpu, run_oracle

pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 8


@pytest.fixture(scope="module")
def sess():
    s = _native.Session(0)
    yield s
    s.close()


def _config(shape, g0, g1, *, decimals=None, n=None):
    dims = dict(synth.SHAPES[shape])
    n = n or dims["n_samples"]
    ann = synth.annotation(dims["n_probes"], dims["n_genes"])
    idx = synth.probe_index(ann, g0, g1)
    rows, local = distributed.slice_probes(idx.probe_idx)
    y = np.empty((rows.size, n))
    step = max(16, 2_000_000 // n)
    from concurrent.futures import ThreadPoolExecutor

    def work(i):
        y[i:i + step] = synth.methylation_rows(rows[i:i + step], n, ann, decimals=decimals)

    with ThreadPoolExecutor(max_workers=min(16, CORES)) as ex:
        list(ex.map(work, range(0, rows.size, step)))
    x = synth.expression(g1 - g0, n, g0=g0)
    return x, y, replace(idx, probe_idx=local)


def test_config2_skcm_474_full_genome(sess):
    """BASELINE configs[2] at the nearest sample count the reference defines (n % 3 == 0): all 20,531 genes."""
    x, y, idx = _config("skcm474", 0, 20531)
    assert idx.n_pairs > 450_000
    got = run_gpu(sess, x, y, idx, te=True, tm=False)
    assert_parity(got, run_oracle(x, y, idx, te=True, tm=False, threads=CORES), "skcm474 full genome")


def test_config2_skcm_473_full_genome_extension(sess):
    """BASELINE configs[2] as quoted (473 samples: the n mod 3 extension), every one of the 474,611 pairs."""
    x, y, idx = _config("skcm", 0, 20531)
    assert idx.n_pairs == 474_611
    got = run_gpu(sess, x, y, idx, te=True, tm=False)
    assert_parity(got, run_oracle(x, y, idx, te=True, tm=False, threads=CORES), "skcm 473 full genome")


def test_config1_skcm_genes_1_to_1000_t_test_on_both(sess):
    """BASELINE configs[1]: gene range 1-1000 (the README's maximum), 473 samples, HM450 annotation, t-test on both."""
    x, y, idx = _config("skcm", 0, 1000)
    got = run_gpu(sess, x, y, idx, te=True, tm=True)
    assert_parity(got, run_oracle(x, y, idx, te=True, tm=True, threads=CORES), "skcm genes 1-1000 tt")


@pytest.mark.parametrize("shape", ["epic999", "epic"])
def test_config3_epic_ks_on_both(sess, shape):
    """BASELINE configs[3]: 866k probes, 999 / 1,000 samples, KS on expression and on methylation; a 3,000-gene range."""
    x, y, idx = _config(shape, 4000, 7000)
    assert idx.n_pairs > 100_000
    got = run_gpu(sess, x, y, idx, te=False, tm=False)
    assert_parity(got, run_oracle(x, y, idx, te=False, tm=False, threads=CORES), f"{shape} genes 4001-7000 ks/ks")


@pytest.mark.parametrize("decimals", [None, 3])
@pytest.mark.parametrize("te,tm", [(True, False), (True, True), (False, False)])
def test_config4_pancan_rows_of_9000_samples(sess, te, tm, decimals):
    """BASELINE configs[4]: 9,000-sample rows (CTA-per-row sort, CTA-per-pair statistics); >= 200 pairs, with and
    without Xena-style rounding (3 decimals: every row is full of ties)."""
    x, y, idx = _config("pancan", 100, 120, decimals=decimals)
    assert idx.n_pairs >= 200
    got = run_gpu(sess, x, y, idx, te=te, tm=tm)
    assert_parity(got, run_oracle(x, y, idx, te=te, tm=tm, threads=CORES), f"pancan n=9000 te={te} tm={tm} dec={decimals}")


def test_nan_and_infinite_inputs_are_refused(sess):
    """include/epx.h: NaN / infinite values are an EPX_EINVAL (the reference drops such probes and samples with
    na.omit, R/EpimethexAnalysis.R:114,278), through every upload path."""
    rng = np.random.default_rng(5)
    y = rng.random((50, 21))
    x = rng.normal(5, 1, (4, 21))
    for bad in (np.nan, np.inf, -np.inf):
        for shape_ld in (21, 24):                      # contiguous rows (re-spaced on the device) and padded rows
            yy = np.zeros((50, shape_ld)); yy[:, :21] = y
            yy[37, 5] = bad
            with pytest.raises(_native.EpxError) as e:
                sess.set_methylation(yy[:, :21])
            assert e.value.code == -1 and "probe 37" in str(e.value) and "sample 5" in str(e.value)
        cols = [np.ascontiguousarray(y[:, j]) for j in range(21)]
        cols[7] = cols[7].copy(); cols[7][12] = bad
        with pytest.raises(_native.EpxError) as e:
            sess.set_methylation_columns(cols)
        assert e.value.code == -1 and "probe 12" in str(e.value) and "sample 7" in str(e.value)
        xx = x.copy(); xx[2, 20] = bad
        with pytest.raises(_native.EpxError) as e:
            sess.set_expression(xx)
        assert e.value.code == -1 and "gene 2" in str(e.value)
    # after a refused upload there is no matrix: the sequence has to start again
    with pytest.raises(_native.EpxError) as e:
        sess.set_index(np.array([0, 1, 2, 3, 4]), np.array([0, 1, 2, 3], np.int32))
    assert e.value.code == -4
    sess.set_methylation(y)
    got = sess.analyze(x, np.array([0, 1, 2, 3, 4]), np.array([0, 1, 2, 3], np.int32))
    assert np.isfinite(got.pair[:, :6]).all()


def test_new_expression_invalidates_results(sess):
    """after set_expression the tables of the previous matrix must not be fetchable (their sizes follow
    G and n of the new one)."""
    rng = np.random.default_rng(6)
    y = rng.random((30, 12))
    sess.set_methylation(y)
    sess.analyze(rng.normal(5, 1, (3, 12)), np.array([0, 1, 2, 3]), np.array([0, 1, 2], np.int32))
    sess.fetch()
    sess.set_expression(rng.normal(5, 1, (9, 12)))
    for call in (sess.fetch, lambda: sess.result_device_ptr(0), lambda: sess.run_pairs(0, 1),
                 lambda: sess.run_groups(np.array([0, 1]), np.array([0], np.int32))):
        with pytest.raises(_native.EpxError) as e:
            call()
        assert e.value.code == -4
    with pytest.raises(_native.EpxError) as e:       # the index still covers 3 genes
        sess.run()
    assert e.value.code == -4
