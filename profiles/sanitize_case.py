#!/usr/bin/env python
"""Workload for the compute-sanitizer passes (one tool per gpurun call, B200_PROFILING.md): the README fixture, slices at
n = 473 / 1,000 / 1,500 samples (warp kernels with the TMA ring, the two-half rank kernel, the CTA-per-row kernels) in both
methylation-test modes, ties included, and one pooled-group pass; every result is still compared with the oracle.
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python profiles/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import oracle  # noqa: E402
from epimethex_b200 import _native, analysis, synth  # noqa: E402
from epimethex_b200.annotation import build_groups  # noqa: E402


def check(got, ref):
    assert np.array_equal(got.perm.astype(np.int32), ref.perm) and np.array_equal(got.ks_dnum, ref.ks_dnum)
    assert np.array_equal(got.pair_na, ref.pair_na)
    assert np.isclose(got.pair, ref.pair, rtol=1e-8, atol=1e-10, equal_nan=True).all()
    assert np.isclose(got.gene, ref.gene, rtol=1e-8, atol=1e-10, equal_nan=True).all()


def main():
    small = "--small" in sys.argv
    with _native.Session(0) as sess:
        dfe, dfa, dfm = ge.analysis_fixture()
        prep = analysis.prepare(dfe, dfa, dfm, 1, 3)
        sess.set_methylation(prep.meth)
        check(sess.analyze(prep.expr, prep.index.row_ptr, prep.index.probe_idx, _native.make_options()),
              oracle.analyze(prep.expr, prep.meth, prep.index.row_ptr, prep.index.probe_idx))
        for n, dec in ((473, None), (1000, 3), (1500, 2)):
            G, P = (12, 120) if small else (24, 300)
            ann = synth.annotation(P, G)
            x = synth.expression(G, n)
            y = synth.methylation(P, n, ann, decimals=dec, threads=2)
            has = np.ones(P, bool); has[::11] = False
            idx = synth.probe_index(ann, 0, G, probe_has_data=has)
            sess.set_methylation(y)
            for te, tm in ((True, False), (False, True)):
                got = sess.analyze(x, idx.row_ptr, idx.probe_idx, _native.make_options(True, te, tm))
                check(got, oracle.analyze(x, y, idx.row_ptr, idx.probe_idx, test_expr_t=te, test_meth_t=tm, threads=4))
            if n == 473:
                gi = build_groups(idx, "gene")
                gg = sess.analyze_groups(gi.group_ptr, gi.group_pairs)
                gr = oracle.analyze_groups(x, y, idx.row_ptr, idx.probe_idx, gi.group_ptr, gi.group_pairs, test_meth_t=True, threads=4)
                assert np.array_equal(gg.group_dnum, gr.group_dnum)
                assert np.isclose(gg.group, gr.group, rtol=1e-8, atol=1e-10, equal_nan=True).all()
            print(f"n={n}: {idx.n_pairs} pairs ok", flush=True)
    print("sanitize_case ok")


if __name__ == "__main__":
    main()
