"""The CUDA path (through the C ABI) against the golden tables of the independent second pipeline
(tests/golden/rpipeline.py; see tests/golden/make_golden.py).  Integers identical, doubles within the BASELINE bar
(1e-10 absolute or 1e-8 relative).  Run with -m gpu on a B200."""
import numpy as np
import pytest

import goldens
from epimethex_b200 import _native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sess():
    s = _native.Session(0)
    yield s
    s.close()


@pytest.mark.parametrize("name", goldens.CASES)
def test_gpu_matches_the_independent_pipeline(sess, name):
    g, x, y, row_ptr, probe_idx = goldens.load(name)
    sess.set_methylation(y)
    gp, gl = goldens.gene_groups(row_ptr, probe_idx)
    for tm in (False, True):
        r = sess.analyze(x, row_ptr, probe_idx, _native.make_options(True, True, tm))
        goldens.check_pairs(g, r, tm, "gpu")
        goldens.check_groups(g, sess.analyze_groups(gp, gl), tm, "gpu")
    r = sess.analyze(x, row_ptr, probe_idx, _native.make_options(True, False, False))
    goldens.check_genes(g, r, False, "gpu")
    r = sess.analyze(x, row_ptr, probe_idx, _native.make_options(True, True, False, fc_from_means=True))
    goldens._close(r.gene[:, :3], g["gene_fc_means_linear"], "fold change of the group means", **goldens.TOL["gpu"])
    r = sess.analyze(x, row_ptr, probe_idx, _native.make_options(False, True, False))
    goldens._close(r.gene[:, :3], g["gene_fc_quantiles_log"], "log fold change of the quantile rows", **goldens.TOL["gpu"])


def test_readme_fixture_values(sess, readme_frames):
    """BASELINE configs[0] against the fixture written by the independent pipeline"""
    import json
    import os
    from epimethex_b200 import analysis
    gold = json.load(open(os.path.join(goldens.HERE, "golden", "readme_fixture.json")))
    dfe, dfa, dfm = readme_frames
    prep = analysis.prepare(dfe, dfa, dfm, 1, 3)
    sess.set_methylation(prep.meth)
    got = sess.analyze(prep.expr, prep.index.row_ptr, prep.index.probe_idx, _native.make_options(True, True, False))
    df = analysis.assemble_table(prep, got)
    for name, want in gold["rows"].items():
        np.testing.assert_allclose(df.loc[name].to_numpy()[:17].astype(float), want, rtol=1e-8, atol=1e-10)
