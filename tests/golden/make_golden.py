"""Generates tests/golden/readme_fixture.json: expected values for the reference's only executable
example (R/EpimethexAnalysis.R:28-63; call epimethex.analysis(Expr, Annot, Meth, 1, 3, 2, TRUE, TRUE, FALSE)).

R is not installable in the build environment, so these are NOT R output: they are computed here with
scipy/numpy only (independent of oracle/ and of the CUDA code) following the semantics written down in
SURVEY.md 8a/8c, and they reproduce SURVEY.md 8c-4's provisional table.  Replace with real R output the
first time R is available.   Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402  (only for the fixture frames)


def linear_fc(a, b):
    fc = max(abs(a), abs(b)) / min(abs(a), abs(b)) if np.sign(a) == np.sign(b) else max(a, b) - min(a, b)
    return fc if b < a else -fc


def main():
    dfe, dfa, dfm = ge.analysis_fixture()
    samples = list(dfe.columns[1:])
    rows = {}
    order_out = {}
    for gi, gene in enumerate(dfe["sample"]):
        x = dfe.iloc[gi, 1:].to_numpy(float)
        perm = np.argsort(-x, kind="stable")
        xs = x[perm]
        q33, q66, q99 = np.quantile(x, [0.33, 0.66, 0.99])  # numpy 'linear' = R type 7
        fc = [linear_fc(q99, q66), linear_fc(q99, q33), linear_fc(q66, q33)]  # the quantile quirk (8g-5)
        up, mid, down = xs[0:2], xs[2:4], xs[4:6]
        pg = [stats.ttest_ind(a, b, equal_var=False).pvalue for a, b in ((up, mid), (up, down), (mid, down))]
        pi = list(dfa["UCSC_RefGene_Name"]).index(gene)
        cg = dfa["ID"][pi]
        y = dfm[dfm["sample"] == cg].iloc[0, 1:].to_numpy(float)[perm]
        yu, ym, yd = y[0:2], y[2:4], y[4:6]
        med = [np.median(yd), np.median(ym), np.median(yu)]
        bd = [med[2] - med[1], med[2] - med[0], med[1] - med[0]]
        pk = [stats.ks_2samp(a, b, method="exact").pvalue for a, b in ((yu, ym), (yu, yd), (ym, yd))]
        r, pr = stats.pearsonr(xs, y)
        name = f"{cg}_{dfa['UCSC_RefGene_Group'][pi]}_{gene}"
        rows[name] = [*med, *bd, *pk, float(r), float(pr), *fc, *pg]
        order_out[name] = [samples[i] for i in perm]
    out = {
        "call": "epimethex.analysis(Expressions, Annotations, Methylation, 1, 3, 2, TRUE, TRUE, FALSE)",
        "provenance": "scipy/numpy restatement, NOT R output (parity unpinned)",
        "columns": ["medianDown", "medianMedium", "medianUP", "bd_UPvsMID", "bd_UPvsDOWN", "bd_MIDvsDOWN",
                    "pvalue_UPvsMID", "pvalue_UPvsDOWN", "pvalue_MIDvsDOWN", "pearson_correlation",
                    "pvalue_pearson_correlation", "fc_UPvsMID(gene)", "fc_UPvsDOWN(gene)", "fc_MIDvsDOWN(gene)",
                    "pvalue_UPvsMID(gene)", "pvalue_UPvsDOWN(gene)", "pvalue_MIDvsDOWN(gene)"],
        "rows": {k: [float(v) for v in vals] for k, vals in rows.items()},
        "sample_order": order_out,
        "islands": {"cg11663302_Body_ARHGEF10L": "Island_chr1:18023481-18023792",
                    "cg01552731_1stExon_HIF3A": "N_Shore_chr19:46806998-46807617",
                    "cg09081385_TSS200_RNF10": "N_Shore_chr12:120972167-120972447"},
    }
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "readme_fixture.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["rows"], indent=1))


if __name__ == "__main__":
    main()
