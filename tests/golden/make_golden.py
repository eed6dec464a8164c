"""Generates the golden tables of tests/golden/:

  readme_fixture.json   the reference's only executable example (R/EpimethexAnalysis.R:28-63; README.md:11-69):
                        epimethex.analysis(Expressions, Annotations, Methylation, 1, 3, 2, TRUE, TRUE, FALSE)
  golden_<case>.npz     40-gene synthetic slices through tests/golden/rpipeline.py, the independent second restatement
                        of the path (numpy / fractions / big integers / mpmath, written from the R sources; imports
                        neither oracle/ nor libepx): exact KS (n = 99), asymptotic KS (n = 474), the n mod 3 extension
                        (n = 473), 3-decimal rounded betas (ties: exact falls back to asymptotic), both methylation tests,
                        both expression tests, and the pooled CG_of_a_genes rows; four more at 600 / 1,000 / 1,500 / 1,501
                        samples, where the CUDA path runs other kernels (the statistics are the same).

R is not installable in the build environment, so NONE of this is R output (parity unpinned); it is the strongest pin
available without R: tests/test_golden_cpu.py holds the oracle to these tables and tests/test_gpu_golden.py the CUDA
path.  Inputs are not stored: they are regenerated bit-exactly by epimethex_b200.synth (counter-based, libm-free) and
checked against the SHA-256 kept in each file.      Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import rpipeline  # noqa: E402

CASES = {
    # name: genes, samples, probes, decimals
    "n99_exact": dict(G=40, n=99, P=170, decimals=None),
    "n474_asymptotic": dict(G=40, n=474, P=170, decimals=None),
    "n473_extension": dict(G=40, n=473, P=170, decimals=None),
    "n99_ties3": dict(G=40, n=99, P=170, decimals=3),
    "n474_ties3": dict(G=40, n=474, P=170, decimals=3),
    "n30_tiny_ties2": dict(G=24, n=30, P=90, decimals=2),
    # longer rows (the CUDA path changes kernels at 513 and 1025 samples; the statistics do not change with n)
    "n600_ties2": dict(G=24, n=600, P=100, decimals=2),          # two-halves rank kernel, long tie runs
    "n1000_epic": dict(G=30, n=1000, P=130, decimals=None),      # BASELINE configs[3] sample count, n mod 3 = 1
    "n1500_long": dict(G=24, n=1500, P=100, decimals=None),      # CTA-per-row kernels
    "n1501_long_ties3": dict(G=24, n=1501, P=100, decimals=3),   # the same with ties and the n mod 3 extension
}


def case_inputs(name):
    """the inputs of a golden case, regenerated from the counter-based generator (the data only: no statistics)"""
    from epimethex_b200 import synth
    c = CASES[name]
    ann = synth.annotation(c["P"], c["G"])
    x = synth.expression(c["G"], c["n"])
    y = synth.methylation(c["P"], c["n"], ann, decimals=c["decimals"], threads=1)
    has = np.ones(c["P"], bool)
    has[::7] = False                                       # annotated probes without a data row -> all-NA pair rows
    idx = synth.probe_index(ann, 0, c["G"], probe_has_data=has)
    used = np.unique(idx.probe_idx[idx.probe_idx >= 0])
    y[used[0]] = 0.25                                      # constant probe: guard NA x 3, sd 0 -> Pearson NA
    y[used[1]] = np.resize([0.125, 0.25, 0.5], c["n"])     # few distinct values: long tie runs, possible constant differences
    j = int(np.nonzero(idx.probe_idx == used[2])[0][0])    # UP - Medium exactly constant for this pair's gene: guard hits ONE test
    perm = np.argsort(-x[idx.pair_gene[j]], kind="stable")
    m = c["n"] // 3
    v = np.empty(c["n"])
    v[perm] = np.concatenate([np.arange(m) / 1024.0 + 0.5, np.arange(m) / 1024.0 + 0.25, np.linspace(0.05, 0.95, c["n"] - 2 * m)])
    y[used[2]] = v
    x = x.copy()
    x[c["G"] - 1] = 3.5                                    # a constant gene: Pearson NA, expression tests NA
    return x, y, np.asarray(idx.row_ptr, np.int64), np.asarray(idx.probe_idx, np.int32)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def readme_fixture():
    import __graft_entry__ as ge
    dfe, dfa, dfm = ge.analysis_fixture()
    samples = list(dfe.columns[1:])
    genes = list(dfe["sample"])
    x = dfe.iloc[:, 1:].to_numpy(float)
    y = dfm.iloc[:, 1:].to_numpy(float)
    # dfCGunique of the example: one probe per gene, in gene-name order; probe rows by id
    row_ptr, probe_idx, names = [0], [], []
    for gene in genes:
        k = list(dfa["UCSC_RefGene_Name"]).index(gene)
        probe_idx.append(list(dfm["sample"]).index(dfa["ID"][k]))
        row_ptr.append(len(probe_idx))
        names.append(f"{dfa['ID'][k]}_{dfa['UCSC_RefGene_Group'][k]}_{gene}")
    res = rpipeline.run(x.tolist(), y.tolist(), row_ptr, probe_idx, data_linear=True, test_expr_t=True, test_meth_t=False)
    rows = {nm: [float(v) for v in res["pair"][j] + res["gene"][j][:6]] for j, nm in enumerate(names)}
    out = {
        "call": "epimethex.analysis(Expressions, Annotations, Methylation, 1, 3, 2, TRUE, TRUE, FALSE)",
        "provenance": "tests/golden/rpipeline.py (independent numpy/mpmath restatement), NOT R output (parity unpinned)",
        "columns": ["medianDown", "medianMedium", "medianUP", "bd_UPvsMID", "bd_UPvsDOWN", "bd_MIDvsDOWN",
                    "pvalue_UPvsMID", "pvalue_UPvsDOWN", "pvalue_MIDvsDOWN", "pearson_correlation",
                    "pvalue_pearson_correlation", "fc_UPvsMID(gene)", "fc_UPvsDOWN(gene)", "fc_MIDvsDOWN(gene)",
                    "pvalue_UPvsMID(gene)", "pvalue_UPvsDOWN(gene)", "pvalue_MIDvsDOWN(gene)"],
        "rows": rows,
        "sample_order": {nm: [samples[i] for i in res["perm"][j]] for j, nm in enumerate(names)},
        "islands": {"cg11663302_Body_ARHGEF10L": "Island_chr1:18023481-18023792",
                    "cg01552731_1stExon_HIF3A": "N_Shore_chr19:46806998-46807617",
                    "cg09081385_TSS200_RNF10": "N_Shore_chr12:120972167-120972447"},
    }
    with open(os.path.join(HERE, "readme_fixture.json"), "w") as f:
        json.dump(out, f, indent=1)
    return out


def golden_case(name):
    x, y, row_ptr, probe_idx = case_inputs(name)
    xl, yl, rp, pi = x.tolist(), y.tolist(), row_ptr.tolist(), probe_idx.tolist()
    out = dict(input_sha256=np.array(digest(x, y, row_ptr, probe_idx)), row_ptr=row_ptr, probe_idx=probe_idx,
               params=np.array(json.dumps(CASES[name])))
    for tm in (False, True):                      # methylation side: the pair table and the pooled rows
        r = rpipeline.run(xl, yl, rp, pi, test_expr_t=True, test_meth_t=tm, pooled_by_gene=True)
        k = "t" if tm else "ks"
        out[f"pair_{k}"] = np.array(r["pair"], float)
        out[f"pair_stat_{k}"] = np.array(r["pair_stat"], float)
        out[f"pair_dnum_{k}"] = np.array(r["ks_dnum"], np.int64)
        out[f"group_{k}"] = np.array(r["group"], float)
        out[f"group_dnum_{k}"] = np.array(r["group_dnum"], np.int64)
        if not tm:
            out["perm"] = np.array(r["perm"], np.int32)
            out["counts"] = np.array(r["counts"], np.int32)
            out["ref_gene"] = np.array(r["ref_gene"], np.int32)
            out["gene_t"] = np.array(r["gene"], float)
            out["gene_dnum_t"] = np.array(r["gene_ks_dnum"], np.int64)
    r = rpipeline.run(xl, yl, [0] * (len(xl) + 1), [], test_expr_t=False)     # expression side in KS mode (no pairs needed)
    out["gene_ks"] = np.array(r["gene"], float)
    out["gene_dnum_ks"] = np.array(r["gene_ks_dnum"], np.int64)
    for fcm, lin, key in ((True, True, "gene_fc_means_linear"), (False, False, "gene_fc_quantiles_log")):
        r = rpipeline.run(xl, yl, [0] * (len(xl) + 1), [], data_linear=lin, fc_from_means=fcm)
        out[key] = np.array(r["gene"], float)[:, :3]
    np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **out)
    return out


def main():
    rf = readme_fixture()
    print("readme fixture:", json.dumps(rf["rows"], indent=1)[:300], "...")
    for name in (sys.argv[1:] or CASES):
        o = golden_case(name)
        print(name, "pairs", o["pair_ks"].shape[0], "exact-range" if CASES[name]["n"] // 3 < 100 else "asymptotic",
              "sha", str(o["input_sha256"])[:12])


if __name__ == "__main__":
    main()
