"""Host-side logic (annotation join -> CSR, input filters, table assembly) and the C-ABI surface. CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pandas as pd
import pytest

from epimethex_b200 import _native, analysis, annotation, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ann(rows):
    cols = ["ID", "Relation_to_UCSC_CpG_Island", "UCSC_CpG_Islands_Name", "UCSC_RefGene_Accession", "Chromosome_36",
            "Coordinate_36", "UCSC_RefGene_Name", "UCSC_RefGene_Group"]
    return pd.DataFrame(rows, columns=cols)


def test_library_exports_every_declared_symbol():
    """every function include/epx.h declares is exported by libepx.so (loading only; no compute without a GPU)"""
    header = open(os.path.join(ROOT, "include", "epx.h")).read()
    declared = set(re.findall(r"\b(epx_[a-z_0-9]+)\s*\(", header)) - {"epx_session", "epx_options", "epx_result", "epx_timing"}
    assert declared == set(_native.EXPORTS)
    L = C.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _native.lib().epx_abi_version() == 1


def test_no_gpu_is_an_error_not_a_fallback():
    try:
        n = _native.device_count()
    except _native.EpxError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_native.EpxError):
        _native.Session(0)


def test_one_shot_calls_refuse_null_handles_before_touching_a_device():
    """epx_analyze / epx_analyze_host / epx_multi_analyze(_host): argument checks come first, with a message, on any box"""
    L = _native.lib()
    for name, nargs in (("epx_analyze", 8), ("epx_analyze_host", 11), ("epx_multi_analyze", 8), ("epx_multi_analyze_host", 11)):
        b = C.create_string_buffer(256)
        args = [None if t is C.c_void_p or hasattr(t, "contents") else 0 for t in getattr(L, name).argtypes[:nargs]]
        assert getattr(L, name)(*args, b, len(b)) == -1, name
        assert b.value, f"{name}: no message"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "epimethex_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "epx_oracle" not in src and "from oracle" not in src, f


def test_explode_sort_dedup_semantics():
    """R/EpimethexAnalysis.R:248-257, :287-291 incl. SURVEY 8g-6 (dedup on (cg, position) ignores the gene)."""
    dfa = _ann([
        ["cg1", "Island", "chr1:1-2", "NM_1;NM_2", "1", "300", "B;A", "Body;Body"],      # shared probe, same group
        ["cg2", "N_Shore", "chr1:5-9", "NM_3;NM_4", "1", "100", "A;A", "TSS200;Body"],   # two groups in one gene
        ["cg3", "", "", "NM_5", "1", "1000", "A", "Body"],                               # "1000" < "300" as strings
        ["cg4", "Island", "chr2:1-2", "NM_6", "2", "50", "", "Body"],                    # no gene -> dropped
        ["cg5", "S_Shelf", "chr3:1-2", "NM_7", "3", "70", "C", "3'UTR"],                 # gene not analysed
        ["cg6", "Island", "chr4:1-2", "NM_8", "4", "10", "B", "1stExon"],                # annotated, no data row
    ])
    idx = annotation.build_probe_index(dfa, ["B", "A", "Z"], ["cg1", "cg2", "cg3", "cg5"])
    # reference order: gene A (sorted names) by coordinate STRING: "100","100","1000","300"; then gene B: "10","300"
    ref_rows = [(idx.cg[i], idx.position[i], ["B", "A", "Z"][idx.pair_gene[i]]) for i in idx.output_order]
    assert ref_rows == [("cg2", "TSS200", "A"), ("cg2", "Body", "A"), ("cg3", "Body", "A"), ("cg1", "Body", "A"),
                        ("cg6", "1stExon", "B")]          # cg1/Body survives only under A (alphabetically first)
    assert idx.row_ptr.tolist() == [0, 1, 5, 5]           # CSR is in EXPRESSION order: B, A, Z
    assert idx.probe_idx.tolist() == [-1, 1, 1, 2, 0]     # cg6 has no methylation row
    assert idx.island[idx.output_order].tolist()[:3] == ["N_Shore_chr1:5-9", "N_Shore_chr1:5-9", "_"]


def test_strsplit_trailing_separator_and_frame_wide_pairing():
    """base::strsplit drops ONE trailing empty field ("A;B;" has two elements) and data.frame(gene = unlist(s),
    position = unlist(s2)) pairs entries by their place in the whole frame (R/EpimethexAnalysis.R:248-257)."""
    assert annotation._strsplit("A;B;") == ["A", "B"] and annotation._strsplit(";A") == ["", "A"]
    assert annotation._strsplit("A;;B") == ["A", "", "B"] and annotation._strsplit("") == [] and annotation._strsplit("A") == ["A"]
    dfa = _ann([["cg1", "Island", "x", "NM;NM", "1", "5", "A;B;", "Body;TSS200;"],       # trailing ';' on both: 2 entries
                ["cg2", "Island", "x", "NM", "1", "7", "A;", "Body"]])                    # on one only: still 1 entry
    idx = annotation.build_probe_index(dfa, ["A", "B"], ["cg1", "cg2"])
    rows = [(idx.cg[i], idx.position[i], ["A", "B"][idx.pair_gene[i]]) for i in idx.output_order]
    assert rows == [("cg1", "Body", "A"), ("cg2", "Body", "A"), ("cg1", "TSS200", "B")]
    # per-probe counts differ but the totals agree: R pairs them across probes without complaint
    dfa = _ann([["cg1", "Island", "x", "NM", "1", "5", "A;B", "Body"], ["cg2", "Island", "x", "NM", "1", "7", "A", "TSS200;3'UTR"]])
    cg, _, gene, pos, _ = annotation.explode_annotation(dfa)
    assert list(zip(cg, gene, pos)) == [("cg1", "A", "Body"), ("cg1", "B", "TSS200"), ("cg2", "A", "3'UTR")]
    with pytest.raises(ValueError, match="differing number of rows"):
        annotation.explode_annotation(_ann([["cg1", "Island", "x", "NM", "1", "5", "A;B", "Body"]]))


def test_numeric_coordinates_sort_numerically():
    dfa = _ann([["cg1", "Island", "x", "NM", "1", 300, "A", "Body"], ["cg2", "Island", "x", "NM", "1", 1000, "A", "Body"]])
    idx = annotation.build_probe_index(dfa, ["A"], ["cg1", "cg2"])
    assert idx.cg.tolist() == ["cg1", "cg2"]
    dfa["Coordinate_36"] = dfa["Coordinate_36"].astype(str)
    assert annotation.build_probe_index(dfa, ["A"], ["cg1", "cg2"]).cg.tolist() == ["cg2", "cg1"]


def test_expression_filters():
    """R/EpimethexAnalysis.R:86, :103-114"""
    df = pd.DataFrame({"sample": ["g1", "g2", "g3", "g4", "g5"],
                       "s1": [1.0, 0.0, 2.0, 0.0, 9.0], "s2": [2.0, 0.0, np.nan, np.nan, 9.0],
                       "s3": [3.0, 0.0, 4.0, 0.0, 9.0], "s4": [4.0, 0.0, 5.0, 0.0, 9.0]})
    x, genes, samples = analysis.prepare_expression(df, 1, 4)      # g5 is outside the range
    assert genes == ["g1", "g3"]                                    # g2 all zero; g4 zero-or-NA -> dropped
    assert samples == ["s1", "s3", "s4"]                            # s2 has an NA in a kept gene
    assert x.tolist() == [[1.0, 3.0, 4.0], [2.0, 4.0, 5.0]]


def test_methylation_matched_by_sample_name_and_na_probes_dropped():
    dfm = pd.DataFrame({"sample": ["cg1", "cg2", "cg3"], "s3": [0.3, 0.6, 0.9], "s1": [0.1, np.nan, 0.7], "s2": [0.2, 0.5, 0.8]})
    m, ids = analysis.prepare_methylation(dfm, ["s1", "s3"])
    assert ids.tolist() == ["cg1", "cg3"] and m.tolist() == [[0.1, 0.3], [0.7, 0.9]]
    with pytest.raises(KeyError):
        analysis.prepare_methylation(dfm, ["s1", "s9"])


def test_synthetic_index_agrees_with_string_path():
    dfe, dfa, dfm, ann = synth.frames(50, 12, 400)
    prep = analysis.prepare(dfe, dfa, dfm, 1, 50)
    assert len(prep.gene_names) == 50
    idx = synth.probe_index(ann, 0, 50)
    for f in ("row_ptr", "probe_idx", "pair_gene", "output_order", "cg", "position", "island"):
        assert np.array_equal(getattr(prep.index, f), getattr(idx, f)), f
    sub = analysis.prepare(dfe, dfa, dfm, 11, 30)                   # a gene range, as the README recommends
    idx2 = synth.probe_index(ann, 10, 30)
    assert np.array_equal(sub.index.row_ptr, idx2.row_ptr) and np.array_equal(sub.index.probe_idx, idx2.probe_idx)


def test_synth_is_counter_based():
    ann = synth.annotation(300, 20)
    full = synth.methylation(300, 17, ann, threads=2, chunk=64)
    part = synth.methylation_rows([5, 250, 7], 17, ann)
    assert np.array_equal(full[[5, 250, 7]], part)
    assert np.array_equal(synth.expression(20, 17)[3:9], synth.expression(6, 17, g0=3))
    assert (synth.expression(2000, 30) == 0).mean() > 0.01          # genuine zero ties exist


def test_assemble_and_csv(tmp_path, readme_frames):
    """table assembly (R/EpimethexAnalysis.R:425-470) from an engine result (here: canned numbers)"""
    dfe, dfa, dfm = readme_frames
    prep = analysis.prepare(dfe, dfa, dfm, 1, 3)
    res = _native.Result(pair=np.arange(33, dtype=float).reshape(3, 11) / 7, pair_na=np.array([0, 1 << 6, 0], np.uint16),
                         pair_stat=np.zeros((3, 3)), ks_dnum=np.zeros((3, 3), np.int32), gene=np.ones((3, 15)) * 0.5,
                         gene_na=np.array([0, 0, 1 << 3], np.uint16), gene_ks_dnum=np.zeros((3, 3), np.int32),
                         perm=np.zeros((3, 6), np.uint16))
    df = analysis.assemble_table(prep, res)
    assert list(df.columns) == analysis.OUTPUT_COLUMNS
    assert list(df.index) == ["cg11663302_Body_ARHGEF10L", "cg01552731_1stExon_HIF3A", "cg09081385_TSS200_RNF10"]
    assert np.isnan(df.iloc[1, 6]) and np.isnan(df.iloc[2, 14]) and df.iloc[0, 17] == "Island_chr1:18023481-18023792"
    p = tmp_path / "t.csv"
    analysis.write_csv_like_r(df, str(p))
    lines = p.read_text().splitlines()
    assert lines[0].startswith('"","medianDown","medianMedium"') and lines[0].endswith('"island"')
    assert lines[1].split(",")[2] == '"0.142857142857143"'          # 15 significant digits, quoted
    assert "NA" in lines[2].split(",")
    analysis.write_csv_like_r(df.iloc[:, :17], str(p), numeric=True)    # the pooled tables: a numeric matrix
    lines = p.read_text().splitlines()
    assert lines[1].split(",")[2] == "0.142857142857143" and lines[0].count(",") == 17


@pytest.mark.parametrize("x,want", [
    (0.1, "0.1"), (1e-4, "1e-04"), (100000.0, "1e+05"), (123456.0, "123456"), (0.00012345, "0.00012345"),
    (1 / 3, "0.333333333333333"), (2 / 3, "0.666666666666667"), (1e15, "1e+15"), (123456789012.0, "123456789012"),
    (-0.5, "-0.5"), (1e-300, "1e-300"), (0.9841, "0.9841"), (1234567.1, "1234567.1"), (100001.0, "100001"),
    (3.0, "3"), (1.5e-10, "1.5e-10"), (0.1 + 0.2, "0.3"), (1e22, "1e+22"), (0.0, "0"), (-1e-5, "-1e-05"),
    (float("inf"), "Inf"), (5.066e-17, "5.066e-17"), (0.05, "0.05"), (0.000123456789012345, "0.000123456789012345")])
def test_numbers_are_rendered_like_r(x, want):
    """R's formatReal with digits = 15 on one value (as.character / write.csv): fewest significant digits, fixed
    notation unless scientific is strictly narrower.  Expected strings are what R prints for these literals."""
    assert _native.format_real15(x) == want


def test_native_csv_writer_quotes_and_escapes(tmp_path):
    p = tmp_path / "q.csv"
    _native.write_csv(str(p), ['a"b', "c"], ["x", "y"], np.array([[1.0, np.nan], [2.5, -1e-7]]),
                      text_col=["i,1", None], text_col_name="island", quote_numbers=True, threads=3)
    assert p.read_text() == '"","x","y","island"\n"a""b","1",NA,"i,1"\n"c","2.5","-1e-07",NA\n'


# ---------------------------------------------------------------- pooled probe groups (host logic + oracle)

def _brute_groups(idx, by):
    import re
    po = idx.output_order
    genes = []
    for p in po:
        if idx.pair_gene[p] not in genes:
            genes.append(idx.pair_gene[p])
    out = []
    if by == "gene":
        return [(g, None, [p for p in po if idx.pair_gene[p] == g and idx.probe_idx[p] >= 0]) for g in genes]
    vals = idx.position if by == "position" else idx.island
    keys = []
    for p in po:
        if vals[p] not in keys:
            keys.append(vals[p])
    if by == "island":
        keys = [k for k in keys if not re.search("NA_NA|^_", k)]
    for g in genes:
        pg = [p for p in po if idx.pair_gene[p] == g]
        for k in keys:
            members = [p for p in pg if vals[p] == k and idx.probe_idx[p] >= 0]
            if len(members) > 1:                       # R/Functions.R:153
                out.append((g, k, members))
    return out


@pytest.mark.parametrize("by", ["gene", "position", "island"])
def test_build_groups_follows_the_reference_loops(by):
    """R/EpimethexAnalysis.R:476-620 / R/Functions.R:136-188: which probes are pooled, and in which column order."""
    from epimethex_b200 import synth
    from epimethex_b200.annotation import build_groups
    ann = synth.annotation(1500, 90)
    has = np.ones(1500, bool); has[::7] = False
    idx = synth.probe_index(ann, 0, 90, has)
    gi = build_groups(idx, by)
    want = _brute_groups(idx, by)
    assert gi.n_groups == len(want) > 0
    for q, (g, k, members) in enumerate(want):
        assert gi.gene[q] == g and gi.key[q] == k
        assert gi.group_pairs[gi.group_ptr[q]:gi.group_ptr[q + 1]].tolist() == members


def test_oracle_groups_of_one_probe_equal_the_pair_table():
    import oracle
    rng = np.random.default_rng(0)
    G, n, P = 12, 30, 40
    expr = rng.normal(5, 2, (G, n)); meth = np.round(rng.random((P, n)), 2)
    row_ptr = np.arange(0, G * 5 + 1, 5, dtype=np.int64)
    probe_idx = rng.integers(0, P, G * 5).astype(np.int32)
    for tm in (False, True):
        res = oracle.analyze(expr, meth, row_ptr, probe_idx, test_meth_t=tm)
        one = oracle.analyze_groups(expr, meth, row_ptr, probe_idx, np.arange(G * 5 + 1, dtype=np.int64),
                                    np.arange(G * 5, dtype=np.int32), test_meth_t=tm)
        assert np.array_equal(one.group, res.pair, equal_nan=True)
        assert np.array_equal(one.group_na, res.pair_na)
        assert np.array_equal(one.group_dnum, res.ks_dnum)


def test_oracle_pooled_group_by_hand():
    """two probes pooled: medians, KS numerator and Pearson against a direct numpy/scipy computation"""
    import oracle
    from scipy import stats
    rng = np.random.default_rng(3)
    n, m = 60, 20
    expr = rng.normal(5, 2, (1, n)); meth = rng.random((2, n))
    row_ptr = np.array([0, 2], np.int64); probe_idx = np.array([0, 1], np.int32)
    r = oracle.analyze_groups(expr, meth, row_ptr, probe_idx, row_ptr, probe_idx)
    order = np.argsort(-expr[0], kind="stable")
    y = meth[:, order]
    up, mid, down = y[:, :m].ravel(), y[:, m:2 * m].ravel(), y[:, 2 * m:].ravel()
    assert r.group[0, 0] == np.median(down) and r.group[0, 1] == np.median(mid) and r.group[0, 2] == np.median(up)
    d = stats.ks_2samp(up, mid, method="asymp").statistic
    assert r.group_dnum[0, 0] == round(d * 40 * 40)
    pr = stats.pearsonr(np.tile(expr[0][order], 2), y.ravel())
    np.testing.assert_allclose(r.group[0, 9:11], [pr.statistic, pr.pvalue], rtol=1e-9)
    with pytest.raises(ValueError):      # a group may not mix genes or name a pair without data
        oracle.analyze_groups(np.vstack([expr, expr]), meth, np.array([0, 1, 2], np.int64), probe_idx,
                              np.array([0, 2], np.int64), np.array([0, 1], np.int32))
