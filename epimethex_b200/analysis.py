"""``epimethex_analysis`` — host-side mirror of the reference's only exported function,

    epimethex.analysis(dfExpressions, dfGpl, dfMethylation, minRangeGene, maxRangeGene,
                       numCores, dataGenesLinear, testExpression, testMethylation)

(R/EpimethexAnalysis.R:70-72), for the ``CG_data_individually.csv`` output (:351-471).  Same
argument names, meaning and order; same output directory ``"<min>-<max>"``, file name and column
names.  R is not available in this build environment, so this Python function plays the role of the
R wrapper in r_package/R/epimethex_analysis.R: it does the string work (annotation join, name
matching) and hands integers and doubles to libepx.so.  All statistics run on the GPU; there is no
CPU fallback.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import _native
from .annotation import GroupIndex, ProbeIndex, build_groups, build_probe_index

OUTPUT_COLUMNS = [
    "medianDown", "medianMedium", "medianUP", "bd_UPvsMID", "bd_UPvsDOWN", "bd_MIDvsDOWN",
    "pvalue_UPvsMID", "pvalue_UPvsDOWN", "pvalue_MIDvsDOWN", "pearson_correlation",
    "pvalue_pearson_correlation", "fc_UPvsMID(gene)", "fc_UPvsDOWN(gene)", "fc_MIDvsDOWN(gene)",
    "pvalue_UPvsMID(gene)", "pvalue_UPvsDOWN(gene)", "pvalue_MIDvsDOWN(gene)", "island",
]  # R/Functions.R:224-229 + the `island` row added at R/EpimethexAnalysis.R:463-467


@dataclass
class PreparedInputs:
    expr: np.ndarray          # [G, n] gene-major, original sample order
    gene_names: list
    sample_names: list
    meth: np.ndarray          # [P, n] probe-major, columns in expression sample order
    meth_ids: np.ndarray
    index: ProbeIndex


def prepare_expression(dfExpressions, minRangeGene: int, maxRangeGene: int):
    """R/EpimethexAnalysis.R:86 and :103-114: slice rows [min:max] (1-based, inclusive), drop the name
    column, drop genes whose non-NA values are all zero, drop samples with an NA in any kept gene."""
    df = dfExpressions.iloc[int(minRangeGene) - 1:int(maxRangeGene)]
    if "sample" in df.columns:
        gene_names = [str(v) for v in df["sample"].tolist()]
    else:
        gene_names = [str(v) for v in df.index.tolist()]
    values = df.iloc[:, 1:]                       # dfExpressions[,-1]
    sample_names = [str(c) for c in values.columns]
    x = values.to_numpy(dtype=np.float64)         # [G, n]
    with np.errstate(invalid="ignore"):
        is_zero_or_na = (x == 0) | np.isnan(x)
    keep_gene = ~is_zero_or_na.all(axis=1)        # which(!apply(df == 0, 2, all)) drops NA results too
    x = x[keep_gene]
    gene_names = [g for g, k in zip(gene_names, keep_gene) if k]
    keep_sample = ~np.isnan(x).any(axis=0)        # na.omit on the samples x genes frame
    x = np.ascontiguousarray(x[:, keep_sample])
    sample_names = [s for s, k in zip(sample_names, keep_sample) if k]
    return x, gene_names, sample_names


def prepare_methylation(dfMethylation, sample_names):
    """R/EpimethexAnalysis.R:278-282 and :299-301: na.omit drops probes with any NA; columns are matched to
    the expression samples BY NAME."""
    ids = dfMethylation.iloc[:, 0].astype(str).to_numpy(dtype=object)
    complete = ~dfMethylation.isna().any(axis=1).to_numpy()
    missing = [s for s in sample_names if s not in dfMethylation.columns]
    if missing:
        raise KeyError(f"undefined columns selected: methylation lacks samples {missing[:5]}")
    vals = dfMethylation.loc[:, sample_names].to_numpy(dtype=np.float64)
    return np.ascontiguousarray(vals[complete]), ids[complete]


def prepare(dfExpressions, dfGpl, dfMethylation, minRangeGene, maxRangeGene) -> PreparedInputs:
    expr, gene_names, sample_names = prepare_expression(dfExpressions, minRangeGene, maxRangeGene)
    if expr.shape[0] == 0:
        raise ValueError("no analysable gene in the requested range")
    meth, meth_ids = prepare_methylation(dfMethylation, sample_names)
    index = build_probe_index(dfGpl, gene_names, meth_ids)
    # :281-282 keeps only annotated probes; rows never referenced are simply not touched on the device
    return PreparedInputs(expr, gene_names, sample_names, meth, meth_ids, index)


def assemble_table(prep: PreparedInputs, res):
    """R/EpimethexAnalysis.R:425-469: 17 numeric rows + `island`, one column per pair named
    cg_position_gene, transposed; genes in sorted-name order, probes in dfCGunique order."""
    import pandas as pd

    idx = prep.index
    order = idx.output_order
    pair = res.pair.copy()
    na = res.pair_na
    for c in range(_native.PAIR_COLS):
        pair[(na >> c) & 1 == 1, c] = np.nan
    gene6 = res.gene[:, :6].copy()
    gna = res.gene_na
    for c in range(6):
        gene6[(gna >> c) & 1 == 1, c] = np.nan
    g = idx.pair_gene[order]
    table = np.concatenate([pair[order], gene6[g]], axis=1)
    names = [f"{cg}_{pos}_{prep.gene_names[gi]}" for cg, pos, gi in zip(idx.cg[order], idx.position[order], g)]
    df = pd.DataFrame(table, index=names, columns=OUTPUT_COLUMNS[:-1])
    df["island"] = idx.island[order]
    return df


GROUP_TABLES = {  # output file, grouping, column-name prefix rule
    "genes": ("CG_of_a_genes.csv", "gene"),          # R/EpimethexAnalysis.R:476-545
    "position": ("CG_by_position.csv", "position"),  # :549-582
    "islands": ("CG_by_islands.csv", "island"),      # :585-620
}


def assemble_group_table(prep: PreparedInputs, res, gi: GroupIndex, gres, by: str):
    """The pooled tables: 17 numeric rows per group, transposed (R/EpimethexAnalysis.R:543-545, :573-577). Names:
    ``CG~gene`` (:529) for the per-gene pool, ``<position>_<gene>`` / ``<island>_<gene>`` (R/Functions.R:181-182)."""
    import pandas as pd

    grp = gres.group.copy()
    for c in range(_native.PAIR_COLS):
        grp[(gres.group_na >> c) & 1 == 1, c] = np.nan
    gene6 = res.gene[:, :6].copy()
    for c in range(6):
        gene6[(res.gene_na >> c) & 1 == 1, c] = np.nan
    g = gi.gene.astype(np.int64)
    empty = (gi.group_ptr[1:] - gi.group_ptr[:-1]) == 0
    tail = gene6[g] if g.size else np.zeros((0, 6))
    tail[empty] = np.nan   # :536-541: the NA column of a gene without data carries no gene-level rows either
    table = np.concatenate([grp, tail], axis=1)
    if by == "gene":
        names = [f"CG~{prep.gene_names[gi_]}" for gi_ in g]
    else:
        names = [f"{k}_{prep.gene_names[gi_]}" for k, gi_ in zip(gi.key, g)]
    return pd.DataFrame(table, index=names, columns=OUTPUT_COLUMNS[:-1])


def write_csv_like_r(df, path: str, numeric: bool = False) -> None:
    """write.csv of the result table through the native writer (epx_write_csv, multithreaded).
    Default: a character matrix (the `island` row coerces every column, R/EpimethexAnalysis.R:463-470): numbers with
    15 significant digits as R's as.character renders them, NA as NA, everything quoted.  ``numeric=True``: a numeric
    matrix (the pooled tables), whose numbers write.csv leaves unquoted."""
    cols = list(df.columns)
    text = None
    import pandas as pd

    if cols and not pd.api.types.is_numeric_dtype(df[cols[-1]]):
        text = [None if (v is None or v is pd.NA or (isinstance(v, float) and np.isnan(v))) else str(v)
                for v in df[cols[-1]].tolist()]
        num_cols = cols[:-1]
    else:
        num_cols = cols
    values = df[num_cols].to_numpy(dtype=np.float64) if num_cols else np.zeros((len(df), 0))
    _native.write_csv(path, [str(v) for v in df.index], num_cols, values, text_col=text,
                      text_col_name=cols[-1] if text is not None else "", quote_numbers=not numeric)


def epimethex_analysis(dfExpressions, dfGpl, dfMethylation, minRangeGene, maxRangeGene, numCores,
                       dataGenesLinear, testExpression, testMethylation, *, device: int = 0, session=None,
                       output_dir: str | None = None, write_csv: bool = True, compat_fc: str = "reference",
                       strict: bool = False, tables=("individually", "genes", "position", "islands"),
                       return_all: bool = False):
    """Drop-in for the reference call; returns the ``CG_data_individually`` table as a DataFrame and
    (by default) writes ``<output_dir>/<min>-<max>/CG_data_individually.csv``.

    ``numCores`` is accepted for signature compatibility (it sized the reference's PSOCK cluster,
    R/EpimethexAnalysis.R:177,333); the GPU path does not use it.
    ``compat_fc="reference"`` reproduces the reference's gene fold changes (computed from the
    0.33/0.66/0.99 quantiles); ``"means"`` uses the group means.
    ``strict=True`` raises for sample counts that are not a multiple of 3, like the reference.
    ``tables``: which of the reference's four outputs to compute: "individually" (always), "genes"
    (``CG_of_a_genes.csv``), "position" (``CG_by_position.csv``), "islands" (``CG_by_islands.csv``).
    ``return_all=True`` returns a dict of the computed tables instead of the first one.
    """
    del numCores
    prep = prepare(dfExpressions, dfGpl, dfMethylation, minRangeGene, maxRangeGene)
    opts = _native.make_options(bool(dataGenesLinear), bool(testExpression), bool(testMethylation),
                                fc_from_means=(compat_fc == "means"), strict_n3=strict)
    own = session is None
    sess = session if session is not None else _native.Session(device)
    try:
        sess.set_methylation(prep.meth)
        res = sess.analyze(prep.expr, prep.index.row_ptr, prep.index.probe_idx, opts)
        pooled = {}
        for t in tables:
            if t == "individually":
                continue
            if t not in GROUP_TABLES:
                raise ValueError(f"unknown table {t!r}")
            gi = build_groups(prep.index, GROUP_TABLES[t][1])
            pooled[t] = (gi, sess.analyze_groups(gi.group_ptr, gi.group_pairs))
    finally:
        if own:
            sess.close()
    df = assemble_table(prep, res)
    df.attrs["timing"] = res.timing
    df.attrs["counts"] = res.counts
    if write_csv:
        name_fd = f"{int(minRangeGene)}-{int(maxRangeGene)}"
        out = os.path.join(output_dir or os.getcwd(), name_fd)
        os.makedirs(out, exist_ok=True)
        write_csv_like_r(df, os.path.join(out, "CG_data_individually.csv"))
    out_tables = {"individually": df}
    for t, (gi, gres) in pooled.items():
        fname, by = GROUP_TABLES[t]
        if gi.n_groups == 0:
            import warnings
            warnings.warn(f"{fname} is not generated")     # R/EpimethexAnalysis.R:580, :618
            continue
        dft = assemble_group_table(prep, res, gi, gres, by)
        out_tables[t] = dft
        if write_csv:
            write_csv_like_r(dft, os.path.join(out, fname), numeric=True)
    return out_tables if return_all else df
