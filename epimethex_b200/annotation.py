"""Annotation join: Illumina probe annotation -> CSR gene->probe index (``dfCGunique``).

Host-side restatement of R/EpimethexAnalysis.R:93-100 (drop probes without a gene, build the
``island`` string), :248-257 (explode the ';'-separated RefGene name/accession/group lists),
:264-265 (keep analysed genes), :287-288 (``plyr::arrange(gene, Coordinate_36)``) and :291
(``!duplicated(dfAnnotations[, c(1, 6)])`` -- de-duplication on (cg, position) only, NOT on gene:
SURVEY 8g-6, reproduced on purpose).  Strings in, integers out: the device only ever sees the CSR.

Collation: R orders the ``gene`` / ``Coordinate_36`` factors by the session locale. This module
orders strings by Unicode code point (R's ``C`` locale); the R shim keeps R's own ordering because it
builds the CSR with R's ``order()``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

REQUIRED_COLUMNS = [
    "ID", "Relation_to_UCSC_CpG_Island", "UCSC_CpG_Islands_Name", "UCSC_RefGene_Accession",
    "Chromosome_36", "Coordinate_36", "UCSC_RefGene_Name", "UCSC_RefGene_Group",
]


@dataclass
class ProbeIndex:
    """``dfCGunique`` as integers. Pairs are stored gene-major in EXPRESSION order of the genes;
    ``output_order`` lists the pairs in the reference's output order (genes by sorted name)."""
    row_ptr: np.ndarray        # int64 [G+1]
    probe_idx: np.ndarray      # int32 [N_p] methylation row or -1 (annotated, no data row)
    pair_gene: np.ndarray      # int32 [N_p] gene (expression order)
    cg: np.ndarray             # object [N_p]
    position: np.ndarray       # object [N_p] RefGene group
    island: np.ndarray         # object [N_p]
    output_order: np.ndarray   # int64 [N_p]

    @property
    def n_pairs(self) -> int:
        return int(self.row_ptr[-1])


def _is_missing(v) -> bool:
    return v is None or (isinstance(v, float) and np.isnan(v))


def _paste(a, b) -> str:
    # R's paste() renders NA as the string "NA" (-> the "NA_NA" islands filtered at :321)
    sa = "NA" if _is_missing(a) else str(a)
    sb = "NA" if _is_missing(b) else str(b)
    return sa + "_" + sb


def _strsplit(s: str) -> list:
    """base::strsplit(s, ";") for one string: "" gives no element at all, and ONE trailing empty field is dropped
    ("A;B;" -> A, B) while leading and inner empty fields stay (";A" -> "", A)."""
    if s == "":
        return []
    parts = s.split(";")
    if parts[-1] == "":
        parts.pop()
    return parts


def explode_annotation(dfGpl):
    """R/EpimethexAnalysis.R:93-100 and :248-257. Returns parallel lists (cg, coord, gene, position, island)
    in exploded order."""
    import pandas as pd

    missing = [c for c in REQUIRED_COLUMNS if c not in dfGpl.columns]
    if missing:
        raise KeyError(f"annotation is missing columns {missing}")
    names = dfGpl["UCSC_RefGene_Name"]
    keep = ~(names.isna() | (names.astype(str) == ""))
    df = dfGpl.loc[keep]
    cg_out, coord_out, gene_out, pos_out, island_out = [], [], [], [], []
    ids = df["ID"].to_numpy(dtype=object)
    rel = df["Relation_to_UCSC_CpG_Island"].to_numpy(dtype=object)
    isl = df["UCSC_CpG_Islands_Name"].to_numpy(dtype=object)
    coord = df["Coordinate_36"].to_numpy(dtype=object)
    gname = df["UCSC_RefGene_Name"].astype(str).to_numpy(dtype=object)
    ggroup = df["UCSC_RefGene_Group"].astype(str).to_numpy(dtype=object)
    # data.frame(cg = rep(ID, lengths(s)), gene = unlist(s), position = unlist(s2), ...): the gene and position columns
    # are the two flattened lists laid side by side, so entries pair up by their place in the WHOLE frame, not per probe
    for i in range(len(df)):
        genes = _strsplit(gname[i])
        island = _paste(rel[i], isl[i])
        for g in genes:
            cg_out.append(str(ids[i])); coord_out.append(coord[i]); gene_out.append(g); island_out.append(island)
        pos_out.extend(_strsplit(ggroup[i]))
    if len(pos_out) != len(gene_out):
        raise ValueError(f"arguments imply differing number of rows: {len(gene_out)}, {len(pos_out)} "
                         "(UCSC_RefGene_Name and UCSC_RefGene_Group list different numbers of entries)")
    return cg_out, coord_out, gene_out, pos_out, island_out


def _rank_strings(values) -> np.ndarray:
    """dense ranks of strings in code-point order"""
    arr = np.asarray(values, dtype=object)
    uniq = sorted(set(arr.tolist()))
    lut = {v: i for i, v in enumerate(uniq)}
    return np.fromiter((lut[v] for v in arr), dtype=np.int64, count=arr.size)


def build_index_core(cg_id, gene_rank, coord_key, position_id, gene_of_rank, probe_row_of_cg, n_genes):
    """Integer core shared by the string path and the synthetic generator.

    cg_id[i], position_id[i]: integer codes of probe and RefGene group of exploded entry i;
    gene_rank[i]: rank of the entry's gene in sorted-name order; coord_key[i]: sortable key of
    Coordinate_36; gene_of_rank[r]: expression-order index of the gene with name rank r;
    probe_row_of_cg[c]: methylation row of probe code c, or -1.
    Returns (row_ptr, probe_idx, pair_gene, kept_entry_indices_in_csr_order, output_order).
    """
    cg_id = np.asarray(cg_id, np.int64); gene_rank = np.asarray(gene_rank, np.int64)
    coord_key = np.asarray(coord_key); position_id = np.asarray(position_id, np.int64)
    n_entries = cg_id.size
    # plyr::arrange(gene, Coordinate_36): order() is stable -> ties keep the exploded order
    order = np.lexsort((np.arange(n_entries), coord_key, gene_rank))
    # !duplicated(cbind(cg, position)): first occurrence in the sorted table survives
    key = cg_id[order] * (int(position_id.max()) + 1 if n_entries else 1) + position_id[order]
    _, first = np.unique(key, return_index=True)
    keep_sorted = np.zeros(n_entries, bool)
    keep_sorted[first] = True
    kept = order[keep_sorted]                      # dfCGunique rows, reference order (gene name, coord)
    g_expr = np.asarray(gene_of_rank, np.int64)[gene_rank[kept]]
    # device layout: gene-major in expression order, dfCGunique order inside a gene
    csr_order = np.argsort(g_expr, kind="stable")
    kept_csr = kept[csr_order]
    counts = np.bincount(g_expr, minlength=n_genes)
    row_ptr = np.zeros(n_genes + 1, np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    probe_idx = np.asarray(probe_row_of_cg, np.int64)[cg_id[kept_csr]].astype(np.int32)
    pair_gene = g_expr[csr_order].astype(np.int32)
    output_order = np.empty(kept.size, np.int64)
    output_order[:] = np.argsort(csr_order, kind="stable")  # reference row r -> CSR pair index
    return row_ptr, probe_idx, pair_gene, kept_csr, output_order


def build_probe_index(dfGpl, gene_names, methylation_ids, coord_numeric=None) -> ProbeIndex:
    """gene_names: analysed genes in expression order (first occurrence wins for duplicates);
    methylation_ids: probe id of each methylation row kept after na.omit."""
    cg, coord, gene, pos, island = explode_annotation(dfGpl)
    gene_first = {}
    for i, g in enumerate(gene_names):
        gene_first.setdefault(str(g), i)
    sel = [i for i, g in enumerate(gene) if g in gene_first]          # :264-265
    cg = [cg[i] for i in sel]; coord = [coord[i] for i in sel]; gene = [gene[i] for i in sel]
    pos = [pos[i] for i in sel]; island = [island[i] for i in sel]
    n_genes = len(gene_names)
    if not cg:
        z = np.zeros(0, object)
        return ProbeIndex(np.zeros(n_genes + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32), z, z, z,
                          np.zeros(0, np.int64))
    if coord_numeric is None:
        coord_numeric = all(isinstance(c, (int, float, np.integer, np.floating)) and not isinstance(c, bool)
                            for c in coord)
    coord_key = np.asarray(coord, float) if coord_numeric else _rank_strings([str(c) for c in coord])
    names_sorted = sorted(set(gene))
    rank_of_name = {g: r for r, g in enumerate(names_sorted)}
    gene_rank = np.fromiter((rank_of_name[g] for g in gene), np.int64, len(gene))
    gene_of_rank = np.array([gene_first[g] for g in names_sorted], np.int64)
    cg_codes = {}
    cg_id = np.fromiter((cg_codes.setdefault(c, len(cg_codes)) for c in cg), np.int64, len(cg))
    pos_codes = {}
    pos_id = np.fromiter((pos_codes.setdefault(p, len(pos_codes)) for p in pos), np.int64, len(pos))
    row_of = {}
    for r, pid in enumerate(methylation_ids):
        row_of.setdefault(str(pid), r)
    probe_row_of_cg = np.full(len(cg_codes), -1, np.int64)
    for c, code in cg_codes.items():
        probe_row_of_cg[code] = row_of.get(c, -1)
    row_ptr, probe_idx, pair_gene, kept_csr, output_order = build_index_core(
        cg_id, gene_rank, coord_key, pos_id, gene_of_rank, probe_row_of_cg, n_genes)
    take = lambda lst: np.asarray(lst, dtype=object)[kept_csr]
    return ProbeIndex(row_ptr, probe_idx, pair_gene, take(cg), take(pos), take(island), output_order)


@dataclass
class GroupIndex:
    """Probe groups of the pooled outputs, in the reference's column order (genes by sorted name; inside a gene
    the positions / islands in order of first appearance in ``dfCGunique``)."""
    group_ptr: np.ndarray     # int64 [Q+1]
    group_pairs: np.ndarray   # int32: CSR pair indices, only pairs with a data row
    gene: np.ndarray          # int32 [Q] gene (expression order)
    key: np.ndarray           # object [Q]: position / island value, None for by="gene"

    @property
    def n_groups(self) -> int:
        return int(self.group_ptr.size - 1)


def build_groups(idx: ProbeIndex, by: str = "gene") -> GroupIndex:
    """Which probes are pooled together.

    by="gene": every gene of ``dfCGunique`` (R/EpimethexAnalysis.R:476-545); a gene whose probes all lack data keeps
    an (empty) group -> NA column (:536-541).
    by="position" / by="island": (gene, value) sets with MORE THAN ONE data-bearing probe (R/Functions.R:146-153);
    island values "NA_NA" or starting with "_" are skipped (R/EpimethexAnalysis.R:321).  The reference selects the
    columns by a name it never assigned ("cg_gene" vs "cg~gene") and stops; here they are selected by
    index, which is what the loop is written to do.
    """
    import re

    if by not in ("gene", "position", "island"):
        raise ValueError("by must be 'gene', 'position' or 'island'")
    po = np.asarray(idx.output_order, np.int64)
    if po.size == 0:
        return GroupIndex(np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, object))
    g = np.asarray(idx.pair_gene, np.int64)[po]
    _, first = np.unique(g, return_index=True)
    gene_order = g[np.sort(first)]                      # genes in output order
    grank = np.full(int(g.max()) + 1, -1, np.int64)
    grank[gene_order] = np.arange(gene_order.size)
    has = np.asarray(idx.probe_idx)[po] >= 0
    if by == "gene":
        sel = np.nonzero(has)[0]
        counts = np.bincount(grank[g[sel]], minlength=gene_order.size)
        order = sel[np.argsort(grank[g[sel]], kind="stable")]
        ptr = np.zeros(gene_order.size + 1, np.int64)
        np.cumsum(counts, out=ptr[1:])
        return GroupIndex(ptr, po[order].astype(np.int32), gene_order.astype(np.int32),
                          np.full(gene_order.size, None, object))
    vals = (idx.position if by == "position" else idx.island)[po]
    codes, names = {}, []
    kc = np.empty(po.size, np.int64)
    for i, v in enumerate(vals.tolist()):
        c = codes.get(v)
        if c is None:
            c = codes[v] = len(names)
            names.append(v)
        kc[i] = c
    valid = np.ones(len(names), bool)
    if by == "island":
        valid = np.array([re.search(r"NA_NA|^_", str(v)) is None for v in names], bool)
    sel = np.nonzero(has & valid[kc])[0]
    comp = grank[g[sel]] * len(names) + kc[sel]
    order = sel[np.argsort(comp, kind="stable")]
    uniq, start, cnt = np.unique(comp[np.argsort(comp, kind="stable")], return_index=True, return_counts=True)
    keep = cnt > 1
    pieces = [order[s:s + c] for s, c in zip(start[keep], cnt[keep])]
    ptr = np.zeros(int(keep.sum()) + 1, np.int64)
    np.cumsum(cnt[keep], out=ptr[1:])
    pairs = po[np.concatenate(pieces)] if pieces else np.zeros(0, np.int64)
    genes = gene_order[uniq[keep] // len(names)].astype(np.int32)
    keys = np.asarray(names, dtype=object)[uniq[keep] % len(names)] if len(names) else np.zeros(0, object)
    return GroupIndex(ptr, pairs.astype(np.int32), genes, keys)
