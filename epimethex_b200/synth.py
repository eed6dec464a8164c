"""Synthetic TCGA-shaped inputs: counter-based, so any slice can be regenerated bit-exactly
anywhere from (seed, stream, row, col) without files.  Only +, -, *, / and integer hashing are used, so
the values do not depend on a libm.

Shapes: SKCM (20,531 genes x 473 samples, 485,577 HM450 probes), EPIC (866k probes x 1,000 samples),
pan-cancer (9,000 samples); see ``SHAPES``.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from .annotation import ProbeIndex, build_index_core

SEED_EXPR, SEED_METH, SEED_ANNOT = 1001, 1002, 1003

SHAPES = {
    "skcm": dict(n_genes=20531, n_samples=473, n_probes=485577),
    "skcm474": dict(n_genes=20531, n_samples=474, n_probes=485577),
    "epic": dict(n_genes=20531, n_samples=1000, n_probes=866000),
    "epic999": dict(n_genes=20531, n_samples=999, n_probes=866000),
    "pancan": dict(n_genes=20531, n_samples=9000, n_probes=485577),
}

GROUPS = ["TSS1500", "TSS200", "5'UTR", "1stExon", "Body", "3'UTR"]
RELATIONS = ["Island", "N_Shore", "S_Shore", "N_Shelf", "S_Shelf", ""]

_U = np.uint64


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = x + _U(0x9E3779B97F4A7C15)
        x = (x ^ (x >> _U(30))) * _U(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> _U(27))) * _U(0x94D049BB133111EB)
        return x ^ (x >> _U(31))


def hash64(seed: int, stream: int, row, col) -> np.ndarray:
    row = np.asarray(row, dtype=np.uint64)
    col = np.asarray(col, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _mix(np.asarray(_U(seed) * _U(0xD1342543DE82EF95) + _U(stream), dtype=np.uint64))
        h = _mix(h + row * _U(0x9E3779B97F4A7C15))
        return _mix(h ^ (col * _U(0xC2B2AE3D27D4EB4F)))


def uniform(seed, stream, row, col) -> np.ndarray:
    """U[0,1) with 53 bits"""
    return (hash64(seed, stream, row, col) >> _U(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normalish(seed, stream, row, col) -> np.ndarray:
    """zero-mean, unit-variance bell (sum of four 32-bit uniforms); exact arithmetic only"""
    h1 = hash64(seed, stream, row, col)
    h2 = hash64(seed, stream + 7919, row, col)
    m = _U(0xFFFFFFFF)
    s = ((h1 & m).astype(np.float64) + (h1 >> _U(32)).astype(np.float64)
         + (h2 & m).astype(np.float64) + (h2 >> _U(32)).astype(np.float64))
    return (s * (1.0 / 4294967296.0) - 2.0) * 1.7320508075688772


def expression(n_genes: int, n_samples: int, g0: int = 0, *, ties: bool = True, seed: int = SEED_EXPR):
    """log2-RSEM-like: max(0, mu_g + sigma_g * z); clamping makes genuine zero ties. [G, n] gene-major."""
    g = np.arange(g0, g0 + n_genes, dtype=np.uint64)[:, None]
    s = np.arange(n_samples, dtype=np.uint64)[None, :]
    mu = -1.0 + 13.0 * uniform(seed, 0, g, 0)
    sigma = 0.3 + 1.7 * uniform(seed, 1, g, 0)
    x = mu + sigma * normalish(seed, 2, g, s)
    return np.maximum(x, 0.0) if ties else x


def expression_z(genes, n_samples: int, seed: int = SEED_EXPR):
    g = np.asarray(genes, dtype=np.uint64)[:, None]
    s = np.arange(n_samples, dtype=np.uint64)[None, :]
    return normalish(seed, 2, g, s)


@dataclass
class SynthAnnotation:
    """integer annotation entries (exploded form) + per-probe attributes"""
    entry_probe: np.ndarray   # int64 [E]
    entry_gene: np.ndarray    # int64 [E] global gene id
    entry_group: np.ndarray   # int64 [E] index into GROUPS
    probe_coord: np.ndarray   # int64 [P]
    probe_relation: np.ndarray  # int64 [P] index into RELATIONS
    probe_gene0: np.ndarray   # int64 [P] first gene (or -1): carrier of the planted slope
    probe_slope: np.ndarray   # float64 [P]
    n_probes: int
    n_genes: int


def annotation(n_probes: int, n_genes: int, seed: int = SEED_ANNOT) -> SynthAnnotation:
    """HM450-like: ~25% of probes have no gene; the rest list 1-3 genes (with transcript duplicates);
    probes per gene are log-normal."""
    p = np.arange(n_probes, dtype=np.uint64)
    gidx = np.arange(n_genes, dtype=np.uint64)
    w = np.exp(1.2 * normalish(seed, 0, gidx, 0))            # host-only bookkeeping, libm is fine here
    cumw = np.cumsum(w)
    total = cumw[-1]
    annotated = uniform(seed, 1, p, 0) >= 0.25
    uk = uniform(seed, 2, p, 0)
    k = np.where(uk < 0.70, 1, np.where(uk < 0.92, 2, 3))
    k = np.where(annotated, k, 0)
    entry_probe, entry_gene, entry_group = [], [], []
    gene_prev = None
    for j in range(3):
        sel = np.nonzero(k > j)[0]
        pj = sel.astype(np.uint64)
        gj = np.searchsorted(cumw, uniform(seed, 10 + j, pj, 0) * total, side="right").clip(0, n_genes - 1)
        if j > 0:
            same = uniform(seed, 20 + j, pj, 0) < 0.5          # transcript duplicate of the previous gene
            gj = np.where(same, gene_prev[sel], gj)
        gene_full = np.full(n_probes, -1, np.int64)
        gene_full[sel] = gj
        if gene_prev is None:
            gene0 = gene_full.copy()
        gene_prev = gene_full
        grp = (uniform(seed, 30 + j, pj, 0) * len(GROUPS)).astype(np.int64).clip(0, len(GROUPS) - 1)
        entry_probe.append(sel.astype(np.int64)); entry_gene.append(gj.astype(np.int64)); entry_group.append(grp)
    ep = np.concatenate(entry_probe); eg = np.concatenate(entry_gene); egr = np.concatenate(entry_group)
    # exploded order: annotation row, then position inside the ';' list
    jj = np.concatenate([np.full(a.size, j) for j, a in enumerate(entry_probe)])
    o = np.lexsort((jj, ep))
    coord = 1 + (uniform(seed, 3, p, 0) * 2.4e8).astype(np.int64)
    rel = (uniform(seed, 4, p, 0) * len(RELATIONS)).astype(np.int64).clip(0, len(RELATIONS) - 1)
    us = uniform(seed, 5, p, 0)
    slope = np.where(us < 0.2, (0.02 + 0.3 * us) * np.where(uniform(seed, 6, p, 0) < 0.5, -1.0, 1.0), 0.0)
    return SynthAnnotation(ep[o], eg[o], egr[o], coord, rel, gene0, slope, n_probes, n_genes)


def methylation_rows(probes, n_samples: int, ann: SynthAnnotation, *, decimals: int | None = None,
                     seed: int = SEED_METH, out: np.ndarray | None = None) -> np.ndarray:
    """beta values for the given probe ids, [len(probes), n] probe-major."""
    probes = np.asarray(probes, dtype=np.int64)
    pu = probes.astype(np.uint64)[:, None]
    s = np.arange(n_samples, dtype=np.uint64)[None, :]
    um = uniform(seed, 0, pu, 0)
    mode = np.where(um < 0.4, 0.08, np.where(um < 0.6, 0.5, 0.9))
    sd = 0.01 + 0.07 * uniform(seed, 1, pu, 0)
    y = mode + sd * normalish(seed, 2, pu, s)
    slope = ann.probe_slope[probes]
    planted = np.nonzero(slope != 0.0)[0]
    if planted.size:
        g0 = ann.probe_gene0[probes[planted]]
        ok = g0 >= 0
        if ok.any():
            y[planted[ok]] += slope[planted[ok], None] * expression_z(g0[ok], n_samples)
    np.clip(y, 0.002, 0.998, out=y)
    if decimals is not None:
        scale = float(10 ** decimals)
        y = np.rint(y * scale) / scale
    if out is not None:
        out[...] = y
        return out
    return y


def methylation(n_probes: int, n_samples: int, ann: SynthAnnotation, *, decimals=None, out=None,
                chunk: int = 8192, threads: int = 8) -> np.ndarray:
    """whole matrix [P, n]; generated in probe chunks on a thread pool (numpy releases the GIL)."""
    if out is None:
        out = np.empty((n_probes, n_samples), np.float64)
    starts = list(range(0, n_probes, chunk))

    def work(a):
        b = min(a + chunk, n_probes)
        methylation_rows(np.arange(a, b), n_samples, ann, decimals=decimals, out=out[a:b])

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, starts))
    return out


def probe_index(ann: SynthAnnotation, g0: int, g1: int, probe_has_data: np.ndarray | None = None) -> ProbeIndex:
    """CSR for the analysed gene range [g0, g1) (0-based, half-open) in dfCGunique order. Gene names are
    zero-padded ("G00017"), so name order = id order; Coordinate_36 is a decimal STRING and is ordered
    as one (code-point collation), as in the reference's README example."""
    sel = np.nonzero((ann.entry_gene >= g0) & (ann.entry_gene < g1))[0]      # :264-265
    ep, eg, egr = ann.entry_probe[sel], ann.entry_gene[sel] - g0, ann.entry_group[sel]
    n_genes = g1 - g0
    coord_str = np.char.mod("%d", ann.probe_coord[ep])
    _, coord_rank = np.unique(coord_str, return_inverse=True)
    row_of = np.arange(ann.n_probes, dtype=np.int64)
    if probe_has_data is not None:
        row_of = np.where(probe_has_data, row_of, -1)
    row_ptr, probe_idx, pair_gene, kept, output_order = build_index_core(
        ep, eg, coord_rank, egr, np.arange(n_genes), row_of, n_genes)
    cg = np.char.mod("cg%08d", ep[kept]).astype(object)
    pos = np.asarray(GROUPS, dtype=object)[egr[kept]]
    rel = np.asarray(RELATIONS, dtype=object)[ann.probe_relation[ep[kept]]]
    isl = np.array([f"{r}_chr1:{c}-{c + 300}" if r else "_" for r, c in zip(rel, ann.probe_coord[ep[kept]])],
                   dtype=object)
    return ProbeIndex(row_ptr, probe_idx, pair_gene, cg, pos, isl, output_order)


def gene_names(n_genes: int, g0: int = 0):
    return [f"G{g:05d}" for g in range(g0, g0 + n_genes)]


def frames(n_genes: int, n_samples: int, n_probes: int, *, decimals=None, ties=True):
    """Small-scale pandas inputs in the reference's format (expression, annotation, methylation)."""
    import pandas as pd

    ann = annotation(n_probes, n_genes)
    x = expression(n_genes, n_samples, ties=ties)
    y = methylation(n_probes, n_samples, ann, decimals=decimals, threads=1)
    samples = [f"TCGA-{i:04d}" for i in range(n_samples)]
    dfe = pd.DataFrame(x, columns=samples)
    dfe.insert(0, "sample", gene_names(n_genes))
    dfm = pd.DataFrame(y, columns=samples)
    dfm.insert(0, "sample", [f"cg{p:08d}" for p in range(n_probes)])
    names = [[] for _ in range(n_probes)]
    groups = [[] for _ in range(n_probes)]
    for p, g, gr in zip(ann.entry_probe, ann.entry_gene, ann.entry_group):
        names[p].append(f"G{g:05d}"); groups[p].append(GROUPS[gr])
    dfa = pd.DataFrame({
        "ID": [f"cg{p:08d}" for p in range(n_probes)],
        "Relation_to_UCSC_CpG_Island": [RELATIONS[r] for r in ann.probe_relation],
        "UCSC_CpG_Islands_Name": [f"chr1:{c}-{c + 300}" if RELATIONS[r] else "" for r, c in
                                  zip(ann.probe_relation, ann.probe_coord)],
        "UCSC_RefGene_Accession": [";".join("NM_%06d" % 1 for _ in nm) for nm in names],
        "Chromosome_36": ["1"] * n_probes,
        "Coordinate_36": [str(c) for c in ann.probe_coord],
        "UCSC_RefGene_Name": [";".join(nm) for nm in names],
        "UCSC_RefGene_Group": [";".join(gr) for gr in groups],
    })
    return dfe, dfa, dfm, ann
