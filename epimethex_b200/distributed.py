"""Gene-range sharding across the GPUs of one box and the gather of the result tables.

The reference parallelises over genes with ``foreach %dopar%`` on a PSOCK cluster and folds the per-gene
results with ``.combine = cbind`` on the master (R/EpimethexAnalysis.R:333-352).  Here: one process per GPU
(``torch.distributed``, NCCL over NVLink), contiguous gene ranges cut at equal shares of the pair count, and one
gather of the per-rank pair tables to rank 0 -- the only collective of the path.  Every rank keeps the whole
expression matrix (78 MB for SKCM) and runs the gene-level kernels for all genes, so the single global coupling
of the path -- the bin.var reference gene and its three group sizes (R/EpimethexAnalysis.R:134-139) -- comes out
identical on every rank without a collective; only the (gene, probe) pairs are sharded.
"""
from __future__ import annotations

from dataclasses import replace

import numpy as np

from .annotation import ProbeIndex


def partition_genes(row_ptr: np.ndarray, world: int):
    """Contiguous gene ranges [g0, g1) with (nearly) equal numbers of pairs. Returns a list of `world` tuples."""
    row_ptr = np.asarray(row_ptr, np.int64)
    G = row_ptr.size - 1
    total = int(row_ptr[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        g = int(np.searchsorted(row_ptr, target, side="left"))
        # choose the boundary closest to the target share
        if g > 0 and abs(row_ptr[g - 1] - target) <= abs(row_ptr[min(g, G)] - target):
            g -= 1
        cuts.append(min(max(g, cuts[-1]), G))
    cuts.append(G)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_index(idx: ProbeIndex, g0: int, g1: int) -> tuple[np.ndarray, np.ndarray, int, int]:
    """CSR of one rank: all genes keep their row (the gene kernels run everywhere), only genes in [g0, g1) keep
    their pairs. Returns (row_ptr_local, probe_idx_local, first_pair, n_local_pairs)."""
    rp = np.asarray(idx.row_ptr, np.int64)
    lo, hi = int(rp[g0]), int(rp[g1])
    local = np.clip(rp, lo, hi) - lo
    return local, np.ascontiguousarray(idx.probe_idx[lo:hi]), lo, hi - lo


def slice_probes(probe_idx: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Probe-sliced methylation copy for one rank (pan-cancer: 35 GB, of which a rank needs only the rows
    its own genes reference). Returns (rows, local_idx): `rows` = ascending distinct methylation rows named by
    `probe_idx` (-1 = annotated probe without data, kept as -1), `local_idx` = the same index against
    `methylation[rows]`. Uploading `methylation[rows]` on every rank moves about one matrix over PCIe in total
    instead of one per GPU. (`epx_multi_set_methylation_sliced` is the one-process twin of this.)"""
    probe_idx = np.asarray(probe_idx, np.int32)
    have = probe_idx >= 0
    rows = np.unique(probe_idx[have])
    local = np.full(probe_idx.shape, -1, np.int32)
    local[have] = np.searchsorted(rows, probe_idx[have]).astype(np.int32)
    return rows, local


def _default_compute(sess, expr, row_ptr, probe_idx, opts):
    return sess.analyze(expr, row_ptr, probe_idx, opts)


def analyze_sharded(sess, expr, idx: ProbeIndex, opts, *, group=None, compute=_default_compute, device=None):
    """Run this rank's shard and gather the pair tables to rank 0.

    `compute(sess, expr, row_ptr, probe_idx, opts)` returns an object with .pair [N,11], .pair_na, .pair_stat,
    .ks_dnum, .gene, .gene_na, .gene_ks_dnum, .perm, .counts, .ref_gene (the GPU session by default; the CPU tests
    inject the oracle to exercise the partition + gather logic without a GPU).
    Returns the full result on rank 0 and None elsewhere.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = partition_genes(idx.row_ptr, world)
    g0, g1 = parts[rank]
    rp_local, pi_local, first, count = shard_index(idx, g0, g1)
    res = compute(sess, expr, rp_local, pi_local, opts)
    assert res.pair.shape[0] == count
    counts = [int(idx.row_ptr[b] - idx.row_ptr[a]) for a, b in parts]
    maxc = max(counts) if counts else 0
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")

    def gather(arr: np.ndarray, dtype):
        """padded gather of a [count, k] table (counts are known on every rank from the CSR: no size exchange)"""
        k = arr.shape[1] if arr.ndim == 2 else 1
        buf = torch.zeros((maxc, k), dtype=dtype, device=dev)
        if count:
            buf[:count] = torch.as_tensor(np.ascontiguousarray(arr).reshape(count, k)).to(dev)
        out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, out, dst=0, group=group)
        if rank != 0:
            return None
        return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0).reshape((-1, k) if arr.ndim == 2 else (-1,))

    pair = gather(res.pair, torch.float64)
    pair_na = gather(res.pair_na.astype(np.int32), torch.int32)
    pair_stat = gather(res.pair_stat, torch.float64)
    ks_dnum = gather(res.ks_dnum, torch.int32)
    if rank != 0:
        return None
    return replace(res, pair=pair, pair_na=pair_na.astype(np.uint16), pair_stat=pair_stat, ks_dnum=ks_dnum)
