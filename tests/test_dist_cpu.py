"""N > 1 host logic on CPU: world_size-2 (and 3) gloo processes exercise the gene partition and the padded
gather of epimethex_b200.distributed; the per-shard compute is the oracle (this is a test of the plumbing)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from epimethex_b200 import distributed, synth


def test_partition_is_contiguous_balanced_and_complete():
    ann = synth.annotation(5000, 400)
    idx = synth.probe_index(ann, 0, 400)
    for world in (1, 2, 3, 8):
        parts = distributed.partition_genes(idx.row_ptr, world)
        assert parts[0][0] == 0 and parts[-1][1] == 400
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [int(idx.row_ptr[b] - idx.row_ptr[a]) for a, b in parts]
        assert sum(sizes) == idx.n_pairs
        assert max(sizes) - min(sizes) <= 2 * int(np.diff(idx.row_ptr).max())   # within one big gene of even
    rp, pi, first, count = distributed.shard_index(idx, 100, 250)
    assert rp[0] == 0 and rp[100] == 0 and rp[250] == count and rp[-1] == count
    assert np.array_equal(pi, idx.probe_idx[first:first + count])


def test_probe_slices_cover_each_rank_and_total_one_matrix():
    """probe-sliced copies: a rank's slice holds exactly the rows its pairs name; the oracle on the slice
    with the translated index equals the oracle on the whole matrix"""
    ann = synth.annotation(3000, 200)
    has = np.ones(3000, bool); has[::41] = False
    idx = synth.probe_index(ann, 0, 200, probe_has_data=has)
    x = synth.expression(200, 30)
    y = synth.methylation(3000, 30, ann, threads=1)
    whole = oracle.analyze(x, y, idx.row_ptr, idx.probe_idx)
    total_rows = 0
    for g0, g1 in distributed.partition_genes(idx.row_ptr, 3):
        rp, pi, first, count = distributed.shard_index(idx, g0, g1)
        rows, local = distributed.slice_probes(pi)
        assert np.all(np.diff(rows) > 0) and np.array_equal(local < 0, pi < 0)
        assert np.array_equal(rows[local[local >= 0]], pi[pi >= 0])
        total_rows += rows.size
        part = oracle.analyze(x, np.ascontiguousarray(y[rows]), rp, local)
        assert np.array_equal(part.pair, whole.pair[first:first + count], equal_nan=True)
        assert np.array_equal(part.ks_dnum, whole.ks_dnum[first:first + count])
    used = np.unique(idx.probe_idx[idx.probe_idx >= 0]).size
    assert used <= total_rows <= used + 3 * int(np.diff(idx.row_ptr).max())    # ~ one matrix in total, not three


def _oracle_compute(sess, expr, row_ptr, probe_idx, opts):
    return oracle.analyze(expr, sess, row_ptr, probe_idx, test_meth_t=opts["tm"])


def _worker(rank, world, port, tm, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ann = synth.annotation(1500, 90)
        x = synth.expression(90, 30)
        y = synth.methylation(1500, 30, ann, threads=1)
        idx = synth.probe_index(ann, 0, 90)
        res = distributed.analyze_sharded(y, x, idx, {"tm": tm}, compute=_oracle_compute)
        if rank == 0:
            np.savez(out_path, pair=res.pair, pair_na=res.pair_na, ks_dnum=res.ks_dnum, pair_stat=res.pair_stat,
                     gene=res.gene, counts=np.array(res.counts))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,tm", [(2, False), (3, True)])
def test_sharded_equals_single_process(tmp_path, world, tm):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, port, tm, out), nprocs=world, join=True)
    got = np.load(out)
    ann = synth.annotation(1500, 90)
    x = synth.expression(90, 30); y = synth.methylation(1500, 30, ann, threads=1)
    idx = synth.probe_index(ann, 0, 90)
    ref = oracle.analyze(x, y, idx.row_ptr, idx.probe_idx, test_meth_t=tm)
    assert np.array_equal(got["pair"], ref.pair, equal_nan=True)          # same code, same inputs: bit-identical
    assert np.array_equal(got["pair_na"], ref.pair_na) and np.array_equal(got["ks_dnum"], ref.ks_dnum)
    assert np.array_equal(got["pair_stat"], ref.pair_stat, equal_nan=True)
    assert np.array_equal(got["gene"], ref.gene, equal_nan=True) and tuple(got["counts"]) == ref.counts
