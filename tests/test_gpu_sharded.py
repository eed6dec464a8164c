"""Genes sharded over 2 GPUs of one box with the NCCL gather of the pair tables (epimethex_b200.distributed).
Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest

from epimethex_b200 import _native, synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    from epimethex_b200 import distributed
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ann = synth.annotation(4000, 240)
        x = synth.expression(240, 99)
        y = synth.methylation(4000, 99, ann, threads=2)
        idx = synth.probe_index(ann, 0, 240)
        sess = _native.Session(rank)
        sess.set_methylation(y)
        res = distributed.analyze_sharded(sess, x, idx, _native.make_options(True, True, False))
        if rank == 0:
            np.savez(out_path, pair=res.pair, pair_na=res.pair_na, ks_dnum=res.ks_dnum, gene=res.gene)
        sess.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_shards_equal_one_gpu(tmp_path):
    if _native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    ann = synth.annotation(4000, 240)
    x = synth.expression(240, 99); y = synth.methylation(4000, 99, ann, threads=2)
    idx = synth.probe_index(ann, 0, 240)
    with _native.Session(0) as sess:
        sess.set_methylation(y)
        one = sess.analyze(x, idx.row_ptr, idx.probe_idx, _native.make_options(True, True, False))
    assert np.array_equal(got["pair"], one.pair, equal_nan=True)     # same kernels, same inputs: bit-identical
    assert np.array_equal(got["pair_na"], one.pair_na) and np.array_equal(got["ks_dnum"], one.ks_dnum)
    assert np.array_equal(got["gene"], one.gene, equal_nan=True)


def test_one_process_multi_gpu_equals_one_gpu():
    """epx_multi_*: what the R `.Call` shim uses to drive all GPUs of the box from R's single process"""
    nd = min(_native.device_count(), 4)
    if nd < 2:
        pytest.skip("needs 2 GPUs")
    ann = synth.annotation(5000, 300)
    x = synth.expression(300, 99); y = synth.methylation(5000, 99, ann, threads=2)
    idx = synth.probe_index(ann, 0, 300)
    with _native.Session(0) as sess:
        sess.set_methylation(y)
        one = sess.analyze(x, idx.row_ptr, idx.probe_idx)
    with _native.MultiSession(list(range(nd))) as ms:
        ms.set_methylation(y)
        for _ in range(2):      # twice: the sessions are reused across calls
            many = ms.analyze(x, idx.row_ptr, idx.probe_idx)
            assert many.counts == one.counts and many.ref_gene == one.ref_gene
            assert np.array_equal(many.pair, one.pair, equal_nan=True) and np.array_equal(many.pair_na, one.pair_na)
            assert np.array_equal(many.ks_dnum, one.ks_dnum) and np.array_equal(many.pair_stat, one.pair_stat, equal_nan=True)
            assert np.array_equal(many.gene, one.gene, equal_nan=True) and np.array_equal(many.perm, one.perm)


def test_probe_sliced_copies_equal_replicas():
    """epx_multi_set_methylation_sliced: each device holds only the rows its genes reference (pan-cancer).
    On a one-GPU box the two sessions share device 0 -- the slicing and the row translation are the same code."""
    nd = min(_native.device_count(), 4)
    devices = list(range(nd)) if nd >= 2 else [0, 0, 0]
    ann = synth.annotation(5000, 300)
    x = synth.expression(300, 99); y = synth.methylation(5000, 99, ann, threads=2)
    has = np.ones(5000, bool); has[::37] = False           # some annotated probes without data (-1 in the index)
    idx = synth.probe_index(ann, 0, 300, probe_has_data=has)
    assert (idx.probe_idx < 0).any()
    with _native.Session(0) as sess:
        sess.set_methylation(y)
        one = sess.analyze(x, idx.row_ptr, idx.probe_idx)
    with _native.MultiSession(devices) as ms:
        ms.set_methylation_sliced(y, idx.row_ptr, idx.probe_idx)
        for _ in range(2):
            many = ms.analyze(x, idx.row_ptr, idx.probe_idx)
            assert many.counts == one.counts and many.ref_gene == one.ref_gene
            assert np.array_equal(many.pair, one.pair, equal_nan=True) and np.array_equal(many.pair_na, one.pair_na)
            assert np.array_equal(many.ks_dnum, one.ks_dnum) and np.array_equal(many.pair_stat, one.pair_stat, equal_nan=True)
            assert np.array_equal(many.gene, one.gene, equal_nan=True) and np.array_equal(many.perm, one.perm)
        # an index that needs rows outside the slices is refused, not silently mis-read
        other = synth.probe_index(ann, 0, 300)
        with pytest.raises(_native.EpxError, match="EPX_ESTATE"):
            ms.analyze(x, other.row_ptr, other.probe_idx)
        # back to replicas: any index works again
        ms.set_methylation(y)
        full = ms.analyze(x, other.row_ptr, other.probe_idx)
        assert full.pair.shape[0] == other.n_pairs


def test_multi_gpu_pooled_groups_and_column_upload():
    """epx_multi_groups + epx_multi_set_methylation_cols: the pooled tables of CG_of_a_genes / CG_by_position, each group
    on the device that owns its gene, equal the single-session result; methylation given as data.frame columns"""
    from epimethex_b200.annotation import build_groups
    nd = min(_native.device_count(), 4)
    devices = list(range(nd)) if nd >= 2 else [0, 0, 0]
    ann = synth.annotation(5000, 300)
    x = synth.expression(300, 99); y = synth.methylation(5000, 99, ann, threads=2)
    has = np.ones(5000, bool); has[::29] = False
    idx = synth.probe_index(ann, 0, 300, probe_has_data=has)
    with _native.Session(0) as sess:
        sess.set_methylation(y)
        one = sess.analyze(x, idx.row_ptr, idx.probe_idx)
        ref = {by: (gi, sess.analyze_groups(gi.group_ptr, gi.group_pairs))
               for by in ("gene", "position") for gi in [build_groups(idx, by)]}
    with _native.MultiSession(devices) as ms:
        with pytest.raises(_native.EpxError, match="EPX_ESTATE"):
            ms.analyze_groups(np.zeros(1, np.int64), np.zeros(0, np.int32))     # before analyze
        ms.set_methylation_columns([np.ascontiguousarray(y[:, j]) for j in range(y.shape[1])])
        many = ms.analyze(x, idx.row_ptr, idx.probe_idx)
        assert np.array_equal(many.pair, one.pair, equal_nan=True)
        for by, (gi, want) in ref.items():
            assert gi.n_groups > 10
            got = ms.analyze_groups(gi.group_ptr, gi.group_pairs)
            assert np.array_equal(got.group, want.group, equal_nan=True), by
            assert np.array_equal(got.group_na, want.group_na) and np.array_equal(got.group_dnum, want.group_dnum)
            assert np.array_equal(got.group_stat, want.group_stat, equal_nan=True)
