#!/usr/bin/env python
"""bench.py — gene-probe pairs/s of the EpiMethEx association path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload skcm_full] [--impl ours|reference]

A "step" is one pass of the whole device pipeline (gene stratification + tests, per-probe rank, fused
gather/moments/medians/KS per pair) over the synthetic workload, inputs resident in HBM.  One JSON line
is printed by rank 0.  `value` is device-timed (CUDA events, max over ranks); `e2e` is the same metric
through the blocking C-ABI calls with pinned HOST buffers (methylation + expression + index H2D, kernels,
result tables D2H inside the timed region), measured two ways -- epx_session_set_methylation_rows + epx_analyze
(the whole matrix is uploaded) and the one call epx_analyze_host (the device fetches only the rows the index
names) -- both in `e2e.flows`, the faster one as the headline with its own byte counts; `roofline` is the dominant kernel's algorithmic
bytes over its CUDA-event time against MEASURED_PEAKS.json; `cpu_baseline` is the CPU oracle (the restated
reference, C + OpenMP, all host cores) on a bounded gene sample of the same workload.

Under torchrun (N > 1) every rank runs its own cohort of the same shape (weak scaling: genes/probes shard
with no data-path collective) and the N result tables of a step are gathered over NCCL inside the timed
region, as the path's only exchange step, on one rank that rotates with the step (rank step mod N).  `--workload pancan_full` (BASELINE configs[4]) is the one
strong-scaling workload: one 9,000-sample cohort, genes cut into N ranges, probe-sliced methylation.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the north_star target shape (fits one GPU; inputs 1.8 GB >> 126 MB L2)
    "skcm_full": dict(shape="skcm", genes=(0, 20531), te=True, tm=False,
                      desc="synthetic SKCM full genome: 20,531 genes x 473 samples, 485,577 probes, t-test expression + KS methylation"),
    "skcm_full_tt": dict(shape="skcm", genes=(0, 20531), te=True, tm=True,
                         desc="synthetic SKCM full genome, t-test on both"),
    # BASELINE.json configs[1]
    "skcm_1000": dict(shape="skcm", genes=(0, 1000), te=True, tm=True,
                      desc="synthetic SKCM shape, gene range 1-1000, 473 samples, HM450 annotation, t-test on both"),
    "skcm474_full": dict(shape="skcm474", genes=(0, 20531), te=True, tm=False,
                         desc="SKCM full genome at 474 samples (nearest multiple of 3: the reference's defined case)"),
    "epic": dict(shape="epic", genes=(0, 20531), te=False, tm=False,
                 desc="synthetic EPIC-array shape: 866k probes x 1,000 samples, KS on both"),
    # BASELINE.json configs[4] at single-GPU scale: 9,000 samples, a 2,000-gene slice (the full 35 GB matrix is an
    # 8-GPU job); only the rows the slice references are generated
    "pancan_slice": dict(shape="pancan", genes=(0, 2000), te=True, tm=False,
                         desc="synthetic pan-cancer shape, 9,000 samples, gene range 1-2000 of 20,531, 485,577 probes"),
    # BASELINE.json configs[4] in full: all 20,531 genes at 9,000 samples. STRONG scaling: the genes are cut into N
    # contiguous ranges with equal pair counts (distributed.partition_genes) and rank r analyses range r as one
    # (from, to) call of the reference would -- the unit the reference itself is run in (README cap: 1000 genes per
    # call). Methylation is probe-sliced: a rank generates and uploads only the rows its genes reference.
    "pancan_full": dict(shape="pancan", genes=(0, 20531), te=True, tm=False, shard=True,
                        desc="synthetic pan-cancer shape: 20,531 genes x 9,000 samples x 485,577 probes, genes sharded over the GPUs, probe-sliced methylation"),
    "tiny": dict(shape=None, genes=(0, 300), te=True, tm=False, n_samples=99, n_probes=6000, n_genes=300,
                 desc="tiny self-test workload"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v == "Active":
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _host_matrix(shape, pinned: bool):
    """page-locked (epx_host_alloc) for the GPU arm; plain numpy for the reference arm, which must not load libepx"""
    if pinned:
        from epimethex_b200 import _native
        try:
            return _native.pinned_empty(shape), True
        except Exception:
            pass
    return np.empty(shape), False


def make_workload(name: str, rank: int, world: int = 1, pinned: bool = True):
    """Synthetic inputs for one rank. Rank r draws its own cohort (data seeds + r) over the same annotation; a
    workload with shard=True is ONE cohort whose genes are cut into `world` ranges (strong scaling)."""
    from epimethex_b200 import synth

    w = WORKLOADS[name]
    if w.get("shard"):
        return make_sharded_workload(w, rank, world, pinned)
    dims = dict(synth.SHAPES[w["shape"]]) if w["shape"] else dict(n_genes=w["n_genes"], n_samples=w["n_samples"], n_probes=w["n_probes"])
    n, P = dims["n_samples"], dims["n_probes"]
    g0, g1 = w["genes"]
    t0 = time.time()
    ann = synth.annotation(P, dims["n_genes"])
    idx = synth.probe_index(ann, g0, g1)
    expr = synth.expression(g1 - g0, n, g0=g0, seed=synth.SEED_EXPR + rank)
    meth, pinned = _host_matrix((P, n), pinned)
    threads = max(1, min(16, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    # only rows some pair references are ever read; generate those, zero the rest
    used = np.unique(idx.probe_idx[idx.probe_idx >= 0])
    meth[...] = 0.0
    from concurrent.futures import ThreadPoolExecutor

    def work(chunk):
        meth[chunk] = synth.methylation_rows(chunk, n, ann, seed=synth.SEED_METH + rank)

    step = max(16, 1_000_000 // n)      # ~8 MB of temporaries per numpy pass: stays in the host caches
    chunks = [used[i:i + step] for i in range(0, used.size, step)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, chunks))
    return dict(w=w, n=n, P=P, G=g1 - g0, ann=ann, idx=idx, expr=np.ascontiguousarray(expr), meth=meth,
                gen_s=time.time() - t0, n_unique=int(used.size), pinned=pinned, rows=P,
                total_pairs=int(idx.n_pairs) * world)


def make_sharded_workload(w, rank: int, world: int, pinned: bool = True):
    """one cohort, genes cut into `world` ranges of equal pair counts; this rank's range with a probe-sliced
    methylation copy (only the rows its pairs name, re-indexed)"""
    from concurrent.futures import ThreadPoolExecutor
    from dataclasses import replace

    from epimethex_b200 import distributed, synth

    dims = dict(synth.SHAPES[w["shape"]])
    n, P = dims["n_samples"], dims["n_probes"]
    t0 = time.time()
    ann = synth.annotation(P, dims["n_genes"])
    g_lo, g_hi = w["genes"]
    whole = synth.probe_index(ann, g_lo, g_hi)
    parts = distributed.partition_genes(whole.row_ptr, world)
    a, b = parts[rank]
    g0, g1 = g_lo + a, g_lo + b
    idx = synth.probe_index(ann, g0, g1)
    # every rank's (from, to) call builds its own index (the reference's dedup acts inside one call only)
    total_pairs = sum(int(synth.probe_index(ann, g_lo + pa, g_lo + pb).n_pairs) if (pa, pb) != (a, b) else int(idx.n_pairs)
                      for pa, pb in parts)
    rows, local = distributed.slice_probes(idx.probe_idx)
    idx = replace(idx, probe_idx=local)
    expr = synth.expression(g1 - g0, n, g0=g0, seed=synth.SEED_EXPR)
    meth, pinned = _host_matrix((max(1, rows.size), n), pinned)
    threads = max(1, min(16, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    step = max(16, 1_000_000 // n)

    def work(i):
        meth[i:i + step] = synth.methylation_rows(rows[i:i + step], n, ann, seed=synth.SEED_METH)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(0, rows.size, step)))
    return dict(w=w, n=n, P=P, G=g1 - g0, ann=ann, idx=idx, expr=np.ascontiguousarray(expr), meth=meth,
                gen_s=time.time() - t0, n_unique=int(rows.size), pinned=pinned, rows=int(meth.shape[0]),
                gene_range=(g0, g1), total_pairs=total_pairs)


_CPU_RATE = {}


def gene_sample(idx, G: int, stride: int, offset: int = 0):
    """every stride-th gene of the workload as its own CSR: (selected genes, row_ptr, probe_idx)"""
    sel = np.arange(offset % stride, G, stride)
    counts = idx.row_ptr[sel + 1] - idx.row_ptr[sel]
    ptr = np.zeros(sel.size + 1, np.int64)
    np.cumsum(counts, out=ptr[1:])
    pi = (np.concatenate([idx.probe_idx[idx.row_ptr[g]:idx.row_ptr[g + 1]] for g in sel]) if sel.size
          else np.zeros(0, np.int32))
    return sel, ptr, np.ascontiguousarray(pi, np.int32)


def cpu_baseline(wl, target_s: float, threads: int, keep: bool = False):
    """oracle (restated reference, C + OpenMP over genes) on every k-th gene of the same workload, sized to take
    about target_s seconds (k = 1, the whole workload, when that fits). keep=True also returns the oracle's tables
    and the gene sample, which the parity check of the GPU arm compares with the GPU's tables for the same sample."""
    import oracle

    idx, expr, meth = wl["idx"], wl["expr"], wl["meth"]
    G = wl["G"]

    def run(stride, offset=0):
        sel, ptr, pi = gene_sample(idx, G, stride, offset)
        t = time.perf_counter()
        res = oracle.analyze(expr[sel], meth, ptr, pi, test_expr_t=wl["w"]["te"], test_meth_t=wl["w"]["tm"], threads=threads)
        return int(ptr[-1]), time.perf_counter() - t, sel, ptr, pi, res

    key = id(wl)
    if key not in _CPU_RATE:
        pairs, dt = run(max(1, G // 64))[:2]
        _CPU_RATE[key] = [pairs / max(dt, 1e-6), 0]
    rate = _CPU_RATE[key][0]
    total_pairs = int(idx.row_ptr[-1])
    want_pairs = min(total_pairs, max(500, int(rate * target_s)))
    stride = max(1, int(total_pairs / max(want_pairs, 1) + 0.999))
    _CPU_RATE[key][1] += 1
    pairs, dt, sel, ptr, pi, res = run(stride, _CPU_RATE[key][1])
    _CPU_RATE[key][0] = pairs / max(dt, 1e-6)
    out = {"value": pairs / dt, "unit": "pairs/s", "cores": threads, "kind": "port", "pairs": pairs, "seconds": dt,
           "sample": f"every {stride}th gene of the workload ({sel.size} genes, {pairs} pairs) in {dt:.2f} s; "
                     "C/OpenMP restatement of the R reference (R is not installable here)"}
    if keep:
        out["oracle"] = (sel, ptr, pi, res)
    return out


RTOL, ATOL = 1e-8, 1e-10   # BASELINE.json north_star: doubles within 1e-10 absolute or 1e-8 relative


def parity_report(got, ref, what: str):
    """GPU tables against the oracle's on the same inputs: integers (sample order, KS numerators, NA masks, bin.var
    group sizes, reference gene) must be identical, doubles within RTOL/ATOL (NaN == NaN)."""
    ints = 0
    ints += int(tuple(got.counts) != tuple(ref.counts)) + int(got.ref_gene != ref.ref_gene)
    for a, b in ((got.perm.astype(np.int32), ref.perm), (got.ks_dnum, ref.ks_dnum), (got.gene_ks_dnum, ref.gene_ks_dnum)):
        ints += int((np.asarray(a) != np.asarray(b)).sum())
    na_bad = int((got.pair_na != ref.pair_na).sum() + (got.gene_na != ref.gene_na).sum())
    max_abs, max_rel, bad = 0.0, 0.0, 0
    for a, b in ((got.pair, ref.pair), (got.pair_stat, ref.pair_stat), (got.gene, ref.gene)):
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        bad += int((nan_a != nan_b).sum())
        fin = ~(nan_a | nan_b)
        with np.errstate(invalid="ignore", divide="ignore"):
            d = np.abs(a[fin] - b[fin])
            d = np.where(a[fin] == b[fin], 0.0, d)      # equal infinities
            bad += int((d > ATOL + RTOL * np.abs(b[fin])).sum())
            if d.size:
                big = d > ATOL                          # the relative bar only matters where the absolute one fails
                max_abs = max(max_abs, float(d.max()))
                if big.any():
                    max_rel = max(max_rel, float((d[big] / np.abs(b[fin][big])).max()))
    return {"against": "oracle (C restatement of the R reference; parity unpinned: no R available)", "sample": what,
            "pairs_checked": int(ref.pair.shape[0]), "genes_checked": int(ref.gene.shape[0]),
            "int_mismatches": ints, "na_mismatches": na_bad, "double_mismatches": bad,
            "max_abs": max_abs, "max_rel_where_abs_exceeds_atol": max_rel, "rtol": RTOL, "atol": ATOL,
            "ok": bool(ints == 0 and na_bad == 0 and bad == 0)}


def config_of(wl, name: str, world: int, total_pairs: int):
    """the `config` object: identical keys and values in the GPU arm and in the reference arm"""
    idx, n, G, P = wl["idx"], wl["n"], wl["G"], wl["P"]
    return {"workload": wl["w"]["desc"], "name": name, "pairs_per_gpu": int(idx.n_pairs), "genes": int(G), "samples": int(n),
            "probes": int(P), "unique_probes": int(wl["n_unique"]), "methylation_rows_on_gpu": int(wl["rows"]),
            "total_pairs": int(total_pairs),
            "l2": "inputs (%.2f GB methylation on this GPU) exceed the 126 MB L2; no flush" % (wl["rows"] * n * 8 / 1e9),
            "parallelism": f"genes sharded, {world} rank(s), NCCL gather of result tables" if world > 1 else "single GPU"}


def bind_to_gpu_numa_node(torch, local_rank):
    """N > 1: run this rank (and first-touch its page-locked host buffers) on the CPUs of the NUMA node its GPU hangs
    off, so that N ranks pushing a matrix each through the host at the same time do not all cross the socket link
    (round 1: per-GPU H2D fell from 51 to 22 GB/s at N = 8). Returns the node, or None when the topology is not exposed."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="skcm_full", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-analyze-host", action="store_true", help="e2e through the two-call path only")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-one-process", dest="one_process", action="store_false",
                    help="N > 1: skip the one-process epx_multi_analyze measurement on rank 0")
    ap.add_argument("--no-other-mode", dest="other_mode", action="store_false",
                    help="skip the second timing with the other methylation test (KS <-> Welch t)")
    ap.add_argument("--pooled", default="", choices=["", "gene", "position", "island"],
                    help="also time the pooled probe-group pass (CG_of_a_genes / by position / by island rows)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds for warmup + steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line (NCCL's version banner goes to stdout)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        # the reference's CPU implementation of the path (the oracle port: the reference is pure R and R is not
        # installable) on this host's cores; rank 0 only. Every step is the whole workload when warmup + steps of that
        # fit --ref-budget seconds, else a bounded gene sample (every k-th gene). No libepx, no CUDA context, no torch.
        if rank != 0:
            return 0
        wl = make_workload(args.workload, 0, args.gpus, pinned=False)
        total = args.warmup + args.steps
        per_step = max(0.05, args.ref_budget / max(1, total))
        runs = [cpu_baseline(wl, per_step, cores) for _ in range(total)][args.warmup:]
        pairs = sum(r["pairs"] for r in runs)
        secs = sum(r["seconds"] for r in runs)
        v = pairs / secs
        # at N > 1 the GPU arm runs one cohort of this shape per GPU (weak scaling): same per-GPU config, N x the pairs
        strong = bool(wl["w"].get("shard"))
        cfg = config_of(wl, args.workload, args.gpus, wl["total_pairs"])
        line = {"impl": "reference", "metric": "gene-probe pairs/s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(runs)),
                "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": runs[-1]["sample"]},
                "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    from epimethex_b200 import _native

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # a CPU-side group for parking ranks: an NCCL barrier is a kernel that spins on the waiting ranks' GPUs
        cpu_group = dist.new_group(backend="gloo")
    wl = make_workload(args.workload, rank, world)
    idx, n, G, P = wl["idx"], wl["n"], wl["G"], wl["P"]
    strong = bool(wl["w"].get("shard"))
    Np = idx.n_pairs
    opts = _native.make_options(True, wl["w"]["te"], wl["w"]["tm"])
    sess = _native.Session(local_rank)
    wl["expr_pinned"] = _native.pinned_empty(wl["expr"].shape)
    wl["expr_pinned"][...] = wl["expr"]
    sess.set_methylation(wl["meth"])
    sess.set_expression(wl["expr"])
    sess.set_index(idx.row_ptr, idx.probe_idx)
    # an explicit (non-default) torch stream: its handle is what the library launches on, so
    # torch.cuda.Event timings below see exactly the library's kernels
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # Result gather (the path's only exchange step): per-rank pair/gene tables -> the step's collecting rank over
    # NCCL, on a second stream. The gather of step i overlaps the gene kernels and the probe ranking of step i+1 (they do not touch
    # the pair table); the pair kernel of step i+1 waits for it. The timed region ends only when the last gather
    # has landed.
    gather_plan = None
    gstream = torch.cuda.Stream() if world > 1 else None
    step_no, last_root, rotate = [0], [0], [True]
    # "lanes": one NCCL communicator + stream + table copies per collecting rank, so that the gathers of consecutive steps
    # (different collectors) are in flight at the same time. One communicator serialises its collectives, and one gather
    # lasts as long as its collector needs to take in N - 1 tables (1.6 ms at N = 8 whoever collects: measured, r02X).
    lanes = {"on": False, "pg": [], "st": [], "pc": [], "gc": []}
    if world > 1 and not strong:
        lanes["pg"] = [dist.new_group(backend="nccl") for _ in range(world)]
        lanes["st"] = [torch.cuda.Stream() for _ in range(world)]

    def step():
        nonlocal gather_plan
        if world == 1:
            sess.run(opts, stream)
            return
        if gather_plan is not None:
            tstream.wait_event(gather_plan[4])    # the previous gene-table gather (small) has read the gene table
        sess.run_prepare(opts, stream)
        if gather_plan is None:
            ptr, _ = sess.result_device_ptr(0)
            pair_t = torch.as_tensor(_DevArray(ptr, (Np, 11)), device="cuda")   # the library's table, no copy
            gptr, _ = sess.result_device_ptr(4)
            gene_t = torch.as_tensor(_DevArray(gptr, (G, 15)), device="cuda")
            if strong:
                # ragged shards: every rank's table sizes, once; rank 0 then receives exact-size tables point to point
                mine = torch.tensor([Np, G], device="cuda", dtype=torch.int64)
                sizes = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(sizes, mine)
                sizes = [(int(c[0]), int(c[1])) for c in sizes]
                pdst = [torch.empty((a, 11), dtype=torch.float64, device="cuda") for a, _ in sizes] if rank == 0 else None
                gdst = [torch.empty((b, 15), dtype=torch.float64, device="cuda") for _, b in sizes] if rank == 0 else None
            else:
                # every rank can collect: the collector of a step is rank (step mod N), see below
                pdst = [torch.empty_like(pair_t) for _ in range(world)]
                gdst = [torch.empty_like(gene_t) for _ in range(world)]
            # the gather reads a COPY of the pair table (41.8 MB device to device, ~15 us): the pair kernel of the next
            # step then never waits for the previous gather, only the next copy does -- a whole step later
            gather_plan = (pair_t, pdst, gene_t, gdst, torch.cuda.Event(), torch.empty_like(pair_t))
        pair_t, pdst, gene_t, gdst, ev_gene, pair_copy = gather_plan
        sess.run_pairs(0, 1, stream)
        if lanes["on"]:
            ln = step_no[0] % world
            step_no[0] += 1
            last_root[0] = ln
            if not lanes["pc"]:
                lanes["pc"] = [torch.empty_like(pair_t) for _ in range(world)]
                lanes["gc"] = [torch.empty_like(gene_t) for _ in range(world)]
            ls = lanes["st"][ln]
            tstream.wait_stream(ls)               # the lane's previous gather (N steps ago) has read its copies
            lanes["pc"][ln].copy_(pair_t)
            lanes["gc"][ln].copy_(gene_t)
            ev_gene.record(tstream)               # the gene table is free for the next run_prepare
            ls.wait_stream(tstream)
            with torch.cuda.stream(ls):
                dist.gather(lanes["gc"][ln], gdst if rank == ln else None, dst=ln, group=lanes["pg"][ln])
                dist.gather(lanes["pc"][ln], pdst if rank == ln else None, dst=ln, group=lanes["pg"][ln])
            return
        tstream.wait_stream(gstream)              # the previous gather has read the copy
        pair_copy.copy_(pair_t)
        pair_t = pair_copy
        gstream.wait_stream(tstream)
        with torch.cuda.stream(gstream):
            if strong:
                if rank == 0:
                    ops = [dist.P2POp(dist.irecv, gdst[r], r) for r in range(1, world)]
                    ops += [dist.P2POp(dist.irecv, pdst[r], r) for r in range(1, world)]
                    gdst[0].copy_(gene_t)
                    pdst[0].copy_(pair_t)
                else:
                    ops = [dist.P2POp(dist.isend, gene_t, 0), dist.P2POp(dist.isend, pair_t, 0)]
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                ev_gene.record(gstream)
            else:
                # weak scaling = one cohort per GPU and step; the N tables of a step are collected on ONE rank, and that
                # rank rotates with the step: with rank 0 collecting every step its inbound 7 x 41.8 MB per 1.45 ms
                # step (~200 GB/s through NCCL send/recv) set the pace (N = 8: 85 % of 8 x one GPU, r02D)
                root = step_no[0] % world if rotate[0] else 0
                step_no[0] += 1
                last_root[0] = root
                dist.gather(gene_t, gdst if rank == root else None, dst=root)
                ev_gene.record(gstream)
                dist.gather(pair_t, pdst if rank == root else None, dst=root)

    def finish():
        if world > 1:
            tstream.wait_stream(gstream)
            for ls in lanes["st"]:
                tstream.wait_stream(ls)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    # the collecting rank rotates: every rank must have collected once before the clock starts (NCCL sets up the
    # send/recv connections of a (sender, collector) pair at their first use: ~1 s each, 131 ms per step over 50 steps
    # when five of the eight collectors met theirs inside the timed region, r02T)
    while world > 1 and not strong and step_no[0] < world:
        step()
    finish()
    torch.cuda.synchronize()
    collector = None
    if world > 1 and not strong:
        # untimed trial of the three ways to collect (N warm + N timed steps each); the fastest one is used for the
        # measurement, so a fabric on which one of them misbehaves cannot decide the figure
        trial = {}
        for mode in ("rank 0", "rotating (rank step mod N)", "rotating, one communicator per collector"):
            rotate[0] = mode != "rank 0"
            lanes["on"] = mode.endswith("collector")
            for _ in range(world):
                step()
            finish()
            torch.cuda.synchronize()
            dist.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(2 * world):
                step()
            finish()
            t1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([t0.elapsed_time(t1)], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            trial[mode] = float(tt.item()) / (2 * world)
        best = min(trial, key=lambda k: trial[k])
        rotate[0] = best != "rank 0"
        lanes["on"] = best.endswith("collector")
        collector = {"trial_ms_per_step": {k: round(v, 4) for k, v in trial.items()}, "chosen": best}
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern = {"gene_sort_ms": 0.0, "gene_tests_ms": 0.0, "probe_rank_ms": 0.0, "pair_stats_ms": 0.0}
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step()
    finish()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    # N > 1: rank 0 checks every gathered shard against a bit-exact checksum taken by the rank that computed it
    # (64-bit wrap-around sum of the table's bit patterns: order-independent, so it is exact)
    shards = None
    if world > 1:
        pair_t, pdst, gene_t, gdst = gather_plan[:4]
        mine = torch.stack([pair_t.view(torch.int64).sum(), gene_t.view(torch.int64).sum(),
                            torch.tensor(pair_t.shape[0], device="cuda"), torch.tensor(gene_t.shape[0], device="cuda")])
        sums = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        verdict = torch.zeros(1, dtype=torch.int64, device="cuda")
        if rank == last_root[0]:             # the collector of the last step checks what it received
            bad = []
            for r in range(world):
                want = [int(v) for v in sums[r].tolist()]
                have = [int(pdst[r].view(torch.int64).sum()), int(gdst[r].view(torch.int64).sum()), pdst[r].shape[0], gdst[r].shape[0]]
                if want != have:
                    bad.append(r)
            verdict[0] = sum(1 << r for r in bad)
        dist.broadcast(verdict, src=last_root[0])
        bad = [r for r in range(world) if (int(verdict[0]) >> r) & 1]
        shards = {"ranks_checked": world, "mismatched_ranks": bad, "ok": not bad, "collector_of_the_last_step": last_root[0],
                  "how": "64-bit wrap-around sum of the bit patterns of the pair and gene tables, sender vs the collector's copy"}
    # per-kernel CUDA-event times of the last step (events recorded by the library on the same stream)
    tim = sess.timing()
    # per-kernel times: a few more steps with the kernels serialised (no overlap of the gene kernels with the probe
    # ranking), each read back right after it ran
    reps = min(5, args.steps)
    opts_serial = _native.make_options(True, wl["w"]["te"], wl["w"]["tm"], serialize=True)
    for _ in range(reps):
        sess.run(opts_serial, stream)
        t = sess.timing()
        for k in kern:
            kern[k] += t[k] / reps
    pooled = None
    if args.pooled and world == 1:
        from epimethex_b200.annotation import build_groups
        gi = build_groups(idx, args.pooled)
        sess.run(opts, stream)
        for _ in range(2):
            sess.run_groups(gi.group_ptr, gi.group_pairs, stream)
        torch.cuda.synchronize()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        psteps = 5
        pe0.record()
        for _ in range(psteps):
            sess.run_groups(gi.group_ptr, gi.group_pairs, stream)
        pe1.record()
        torch.cuda.synchronize()
        pms = pe0.elapsed_time(pe1) / psteps
        sizes = np.diff(gi.group_ptr)
        pooled = {"by": args.pooled, "groups": int(gi.n_groups), "stacked_probes": int(gi.group_pairs.size),
                  "largest_group": int(sizes.max()) if sizes.size else 0, "ms_per_pass": round(pms, 3),
                  "stacked_probes_per_s": gi.group_pairs.size / (pms * 1e-3), "includes": "host descriptor build + upload"}
        if not args.no_cpu_baseline:
            # the oracle's pooled pass on every s-th group (bounded), all host cores
            import oracle
            stride = max(1, int(gi.group_pairs.size // 60000))
            pick = np.arange(0, gi.n_groups, stride)
            gp = np.zeros(pick.size + 1, np.int64)
            np.cumsum(sizes[pick], out=gp[1:])
            gl = np.concatenate([gi.group_pairs[gi.group_ptr[q]:gi.group_ptr[q + 1]] for q in pick]).astype(np.int32)
            t0 = time.perf_counter()
            oracle.analyze_groups(wl["expr"], wl["meth"], idx.row_ptr, idx.probe_idx, gp, gl, test_meth_t=wl["w"]["tm"], threads=cores)
            dt = time.perf_counter() - t0
            pooled["cpu_baseline"] = {"value": gl.size / dt, "unit": "stacked probes/s", "cores": cores, "kind": "port",
                                      "sample": f"every {stride}th group: {pick.size} groups, {gl.size} stacked probes, {dt:.1f} s"}
    clocks = sampler.stop()
    if world > 1:
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
        tot = torch.tensor([float(Np)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot)
        total_pairs = float(tot.item())
    else:
        total_pairs = float(Np)
    ms_per_step = ms / args.steps
    value = total_pairs / (ms_per_step * 1e-3)

    # ---- e2e through the blocking C-ABI call with host buffers ----
    # methylation + expression + index go over PCIe every step. The reference is called once per gene range with a new
    # annotation join every time (README.md:90), so consecutive steps alternate between two indices (the workload's, and
    # the same one without the pairs of its last gene that has any): epx_session_set_index sees a different index on
    # every call and rebuilds its unique-probe slots and work items inside the timed region.
    rp_alt = idx.row_ptr.copy()
    last = int(np.nonzero(np.diff(idx.row_ptr) > 0)[0][-1]) if Np else 0
    if Np:
        rp_alt[last + 1:] = rp_alt[last]
    pi_alt = np.ascontiguousarray(idx.probe_idx[:int(rp_alt[-1])])
    variants = [(idx.row_ptr, idx.probe_idx), (rp_alt, pi_alt)]
    h2d = wl["rows"] * n * 8 + G * n * 8 + (G + 1) * 8 + Np * 4
    d2h = Np * (11 * 8 + 2) + G * (15 * 8 + 2)
    e2e_ms = []
    e2e_upload_ms = []
    e2e_pairs = []
    e2e_call_ms = []
    for i in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):
        rp_i, pi_i = variants[i & 1]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sess.set_methylation(wl["meth"])
        t1 = time.perf_counter()
        sess.analyze(wl["expr_pinned"], rp_i, pi_i, opts, full=False, pinned_results=True)
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_ms.append(dt * 1e3)
            e2e_upload_ms.append((t1 - t0) * 1e3)
            e2e_pairs.append(int(rp_i[-1]))
            e2e_call_ms.append((dt - (t1 - t0)) * 1e3)
    e2e_t = float(np.median(e2e_ms)) if e2e_ms else float("nan")  # median: host-side outliers (page faults, PCIe neighbours)
    if args.e2e_steps <= 0:
        e2e_t = float("nan")
    e2e_frac = float(np.mean(e2e_pairs)) / Np if e2e_pairs and Np else 1.0   # pairs actually analysed per e2e step
    e2e_analyze_ms = float(np.median(e2e_call_ms)) if e2e_call_ms else float("nan")   # the epx_analyze call alone (matrix resident)
    # The same step through ONE call, epx_analyze_host: the matrix stays in (page-locked) host memory and the device
    # fetches only the rows the index names -- the reference's own methylation filter (R/EpimethexAnalysis.R:278-282)
    # folded into the upload. Same inputs, same alternating indices, same result tables; the bytes that cross the bus
    # are counted from the rows each index names.
    host_ms = []
    host_h2d = float("nan")
    host_error = None
    if args.e2e_steps > 0 and wl["pinned"] and not wl["w"].get("shard") and not args.no_analyze_host:
        uniq_rows = [int(np.unique(pi[pi >= 0]).size) for _, pi in variants]
        try:
            for i in range(args.e2e_steps + 1):
                rp_i, pi_i = variants[i & 1]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                sess.analyze_host(wl["meth"], wl["expr_pinned"], rp_i, pi_i, opts, full=False, pinned_results=True)
                if i > 0:
                    host_ms.append((time.perf_counter() - t0) * 1e3)
            host_h2d = float(np.mean([uniq_rows[i & 1] for i in range(1, args.e2e_steps + 1)])) * n * 8 + G * n * 8 + (G + 1) * 8 + Np * 4
        except _native.EpxError as ex:      # the two-call figure stands; the line says why this one is missing
            host_ms, host_error = [], str(ex)
        sess.set_methylation(wl["meth"])   # untimed: the legs below expect the whole matrix on the device
    host_t = float(np.median(host_ms)) if host_ms else float("nan")
    if world > 1:
        tmax = torch.tensor([e2e_t, host_t if host_t == host_t else 1e30], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_t = float(tmax[0].item())
        host_t = float(tmax[1].item()) if float(tmax[1].item()) < 1e29 else float("nan")
    e2e_flows = {"epx_session_set_methylation_rows + epx_analyze": {
        "ms_per_step": round(e2e_t, 3) if e2e_t == e2e_t else None, "h2d_bytes_per_step": int(h2d),
        "ms_each": [round(v, 2) for v in e2e_ms],
        "methylation_upload_ms": round(float(np.median(e2e_upload_ms)), 2) if e2e_upload_ms else None,
        "analyze_call_ms": round(e2e_analyze_ms, 2) if e2e_analyze_ms == e2e_analyze_ms else None}}
    e2e_call = "epx_session_set_methylation_rows + epx_analyze"
    if host_t == host_t:
        e2e_flows["epx_analyze_host"] = {"ms_per_step": round(host_t, 3), "h2d_bytes_per_step": int(host_h2d),
                                         "ms_each": [round(v, 2) for v in host_ms],
                                         "rows_fetched_gbs": round((host_h2d - G * n * 8) / max(host_t - e2e_analyze_ms, 1e-3) / 1e6, 1)}
        if host_t < e2e_t:   # the headline is the faster of the two public calls; both are in `flows`
            e2e_call, e2e_t, h2d, e2e_ms = "epx_analyze_host", host_t, host_h2d, host_ms
    elif host_error:
        e2e_flows["epx_analyze_host"] = {"error": host_error}
    e2e_value = total_pairs * e2e_frac / (e2e_t * 1e-3)

    # ---- parity: the GPU's tables against the oracle's on the same inputs (the whole workload when the oracle does it
    # within --cpu-seconds, else every k-th gene), through the same blocking C-ABI call ----
    cb = parity = None
    if not args.no_cpu_baseline and rank == 0:
        cb = cpu_baseline(wl, args.cpu_seconds if world == 1 else min(args.cpu_seconds, 4.0), cores, keep=True)
        sel, ptr_s, pi_s, ref = cb.pop("oracle")
        sess.set_methylation(wl["meth"])
        got = sess.analyze(wl["expr"][sel], ptr_s, pi_s, opts, full=True)
        parity = parity_report(got, ref, cb["sample"].split(";")[0])
        if wl["pinned"] and not wl["w"].get("shard") and not args.no_analyze_host:
            # the one-call path on the same inputs: its kernels see the same rows, so the tables must be the same bits
            got_h = sess.analyze_host(wl["meth"], wl["expr"][sel], ptr_s, pi_s, opts, full=True)
            parity["analyze_host_tables_identical"] = bool(
                got_h.counts == got.counts and got_h.ref_gene == got.ref_gene and
                all(np.array_equal(getattr(got_h, k), getattr(got, k), equal_nan=True) for k in ("pair", "pair_stat", "gene")) and
                all(np.array_equal(getattr(got_h, k), getattr(got, k)) for k in ("ks_dnum", "gene_ks_dnum", "pair_na", "gene_na", "perm")))
            parity["ok"] = bool(parity["ok"] and parity["analyze_host_tables_identical"])
            del got_h
            sess.set_methylation(wl["meth"])
        del got, ref
    # the other mode of the methylation test on the same inputs (north_star: "t-test and KS modes")
    modes = None
    if world == 1 and not wl["w"].get("shard") and args.other_mode:
        tm2 = not wl["w"]["tm"]
        opts2 = _native.make_options(True, wl["w"]["te"], tm2, serialize=True)
        sess.set_methylation(wl["meth"])
        sess.set_expression(wl["expr"])
        sess.set_index(idx.row_ptr, idx.probe_idx)
        k2 = {"probe_rank_ms": 0.0, "pair_stats_ms": 0.0, "total_ms": 0.0}
        for i in range(3 + 5):
            sess.run(opts2, stream)
            t = sess.timing()
            if i >= 3:
                for k in k2:
                    k2[k] += t[k] / 5
        pk = peaks()[0]
        modes = {("tt" if tm2 else "ks"): {
            "methylation_test": "Welch t" if tm2 else "Kolmogorov-Smirnov", "ms_per_step_serialised": round(k2["total_ms"], 4),
            "pairs_per_s": Np / (k2["total_ms"] * 1e-3), "pair_stats_ms": round(k2["pair_stats_ms"], 4),
            "pair_stats_frac": round(Np * (8 * n + 88) / (k2["pair_stats_ms"] * 1e-3) / 1e9 / pk, 4),
            "probe_rank_ms": round(k2["probe_rank_ms"], 4)}}
        if not args.no_cpu_baseline:
            sel2, ptr2, pi2 = gene_sample(idx, G, max(1, G // 400))
            import oracle
            ref2 = oracle.analyze(wl["expr"][sel2], wl["meth"], ptr2, pi2, test_expr_t=wl["w"]["te"], test_meth_t=tm2, threads=cores)
            got2 = sess.analyze(wl["expr"][sel2], ptr2, pi2, _native.make_options(True, wl["w"]["te"], tm2), full=True)
            modes["tt" if tm2 else "ks"]["parity"] = parity_report(got2, ref2, f"{sel2.size} genes, {int(ptr2[-1])} pairs")

    # ---- the multi-GPU path the R API can reach: ONE process drives all N GPUs (R is a single process;
    # epx_multi_analyze replaces the PSOCK cluster of R/EpimethexAnalysis.R:333-352). The ranks have finished their own
    # measurement; they free their GPUs and wait while rank 0 analyses ITS cohort once more across all N devices: genes
    # cut into N ranges of equal pair counts, every device writes its rows straight into the caller's host tables. ----
    one_process = None
    if world > 1 and not strong and args.one_process:
        single = sess.analyze(wl["expr"], idx.row_ptr, idx.probe_idx, opts, full=True) if rank == 0 else None
        sess.close()
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)    # the other ranks wait on the CPU: their GPUs must be idle for rank 0's measurement
        if rank == 0:
            try:
                os.sched_setaffinity(0, range(os.cpu_count() or 1))   # one process, N GPUs: no NUMA binding any more
            except OSError:
                pass
            t0 = time.perf_counter()
            ms_ = _native.MultiSession(list(range(world)))
            ms_.set_methylation(wl["meth"])
            up_s = time.perf_counter() - t0
            calls = []
            for i in range(1 + 6):   # the same call as the single-GPU e2e leg: two result tables into page-locked buffers,
                rp_i, pi_i = variants[i & 1]   # and a different index on every call (rebuilt inside the call)
                t0 = time.perf_counter()
                ms_.analyze(wl["expr_pinned"], rp_i, pi_i, opts, full=False, pinned_results=True)
                calls.append((time.perf_counter() - t0) * 1e3)
            multi = ms_.analyze(wl["expr_pinned"], idx.row_ptr, idx.probe_idx, opts)   # every table, for the comparison
            same = all(np.array_equal(getattr(single, k), getattr(multi, k), equal_nan=True)
                       for k in ("pair", "pair_stat", "gene")) and all(
                np.array_equal(getattr(single, k), getattr(multi, k)) for k in ("ks_dnum", "pair_na", "gene_na", "perm"))
            call_ms = float(np.median(calls[1:]))
            one_process = {"what": "epx_multi_analyze: one host process, %d GPUs, one cohort (strong scaling of a single call), "
                                   "host buffers in and out, methylation replicas resident" % world,
                           "ms_per_call": round(call_ms, 3), "pairs_per_s": Np / (call_ms * 1e-3), "calls_ms": [round(v, 2) for v in calls[1:]],
                           "methylation_upload_s": round(up_s, 3), "single_gpu_call_ms": None,
                           "identical_to_single_gpu_tables": bool(same)}
            # the same whole-genome call with the matrix still in host memory: every device fetches, over its own PCIe
            # link, the rows its gene range names (epx_multi_analyze_host) -- end to end for ONE cohort on N GPUs
            if wl["pinned"] and not args.no_analyze_host:
                try:
                    hcalls = []
                    for i in range(1 + 5):
                        rp_i, pi_i = variants[i & 1]
                        t0 = time.perf_counter()
                        ms_.analyze_host(wl["meth"], wl["expr_pinned"], rp_i, pi_i, opts, full=False, pinned_results=True)
                        hcalls.append((time.perf_counter() - t0) * 1e3)
                    mh = ms_.analyze_host(wl["meth"], wl["expr_pinned"], idx.row_ptr, idx.probe_idx, opts)
                    same_h = all(np.array_equal(getattr(single, k), getattr(mh, k), equal_nan=True)
                                 for k in ("pair", "pair_stat", "gene")) and all(
                        np.array_equal(getattr(single, k), getattr(mh, k)) for k in ("ks_dnum", "pair_na", "gene_na", "perm"))
                    h_ms = float(np.median(hcalls[1:]))
                    one_process["analyze_host"] = {
                        "what": "epx_multi_analyze_host: the same call with the methylation matrix in page-locked host memory, "
                                "every device fetches the rows of its own gene range (end to end, nothing resident)",
                        "ms_per_call": round(h_ms, 3), "pairs_per_s": Np / (h_ms * 1e-3), "calls_ms": [round(v, 2) for v in hcalls[1:]],
                        "single_gpu_call_ms": round(host_t, 3) if host_t == host_t else None,
                        "identical_to_single_gpu_tables": bool(same_h)}
                    del mh
                except _native.EpxError as ex:
                    one_process["analyze_host"] = {"error": str(ex)}
            ms_.close()
        dist.barrier(group=cpu_group)

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes per launch: pair kernel N_p*(8n+88); rank kernel U*(8n read + 2n written)
        alg = {"pair_stats_ms": Np * (8 * n + 88), "probe_rank_ms": wl["n_unique"] * (8 * n + 2 * n),
               "gene_sort_ms": G * (8 * n + 2 * n + 8 * n + n), "gene_tests_ms": G * (10 * n + 15 * 8)}
        kernels = {}
        for k, t in kern.items():
            gbs = alg[k] / (t * 1e-3) / 1e9 if t > 0 else 0.0
            kernels[k[:-3]] = {"ms": round(t, 4), "algorithmic_bytes": int(alg[k]), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        # rows whose bucket placement failed its check and went through the full sorting network (a performance
        # counter of the rank kernel, cumulative over the session's runs; the results are the same either way)
        runs = args.warmup + args.steps + reps
        kernels["probe_rank"]["rows_through_full_network_per_run"] = round(int(tim.get("sort_full_rows", 0)) / max(runs, 1), 1)
        dom = max(kern, key=lambda k: kern[k])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload, {}).get(dom[:-3])
        line = {
            "metric": "gene-probe pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, args.workload, world, int(total_pairs)),
            "roofline": {"bound": "hbm", "kernel": dom[:-3], "achieved": kernels[dom[:-3]]["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kernels[dom[:-3]]["frac"], "traffic": traffic, "peak_source": peak_src,
                         "traffic_source": "ncu --set full, profiles/traffic.json" if traffic else None},
            "kernels": kernels,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "call": e2e_call, "flows": e2e_flows, "index_rebuilt_every_step": True,
                    "ms_per_step": e2e_t, "steps": len(e2e_ms), "ms_each": [round(v, 2) for v in e2e_ms],
                    "pinned_host_buffers": bool(wl["pinned"]),
                    "h2d_gbs": round(h2d / (e2e_t * 1e-3) / 1e9, 1) if e2e_t == e2e_t else None},
            "gpu_launches": int(tim["kernel_launches"]) * args.steps,
            "numa_node_of_rank0": numa,
            "clocks": clocks,
            "setup_s": round(wl["gen_s"], 1),
        }
        if pooled:
            line["pooled"] = pooled
        if parity is not None:
            line["parity"] = parity
        if shards is not None:
            line["gathered_shards"] = shards
        if collector:
            line["collector"] = collector
        if modes:
            line["modes"] = modes
        if one_process:
            one_process["single_gpu_call_ms"] = round(e2e_analyze_ms, 3) if e2e_analyze_ms == e2e_analyze_ms else None
            line["one_process"] = one_process
        if cb is not None and world == 1:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        assert int(total_pairs) == wl["total_pairs"], (total_pairs, wl["total_pairs"])
        print(json.dumps(line))
    sess.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


class _DevArray:
    """minimal __cuda_array_interface__ holder so torch can view a libepx device table without copying"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


if __name__ == "__main__":
    sys.exit(main())
