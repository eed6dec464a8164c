# epx_analyze_host evidence for one tag: the GPU suite (the new tests first), the default bench (e2e flows), the phase
# times of three calls. usage: bash profiles/host_call.sh r02H
TAG=${1:-r02H}
timeout 120 python -m pytest tests/test_gpu_analyze_host.py -m gpu -q > gpurun_out/${TAG}_pytest_host.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_host.log
timeout 200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_analyze_host.py > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.err
TAG=$TAG EPX_TRACE=1 python - <<'PY' 2>&1 | tail -12
import json, os
try:
    d = json.loads(open("gpurun_out/%s_bench_default.json" % os.environ["TAG"]).read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step")}), json.dumps(d["e2e"]["flows"]), d["e2e"]["call"], d.get("parity", {}).get("ok"), d.get("parity", {}).get("analyze_host_tables_identical"))
except Exception as e:
    print("no bench line:", e)
# phase times of the one-call path (EPX_TRACE): whole genome, then a 1,000-gene range (the reference's unit of work)
import numpy as np, sys, time
sys.path.insert(0, ".")
import bench
from epimethex_b200 import _native, synth
wl = bench.make_workload("skcm_full", 0)
idx = wl["idx"]
small = synth.probe_index(wl["ann"], 0, 1000)
with _native.Session(0) as s:
    for what, x, ix in (("genome", wl["expr"], idx), ("genes 1-1000", wl["expr"][:1000], small)):
        xp = _native.pinned_empty(x.shape); xp[...] = x
        for i in range(3):
            t0 = time.perf_counter()
            s.analyze_host(wl["meth"], xp, ix.row_ptr, ix.probe_idx, full=False, pinned_results=True)
            dt = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); s.set_methylation(wl["meth"]); t1 = time.perf_counter()
        s.analyze(xp, ix.row_ptr, ix.probe_idx, full=False, pinned_results=True)
        print("%s: analyze_host %.2f ms; set_methylation %.2f + analyze %.2f ms" % (what, dt, (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3), file=sys.stderr)
PY
