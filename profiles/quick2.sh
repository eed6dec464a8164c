#!/bin/bash
# quick GPU check while iterating on a kernel: a few parity tests first, then the default bench
# usage (under gpurun): bash profiles/quick2.sh <tag> [extra bench args]
TAG=$1; shift
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 50 --e2e-steps 2 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms/step", round(d["ms_per_step"], 4), {k: v["ms"] for k, v in d["kernels"].items()}, "pair frac", d["kernels"]["pair_stats"]["frac"])
print("rank", d["kernels"]["probe_rank"])
m = d.get("modes") or {}
for k, v in m.items():
    print(k, "pair", v["pair_stats_ms"], "frac", v["pair_stats_frac"], "step", v["ms_per_step_serialised"], "parity", v.get("parity", {}).get("ok"))
print("e2e ms", d["e2e"]["ms_per_step"], "parity", d.get("parity", {}).get("ok"), d.get("parity", {}).get("pairs_checked"))
PY
