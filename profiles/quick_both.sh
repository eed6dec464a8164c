# quick check: GPU parity tests, default bench (no e2e / cpu legs), pan-cancer slice. usage: bash profiles/quick_both.sh <tag>
TAG=${1:-r01x}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
for w in ${WL:-skcm_full pancan_slice}; do
timeout 300 python bench.py --workload $w --steps 100 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${TAG}_quick_$w.json 2> gpurun_out/bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_quick_$w.json").read().strip().splitlines()[-1])
print("$w", round(d["ms_per_step"], 4), d["setup_s"], {k: v["ms"] for k, v in d["kernels"].items()})
PY
tail -3 gpurun_out/bench.err
done
