# quick check of the long-row kernels: GPU parity tests + the 2,000-gene pan-cancer slice. usage: bash profiles/quick_pancan.sh <tag>
TAG=${1:-r01x}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload pancan_slice --steps 50 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${TAG}_bench_pancan_slice.json 2> gpurun_out/bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_pancan_slice.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["setup_s"], {k: v["ms"] for k, v in d["kernels"].items()})
PY
tail -3 gpurun_out/bench.err
