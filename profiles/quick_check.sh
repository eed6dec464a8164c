python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
python bench.py --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_default.log 2> gpurun_out/bench.err
python - <<EOF2
import json
d=json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()})
EOF2
