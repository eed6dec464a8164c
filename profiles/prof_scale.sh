# Multi-GPU evidence on one box: bash profiles/prof_scale.sh <tag> <N> [workloads...]
# pancan_full = BASELINE configs[4] (strong scaling, probe-sliced methylation); epic = configs[3] (weak scaling, one cohort per GPU)
TAG=${1:-r01x}; N=${2:-8}; shift 2
WL=${@:-pancan_full epic skcm_full}
nproc > gpurun_out/${TAG}_env_n${N}.log; free -g >> gpurun_out/${TAG}_env_n${N}.log; nvidia-smi -L >> gpurun_out/${TAG}_env_n${N}.log
for w in $WL; do
  steps=100; [ $w = pancan_full ] && steps=30
  if [ $N = 1 ]; then
    timeout 420 python bench.py --workload $w --steps $steps --warmup 3 --e2e-steps 0 --no-cpu-baseline \
      > gpurun_out/${TAG}_bench_${w}_n${N}.json 2> gpurun_out/${TAG}_bench_${w}_n${N}.err
  else
    timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --workload $w --steps $steps --warmup 3 --e2e-steps 0 --no-cpu-baseline \
      > gpurun_out/${TAG}_bench_${w}_n${N}.json 2> gpurun_out/${TAG}_bench_${w}_n${N}.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${w}_n${N}.json").read().strip().splitlines()[-1])
    print("$w", "N=$N", round(d["ms_per_step"], 3), "ms/step", round(d["value"] / 1e6, 2), "M pairs/s", "setup", d["setup_s"])
except Exception as e:
    print("$w failed:", e)
PY
  tail -2 gpurun_out/${TAG}_bench_${w}_n${N}.err
done
