# pooled-group pass: parity tests, timing, per-launch ncu list. usage: bash profiles/prof_pooled.sh [gene|position|island]
BY=${1:-gene}
timeout 900 python -m pytest tests/test_gpu_groups.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --pooled $BY > gpurun_out/pooled_$BY.json 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_group -c 9 --csv --log-file gpurun_out/pooled_${BY}_launches.csv python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --pooled $BY > gpurun_out/pooled_ncu.log 2>&1
grep -o "\"pooled.*" gpurun_out/pooled_$BY.json
