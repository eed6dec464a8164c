timeout 900 python -m pytest tests/test_gpu_groups.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --pooled gene > gpurun_out/pg.json 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_group -c 9 --csv --log-file gpurun_out/pooled_launches.csv python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --pooled gene > gpurun_out/pg_ncu.log 2>&1
grep -o "\"pooled.*" gpurun_out/pg.json
