# Main path evidence for one round tag: parity tests, default bench, reference arm, per-launch ncu list, one ncu --set full
# capture of the four kernels of a step. usage: bash profiles/prof_main.sh r01m
TAG=${1:-r01x}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -1 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench.err || exit 1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(pair|probe|gene|pick)" -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_(pair_stats2|probe_rank_w|gene_tests_w)" -s 12 -c 4 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-300
