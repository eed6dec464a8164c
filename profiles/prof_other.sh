# ncu --set full of the four kernels of one step on another workload (epic: 513..1024-sample kernels; pancan_slice: long-row
# CTA kernels). One workload per call: two reports exceed what gpurun copies back. usage: bash profiles/prof_other.sh <tag> <workload>
TAG=${1:-r01x}; w=${2:-epic}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_(pair_stats|probe_rank|gene_sort|gene_tests)' -s 12 -c 4 \
    -o gpurun_out/prof_${TAG}_$w -f python bench.py --workload $w --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$w.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_$w.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/prof_${TAG}_$w.ncu-rep > gpurun_out/${TAG}_ncu_${w}_summary.txt
ls -la gpurun_out
