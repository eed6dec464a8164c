#!/bin/bash
# usage (under gpurun): bash profiles/prof_kernels.sh <tag> <workload> [kernel-regex]
# plain run first (must exit 0), then the launch list and one `ncu --set full` capture of the named kernels
# (B200_PROFILING.md recipe). Outputs: gpurun_out/<tag>_{plain.json,launches.csv,prof.ncu-rep}
set -u
TAG=$1; WL=$2; KRE=${3:-"k_pair_stats2|k_probe_rank_w|k_gene_sort_w|k_gene_tests_w|k_rows_w"}
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-other-mode"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:"$KRE" -s 2 -c ${NCAP:-1} -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "prof rc=$?"; ls -la gpurun_out/${TAG}_*
