#!/bin/bash
# final evidence call of round 2: ncu launch list of the bench command, `--set full` rows of the kernels added or changed late in the round
set -o pipefail
TAG=${1:-r2i}
NCU="ncu --clock-control none"
B="python bench.py --steps 20 --warmup 3 --timing-frames 5 --no-cpu-baseline --sequences-per-gpu 0"
$B > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err || { echo "plain bench failed"; tail -5 gpurun_out/${TAG}_bench_short.err; exit 1; }
$NCU --metrics gpu__time_duration.sum -s 1200 -c 160 --csv --log-file gpurun_out/${TAG}_launches_ncu.csv $B > /dev/null 2>&1
echo "bench launch list: $(wc -l < gpurun_out/${TAG}_launches_ncu.csv) lines"
W="python tools/profile_workload.py 24"
$W > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain workload failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
$NCU --set full --import-source on -k regex:"voxel_classify|voxel_merge|voxel_rank|grid_count|grid_scatter|dirty|sector_kernel" -s 200 -c 44 -f -o gpurun_out/${TAG}_late $W > gpurun_out/${TAG}_late.log 2>&1
ncu -i gpurun_out/${TAG}_late.ncu-rep --page raw --csv > gpurun_out/${TAG}_late_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_late.ncu-rep
echo "late kernels: $(wc -l < gpurun_out/${TAG}_late_raw.csv) csv lines"
