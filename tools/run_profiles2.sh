#!/bin/bash
# second evidence call of round 2: the launch list of the whole profiling workload and the `--set full` rows of the voxel / sort / grid kernels
set -o pipefail
TAG=${1:-r2}
W="python tools/profile_workload.py 24"
NCU="ncu --clock-control none"
$W > gpurun_out/${TAG}_plain2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain2.log; exit 1; }
$NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv $W > /dev/null 2>&1
echo "launch list: $(wc -l < gpurun_out/${TAG}_launches.csv) lines"
$NCU --set full --import-source on -k regex:"voxel_bbox|voxel_keys|voxel_rank|voxel_reduce|radix_hist|radix_scatter|grid_count|grid_scatter|scan_add|single_block_scan|predict|mail_state" -s 700 -c 56 -f -o gpurun_out/${TAG}_voxel $W > gpurun_out/${TAG}_voxel.log 2>&1
ncu -i gpurun_out/${TAG}_voxel.ncu-rep --page raw --csv > gpurun_out/${TAG}_voxel_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_voxel.ncu-rep
echo "voxel: $(wc -l < gpurun_out/${TAG}_voxel_raw.csv) csv lines"
