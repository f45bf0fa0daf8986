#!/bin/bash
# ncu evidence for one build (run under gpurun; ONE call): plain run first, then the launch list, then `--set full` captures per kernel
# group.  Every report is exported to CSV on the box (raw page = all metrics per launch; the source page of the two kNN variants) and
# the .ncu-rep files are removed again: gpurun only brings back 64 MiB.  tools/ncu_summary.py condenses the CSVs for profiles/.
set -o pipefail
TAG=${1:-r2}
W="python tools/profile_workload.py 24"
NCU="ncu --clock-control none"
$W > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log
$NCU --metrics gpu__time_duration.sum -s 2600 -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv $W > /dev/null 2>&1
FULL="$NCU --set full --import-source on"
cap() {  # name, env, regex, skip, count
  env $2 $FULL -k regex:"$3" -s $4 -c $5 -f -o gpurun_out/${TAG}_$1 $W > gpurun_out/${TAG}_$1.log 2>&1
  if [ -f gpurun_out/${TAG}_$1.ncu-rep ]; then
    ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
    if [ "$6" = "source" ]; then ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv > gpurun_out/${TAG}_$1_source.csv 2>/dev/null; fi
    rm -f gpurun_out/${TAG}_$1.ncu-rep
    echo "$1: $(wc -l < gpurun_out/${TAG}_$1_raw.csv) csv lines"
  else
    echo "$1: no report"; tail -3 gpurun_out/${TAG}_$1.log
  fi
}
cap solve_tma1 FLOAM_KNN_TMA=1 "assoc_knn|assoc_eval|lm_cluster" 246 6 source
cap knn_tma0 FLOAM_KNN_TMA=0 "assoc_knn" 82 2 source
cap front X=1 "ring_count|ring_scatter|sector|feature_offsets|feature_gather|deskew_align|unpack_pc2" 140 7
cap voxel X=1 "voxel_bbox|voxel_keys|voxel_rank|voxel_reduce|radix_hist|radix_scatter|grid_count|grid_scatter|scan_add|single_block_scan|predict|mail_state" 1300 56
cap mapping X=1 "classify_old|partition_kernel|transform_new|keys1|keys2|heads_kernel|^reduce_kernel|commit_kernel|cell_keys|^gather_kernel|crop_flags|crop_scatter|repack|knn5|compensate_velocity" 0 60
du -sh gpurun_out
