"""Writes tests/golden/replica_pose_crc.json: CRC-32 of the first 16 poses of the bench sequences (seeds 0..7) replayed on ONE GPU.
bench.py --gpus N compares every rank's sequence against it (SURVEY.md section 4 "identical per-sequence trajectories to the 1-GPU run").
Regenerate after any change to the device arithmetic:  python tools/replica_crc.py   (needs a B200)"""
import json
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from floam_b200 import capi  # noqa: E402

FRAMES = 16
out = {"frames": FRAMES, "sensor": bench.SENSOR, "odom": bench.ODOM, "seeds": {}}
for seed in range(8):
    seq, scans, off = bench.build_sequence(seed, FRAMES)
    ctx = capi.Context(num_lines=seq.num_lines, max_scan_points=seq.max_points + 1024, max_map_points=1 << 21, max_global_map_points=0, max_grid_cells=1 << 23,
                       **bench.ODOM)
    ctx.stage_scans(scans, off)
    P, _ = ctx.replay_staged(0, FRAMES)
    ctx.close()
    out["seeds"][str(seed)] = zlib.crc32(np.ascontiguousarray(P).tobytes()) & 0xffffffff
path = os.path.join("tests", "golden", "replica_pose_crc.json")
json.dump(out, open(path, "w"), indent=1)
print(open(path).read())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.path.join("gpurun_out", "replica_pose_crc.json"), "w"), indent=1)
