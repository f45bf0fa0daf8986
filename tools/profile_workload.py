"""Profiling workload (for ncu): a short run that launches EVERY kernel class of the library at HDL-64 size —
the fused frame path from raw PointCloud2 bytes with the IMU steps folded in (unpack_pc2, deskew_align, ring_*, sector,
feature_*, voxel_*, radix_*, grid_*, predict, assoc_*, lm_cluster, mail_state), the stage entry points (crop_*, repack, knn5,
compensate_velocity, the merge update's voxel_classify / voxel_merge) and LaserMappingClass (mapping.cu kernels).  FLOAM_KNN_TMA=0/1 selects the kNN variant for the A/B.
Usage: python tools/profile_workload.py [frames=24]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from floam_b200 import capi, synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 24
seq = synth.Sequence("hdl64", seed=0, distort=True)
ext = np.array([0.0, 0.0, 1.0, 6.123233995736766e-17])
ctx = capi.Context(num_lines=64, loss="cauchy", max_scan_points=seq.max_points + 1024, max_map_points=1 << 21, max_global_map_points=1 << 22,
                   max_grid_cells=1 << 23)
ctx.set_graphs(False)     # kernels launched one by one: ncu sees plain launches
for k in range(-40, 20 * frames + 40):
    t = 100.0 + 0.005 * k
    ctx.imu_push(t, seq.imu(max(t - 100.0, 0.0)))
poses = []
for f in range(frames):
    s = seq.scan(f)
    L = capi.pc2_layout(len(s), 22)
    raw = capi.pack_pointcloud2(s, L)
    ctx.process_submit_pc2(raw, L, False, int((100.0 + 0.1 * f) * 1e6), ext)
    poses.append(ctx.process_wait())
    if f >= frames - 4:       # LaserMappingClass on the filtered cloud (edge + surf), like the mapping node
        e, sf = ctx.feature_extract(s)
        T, _ = ctx.odom_get()
        ctx.mapping_update(synth.to_xyzi(np.concatenate([e, sf])), T)
m = ctx.mapping_get_map()
em, sm = ctx.odom_get_map()
c = ctx.crop_box(sm, [-30, -30, -5], [30, 30, 5])
v = ctx.voxel_grid(sm, 0.8)
u = ctx.voxel_grid_update(v, sm[:2000], 0.8)          # the large-map keyframe filter: voxel_classify + voxel_merge
changed, cells = ctx.mapping_get_changed_cells()
ids, d2 = ctx.knn5(sm, sm[:2000])
pts = seq.scan(0); ctx.compensate_velocity(pts, [1.0, 0.0, 0.0])
print("profile workload ok: %d frames, pose[-1] %s, maps %d/%d, global map %d, crop %d, voxel %d" % (frames, np.round(poses[-1], 4), len(em), len(sm), len(m), len(c), len(v)))
ctx.close()
