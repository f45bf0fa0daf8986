"""Development aid: device time of the keyframe map filter by map size, full re-sort (floam_voxel_grid on map + new) against the merge
path (floam_voxel_grid_update), kernel event pairs summed (pair overhead subtracted via the noop calibration is not available here:
both columns carry the same per-launch overhead of ~5 us; launches: 8 vs 9)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from floam_b200 import capi
rng = np.random.default_rng(3)
ctx = capi.Context(num_lines=16, max_scan_points=1 << 21, max_map_points=1 << 21, max_global_map_points=0)


def cloud(n, ext):
    p = np.zeros(n, capi.POINT_I)
    p["x"] = rng.uniform(-ext, ext, n); p["y"] = rng.uniform(-ext, ext, n); p["z"] = rng.uniform(-3, 12, n)
    return p


names_sort = ("voxel_bbox", "voxel_keys", "radix_hist", "radix_scatter", "voxel_rank", "voxel_reduce")
names_merge = ("voxel_bbox", "voxel_classify", "radix_hist", "radix_scatter", "voxel_merge", "voxel_rank", "voxel_reduce")
for m_raw in (40000, 90000, 160000, 260000, 420000, 800000, 1600000):
    filt = ctx.voxel_grid(cloud(m_raw, 100.0), 0.4)
    new = cloud(6000, 60.0)
    both = np.concatenate([filt, new])
    row = []
    for which in (0, 1):
        for _ in range(3):
            out = ctx.voxel_grid(both, 0.4) if which == 0 else ctx.voxel_grid_update(filt, new, 0.4)
        ctx.set_kernel_timing(True)
        for _ in range(10):
            out = ctx.voxel_grid(both, 0.4) if which == 0 else ctx.voxel_grid_update(filt, new, 0.4)
        t = ctx.kernel_timing(); ctx.set_kernel_timing(False)
        names = names_sort if which == 0 else names_merge
        row.append((sum(t[k][0] for k in names if k in t) * 1e3 / 10, sum(t[k][1] for k in names if k in t) / 10))
    print("map %7d + new %d: re-sort %7.1f us (%d launches)   merge %7.1f us (%d launches)" % (len(filt), len(new), row[0][0], row[0][1], row[1][0], row[1][1]), flush=True)
