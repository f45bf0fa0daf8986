"""Development aid: which branch of the keyframe map update ends last? Stamps of the last frame: start of the surf-map and of the edge-map
grid_scatter kernels relative to the write-back (FLOAM_DBG_TIMELINE words 5..7), sampled over frames of a replay."""
import sys
import numpy as np
sys.path.insert(0, ".")
from floam_b200 import capi, synth
seq = synth.Sequence("hdl64", seed=0)
scans, off = seq.scans(0, 120)
ctx = capi.Context(num_lines=64, loss="cauchy", max_scan_points=seq.max_points + 1024, max_map_points=1 << 21, max_global_map_points=0, max_grid_cells=1 << 23)
ctx.stage_scans(scans, off)
ctx.replay_staged(0, 60)
rows = []
for f in range(60, 120):
    ctx.replay_staged(f, 1)
    t = ctx.debug_fetch(capi.DBG_TIMELINE, np.int64)
    rows.append(((t[6] - t[5]) / 1e3, (t[7] - t[5]) / 1e3, (t[5] - t[4]) / 1e3))
r = np.array(rows)
ne, ns = ctx.odom_map_sizes()
print("maps edge %d surf %d; write-back -> start of last kernel: surf branch %.1f us (p50), edge branch %.1f us (p50); edge later in %d of %d frames; solve part %.1f us"
      % (ne, ns, np.median(r[:, 0]), np.median(r[:, 1]), int((r[:, 1] > r[:, 0]).sum()), len(r), np.median(r[:, 2])))
