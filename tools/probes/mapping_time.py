import sys, time
import numpy as np
sys.path.insert(0, ".")
from floam_b200 import capi, synth
seq = synth.Sequence("hdl64", seed=0)
scans, off = seq.scans(0, 40)
ctx = capi.Context(num_lines=64, max_scan_points=seq.max_points + 1024, max_map_points=1 << 20, max_global_map_points=1 << 22, max_grid_cells=1 << 20)
for f in range(40):
    pts = synth.to_xyzi(scans[off[f]:off[f + 1]])
    T = seq.pose(0.1 * f)
    t = time.perf_counter(); ctx.mapping_update(pts, T); dt = time.perf_counter() - t
    if f % 8 == 0 or f == 39:
        print("frame", f, "n", len(pts), "update ms %.3f" % (dt * 1e3), "map", len(ctx.mapping_get_map()))
