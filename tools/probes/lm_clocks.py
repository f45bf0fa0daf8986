"""Development aid: SM-clock breakdown of one step attempt inside lm_cluster_kernel (FLOAM_DBG_CLOCKS)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from floam_b200 import capi, synth
seq = synth.Sequence("hdl64", seed=0)
scans, off = seq.scans(0, 30)
ctx = capi.Context(num_lines=64, loss="cauchy", max_scan_points=seq.max_points + 1024, max_map_points=1 << 21, max_global_map_points=0, max_grid_cells=1 << 23)
ctx.stage_scans(scans, off)
for f in range(30):
    ctx.process_staged(f)
    if f >= 25:
        c = ctx.debug_fetch(capi.DBG_CLOCKS, np.int64)
        d = ctx.debug()
        print("frame", f, "corr", d["n_corr"], "cycles: eval %d  block-reduce %d  sync1 %d  dsmem+LM %d  sync2 %d  total %d" % (
            c[1] - c[0], c[2] - c[1], c[3] - c[2], c[4] - c[3], c[5] - c[4], c[5] - c[0]), " dev ms", ctx.last_frame_ms())
