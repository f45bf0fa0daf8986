"""Device-side timeline of the pose-dependent half (BACK) of a frame, from the globaltimer stamps the kernels leave in the state
(FLOAM_DBG_TIMELINE): average BACK duration (predict start -> last grid scatter), the gap between two consecutive BACKs (graph
boundary), and the solve part (predict start -> write-back). One replay call, so only the first gap spans a host call boundary.
Usage: python tools/timeline.py [sensor=hdl64] [frames=400] > profiles/rN_timeline.json"""
import json
import sys
import numpy as np
sys.path.insert(0, ".")
from floam_b200 import capi, synth
sensor = sys.argv[1] if len(sys.argv) > 1 else "hdl64"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 400
seq = synth.Sequence(sensor, seed=0)
scans, off = seq.scans(0, frames)
ctx = capi.Context(num_lines=seq.num_lines, loss="cauchy", max_scan_points=seq.max_points + 1024, max_map_points=1 << 21, max_global_map_points=0,
                   max_grid_cells=1 << 23)
ctx.stage_scans(scans, off)
warm = 40
ctx.replay_staged(0, warm)
a = ctx.debug_fetch(capi.DBG_TIMELINE, np.int64)
_, ms = ctx.replay_staged(warm, frames - warm)
b = ctx.debug_fetch(capi.DBG_TIMELINE, np.int64)
n = int(b[3] - a[3])
out = {"sensor": sensor, "frames_timed": frames - warm, "back_halves_counted": n,
       "replay_ms_per_frame": ms / (frames - warm),
       "back_us": (b[0] - a[0]) / n / 1e3, "gap_between_backs_us": (b[1] - a[1]) / n / 1e3, "solve_part_us": (b[2] - a[2]) / n / 1e3,
       "map_update_part_us": ((b[0] - a[0]) - (b[2] - a[2])) / n / 1e3,
       "note": "globaltimer (ns) stamps written by predict_kernel, the write-back and the two grid_scatter kernels; sums kept on the device"}
print(json.dumps(out))
