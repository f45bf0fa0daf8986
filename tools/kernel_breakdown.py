"""Development aid: per-kernel-class device time for any synthetic configuration (graph-embedded event pairs, noop-calibrated).
Usage: python tools/kernel_breakdown.py [sensor=hdl64] [map_resolution=0.4] [frames=40] [max_dis=60] [min_dis=2]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from floam_b200 import capi, synth
sensor = sys.argv[1] if len(sys.argv) > 1 else "hdl64"
res = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 40
max_dis = float(sys.argv[4]) if len(sys.argv) > 4 else 60.0
min_dis = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
seq = synth.Sequence(sensor, seed=0)
scans, off = seq.scans(0, frames)
ctx = capi.Context(num_lines=seq.num_lines, loss="cauchy", map_resolution=res, max_scan_points=seq.max_points + 1024, max_map_points=1 << 22,
                   max_global_map_points=0, max_grid_cells=1 << 24, max_distance=max_dis, min_distance=min_dis)
ctx.stage_scans(scans, off)
ctx.replay_staged(0, frames - 10)
_, ms = ctx.replay_staged(frames - 10, 5)
print("replay ms/frame %.4f" % (ms / 5))
ctx.set_kernel_timing(True)
for f in range(frames - 5, frames):
    ctx.process_staged(f)
t = ctx.kernel_timing(); ctx.set_kernel_timing(False)
noop = t.pop("noop", (0, 1)); ov = noop[0] / noop[1] * 1e3
rows = sorted(((k, max(v[0] / v[1] * 1e3 - ov, 0.3), v[1] / 5) for k, v in t.items()), key=lambda r: -r[1] * r[2])
tot = sum(u * n for _, u, n in rows)
d = ctx.debug(); ne, ns = ctx.odom_map_sizes()
print("queries %d corr %d maps %d/%d  event overhead %.2f us  sum of kernels %.1f us/frame" % (len(d["ds_edge"]) + len(d["ds_surf"]), d["n_corr"], ne, ns, ov, tot))
for k, u, n in rows[:16]:
    print("%-16s %5.1f launches/frame  %8.2f us  share %.3f" % (k, n, u, u * n / tot))
