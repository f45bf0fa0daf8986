"""Development aid: device-resident replay of an HDL-64 sequence; prints ms/frame and a pose checksum (A/B runs via env vars)."""
import sys, zlib
import numpy as np
sys.path.insert(0, ".")
from floam_b200 import capi, synth
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 212
seq = synth.Sequence("hdl64", seed=0)
scans, off = seq.scans(0, frames)
ctx = capi.Context(num_lines=64, loss="cauchy", max_scan_points=seq.max_points + 1024, max_map_points=1 << int(__import__("os").environ.get("MAPBITS", "21")), max_global_map_points=0, max_grid_cells=1 << int(__import__("os").environ.get("CELLBITS", "23")))
ctx.stage_scans(scans, off)
p0, _ = ctx.replay_staged(0, 12)
res = []
for rep in range(3):
    lo = 12 + rep * ((frames - 12) // 3); n = (frames - 12) // 3
    p, ms = ctx.replay_staged(lo, n)
    res.append(ms / n)
    p0 = np.concatenate([p0, p])
print("ms/frame per third:", np.round(res, 4), " fps:", np.round(1e3 / np.array(res), 1), " pose crc %08x" % (zlib.crc32(p0.tobytes()) & 0xffffffff), " last t", np.round(p0[-1, 4:], 4))
tl = ctx.debug_fetch(capi.DBG_TIMELINE, np.int64)
if tl[3] > 0:
    print("pose-dependent half: %.1f us/frame (solve part %.1f us), gap to the next frame's %.1f us, over %d frames" % (tl[0] / tl[3] / 1e3, tl[2] / tl[3] / 1e3, tl[1] / tl[3] / 1e3, tl[3]))
