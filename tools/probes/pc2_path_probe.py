"""Development aid: where does the raw-PointCloud2 entry path (floam_process_submit_pc2) spend its time, against the 32-byte point path?"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from floam_b200 import capi, synth  # noqa: E402

frames = 160
seq = synth.Sequence("hdl64", seed=0)
scans, off = seq.scans(0, frames)
prm = dict(num_lines=64, loss="cauchy", max_scan_points=seq.max_points + 1024, max_map_points=1 << 21, max_global_map_points=0, max_grid_cells=1 << 23)
pinned = capi.PinnedBuffer(len(scans)); pinned.array[:] = scans
lay = [capi.pc2_layout(int(off[f + 1] - off[f]), 22) for f in range(frames)]
roff = np.zeros(frames + 1, np.int64); roff[1:] = np.cumsum([L.row_step for L in lay])
rawp = capi.PinnedBuffer(int(roff[-1]) // 32 + 2); raw = rawp.array.view(np.uint8)
for f in range(frames):
    raw[roff[f]:roff[f + 1]] = capi.pack_pointcloud2(scans[off[f]:off[f + 1]], lay[f])


def run(mode, timing=False):
    ctx = capi.Context(**prm)
    if timing:
        ctx.set_kernel_timing(True)
    pend = 0; t0 = None
    for f in range(frames):
        if f == 40:
            while pend:
                ctx.process_wait(); pend -= 1
            t0 = time.perf_counter()
        if pend == 3:
            ctx.process_wait(); pend -= 1
        if mode == "pc2":
            ctx.process_submit_pc2(raw[roff[f]:roff[f + 1]], lay[f])
        else:
            ctx.process_submit(pinned.array[off[f]:off[f + 1]])
        pend += 1
    while pend:
        ctx.process_wait(); pend -= 1
    dt = time.perf_counter() - t0
    out = "%s: %.0f frames/s" % (mode, (frames - 40) / dt)
    if timing:
        kt = ctx.kernel_timing()
        out += "  unpack_pc2 %s" % (kt.get("unpack_pc2"),)
    ctx.close()
    return out


print(run("points")); print(run("pc2")); print(run("pc2", timing=True))
