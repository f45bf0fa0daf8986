"""Development aid: how large does the OS1-128 local map get for a given map_resolution (configs[3] asks for >= 1M points)?"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from floam_b200 import capi, synth  # noqa: E402

for res in [float(a) for a in sys.argv[1:]] or [0.1, 0.08]:
    seq = synth.Sequence("os1-128", seed=1)
    ctx = capi.Context(num_lines=128, loss="cauchy", map_resolution=res, max_distance=90.0, min_distance=0.5, max_scan_points=seq.max_points + 1024,
                       max_map_points=1 << 22, max_global_map_points=0, max_grid_cells=1 << 24)
    t0 = time.time()
    for f in range(240):
        ctx.process_scan(seq.scan(f))
        if f % 20 == 19:
            d = ctx.debug()
            print("res %.2f frame %3d map %s ds %d+%d corr %d frame_ms %.3f" % (res, f, ctx.odom_map_sizes(), len(d["ds_edge"]), len(d["ds_surf"]), d["n_corr"], ctx.last_frame_ms()), flush=True)
    print("wall %.1fs" % (time.time() - t0))
    ctx.close()
