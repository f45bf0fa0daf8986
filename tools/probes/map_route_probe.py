"""Which frame of a back-and-forth route makes the device's global map differ from the reference build's? (debug aid)"""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
from floam_b200 import synth
from oracle import pyoracle as po
try:
    from oracle import pyref as pr
except Exception:
    pr = None
gpu = len(sys.argv) > 1 and sys.argv[1] == "gpu"
if gpu:
    from floam_b200 import capi
seq = synth.Sequence("vlp16", seed=0)
scans, off = seq.scans(0, 12)
route = [0.0, 4.0, 8.0, 8.2, 12.0, 12.0, 6.0, 2.0, 0.5, 9.0, 16.0, 16.1]
ref = (pr or po).Mapping(map_resolution=0.4, total_order=True)
orc = po.Mapping(map_resolution=0.4, total_order=True)
ctx = None
if gpu:
    ctx = capi.Context(num_lines=16, map_resolution=0.4, max_scan_points=300000, max_map_points=1 << 21, max_global_map_points=1 << 21, max_grid_cells=1 << 22)
xyzi = lambda a: np.stack([a["x"], a["y"], a["z"], a["intensity"]], 1)
for f in range(12):
    e, sf, _, _, _ = po.feature_extract(scans[off[f]:off[f + 1]], 16, 2.0, 60.0)
    pts = synth.to_xyzi(np.concatenate([e, sf]))
    T = seq.pose(route[f])
    ref.update(pts, T); orc.update(pts, T)
    a = xyzi(ref.get_map()); b = xyzi(orc.get_map())
    line = f"frame {f} t={route[f]} pos={T[:3,3].round(2)} ref {len(a)} oracle {len(b)} equal {np.array_equal(a, b)}"
    if gpu:
        ctx.mapping_update(pts, T)
        c = xyzi(ctx.mapping_get_map())
        same = np.array_equal(a, c)
        line += f" gpu {len(c)} equal {same}"
        if not same and len(a) == len(c):
            bad = np.nonzero((a != c).any(axis=1))[0]
            sa = a[np.lexsort(a.T)]; sc = c[np.lexsort(c.T)]
            line += f" first bad row {bad[0]} of {len(bad)}; same multiset {np.array_equal(sa, sc)}; cells ref {np.floor(a[bad[0], :3] / 50 + 0.5)} gpu {np.floor(c[bad[0], :3] / 50 + 0.5)}"
    print(line)
