"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/) on the deterministic synthetic generator.

The reference (dan11003/floam) ships no golden vectors and cannot be compiled in the build container (PCL / Ceres / Eigen /
FLANN / ROS absent), so these fixtures are the oracle's own outputs: they pin the oracle against regressions and give the GPU
tests an oracle-independent target.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from floam_b200 import synth          # noqa: E402
from oracle import pyoracle as po     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def sequence_golden(sensor, frames, map_resolution, loss, deskew, distort, name, n_az=None):
    seq = synth.Sequence(sensor, seed=0, distort=distort, n_az=n_az)
    scans, off = seq.scans(0, frames)
    orc = po.Odom(num_lines=seq.num_lines, map_resolution=map_resolution, loss=loss, total_order=True, use_kdtree=False)
    out = {"scan_crc": [], "edge_src_0": None, "edge_crc": [], "surf_crc": [], "n_edge": [], "n_surf": [], "poses": [], "map_sizes": []}
    for f in range(frames):
        s = scans[off[f]:off[f + 1]]
        e, sf, es, ss, ties = po.feature_extract(s, seq.num_lines, 2.0, 60.0, total_order=True)
        assert ties == 0
        if f == 0:
            out["edge_src_0"] = es.copy()
            orc.init_map(synth.to_xyzi(e), synth.to_xyzi(sf))
            pose = np.array([0, 0, 0, 1, 0, 0, 0.0])
        else:
            pose = orc.update(e.copy(), sf.copy(), deskew)
        em, sm = orc.get_map()
        out["scan_crc"].append(crc(s)); out["edge_crc"].append(crc(es)); out["surf_crc"].append(crc(ss))
        out["n_edge"].append(len(es)); out["n_surf"].append(len(ss)); out["poses"].append(pose); out["map_sizes"].append((len(em), len(sm)))
    np.savez_compressed(os.path.join(HERE, name), sensor=sensor, frames=frames, map_resolution=map_resolution, loss=loss, deskew=deskew,
                        distort=distort, n_az=n_az or 0, scan_crc=np.array(out["scan_crc"], np.uint32), edge_src_0=out["edge_src_0"],
                        edge_crc=np.array(out["edge_crc"], np.uint32), surf_crc=np.array(out["surf_crc"], np.uint32),
                        n_edge=np.array(out["n_edge"]), n_surf=np.array(out["n_surf"]), poses=np.array(out["poses"]),
                        map_sizes=np.array(out["map_sizes"]))
    print(name, "poses[-1]", out["poses"][-1])


if __name__ == "__main__":
    sequence_golden("vlp16", 8, 0.4, "cauchy", False, False, "vlp16_vanilla.npz")
    sequence_golden("vlp16", 6, 0.4, "huber", True, True, "vlp16_deskew_huber.npz")
    sequence_golden("hdl64", 4, 0.4, "cauchy", False, False, "hdl64_vanilla.npz")
