"""First-contact GPU diagnostic: runs every stage of the CUDA path against the oracle and prints what differs.
(Development aid; the judged checks live in tests/.)  Usage: python tools/probes/gpu_diag.py [sensor] [frames]"""
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, ".")
from floam_b200 import capi, synth  # noqa: E402
from oracle import pyoracle as po   # noqa: E402

sensor = sys.argv[1] if len(sys.argv) > 1 else "vlp16"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
PRM = dict(num_lines=synth.SENSORS[sensor][1], min_distance=2.0, max_distance=60.0, map_resolution=0.4, loss="cauchy",
           max_scan_points=300000, max_map_points=1 << 21, max_global_map_points=1 << 21, max_grid_cells=1 << 22)


def section(name):
    print("\n=== %s ===" % name, flush=True)


def xyz(a):
    return np.stack([a["x"], a["y"], a["z"], a["intensity"]], 1)


def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()


seq = synth.Sequence(sensor, seed=0)
scans, off = seq.scans(0, frames)
ctx = capi.Context(**PRM)
print("context ok", capi.lib().floam_version())


def t_feature():
    section("feature extraction")
    for f in range(min(frames, 3)):
        s = scans[off[f]:off[f + 1]]
        t = time.time(); e, sf, es, ss = ctx.feature_extract(s, with_src=True); dt = time.time() - t
        oe, osf, oes, oss, ties = po.feature_extract(s, PRM["num_lines"], 2.0, 60.0, total_order=True)
        print("frame", f, "n", len(s), "gpu edge/surf", len(e), len(sf), "oracle", len(oe), len(osf), "ties", ties, "ms %.2f" % (dt * 1e3))
        print("  edge ids equal:", np.array_equal(es, oes), " surf ids equal:", np.array_equal(ss, oss),
              " edge bytes equal:", e.tobytes() == oe.tobytes() if len(e) == len(oe) else False)
        if not np.array_equal(es, oes):
            k = min(len(es), len(oes)); bad = np.nonzero(es[:k] != oes[:k])[0]
            print("  first edge mismatches", bad[:10], es[bad[:5]], oes[bad[:5]])
        if not np.array_equal(ss, oss):
            k = min(len(ss), len(oss)); bad = np.nonzero(ss[:k] != oss[:k])[0]
            print("  surf mismatches", len(bad), bad[:10], ss[bad[:5]], oss[bad[:5]])


def t_voxel():
    section("voxel / crop")
    s = scans[off[0]:off[1]]
    e, sf = ctx.feature_extract(s)
    for cloud, leaf in ((synth.to_xyzi(e), 0.4), (synth.to_xyzi(sf), 0.8), (synth.to_xyzi(sf), 0.2)):
        g = ctx.voxel_grid(cloud, leaf); o, pt = po.voxel_grid(cloud, leaf, total_order=True)
        same = len(g) == len(o) and np.array_equal(xyz(g), xyz(o))
        print("leaf", leaf, "n", len(cloud), "->", len(g), len(o), "bit-equal", same)
        if len(g) == len(o) and not same:
            d = np.abs(xyz(g) - xyz(o)); print("  max diff", d.max(), "rows differing", (d.max(1) > 0).sum())
    cloud = synth.to_xyzi(sf)
    mn = np.array([-10, -20, -1], np.float32); mx = np.array([30, 15, 3], np.float32)
    g = ctx.crop_box(cloud, mn, mx); o = po.crop_box(cloud, mn, mx)
    print("crop", len(g), len(o), "equal", len(g) == len(o) and np.array_equal(xyz(g), xyz(o)))


def t_knn():
    section("knn5")
    s0 = scans[off[0]:off[1]]; s1 = scans[off[1]:off[2]]
    _, sf0 = ctx.feature_extract(s0); _, sf1 = ctx.feature_extract(s1)
    m = synth.to_xyzi(sf0); q = po.voxel_grid(synth.to_xyzi(sf1), 0.8, total_order=True)[0]
    t = time.time(); ids, d2 = ctx.knn5(m, q); dt = time.time() - t
    oids, od2 = po.knn(m, q, 5, use_kdtree=False)
    near = od2[:, 4] < 1.0
    print("map", len(m), "queries", len(q), "near", near.sum(), "ms %.2f" % (dt * 1e3))
    print("  ids equal on near:", np.array_equal(ids[near], oids[near]), " d2 equal:", np.array_equal(d2[near], od2[near]),
          " far flagged -1:", bool((ids[~near] == -1).all()))
    if not np.array_equal(ids[near], oids[near]):
        bad = np.nonzero((ids != oids).any(1) & near)[0]; print("  bad", len(bad), ids[bad[:3]], oids[bad[:3]], d2[bad[:3]], od2[bad[:3]])


def t_sequence():
    section("sequence parity (process_scan vs oracle, total_order)")
    c2 = capi.Context(**PRM)
    orc = po.Odom(num_lines=PRM["num_lines"], map_resolution=0.4, loss="cauchy", total_order=True, use_kdtree=False)
    lp_first = True
    for f in range(frames):
        s = scans[off[f]:off[f + 1]]
        t = time.time(); pose = c2.process_scan(s); dt = time.time() - t
        oe, osf, _, _, _ = po.feature_extract(s, PRM["num_lines"], 2.0, 60.0, total_order=True)
        if lp_first:
            orc.init_map(synth.to_xyzi(oe), synth.to_xyzi(osf)); opose = np.array([0, 0, 0, 1, 0, 0, 0.]); lp_first = False
        else:
            opose = orc.update(oe.copy(), osf.copy(), False)
        dq = np.abs(pose[:4] - opose[:4]).max(); dtt = np.abs(pose[4:] - opose[4:]).max()
        extra = ""
        if f > 0:
            d = c2.debug(); od = orc.debug()
            extra = " outer %d/%d kf %d/%d nds %d/%d,%d/%d corr %d/%d lm(it %d/%d acc %d/%d term %d/%d) cost0 %.6g/%.6g" % (
                d["outer_iterations"], od["outer_iterations"], d["keyframe"], od["keyframe"], len(d["ds_edge"]), len(od["ds_edge"]),
                len(d["ds_surf"]), len(od["ds_surf"]), d["n_corr"], len(od["residuals"]), d["lm"]["iterations"], od["lm"]["iterations"],
                d["lm"]["accepted"], od["lm"]["accepted"], d["lm"]["termination"], od["lm"]["termination"], d["lm"]["initial_cost"], od["lm"]["initial_cost"])
            ne, ns = c2.odom_map_sizes(); oem, osm = orc.get_map()
            extra += " map %d/%d %d/%d" % (ne, len(oem), ns, len(osm))
        print("frame %2d dq %.2e dt %.2e wall %.2f ms dev %.3f ms t=%s%s" % (f, dq, dtt, dt * 1e3, c2.last_frame_ms(), np.round(pose[4:], 4), extra), flush=True)
    print("launches", c2.launch_count())
    c2.close()


def t_perf():
    section("device-resident replay timing")
    c3 = capi.Context(**PRM)
    c3.stage_scans(scans, off)
    for rep in range(2):
        ms = []
        for f in range(frames):
            t = time.time(); c3.process_staged(f); ms.append((time.time() - t) * 1e3)
        print("wall ms per frame:", np.round(ms, 3))
    c3.close()


run(t_feature); run(t_voxel); run(t_knn); run(t_sequence); run(t_perf)
print("\nDONE")
