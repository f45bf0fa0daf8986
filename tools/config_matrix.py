"""BASELINE.json configs[0..3] on one B200: frames/s of the CUDA path, parity against the oracle where the oracle finishes in
seconds, and the oracle's own single-thread frames/s on the same frames. One JSON line per config (-> profiles/).

  0  VLP-16, 100 frames, res 0.4, deskew off          (the reference's CPU-runnable case)
  1  HDL-64, res 0.4, deskew off                      (bench.py's workload; shorter here)
  2  HDL-64 + 200 Hz IMU: deskew + alignment + two-pass velocity deskew, Huber loss
  3  OS1-128 (128 x 2048), map_resolution 0.08 (surf leaf 0.16), max_dis 90 / min_dis 0.5 (launch file): the local map passes 1M points
     around frame 110; frames/s are taken over frames 140..180 (parity of this configuration: tests/test_gpu_reference.py)
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from floam_b200 import capi, synth          # noqa: E402
from oracle import pyoracle as po            # noqa: E402


def run(cfg):
    name, sensor, frames, res, loss, deskew, imu, oracle_frames = cfg[:8]
    extra = cfg[8] if len(cfg) > 8 else {}
    timed_from = extra.get("timed_from", 12)
    seq = synth.Sequence(sensor, seed=0, distort=deskew, speed=extra.get("speed", 10.0))
    scans, off = seq.scans(0, frames)
    nl = seq.num_lines
    prm = dict(num_lines=nl, map_resolution=res, loss=loss, max_scan_points=seq.max_points + 1024, max_map_points=1 << 22,
               max_global_map_points=0, max_grid_cells=1 << 24)
    prm.update(extra.get("params", {}))
    ctx = capi.Context(**prm)
    ext = po.euler2quat(0, 0, 180)
    out = {"config": name, "sensor": sensor, "frames": frames, "map_resolution": res, "loss": loss, "deskew": deskew, "imu": imu,
           "points_per_scan": float(np.mean(np.diff(off)))}
    if imu:
        for k in range(-40, 20 * frames + 40):
            t = 100.0 + 0.005 * k
            ctx.imu_push(t, seq.imu(max(t - 100.0, 0.0)))
        poses = []; ms = []
        for f in range(frames):
            s = scans[off[f]:off[f + 1]].copy()
            t0 = time.perf_counter()
            rc, pose, _ = ctx.process_scan_imu(s, int((100.0 + 0.1 * f) * 1e6), ext, deskew)
            ms.append((time.perf_counter() - t0) * 1e3); poses.append(pose)
            assert rc == capi.OK
        poses = np.array(poses)
        steady = ms[12:]
        out["frames_per_s_e2e_sync"] = 1e3 / float(np.mean(steady)); out["p50_ms"] = float(np.percentile(steady, 50)); out["p99_ms"] = float(np.percentile(steady, 99))
        # the same frames again through the pipelined calls (three in flight) on a fresh context: throughput figure, poses must be identical
        ctx.close()
        ctx = capi.Context(**prm)
        for k in range(-40, 20 * frames + 40):
            t = 100.0 + 0.005 * k
            ctx.imu_push(t, seq.imu(max(t - 100.0, 0.0)))
        keep = [scans[off[f]:off[f + 1]].copy() for f in range(frames)]
        Q = []; pending = 0; t12 = None
        for f in range(frames):
            if f == 12:
                while pending:
                    Q.append(ctx.process_wait()); pending -= 1
                t12 = time.perf_counter()
            if pending == 3:
                Q.append(ctx.process_wait()); pending -= 1
            rc, _ = ctx.process_submit_imu(keep[f], int((100.0 + 0.1 * f) * 1e6), ext, deskew)
            assert rc == capi.OK
            pending += 1
        while pending:
            Q.append(ctx.process_wait()); pending -= 1
        out["frames_per_s_e2e_pipelined"] = (frames - 12) / (time.perf_counter() - t12)
        out["pipelined_poses_identical"] = bool(np.array_equal(np.array(Q), poses))
    else:
        ctx.stage_scans(scans, off)
        p0, _ = ctx.replay_staged(0, timed_from)
        out["map_points_at_start_of_timed_frames"] = list(ctx.odom_map_sizes())
        p1, ms = ctx.replay_staged(timed_from, frames - timed_from)
        poses = np.concatenate([p0, p1])
        out["timed_frames"] = [timed_from, frames]
        out["frames_per_s_device"] = (frames - timed_from) / (ms * 1e-3); out["ms_per_frame"] = ms / (frames - timed_from)
    d = ctx.debug()
    ne, ns = ctx.odom_map_sizes()
    out.update(map_points=[ne, ns], queries_last_frame=len(d["ds_edge"]) + len(d["ds_surf"]), correspondences_last_frame=d["n_corr"])
    out["knn_queries_per_s"] = out["queries_last_frame"] * 2 * out.get("frames_per_s_device", out.get("frames_per_s_e2e_sync", 0.0))
    ctx.close()
    # oracle on a prefix of the same frames: parity + CPU frames/s
    m = min(frames, oracle_frames)
    if m > 1:
        if imu:
            h = po.Imu()
            for k in range(-40, 20 * frames + 40):
                t = 100.0 + 0.005 * k
                h.add(t, seq.imu(max(t - 100.0, 0.0)))
            orc = po.Odom(num_lines=nl, map_resolution=res, loss=loss, total_order=True, use_kdtree=True)
            O = []; t0 = time.perf_counter()
            for f in range(m):
                s = scans[off[f]:off[f + 1]].copy()
                h.deskew_align(s, int((100.0 + 0.1 * f) * 1e6), ext)
                e, sf = po.feature_extract(s, nl, 2.0, 60.0, total_order=True)[:2]
                if f == 0:
                    orc.init_map(synth.to_xyzi(e), synth.to_xyzi(sf)); O.append(np.array([0, 0, 0, 1, 0, 0, 0.0]))
                else:
                    O.append(orc.update(e, sf, deskew))
            cpu_s = time.perf_counter() - t0
            O = np.array(O)
        else:
            orc = po.Odom(num_lines=nl, map_resolution=res, loss=loss, total_order=True, use_kdtree=True)
            O = []; t0 = time.perf_counter()
            for f in range(m):
                e, sf = po.feature_extract(scans[off[f]:off[f + 1]], nl, 2.0, 60.0, total_order=True)[:2]
                if f == 0:
                    orc.init_map(synth.to_xyzi(e), synth.to_xyzi(sf)); O.append(np.array([0, 0, 0, 1, 0, 0, 0.0]))
                else:
                    O.append(orc.update(e, sf, deskew))
            cpu_s = time.perf_counter() - t0
            O = np.array(O)
        out["oracle_frames"] = m
        out["max_pose_diff_vs_oracle"] = float(np.abs(poses[:m] - O).max())
        out["oracle_cpu_frames_per_s_1thread"] = m / cpu_s
    gt = [seq.pose(0.1 * f) for f in range(frames)]
    out["ate_vs_ground_truth_m"] = synth.ate(poses, gt)[0]
    print(json.dumps(out), flush=True)


CONFIGS = [
    ("0: VLP-16 100 frames", "vlp16", 100, 0.4, "cauchy", False, False, 100),
    ("1: HDL-64 300 frames", "hdl64", 300, 0.4, "cauchy", False, False, 40),
    ("2: HDL-64 + IMU deskew, Huber, 60 frames", "hdl64", 60, 0.4, "huber", True, True, 20),
    ("3: OS1-128, >= 1M-point local map (res 0.08, max_dis 90)", "os1-128", 180, 0.08, "cauchy", False, False, 0,
     {"timed_from": 140, "params": {"max_distance": 90.0, "min_distance": 0.5}}),
    # the reference's own regime (README: walking pace): most frames are not keyframes, and the Q2 double prediction of the deskew mode costs
    # centimetres, not metres, of trajectory error
    ("2b: HDL-64 + IMU deskew, Huber, 60 frames at 1 m/s", "hdl64", 60, 0.4, "huber", True, True, 20, {"speed": 1.0}),
    ("1b: HDL-64 120 frames at 1 m/s", "hdl64", 120, 0.4, "cauchy", False, False, 40, {"speed": 1.0}),
]
if __name__ == "__main__":
    which = [int(a) for a in sys.argv[1:]] or list(range(len(CONFIGS)))
    for i in which:
        run(CONFIGS[i])
