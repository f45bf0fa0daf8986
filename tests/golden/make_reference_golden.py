"""Regenerates tests/golden/ref_*.npz from the REFERENCE BUILD (oracle/_ref/libfloam_ref.so: the reference's own class sources
compiled unmodified, oracle/Makefile target `ref`) on the deterministic synthetic generator.  Only runs where /root/reference is
present (the authoring container); the fixtures travel to the GPU box, where both the restatement (CPU test) and the CUDA path
(GPU test) must reproduce them.  Run from the repo root:  python tests/golden/make_reference_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LINES = {"vlp16": 16, "hdl64": 64, "os1-128": 128}
T0 = 300.0   # epoch of frame 0 in the IMU cases


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def ros_stamp(t):
    """sec + 1e-9 * nsec: the doubles ros::Time::toSec() produces (what ImuHandler::AddMsg is fed, src/dataHandler.cpp:29)."""
    ns = int(round(t * 1e9))
    return float(ns // 1000000000) + 1e-9 * float(ns % 1000000000)


def imu_times(frames):
    return [ros_stamp(T0 + 0.005 * k) for k in range(-40, 40 + 20 * frames)]


def run_case(backend, synth, sensor, frames, map_resolution, loss, deskew, speed, contract, imu):
    """One sequence through `backend` (oracle.pyref = the reference's code, oracle.pyoracle = the restatement).
    imu=True: scans are motion-distorted and pass CenterTime + Compensate + IMU alignment first (src/laserProcessingNode.cpp:99-116)."""
    seq = synth.Sequence(sensor, seed=0, distort=deskew or imu, speed=speed)
    nl = LINES[sensor]
    est = backend.Odom(num_lines=nl, map_resolution=map_resolution, loss=loss, total_order=contract, use_kdtree=not contract)
    handler = None
    if imu:
        handler = backend.Imu(); ext = backend.euler2quat(0, 0, 180)
        for t in imu_times(frames):
            handler.add(t, seq.imu(max(t - T0, 0.0)))
    out = {k: [] for k in ("scan_crc", "edge_crc", "surf_crc", "n_edge", "n_surf", "poses", "map_sizes", "keyframe", "stamps")}
    for f in range(frames):
        s = seq.scan(f)
        out["scan_crc"].append(crc(s))
        if imu:
            rc, st = handler.deskew_align(s, int((T0 + 0.1 * f) * 1e6), ext)
            assert rc == 0
            out["stamps"].append(st)
        e, sf, es, ss, _ = backend.feature_extract(s, nl, 2.0, 60.0)
        if f == 0:
            out["edge_src_0"] = es.copy()
            est.init_map(synth.to_xyzi(e), synth.to_xyzi(sf))
            pose = np.array([0, 0, 0, 1, 0, 0, 0.0]); kf = False
        else:
            pose = est.update(e.copy(), sf.copy(), deskew)
            kf = est.debug()["keyframe"]
        em, sm = est.get_map()
        out["edge_crc"].append(crc(es)); out["surf_crc"].append(crc(ss)); out["n_edge"].append(len(es)); out["n_surf"].append(len(ss))
        out["poses"].append(pose); out["map_sizes"].append((len(em), len(sm))); out["keyframe"].append(kf)
    return {"scan_crc": np.array(out["scan_crc"], np.uint32), "edge_crc": np.array(out["edge_crc"], np.uint32),
            "surf_crc": np.array(out["surf_crc"], np.uint32), "edge_src_0": out["edge_src_0"], "n_edge": np.array(out["n_edge"]),
            "n_surf": np.array(out["n_surf"]), "poses": np.array(out["poses"]), "map_sizes": np.array(out["map_sizes"]),
            "keyframe": np.array(out["keyframe"]), "stamps": np.array(out["stamps"], np.uint64)}


CASES = {   # name: sensor, frames, map_resolution, loss, deskew, speed, contract, imu
    "ref_vlp16_vanilla": ("vlp16", 10, 0.4, "cauchy", False, 10.0, True, False),
    "ref_vlp16_vanilla_faithful": ("vlp16", 10, 0.4, "cauchy", False, 10.0, False, False),
    "ref_vlp16_deskew_huber": ("vlp16", 8, 0.4, "huber", True, 10.0, True, False),
    "ref_vlp16_walking_pace": ("vlp16", 16, 0.4, "cauchy", False, 0.3, True, False),
    "ref_vlp16_imu_deskew_huber": ("vlp16", 8, 0.4, "huber", True, 10.0, True, True),
    "ref_hdl64_vanilla": ("hdl64", 5, 0.4, "cauchy", False, 10.0, True, False),
    "ref_hdl64_imu_deskew_cauchy": ("hdl64", 5, 0.4, "Cauchy", True, 10.0, True, True),
}


if __name__ == "__main__":
    from floam_b200 import synth
    from oracle import pyref
    assert pyref.lib().fo_backend() == b"reference"
    for name, (sensor, frames, res, loss, deskew, speed, contract, imu) in CASES.items():
        out = run_case(pyref, synth, sensor, frames, res, loss, deskew, speed, contract, imu)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), sensor=sensor, frames=frames, map_resolution=res, loss=loss, deskew=deskew,
                            speed=speed, contract=contract, imu=imu, **out)
        print(name, "keyframes", out["keyframe"].astype(int).tolist(), "pose[-1]", out["poses"][-1])
