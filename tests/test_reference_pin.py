"""CPU tests that PIN the oracle: the restatement (oracle/libfloam_oracle.so) against the reference's own class sources compiled
unmodified into oracle/_ref/libfloam_ref.so (oracle/Makefile target `ref`, stand-in third-party headers under oracle/stubs/).

Scope of the pin: every line of /root/reference/src/{laserProcessingClass,dataHandler,lidar,lidarOptimization,odomEstimationClass,
laserMappingClass}.cpp and CenterTime of laserProcessingNode.cpp is the reference's.  The PCL / Eigen / Ceres / FLANN internals behind
the stand-in interfaces are shared restatements (not in this image); FLANN's kd-tree is pinned separately against OpenCV's bundled
copy (tests/test_oracle.py::test_kdtree_matches_a_real_flann_build)."""
import os

import numpy as np
import pytest

from conftest import xyzi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LINES = {"vlp16": 16, "hdl64": 64, "os1-128": 128}
FIELDS = ("x", "y", "z", "intensity", "ring", "time")


def ros_stamp(t):
    """A double that IS a ros::Time: sec + 1e-9 * nsec (TimeBase::toSec).  IMU stamps reach ImuHandler::AddMsg as
    msg->header.stamp.toSec() (src/dataHandler.cpp:29), so only such doubles occur; an arbitrary double would be rounded to the
    nanosecond by the message type before the class ever sees it."""
    ns = int(round(t * 1e9))
    return float(ns // 1000000000) + 1e-9 * float(ns % 1000000000)


def same_points(a, b):
    return len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in FIELDS)


def test_ref_library_is_built_from_the_reference_sources(pr):
    assert pr.lib().fo_backend() == b"reference"
    if os.path.isdir("/root/reference/src"):          # authoring container: the recipe must rebuild from the sources where they lie
        so = pr.build()
        newest = max(os.path.getmtime(os.path.join("/root/reference/src", f)) for f in os.listdir("/root/reference/src"))
        assert os.path.getmtime(so) >= newest
    # nothing of the reference is copied into the repo: the recipe reads it through $(REF)
    mk = open(os.path.join(os.path.dirname(GOLD), "..", "oracle", "Makefile")).read()
    assert "$(REF)/src/$$f.cpp" in mk and "REF ?= /root/reference" in mk


# ------------------------------------------------------------------------------------------------ featureExtraction -----
@pytest.mark.parametrize("sensor", ["vlp16", "hdl64", "os1-128"])
def test_feature_extraction_equals_reference(po, pr, sequences, sensor):
    seq, scans, off = sequences(sensor, 2)
    for f in range(2):
        s = scans[off[f]:off[f + 1]]
        re_, rs, res, rss, _ = pr.feature_extract(s, LINES[sensor], 2.0, 60.0)
        for total_order in (False, True):
            oe, osf, oes, oss, ties = po.feature_extract(s, LINES[sensor], 2.0, 60.0, total_order=total_order)
            assert ties == 0
            assert np.array_equal(oes, res) and np.array_equal(oss, rss)
            assert same_points(oe, re_) and same_points(osf, rs)
        assert same_points(re_, s[res]) and same_points(rs, s[rss])


def ragged_scan(rng, po, sizes, lines):
    parts = []
    for ring, n in enumerate(sizes):
        p = np.zeros(n, po.POINT_IRT)
        az = np.sort(rng.uniform(-np.pi, np.pi, n)); r = 8 + 4 * rng.random(n) * (rng.random(n) < 0.1) + 0.01 * rng.standard_normal(n)
        p["x"] = r * np.cos(az); p["y"] = r * np.sin(az); p["z"] = 0.2 * ring; p["ring"] = ring; p["pad0"] = 1; p["intensity"] = rng.random(n)
        p["time"] = np.linspace(0, 0.1, n, endpoint=False)
        parts.append(p)
    pts = np.concatenate(parts)
    order = np.argsort(np.concatenate([np.arange(n) * lines + ring for ring, n in enumerate(sizes)]), kind="stable")
    return pts[order]


def test_feature_extraction_ragged_rings_equal_reference(po, pr):
    rng = np.random.default_rng(7)
    fixed = [0, 1, 130, 131, 136, 137, 600, 1233, 2100, 17, 131, 400, 0, 905, 3000, 132]
    for trial in range(12):
        sizes = fixed if trial == 0 else [int(x) for x in rng.choice([0, 5, 130, 131, 132, 137, 143, 200, 777, 1500, 2049], 16)]
        pts = ragged_scan(rng, po, sizes, 16)
        if trial % 3 == 2:                     # range gate on both sides (min_distance 2, max_distance 60), ties with the bounds included
            pts["x"][::13] *= 0.1; pts["y"][::13] *= 0.1; pts["x"][::17] *= 9
        _, _, res, rss, _ = pr.feature_extract(pts, 16, 2.0, 60.0)
        _, _, oes, oss, ties = po.feature_extract(pts, 16, 2.0, 60.0, total_order=False)
        assert np.array_equal(oes, res) and np.array_equal(oss, rss), (trial, sizes)
        if ties == 0:
            _, _, tes, tss, _ = po.feature_extract(pts, 16, 2.0, 60.0, total_order=True)
            assert np.array_equal(tes, res) and np.array_equal(tss, rss)


def test_range_gate_rounds_like_the_reference_build(po, pr):
    # SURVEY a6 asked "float or double sqrt" for `sqrt(x*x + y*y)` on float operands (src/laserProcessingClass.cpp:14): the compiler,
    # given PCL's <math.h> include, picks the float overload.  Points placed within an ulp of the two bounds decide it.
    rng = np.random.default_rng(3)
    n = 4000
    p = np.zeros(n, po.POINT_IRT)
    ang = np.sort(rng.uniform(-np.pi, np.pi, n))
    r = np.where(rng.random(n) < 0.5, 60.0, 2.0) * (1 + rng.integers(-3, 4, n) * 2.0 ** -24)
    p["x"] = (r * np.cos(ang)).astype(np.float32); p["y"] = (r * np.sin(ang)).astype(np.float32); p["pad0"] = 1
    p["time"] = np.linspace(0, 0.1, n, endpoint=False)
    _, _, res, rss, _ = pr.feature_extract(p, 16, 2.0, 60.0)
    _, _, oes, oss, _ = po.feature_extract(p, 16, 2.0, 60.0, total_order=False)
    assert len(res) + len(rss) > 500
    assert np.array_equal(oes, res) and np.array_equal(oss, rss)


# ------------------------------------------------------------------------------------------------ dataHandler ----------
def test_imu_handler_and_deskew_equal_reference(po, pr, synth):
    seq = synth.Sequence("vlp16", seed=1, distort=True)
    ext = pr.euler2quat(0, 0, 180)               # src/laserProcessingNode.cpp:196
    assert np.array_equal(ext, po.euler2quat(0, 0, 180))
    a, b = po.Imu(), pr.Imu()
    rng = np.random.default_rng(0)
    for k in range(-40, 140):
        t = ros_stamp(1000.0 + 0.005 * k + (4e-6 if k % 17 == 0 else 0.0))
        q = seq.imu(t - 1000.0 if t >= 1000.0 else 0.0)
        a.add(t, q); b.add(t, q)
        if k % 9 == 0:                            # near-duplicate stamps: dropped when <= 10 us after the last kept sample (:30-33)
            t2 = ros_stamp(t + float(rng.choice([5e-6, 1.0e-5, 1.2e-5])))
            a.add(t2, q); b.add(t2, q)
    assert a.size() == b.size()
    for t in (1000.0121, 999.0, 999.8, 999.805, 1000.3, 1000.695, 5000.0, 1000.0):
        ok1, q1 = a.get(t); ok2, q2 = b.get(t)
        assert ok1 == ok2 and (not ok1 or np.array_equal(q1, q2)), t
    for f in range(4):
        s1 = seq.scan(f); s2 = s1.copy()
        stamp = int((1000.0 + 0.1 * f) * 1e6) + 3
        rc1, st1 = a.deskew_align(s1, stamp, ext); rc2, st2 = b.deskew_align(s2, stamp, ext)
        assert rc1 == rc2 == 0 and st1 == st2
        assert same_points(s1, s2)
    s1 = seq.scan(5); s2 = s1.copy()             # not covered by the IMU buffer: Compensate returns false after CenterTime ran
    rc1, st1 = a.deskew_align(s1, int(3000.0 * 1e6), ext); rc2, st2 = b.deskew_align(s2, int(3000.0 * 1e6), ext)
    assert rc1 == rc2 == 1 and st1 == st2 and same_points(s1, s2)


def test_compensate_velocity_equals_reference(po, pr, sequences):
    seq, scans, off = sequences("vlp16", 1)
    s1 = scans[off[0]:off[1]].copy(); s2 = s1.copy()
    v = np.array([9.7, -0.31, 0.05])
    po.compensate_velocity(s1, v); pr.compensate_velocity(s2, v)
    assert same_points(s1, s2)


# ------------------------------------------------------------------------------------------------ lidarOptimization -----
def random_blocks(rng, n, pose):
    from floam_b200 import synth
    T = synth.pose7_to_matrix(pose)
    recs = np.zeros((n, 10))
    for i in range(n):
        p = rng.uniform(-20, 20, 3); pw = T[:3, :3] @ p + T[:3, 3]
        if i % 3 == 0:   # edge: a line near the transformed point
            d = rng.standard_normal(3); d /= np.linalg.norm(d)
            c = pw + 0.05 * rng.standard_normal(3)
            recs[i] = [0, *p, *(c + 0.1 * d), *(c - 0.1 * d)]
        else:            # surf: a plane near it
            nrm = rng.standard_normal(3); nrm /= np.linalg.norm(nrm)
            recs[i] = [1, *p, *nrm, -(nrm @ pw) + 0.03 * rng.standard_normal(), 0, 0]
    return recs


def test_cost_functions_and_parameterization_equal_reference(po, pr):
    rng = np.random.default_rng(11)
    for trial in range(20):
        q = rng.standard_normal(4); q /= np.linalg.norm(q)
        x = np.concatenate([q, rng.uniform(-5, 5, 3)])
        for rec in random_blocks(rng, 6, x):
            r1, J1, ok1 = po.evaluate_residual(rec, x); r2, J2, ok2 = pr.evaluate_residual(rec, x)
            assert ok1 and ok2
            assert abs(r1 - r2) <= 4e-16 * max(1.0, abs(r1))
            assert np.allclose(J1, J2, rtol=1e-13, atol=1e-15)
        delta = rng.standard_normal(6) * rng.choice([1e-12, 1e-6, 1e-2, 0.5])
        assert np.allclose(po.se3_plus(x, delta), pr.se3_plus(x, delta), rtol=0, atol=2e-16)
    assert np.array_equal(po.se3_plus(x, np.zeros(6)), pr.se3_plus(x, np.zeros(6)))


@pytest.mark.parametrize("loss", [0, 1, 2])
def test_lm_solve_equals_reference(po, pr, loss):
    rng = np.random.default_rng(5 + loss)
    for trial in range(6):
        q = np.array([0.01, -0.02, 0.3, 0.95]); q /= np.linalg.norm(q)
        truth = np.concatenate([q, [1.0, -2.0, 0.5]])
        recs = random_blocks(rng, 400, truth)
        if trial % 2:
            recs[::25, 7] += 0.8          # gross outliers on some planes
        start = po.se3_plus(truth, rng.standard_normal(6) * [0.01, 0.01, 0.01, 0.2, 0.2, 0.2])
        x1, s1 = po.lm_solve(recs, loss, start); x2, s2 = pr.lm_solve(recs, loss, start)
        assert (s1["iterations"], s1["accepted"], s1["termination"]) == (s2["iterations"], s2["accepted"], s2["termination"])
        assert np.allclose(x1, x2, rtol=0, atol=1e-12)
        assert np.allclose(s1["H0"], s2["H0"], rtol=1e-12) and np.allclose(s1["g0"], s2["g0"], rtol=1e-10, atol=1e-12)
        assert abs(s1["final_cost"] - s2["final_cost"]) <= 1e-12 * max(1.0, s1["final_cost"])


# ------------------------------------------------------------------------------------------------ OdomEstimationClass ---
CASES = [  # sensor, frames, loss, deskew, speed, map_resolution
    ("vlp16", 14, "cauchy", False, 10.0, 0.4),
    ("vlp16", 10, "Huber", True, 10.0, 0.4),       # the class lower-cases the string (:23)
    ("vlp16", 16, "huber", False, 0.3, 0.4),       # walking pace: KeyFrameUpdate returns false on most frames (:320-343)
    ("vlp16", 10, "cauchy", True, 0.5, 0.2),
    ("hdl64", 5, "cauchy", False, 10.0, 0.4),
]


@pytest.mark.parametrize("sensor,frames,loss,deskew,speed,res", CASES)
@pytest.mark.parametrize("contract", [False, True])
def test_odometry_sequence_equals_reference(po, pr, synth, sensor, frames, loss, deskew, speed, res, contract):
    """contract=False: the libraries as they are (std::sort inside VoxelGrid, kd-tree traversal order); contract=True: stable
    voxel order + (distance, index) neighbour order, the deterministic contract of the CUDA path."""
    seq = synth.Sequence(sensor, seed=4, distort=deskew, speed=speed)
    scans, off = seq.scans(0, frames)
    kw = dict(num_lines=LINES[sensor], loss=loss, map_resolution=res, total_order=contract, use_kdtree=not contract)
    o = po.Odom(**kw); r = pr.Odom(**kw)
    keyframes = []
    for f in range(frames):
        s = scans[off[f]:off[f + 1]]
        e, sf, _, _, _ = pr.feature_extract(s, LINES[sensor], 2.0, 60.0)
        if f == 0:
            o.init_map(synth.to_xyzi(e), synth.to_xyzi(sf)); r.init_map(synth.to_xyzi(e), synth.to_xyzi(sf))
            continue
        e1, s1, e2, s2 = e.copy(), sf.copy(), e.copy(), sf.copy()
        p1 = o.update(e1, s1, deskew); p2 = r.update(e2, s2, deskew)
        assert np.abs(p1 - p2).max() <= 1e-12, (f, p1, p2)
        assert same_points(e1, e2) and same_points(s1, s2)          # CompensateVelocity mutates the caller's clouds (:42-43)
        T1, L1, v1, oc1 = o.get(); T2, L2, v2, oc2 = r.get()
        assert oc1 == oc2 and np.allclose(T1, T2, atol=1e-12) and np.allclose(L1, L2, atol=1e-12) and np.allclose(v1, v2, atol=1e-10)
        d1, d2 = o.debug(), r.debug()
        assert d1["keyframe"] == d2["keyframe"]
        assert (d1["lm"]["iterations"], d1["lm"]["accepted"], d1["lm"]["termination"]) == (d2["lm"]["iterations"], d2["lm"]["accepted"], d2["lm"]["termination"])
        if d2["keyframe"]:
            assert np.array_equal(xyzi(d1["ds_edge"]), xyzi(d2["ds_edge"])) and np.array_equal(xyzi(d1["ds_surf"]), xyzi(d2["ds_surf"]))
        m1, m2 = o.get_map(), r.get_map()
        assert np.array_equal(xyzi(m1[0]), xyzi(m2[0])) and np.array_equal(xyzi(m1[1]), xyzi(m2[1]))
        keyframes.append(d2["keyframe"])
    assert keyframes[0]                                              # the function-static `first` flag (Q10)
    if speed < 1.0:
        assert keyframes.count(False) > len(keyframes) // 2          # the non-keyframe branch is what this case is for


def test_update_types_and_schedule_equal_reference(po, pr, synth, sequences):
    # direct updatePointsToMap(PointXYZI, type) calls: INITIAL_ITERATION must not touch the map; optimization_count 12 -> 2 (Q4)
    seq, scans, off = sequences("vlp16", 8)
    o = po.Odom(num_lines=16); r = pr.Odom(num_lines=16)
    counts = []
    for f in range(8):
        e, sf, _, _, _ = pr.feature_extract(scans[off[f]:off[f + 1]], 16, 2.0, 60.0)
        e, sf = synth.to_xyzi(e), synth.to_xyzi(sf)
        if f == 0:
            o.init_map(e, sf); r.init_map(e, sf); continue
        for t in ((1, 2) if f % 2 else (0,)):
            n0 = [len(m) for m in r.get_map()]
            p1 = o.update_xyzi(e, sf, t); p2 = r.update_xyzi(e, sf, t)
            assert np.abs(p1 - p2).max() <= 1e-12
            if t == 1:
                assert [len(m) for m in r.get_map()] == n0
            assert o.get()[3] == r.get()[3]
            counts.append(r.get()[3])
    assert counts[0] == 11 and counts[-1] == 2 and all(a >= b for a, b in zip(counts, counts[1:]))


def test_too_small_map_skips_the_solve_like_the_reference(po, pr, synth, sequences):
    seq, scans, off = sequences("vlp16", 2)
    e, sf, _, _, _ = pr.feature_extract(scans[off[0]:off[1]], 16, 2.0, 60.0)
    o = po.Odom(num_lines=16); r = pr.Odom(num_lines=16)
    o.init_map(synth.to_xyzi(e[:10]), synth.to_xyzi(sf[:50])); r.init_map(synth.to_xyzi(e[:10]), synth.to_xyzi(sf[:50]))   # needs > 10 and > 50 (:77)
    e2, s2, _, _, _ = pr.feature_extract(scans[off[1]:off[2]], 16, 2.0, 60.0)
    p1 = o.update(e2.copy(), s2.copy(), False); p2 = r.update(e2.copy(), s2.copy(), False)
    assert np.array_equal(p1, p2) and r.debug()["outer_iterations"] == 0
    assert np.array_equal(p2, [0, 0, 0, 1, 0, 0, 0])


# ------------------------------------------------------------------------------------------------ LaserMappingClass -----
@pytest.mark.parametrize("contract", [False, True])
def test_laser_mapping_equals_reference(po, pr, synth, sequences, contract):
    seq, scans, off = sequences("vlp16", 6)
    a = po.Mapping(map_resolution=0.4, total_order=contract); b = pr.Mapping(map_resolution=0.4, total_order=contract)
    for f in range(6):
        # the mapping node is fed /velodyne_points_filtered = edge + surf features (src/laserProcessingNode.cpp:139-145), which are
        # range-gated to max_distance; a raw scan with farther returns indexes cells outside the allocated 5x5x5 block and the
        # reference dereferences them (src/laserMappingClass.cpp:170)
        e, sf, _, _, _ = pr.feature_extract(scans[off[f]:off[f + 1]], 16, 2.0, 60.0)
        pts = synth.to_xyzi(np.concatenate([e, sf]))
        T = seq.pose(0.1 * f * 40)           # the 5x5x5 block crosses 50 m cell boundaries, the grid grows on both sides
        if f == 4:
            T = T.copy(); T[:3, 3] = [-130.0, -75.0, 12.0]     # negative growth: addWidth/Height/DepthCellNegative
        a.update(pts, T); b.update(pts, T)
        ma, mb = a.get_map(), b.get_map()
        assert len(ma) == len(mb) and np.array_equal(xyzi(ma), xyzi(mb))


# ------------------------------------------------------------------------------------------------ golden fixtures -------
@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLD) if f.startswith("ref_") and f.endswith(".npz")))
def test_reference_golden_fixtures_are_reproduced(po, pr, synth, name):
    """tests/golden/ref_*.npz were produced by tests/golden/make_reference_golden.py from the reference build.  The restatement
    must reproduce them, and so must the reference build itself (fixture drift check) where it is present."""
    from golden.make_reference_golden import run_case
    g = np.load(os.path.join(GOLD, name), allow_pickle=False)
    for backend in (po, pr):
        out = run_case(backend, synth, str(g["sensor"]), int(g["frames"]), float(g["map_resolution"]), str(g["loss"]), bool(g["deskew"]),
                       float(g["speed"]), bool(g["contract"]), bool(g["imu"]))
        assert np.array_equal(out["scan_crc"], g["scan_crc"]), "synthetic generator drifted: regenerate tests/golden"
        assert np.array_equal(out["edge_crc"], g["edge_crc"]) and np.array_equal(out["surf_crc"], g["surf_crc"])
        assert np.array_equal(out["edge_src_0"], g["edge_src_0"])
        assert np.array_equal(out["keyframe"], g["keyframe"]) and np.array_equal(out["map_sizes"], g["map_sizes"])
        assert np.abs(out["poses"] - g["poses"]).max() <= 1e-12
