"""ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own class sources (oracle/_ref, tests/test_reference_pin.py); the
third-party internals it restates (PCL / FLANN / Eigen / Ceres) are unpinned.

ctypes driver for oracle/libfloam_oracle.so (the CPU restatement of the reference hot path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

POINT_IRT = np.dtype({"names": ["x", "y", "z", "pad0", "intensity", "ring", "pad1", "time", "pad2"],
                      "formats": ["<f4", "<f4", "<f4", "<f4", "<f4", "<u2", "<u2", "<f4", "<f4"], "itemsize": 32})
POINT_I = np.dtype({"names": ["x", "y", "z", "pad0", "intensity", "p1", "p2", "p3"],
                    "formats": ["<f4"] * 8, "itemsize": 32})


def build(force=False):
    so = os.path.join(_HERE, "libfloam_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.fo_replay_sequence.restype = C.c_double
        L.fo_odom_knn_queries.restype = C.c_long
        for f in ("fo_imu_create", "fo_odom_create", "fo_mapping_create"):
            getattr(L, f).restype = C.c_void_p
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def feature_extract(pts, num_lines, min_dis, max_dis, total_order=False):
    """src/laserProcessingClass.cpp:72. Returns (edge, surf, edge_src, surf_src, ties)."""
    pts = np.ascontiguousarray(pts, dtype=POINT_IRT)
    n = len(pts)
    edge = np.zeros(n, POINT_IRT); surf = np.zeros(n, POINT_IRT)
    es = np.zeros(n, np.int32); ss = np.zeros(n, np.int32)
    ne = C.c_int(); ns = C.c_int(); ties = C.c_long()
    lib().fo_feature_extract(_p(pts), n, int(num_lines), C.c_double(min_dis), C.c_double(max_dis), int(total_order),
                             _p(edge), _p(es), n, C.byref(ne), _p(surf), _p(ss), n, C.byref(ns), C.byref(ties))
    return edge[:ne.value], surf[:ns.value], es[:ne.value], ss[:ns.value], ties.value


def voxel_grid(pts, leaf, total_order=False):
    pts = np.ascontiguousarray(pts, dtype=POINT_I)
    out = np.zeros(max(len(pts), 1), POINT_I)
    pt = C.c_int()
    n = lib().fo_voxel_grid(_p(pts), len(pts), C.c_float(leaf), int(total_order), _p(out), len(out), C.byref(pt))
    return out[:n], bool(pt.value)


def crop_box(pts, mn, mx):
    pts = np.ascontiguousarray(pts, dtype=POINT_I)
    out = np.zeros(max(len(pts), 1), POINT_I)
    mn = np.asarray(mn, np.float32); mx = np.asarray(mx, np.float32)
    n = lib().fo_crop_box(_p(pts), len(pts), _p(mn), _p(mx), _p(out), len(out))
    return out[:n]


def knn(map_pts, queries, k=5, use_kdtree=True):
    map_pts = np.ascontiguousarray(map_pts, dtype=POINT_I); queries = np.ascontiguousarray(queries, dtype=POINT_I)
    ids = np.full((len(queries), k), -1, np.int32); d2 = np.zeros((len(queries), k), np.float32)
    lib().fo_knn(_p(map_pts), len(map_pts), _p(queries), len(queries), k, int(use_kdtree), _p(ids), _p(d2))
    return ids, d2


def eigen3(A):
    A = np.ascontiguousarray(A, np.float64); vals = np.zeros(3); vecs = np.zeros((3, 3))
    lib().fo_eigen3(_p(A), _p(vals), _p(vecs))
    return vals, vecs


def colpiv_qr_solve(A, b):
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64); x = np.zeros(3)
    lib().fo_colpiv_qr_solve(_p(A), _p(b), A.shape[0], _p(x))
    return x


def se3_plus(x, delta):
    x = np.ascontiguousarray(x, np.float64); d = np.ascontiguousarray(delta, np.float64); o = np.zeros(7)
    lib().fo_se3_plus(_p(x), _p(d), _p(o))
    return o


def quat_from_matrix(R):
    R = np.ascontiguousarray(R, np.float64); q = np.zeros(4)
    lib().fo_quat_from_matrix(_p(R), _p(q))
    return q


def euler2quat(roll, pitch, yaw):
    q = np.zeros(4)
    lib().fo_euler2quat(C.c_double(roll), C.c_double(pitch), C.c_double(yaw), _p(q))
    return q


def evaluate_residual(rec, x):
    rec = np.ascontiguousarray(rec, np.float64); x = np.ascontiguousarray(x, np.float64)
    r = C.c_double(); J = np.zeros(7)
    rc = lib().fo_evaluate_residual(_p(rec), _p(x), C.byref(r), _p(J))
    return r.value, J, rc == 0


def lm_solve(recs, loss, x, max_iter=4):
    """ceres::Solve restatement. recs: (n,10) doubles. Returns (x_new, summary dict)."""
    recs = np.ascontiguousarray(recs, np.float64).reshape(-1, 10); x = np.array(x, np.float64)
    s = np.zeros(47)
    lib().fo_lm_solve(_p(recs), len(recs), int(loss), _p(x), int(max_iter), _p(s))
    return x, {"iterations": int(s[0]), "accepted": int(s[1]), "initial_cost": s[2], "final_cost": s[3],
               "termination": int(s[4]), "H0": s[5:41].reshape(6, 6).copy(), "g0": s[41:47].copy()}


class Imu:
    def __init__(self):
        self.h = C.c_void_p(lib().fo_imu_create())

    def __del__(self):
        if self.h:
            lib().fo_imu_destroy(self.h); self.h = None

    def add(self, stamp, q_xyzw):
        q = np.ascontiguousarray(q_xyzw, np.float64)
        lib().fo_imu_add(self.h, C.c_double(stamp), _p(q))

    def set_slerp(self, on):
        """Opt-in fix FLOAM_FIX_IMU_SLERP (restatement only)."""
        lib().fo_imu_set_slerp(self.h, int(on))

    def size(self):
        return lib().fo_imu_size(self.h)

    def get(self, t):
        q = np.zeros(4)
        ok = lib().fo_imu_get(self.h, C.c_double(t), _p(q))
        return bool(ok), q

    def deskew_align(self, pts, stamp_us, extr_xyzw):
        """CenterTime + Compensate + IMU alignment, in place. Returns (status, new_stamp_us)."""
        assert pts.dtype == POINT_IRT and pts.flags.c_contiguous
        st = C.c_ulonglong(int(stamp_us)); ex = np.ascontiguousarray(extr_xyzw, np.float64)
        rc = lib().fo_deskew_align(self.h, _p(pts), len(pts), C.byref(st), _p(ex))
        return rc, st.value


def compensate_velocity(pts, v):
    assert pts.dtype == POINT_IRT and pts.flags.c_contiguous
    v = np.ascontiguousarray(v, np.float64)
    lib().fo_compensate_velocity(_p(pts), len(pts), _p(v))


class Odom:
    """OdomEstimationClass restatement (src/odomEstimationClass.cpp)."""

    def __init__(self, num_lines=64, scan_period=0.1, min_dis=2.0, max_dis=60.0, map_resolution=0.4, loss="cauchy",
                 total_order=False, use_kdtree=True, _lib=None):
        self._L = _lib if _lib is not None else lib()
        self.h = C.c_void_p(self._L.fo_odom_create(int(num_lines), C.c_double(scan_period), C.c_double(min_dis), C.c_double(max_dis),
                                                 C.c_double(map_resolution), loss.encode(), int(total_order), int(use_kdtree)))

    def __del__(self):
        if self.h:
            self._L.fo_odom_destroy(self.h); self.h = None

    def set_fixes(self, fixes):
        """Opt-in algorithmic fixes (floam_fix bits 1 = single prediction, 2 = rotated velocity); the restatement only."""
        self._L.fo_odom_set_fixes(self.h, int(fixes))

    def init_map(self, edge, surf):
        edge = np.ascontiguousarray(edge, POINT_I); surf = np.ascontiguousarray(surf, POINT_I)
        self._L.fo_odom_init_map(self.h, _p(edge), len(edge), _p(surf), len(surf))

    def update(self, edge, surf, deskew=False):
        assert edge.dtype == POINT_IRT and surf.dtype == POINT_IRT
        pose = np.zeros(7)
        self._L.fo_odom_update(self.h, _p(edge), len(edge), _p(surf), len(surf), int(deskew), _p(pose))
        return pose

    def update_xyzi(self, edge, surf, update_type=0):
        edge = np.ascontiguousarray(edge, POINT_I); surf = np.ascontiguousarray(surf, POINT_I)
        pose = np.zeros(7)
        self._L.fo_odom_update_xyzi(self.h, _p(edge), len(edge), _p(surf), len(surf), int(update_type), _p(pose))
        return pose

    def get(self):
        odom = np.zeros(16); last = np.zeros(16); v = np.zeros(3); oc = C.c_int()
        self._L.fo_odom_get(self.h, _p(odom), _p(last), _p(v), C.byref(oc))
        return odom.reshape(4, 4), last.reshape(4, 4), v, oc.value

    def set_state(self, odom, last_odom, optimization_count):
        o = np.ascontiguousarray(odom, np.float64).reshape(16); l = np.ascontiguousarray(last_odom, np.float64).reshape(16)
        self._L.fo_odom_set_state(self.h, _p(o), _p(l), int(optimization_count))

    def set_map(self, edge, surf):
        edge = np.ascontiguousarray(edge, POINT_I); surf = np.ascontiguousarray(surf, POINT_I)
        self._L.fo_odom_set_map(self.h, _p(edge), len(edge), _p(surf), len(surf))

    def get_map(self):
        ne = C.c_int(); ns = C.c_int()
        self._L.fo_odom_map_sizes(self.h, C.byref(ne), C.byref(ns))
        e = np.zeros(max(ne.value, 1), POINT_I); s = np.zeros(max(ns.value, 1), POINT_I)
        self._L.fo_odom_get_map(self.h, _p(e), len(e), _p(s), len(s))
        return e[:ne.value], s[:ns.value]

    def knn_queries(self):
        return self._L.fo_odom_knn_queries(self.h)

    def debug(self):
        L = self._L

        def fetch(what, dtype, per=1):
            n = L.fo_odom_debug(self.h, what, None, 0)
            a = np.zeros(max(n * (per if what == 8 else 1), 1), dtype)
            L.fo_odom_debug(self.h, what, _p(a), n)
            return a[:n * (per if what == 8 else 1)]
        sc = np.zeros(2, np.int32); L.fo_odom_debug(self.h, 10, _p(sc), 2)
        lm = np.zeros(47); L.fo_odom_debug(self.h, 9, _p(lm), 47)
        return {
            "ds_edge": fetch(0, POINT_I), "ds_surf": fetch(1, POINT_I),
            "edge_knn": fetch(2, np.int32).reshape(-1, 5), "surf_knn": fetch(3, np.int32).reshape(-1, 5),
            "edge_d2": fetch(4, np.float32).reshape(-1, 5), "surf_d2": fetch(5, np.float32).reshape(-1, 5),
            "edge_ok": fetch(6, np.uint8), "surf_ok": fetch(7, np.uint8),
            "residuals": fetch(8, np.float64, 10).reshape(-1, 10),
            "lm": {"iterations": int(lm[0]), "accepted": int(lm[1]), "initial_cost": lm[2], "final_cost": lm[3],
                   "termination": int(lm[4]), "H0": lm[5:41].reshape(6, 6).copy(), "g0": lm[41:47].copy()},
            "outer_iterations": int(sc[0]), "keyframe": bool(sc[1]),
        }


class Mapping:
    def __init__(self, map_resolution=0.4, total_order=False):
        self.h = C.c_void_p(lib().fo_mapping_create(C.c_double(map_resolution), int(total_order)))

    def __del__(self):
        if self.h:
            lib().fo_mapping_destroy(self.h); self.h = None

    def update(self, pts, pose):
        pts = np.ascontiguousarray(pts, POINT_I); T = np.ascontiguousarray(pose, np.float64).reshape(16)
        lib().fo_mapping_update(self.h, _p(pts), len(pts), _p(T))

    def get_map(self, cap=1 << 22):
        out = np.zeros(cap, POINT_I)
        n = lib().fo_mapping_get_map(self.h, _p(out), cap)
        assert n <= cap
        return out[:n].copy()


def replay_sequence(scans, offsets, num_lines, scan_period=0.1, min_dis=2.0, max_dis=60.0, map_resolution=0.4, loss="cauchy", deskew=False):
    """featureExtraction + odometry over a whole sequence on one host thread; returns (seconds, poses[n,7], per_frame_ms, knn_queries)."""
    scans = np.ascontiguousarray(scans, POINT_IRT); offsets = np.ascontiguousarray(offsets, np.int64)
    nf = len(offsets) - 1
    poses = np.zeros((nf, 7)); ms = np.zeros(nf); q = C.c_long()
    sec = lib().fo_replay_sequence(_p(scans), _p(offsets), nf, int(num_lines), C.c_double(scan_period), C.c_double(min_dis), C.c_double(max_dis),
                                   C.c_double(map_resolution), loss.encode(), int(deskew), _p(poses), _p(ms), C.byref(q))
    return sec, poses, ms, q.value


STAGES = ("featureExtraction", "downSamplingToMap", "kdtree_build", "association_knn_fit", "ceres_solve", "addPointsToMap")


def replay_sequence_stages(scans, offsets, num_lines, skip, scan_period=0.1, min_dis=2.0, max_dis=60.0, map_resolution=0.4, loss="cauchy", deskew=False):
    """replay_sequence plus wall-clock milliseconds per frame and stage over frames [skip, n): dict stage -> ms."""
    scans = np.ascontiguousarray(scans, POINT_IRT); offsets = np.ascontiguousarray(offsets, np.int64)
    nf = len(offsets) - 1
    poses = np.zeros((nf, 7)); ms = np.zeros(nf); q = C.c_long(); st = np.zeros(6)
    L = lib()
    L.fo_replay_sequence_stages.restype = C.c_double
    sec = L.fo_replay_sequence_stages(_p(scans), _p(offsets), nf, int(num_lines), C.c_double(scan_period), C.c_double(min_dis), C.c_double(max_dis),
                                      C.c_double(map_resolution), loss.encode(), int(deskew), _p(poses), _p(ms), C.byref(q), int(skip), _p(st))
    return sec, poses, ms, q.value, dict(zip(STAGES, [float(x) for x in st]))
