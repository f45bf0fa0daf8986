"""ctypes driver for the synthetic LiDAR workload generator (floam_b200/synth/synth.cpp)."""
import ctypes as C
import os

import numpy as np

from ..build import build_synth

POINT_IRT = np.dtype({"names": ["x", "y", "z", "pad0", "intensity", "ring", "pad1", "time", "pad2"],
                      "formats": ["<f4", "<f4", "<f4", "<f4", "<f4", "<u2", "<u2", "<f4", "<f4"], "itemsize": 32})
POINT_I = np.dtype({"names": ["x", "y", "z", "pad0", "intensity", "p1", "p2", "p3"], "formats": ["<f4"] * 8, "itemsize": 32})

SENSORS = {"vlp16": (0, 16, 1800), "hdl64": (1, 64, 1875), "os1-128": (2, 128, 2048)}
_lib = None


def _L():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_synth())
        _lib.synth_create.restype = C.c_void_p
        _lib.synth_scans.restype = C.c_longlong
    return _lib


class Sequence:
    """One synthetic sequence: scans (PointXYZIRT, azimuth-major firing order), ground truth poses and IMU samples."""

    def __init__(self, sensor="hdl64", seed=0, sigma=0.02, distort=False, speed=10.0, n_az=None):
        sid, rings, az = SENSORS[sensor]
        self.sensor = sensor
        self.num_lines = rings
        self.n_az = n_az or az
        self.scan_period = 0.1
        self.h = C.c_void_p(_L().synth_create(C.c_ulonglong(seed), sid, self.n_az, C.c_double(sigma), int(distort), C.c_double(speed)))
        self.max_points = rings * self.n_az

    def __del__(self):
        if getattr(self, "h", None):
            _L().synth_destroy(self.h); self.h = None

    def scan(self, frame):
        out = np.zeros(self.max_points, POINT_IRT)
        n = _L().synth_scan(self.h, int(frame), out.ctypes.data_as(C.c_void_p), self.max_points)
        return out[:n].copy()

    def scans(self, frame0, n, threads=None, out=None):
        """Frames [frame0, frame0+n) densely packed. Returns (points, offsets[n+1])."""
        threads = threads or min(os.cpu_count() or 1, 16)
        if out is None:
            out = np.zeros(self.max_points * n, POINT_IRT)
        counts = np.zeros(n, np.int32)
        total = _L().synth_scans(self.h, int(frame0), int(n), out.ctypes.data_as(C.c_void_p), self.max_points,
                                 counts.ctypes.data_as(C.c_void_p), int(threads))
        offsets = np.zeros(n + 1, np.int64); offsets[1:] = np.cumsum(counts)
        return out[:total], offsets

    def pose(self, t):
        T = np.zeros(16)
        _L().synth_pose(self.h, C.c_double(t), T.ctypes.data_as(C.c_void_p))
        return T.reshape(4, 4)

    def imu(self, t):
        q = np.zeros(4)
        _L().synth_imu(self.h, C.c_double(t), q.ctypes.data_as(C.c_void_p))
        return q


def to_xyzi(cloud_irt):
    """VelToIntensityCopy (reference src/odomEstimationClass.cpp:308-318) for numpy clouds."""
    out = np.zeros(len(cloud_irt), POINT_I)
    out["x"] = cloud_irt["x"]; out["y"] = cloud_irt["y"]; out["z"] = cloud_irt["z"]; out["pad0"] = 1.0
    out["intensity"] = cloud_irt["intensity"]
    return out


def pose7_to_matrix(p):
    x, y, z, w = p[0:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = p[4:7]
    return T


def ate(poses7, gt_T):
    """Absolute trajectory error (RMSE, m) after aligning the first estimated pose with the first ground-truth pose."""
    T0 = gt_T[0] @ np.linalg.inv(pose7_to_matrix(poses7[0]))
    err = [np.linalg.norm((T0 @ pose7_to_matrix(p))[:3, 3] - g[:3, 3]) for p, g in zip(poses7, gt_T)]
    return float(np.sqrt(np.mean(np.square(err)))), float(np.max(err))
