"""ctypes binding of the C ABI in include/floam_b200.h (libfloam_b200.so) — what tests/ and bench.py drive.

The product library is CUDA-only: importing this module on a machine without the built .so, or creating a Context without an
sm_100 device, fails loudly.  Nothing here falls back to a CPU implementation (and nothing here touches oracle/).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

POINT_IRT = np.dtype({"names": ["x", "y", "z", "pad0", "intensity", "ring", "pad1", "time", "pad2"],
                      "formats": ["<f4", "<f4", "<f4", "<f4", "<f4", "<u2", "<u2", "<f4", "<f4"], "itemsize": 32})
POINT_I = np.dtype({"names": ["x", "y", "z", "pad0", "intensity", "p1", "p2", "p3"], "formats": ["<f4"] * 8, "itemsize": 32})

OK, ERR_NO_DEVICE, ERR_CUDA, ERR_CAPACITY, ERR_ARG, NO_IMU, ERR_NONFINITE = range(7)
LOSS_TRIVIAL, LOSS_HUBER, LOSS_CAUCHY_TRUE = 0, 1, 2
VANILLA, INITIAL_ITERATION, REFINEMENT_AND_UPDATE = 0, 1, 2
FIX_SINGLE_PREDICTION, FIX_ROTATED_VELOCITY, FIX_IMU_SLERP = 1, 2, 4
(DBG_DS_EDGE, DBG_DS_SURF, DBG_EDGE_KNN, DBG_SURF_KNN, DBG_EDGE_D2, DBG_SURF_D2, DBG_EDGE_OK, DBG_SURF_OK, DBG_RESIDUALS, DBG_LM,
 DBG_SCALARS, DBG_FEATURE_SRC_EDGE, DBG_FEATURE_SRC_SURF, DBG_CLOCKS, DBG_TIMELINE) = range(15)


class Params(C.Structure):
    _fields_ = [("num_lines", C.c_int), ("scan_period", C.c_double), ("vertical_angle", C.c_double), ("max_distance", C.c_double),
                ("min_distance", C.c_double), ("map_resolution", C.c_double), ("loss", C.c_int), ("max_scan_points", C.c_int),
                ("max_map_points", C.c_int), ("max_global_map_points", C.c_int), ("max_grid_cells", C.c_int), ("fixes", C.c_int)]


# every symbol include/floam_b200.h declares (tests/test_abi.py checks the header against this list and the .so against both)
SYMBOLS = [
    "floam_params_default", "floam_loss_from_string", "floam_status_string", "floam_version", "floam_create", "floam_destroy",
    "floam_alloc_pinned", "floam_free_pinned", "floam_set_graphs", "floam_set_map_merge", "floam_voxel_grid_update", "floam_imu_push", "floam_imu_get", "floam_imu_size", "floam_imu_time_contained", "floam_deskew_align",
    "floam_feature_extract", "floam_odom_init_map", "floam_odom_update", "floam_odom_update_xyzi", "floam_odom_get", "floam_odom_map_sizes",
    "floam_odom_get_map", "floam_odom_set_state", "floam_odom_get_state", "floam_odom_set_map", "floam_process_scan", "floam_process_submit",
    "floam_process_wait", "floam_stage_scans", "floam_process_staged", "floam_mapping_update", "floam_mapping_get_map", "floam_mapping_get_changed_cells", "floam_voxel_grid",
    "floam_crop_box", "floam_knn5", "floam_debug_fetch", "floam_launch_count", "floam_last_frame_ms", "floam_replay_staged",
    "floam_set_kernel_timing", "floam_kernel_slots", "floam_kernel_name", "floam_kernel_timing", "floam_deskew_align_ex",
    "floam_write_pcd_binary", "floam_save_posegraph", "floam_save_odom", "floam_save_balm", "floam_save_merged",
    "floam_compensate_velocity", "floam_process_submit_imu", "floam_process_scan_imu", "floam_unpack_pointcloud2", "floam_process_submit_pc2",
]

_lib = None


class Pc2Layout(C.Structure):
    """floam_pc2_layout: where the PointXYZIRT fields sit inside a sensor_msgs/PointCloud2 point (offset -1 = field absent)."""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("point_step", C.c_uint32), ("row_step", C.c_uint32),
                ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32), ("off_intensity", C.c_int32), ("off_ring", C.c_int32),
                ("off_time", C.c_int32), ("is_bigendian", C.c_int32)]


def pc2_layout(n, point_step, x=0, y=4, z=8, intensity=12, ring=16, time=18, height=1, row_step=None, bigendian=False):
    """Defaults: the Velodyne driver's XYZIRT message (22 bytes per point)."""
    width = n // height
    return Pc2Layout(width, height, point_step, width * point_step if row_step is None else row_step, x, y, z, intensity, ring, time, int(bigendian))


def pack_pointcloud2(pts, layout):
    """Test/bench helper: the msg.data bytes a driver would publish for `pts` (POINT_IRT) in the given layout; padding bytes are 0xAB."""
    n = len(pts)
    assert n == layout.width * layout.height
    raw = np.full((layout.height, layout.row_step), 0xAB, np.uint8)
    body = np.full((n, layout.point_step), 0xAB, np.uint8)
    order = ">" if layout.is_bigendian else "<"
    for name, off, dt in (("x", layout.off_x, "f4"), ("y", layout.off_y, "f4"), ("z", layout.off_z, "f4"), ("intensity", layout.off_intensity, "f4"),
                          ("ring", layout.off_ring, "u2"), ("time", layout.off_time, "f4")):
        if off >= 0:
            b = np.ascontiguousarray(pts[name].astype(order + dt)).view(np.uint8).reshape(n, -1)
            body[:, off:off + b.shape[1]] = b
    raw[:, :layout.width * layout.point_step] = body.reshape(layout.height, layout.width * layout.point_step)
    return np.ascontiguousarray(raw.reshape(-1))


class FloamError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        super().__init__("%s failed: status %d (%s)" % (where, status, status_string(status)))


def lib_path():
    return _build.CUDA_LIB


def lib():
    """Loads libfloam_b200.so (building it with nvcc if it is missing or stale). No fallback if that fails."""
    global _lib
    if _lib is None:
        path = _build.build_cuda()
        L = C.CDLL(path)
        L.floam_status_string.restype = C.c_char_p
        L.floam_version.restype = C.c_char_p
        L.floam_alloc_pinned.restype = C.c_void_p
        L.floam_alloc_pinned.argtypes = [C.c_size_t]
        L.floam_free_pinned.argtypes = [C.c_void_p]
        L.floam_kernel_name.restype = C.c_char_p
        _lib = L
    return _lib


def status_string(status):
    return lib().floam_status_string(int(status)).decode()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc, where, allow=()):
    if rc != OK and rc not in allow:
        raise FloamError(rc, where)
    return rc


def default_params(**kw):
    p = Params()
    lib().floam_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "loss" and isinstance(v, str):
            v = lib().floam_loss_from_string(v.encode())
        setattr(p, k, v)
    return p


def _dump_args(poses, stamps, clouds):
    """clouds: list of POINT_I arrays -> (poses16, stamps, concatenated cloud, offsets)."""
    P = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(-1, 16))
    st = np.ascontiguousarray(stamps, np.float64)
    off = np.zeros(len(clouds) + 1, np.int64); off[1:] = np.cumsum([len(c) for c in clouds])
    cat = np.ascontiguousarray(np.concatenate(clouds) if len(clouds) else np.zeros(0, POINT_I), POINT_I)
    return P, st, cat, off


def write_pcd_binary(path, pts):
    pts = np.ascontiguousarray(pts, POINT_I)
    _check(lib().floam_write_pcd_binary(str(path).encode(), _p(pts), len(pts)), "floam_write_pcd_binary")


def save_posegraph(directory, poses, stamps, clouds):
    """SavePosegraph (reference src/utils.cpp:3-79). Host I/O only: needs no context."""
    P, st, cat, off = _dump_args(poses, stamps, clouds)
    _check(lib().floam_save_posegraph(str(directory).encode(), _p(P), _p(st), _p(cat), _p(off), len(clouds)), "floam_save_posegraph")


def save_odom(directory, poses, stamps, clouds):
    P, st, cat, off = _dump_args(poses, stamps, clouds)
    _check(lib().floam_save_odom(str(directory).encode(), _p(P), _p(st), _p(cat), _p(off), len(clouds)), "floam_save_odom")


def save_balm(directory, poses, stamps, clouds):
    P, st, cat, off = _dump_args(poses, stamps, clouds)
    _check(lib().floam_save_balm(str(directory).encode(), _p(P), _p(st), _p(cat), _p(off), len(clouds)), "floam_save_balm")


class PinnedBuffer:
    """Page-locked host array of PointXYZIRT (floam_alloc_pinned) for overlapped scan uploads."""

    def __init__(self, n_points):
        self.nbytes = int(n_points) * 32
        self.ptr = lib().floam_alloc_pinned(self.nbytes)
        if not self.ptr:
            raise MemoryError("floam_alloc_pinned(%d)" % self.nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr)).view(POINT_IRT)

    def close(self):
        if self.ptr:
            self.array = None
            lib().floam_free_pinned(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        self.close()


class Context:
    """One floam_ctx: one device, one stream, one sequence (include/floam_b200.h)."""

    def __init__(self, device=0, **params):
        self.params = default_params(**params)
        self.h = C.c_void_p()
        _check(lib().floam_create(C.byref(self.params), int(device), C.byref(self.h)), "floam_create")

    def close(self):
        if getattr(self, "h", None):
            lib().floam_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_graphs(self, enabled):
        _check(lib().floam_set_graphs(self.h, int(enabled)), "floam_set_graphs")

    def set_map_merge(self, mode):
        """Keyframe map update: 0 = full re-sort, 1 = by map size (default), 2 = sort only the out-of-place points and merge. Same maps."""
        _check(lib().floam_set_map_merge(self.h, int(mode)), "floam_set_map_merge")

    # ---- IMU ----
    def imu_push(self, stamp, q_xyzw):
        q = np.ascontiguousarray(q_xyzw, np.float64)
        _check(lib().floam_imu_push(self.h, C.c_double(stamp), _p(q)), "floam_imu_push")

    def imu_get(self, stamp):
        q = np.zeros(4); valid = C.c_int()
        _check(lib().floam_imu_get(self.h, C.c_double(stamp), _p(q), C.byref(valid)), "floam_imu_get")
        return bool(valid.value), q

    def imu_size(self):
        n = C.c_int()
        _check(lib().floam_imu_size(self.h, C.byref(n)), "floam_imu_size")
        return n.value

    def deskew_align(self, pts, stamp_us, extr_xyzw):
        """In place. Returns (status, new_stamp_us); status is OK or NO_IMU like dmapping::Compensate's bool."""
        assert pts.dtype == POINT_IRT and pts.flags.c_contiguous
        st = C.c_uint64(int(stamp_us)); ex = np.ascontiguousarray(extr_xyzw, np.float64)
        rc = _check(lib().floam_deskew_align(self.h, _p(pts), len(pts), C.byref(st), _p(ex)), "floam_deskew_align", allow=(NO_IMU,))
        return rc, st.value

    def compensate_velocity(self, pts, v):
        assert pts.dtype == POINT_IRT and pts.flags.c_contiguous
        v = np.ascontiguousarray(v, np.float64)
        _check(lib().floam_compensate_velocity(self.h, _p(pts), len(pts), _p(v)), "floam_compensate_velocity")

    def deskew_align_ex(self, pts, stamp_us, extr_xyzw, flags):
        assert pts.dtype == POINT_IRT and pts.flags.c_contiguous
        st = C.c_uint64(int(stamp_us)); ex = np.ascontiguousarray(extr_xyzw, np.float64)
        rc = _check(lib().floam_deskew_align_ex(self.h, _p(pts), len(pts), C.byref(st), _p(ex), int(flags)), "floam_deskew_align_ex", allow=(NO_IMU,))
        return rc, st.value

    # ---- feature extraction ----
    def feature_extract(self, pts, with_src=False):
        pts = np.ascontiguousarray(pts, POINT_IRT)
        n = len(pts)
        edge = np.zeros(max(n, 1), POINT_IRT); surf = np.zeros(max(n, 1), POINT_IRT)
        ne = C.c_int(); ns = C.c_int()
        _check(lib().floam_feature_extract(self.h, _p(pts), n, _p(edge), len(edge), C.byref(ne), _p(surf), len(surf), C.byref(ns)), "floam_feature_extract")
        if not with_src:
            return edge[:ne.value], surf[:ns.value]
        es = self.debug_fetch(DBG_FEATURE_SRC_EDGE, np.int32); ss = self.debug_fetch(DBG_FEATURE_SRC_SURF, np.int32)
        return edge[:ne.value], surf[:ns.value], es, ss

    # ---- odometry ----
    def odom_init_map(self, edge, surf):
        edge = np.ascontiguousarray(edge, POINT_I); surf = np.ascontiguousarray(surf, POINT_I)
        _check(lib().floam_odom_init_map(self.h, _p(edge), len(edge), _p(surf), len(surf)), "floam_odom_init_map")

    def odom_set_map(self, edge, surf):
        edge = np.ascontiguousarray(edge, POINT_I); surf = np.ascontiguousarray(surf, POINT_I)
        _check(lib().floam_odom_set_map(self.h, _p(edge), len(edge), _p(surf), len(surf)), "floam_odom_set_map")

    def odom_update(self, edge, surf, deskew=False):
        assert edge.dtype == POINT_IRT and surf.dtype == POINT_IRT and edge.flags.c_contiguous and surf.flags.c_contiguous
        pose = np.zeros(7)
        _check(lib().floam_odom_update(self.h, _p(edge), len(edge), _p(surf), len(surf), int(deskew), _p(pose)), "floam_odom_update")
        return pose

    def odom_update_xyzi(self, edge, surf, update_type=VANILLA):
        edge = np.ascontiguousarray(edge, POINT_I); surf = np.ascontiguousarray(surf, POINT_I)
        pose = np.zeros(7)
        _check(lib().floam_odom_update_xyzi(self.h, _p(edge), len(edge), _p(surf), len(surf), int(update_type), _p(pose)), "floam_odom_update_xyzi")
        return pose

    def odom_get(self):
        T = np.zeros(16); v = np.zeros(3)
        _check(lib().floam_odom_get(self.h, _p(T), _p(v)), "floam_odom_get")
        return T.reshape(4, 4), v

    def odom_get_state(self):
        T = np.zeros(16); L = np.zeros(16); oc = C.c_int()
        _check(lib().floam_odom_get_state(self.h, _p(T), _p(L), C.byref(oc)), "floam_odom_get_state")
        return T.reshape(4, 4), L.reshape(4, 4), oc.value

    def odom_set_state(self, odom, last_odom, optimization_count):
        o = np.ascontiguousarray(odom, np.float64).reshape(16); l = np.ascontiguousarray(last_odom, np.float64).reshape(16)
        _check(lib().floam_odom_set_state(self.h, _p(o), _p(l), int(optimization_count)), "floam_odom_set_state")

    def odom_map_sizes(self):
        ne = C.c_int(); ns = C.c_int()
        _check(lib().floam_odom_map_sizes(self.h, C.byref(ne), C.byref(ns)), "floam_odom_map_sizes")
        return ne.value, ns.value

    def odom_get_map(self):
        ne, ns = self.odom_map_sizes()
        e = np.zeros(max(ne, 1), POINT_I); s = np.zeros(max(ns, 1), POINT_I)
        _check(lib().floam_odom_get_map(self.h, _p(e), len(e), _p(s), len(s)), "floam_odom_get_map")
        return e[:ne], s[:ns]

    # ---- fused frame path ----
    def process_scan(self, pts, deskew=False):
        pts = np.ascontiguousarray(pts, POINT_IRT)
        pose = np.zeros(7)
        _check(lib().floam_process_scan(self.h, _p(pts), len(pts), int(deskew), _p(pose)), "floam_process_scan")
        return pose

    def process_scan_imu(self, pts, stamp_us, extr_xyzw, deskew=False):
        """IMU deskew + alignment + features + odometry in one device pass. Returns (status, pose, new_stamp_us); status NO_IMU = scan skipped."""
        pts = np.ascontiguousarray(pts, POINT_IRT)
        pose = np.zeros(7); st = C.c_uint64(int(stamp_us)); ex = np.ascontiguousarray(extr_xyzw, np.float64)
        rc = _check(lib().floam_process_scan_imu(self.h, _p(pts), len(pts), C.byref(st), _p(ex), int(deskew), _p(pose)), "floam_process_scan_imu", allow=(NO_IMU,))
        return rc, pose, st.value

    def process_submit_imu(self, pts, stamp_us, extr_xyzw, deskew=False):
        """floam_process_submit with the IMU steps folded in; returns (status, new_stamp_us). status NO_IMU = nothing was submitted.
        pts must stay alive until the matching process_wait returns."""
        assert pts.dtype == POINT_IRT and pts.flags.c_contiguous
        st = C.c_uint64(int(stamp_us)); ex = np.ascontiguousarray(extr_xyzw, np.float64)
        rc = _check(lib().floam_process_submit_imu(self.h, _p(pts), len(pts), C.byref(st), _p(ex), int(deskew)), "floam_process_submit_imu", allow=(NO_IMU,))
        return rc, st.value

    def process_submit(self, pts, n=None, deskew=False):
        """pts must stay alive (ideally a PinnedBuffer.array slice) until the matching process_wait returns."""
        assert pts.dtype == POINT_IRT and pts.flags.c_contiguous
        _check(lib().floam_process_submit(self.h, _p(pts), len(pts) if n is None else int(n), int(deskew)), "floam_process_submit")

    def unpack_pointcloud2(self, raw, layout):
        """pcl::fromROSMsg on the device: raw msg.data bytes -> POINT_IRT array."""
        raw = np.ascontiguousarray(raw, np.uint8)
        out = np.zeros(layout.width * layout.height, POINT_IRT)
        _check(lib().floam_unpack_pointcloud2(self.h, _p(raw), C.byref(layout), _p(out)), "floam_unpack_pointcloud2")
        return out

    def process_submit_pc2(self, raw, layout, deskew=False, stamp_us=None, extrinsics_xyzw=None):
        """raw must stay alive until the matching process_wait returns. With stamp_us / extrinsics the IMU steps run too; returns the
        re-centred stamp in that case."""
        assert raw.dtype == np.uint8 and raw.flags.c_contiguous
        if stamp_us is None:
            _check(lib().floam_process_submit_pc2(self.h, _p(raw), C.byref(layout), None, None, int(deskew)), "floam_process_submit_pc2")
            return None
        st = C.c_uint64(int(stamp_us)); ex = np.ascontiguousarray(extrinsics_xyzw, np.float64)
        _check(lib().floam_process_submit_pc2(self.h, _p(raw), C.byref(layout), C.byref(st), _p(ex), int(deskew)), "floam_process_submit_pc2")
        return st.value

    def process_wait(self):
        pose = np.zeros(7)
        _check(lib().floam_process_wait(self.h, _p(pose)), "floam_process_wait")
        return pose

    def stage_scans(self, pts, offsets):
        pts = np.ascontiguousarray(pts, POINT_IRT); offsets = np.ascontiguousarray(offsets, np.int64)
        _check(lib().floam_stage_scans(self.h, _p(pts), _p(offsets), len(offsets) - 1), "floam_stage_scans")

    def process_staged(self, frame, deskew=False):
        pose = np.zeros(7)
        _check(lib().floam_process_staged(self.h, int(frame), int(deskew), _p(pose)), "floam_process_staged")
        return pose

    def replay_staged(self, first, count, deskew=False):
        """Frames [first, first+count) back to back on the device. Returns (poses[count,7], device milliseconds)."""
        poses = np.zeros((count, 7)); ms = C.c_float()
        _check(lib().floam_replay_staged(self.h, int(first), int(count), int(deskew), _p(poses), C.byref(ms)), "floam_replay_staged")
        return poses, ms.value

    def set_kernel_timing(self, enabled):
        _check(lib().floam_set_kernel_timing(self.h, int(enabled)), "floam_set_kernel_timing")

    def kernel_timing(self):
        """{kernel class: (total ms, launches)} accumulated since timing was enabled."""
        out = {}
        for k in range(lib().floam_kernel_slots()):
            ms = C.c_double(); n = C.c_int64()
            _check(lib().floam_kernel_timing(self.h, k, C.byref(ms), C.byref(n)), "floam_kernel_timing")
            if n.value:
                out[lib().floam_kernel_name(k).decode()] = (ms.value, n.value)
        return out

    # ---- LaserMappingClass ----
    def mapping_update(self, pts, pose):
        pts = np.ascontiguousarray(pts, POINT_I); T = np.ascontiguousarray(pose, np.float64).reshape(16)
        _check(lib().floam_mapping_update(self.h, _p(pts), len(pts), _p(T)), "floam_mapping_update")

    def mapping_get_map(self):
        n = C.c_int()
        _check(lib().floam_mapping_get_map(self.h, None, 0, C.byref(n)), "floam_mapping_get_map")
        out = np.zeros(max(n.value, 1), POINT_I)
        _check(lib().floam_mapping_get_map(self.h, _p(out), len(out), C.byref(n)), "floam_mapping_get_map")
        return out[:n.value]

    # ---- stage entry points ----
    def mapping_get_changed_cells(self):
        """Incremental getMap(): (points, cells[n,3]) of every 50 m cell changed since the previous call; clears the change marks."""
        n = C.c_int()
        _check(lib().floam_mapping_get_changed_cells(self.h, None, None, 0, C.byref(n)), "floam_mapping_get_changed_cells")
        out = np.zeros(max(n.value, 1), POINT_I); cells = np.zeros((max(n.value, 1), 3), np.int32)
        _check(lib().floam_mapping_get_changed_cells(self.h, _p(out), _p(cells), len(out), C.byref(n)), "floam_mapping_get_changed_cells")
        return out[:n.value], cells[:n.value]

    def voxel_grid(self, pts, leaf):
        pts = np.ascontiguousarray(pts, POINT_I)
        out = np.zeros(max(len(pts), 1), POINT_I); n = C.c_int()
        _check(lib().floam_voxel_grid(self.h, _p(pts), len(pts), C.c_float(leaf), _p(out), len(out), C.byref(n)), "floam_voxel_grid")
        return out[:n.value]

    def voxel_grid_update(self, map_pts, new_pts, leaf, mn=None, mx=None):
        """VoxelGrid(CropBox(map_pts + new_pts)) through the keyframe update's merge path (floam_voxel_grid_update)."""
        map_pts = np.ascontiguousarray(map_pts, POINT_I); new_pts = np.ascontiguousarray(new_pts, POINT_I)
        out = np.zeros(max(len(map_pts) + len(new_pts), 1), POINT_I); n = C.c_int()
        if mn is not None:
            mn = np.ascontiguousarray(mn, np.float32); mx = np.ascontiguousarray(mx, np.float32)
        _check(lib().floam_voxel_grid_update(self.h, _p(map_pts), len(map_pts), _p(new_pts), len(new_pts), C.c_float(leaf),
                                             _p(mn) if mn is not None else None, _p(mx) if mx is not None else None, _p(out), len(out), C.byref(n)),
               "floam_voxel_grid_update")
        return out[:n.value]

    def crop_box(self, pts, mn, mx):
        pts = np.ascontiguousarray(pts, POINT_I)
        mn = np.ascontiguousarray(mn, np.float32); mx = np.ascontiguousarray(mx, np.float32)
        out = np.zeros(max(len(pts), 1), POINT_I); n = C.c_int()
        _check(lib().floam_crop_box(self.h, _p(pts), len(pts), _p(mn), _p(mx), _p(out), len(out), C.byref(n)), "floam_crop_box")
        return out[:n.value]

    def knn5(self, map_pts, queries):
        map_pts = np.ascontiguousarray(map_pts, POINT_I); queries = np.ascontiguousarray(queries, POINT_I)
        ids = np.full((max(len(queries), 1), 5), -1, np.int32); d2 = np.zeros((max(len(queries), 1), 5), np.float32)
        _check(lib().floam_knn5(self.h, _p(map_pts), len(map_pts), _p(queries), len(queries), _p(ids), _p(d2)), "floam_knn5")
        return ids[:len(queries)], d2[:len(queries)]

    def save_merged(self, directory, poses, clouds, downsample_size):
        """SaveMerged (reference src/odomEstimationNode.cpp:66-92): transform + merge + VoxelGrid on the device, two PCD files."""
        P, _, cat, off = _dump_args(poses, np.zeros(len(clouds)), clouds)
        _check(lib().floam_save_merged(self.h, str(directory).encode(), _p(P), _p(cat), _p(off), len(clouds), C.c_double(downsample_size)), "floam_save_merged")

    # ---- taps / accounting ----
    def debug_fetch(self, what, dtype):
        nb = C.c_size_t()
        _check(lib().floam_debug_fetch(self.h, int(what), None, C.c_size_t(0), C.byref(nb)), "floam_debug_fetch")
        item = np.dtype(dtype).itemsize
        a = np.zeros(max(nb.value // item, 1), dtype)
        _check(lib().floam_debug_fetch(self.h, int(what), _p(a), C.c_size_t(a.nbytes), C.byref(nb)), "floam_debug_fetch")
        return a[:nb.value // item]

    def debug(self):
        lm = self.debug_fetch(DBG_LM, np.float64)
        sc = self.debug_fetch(DBG_SCALARS, np.int32)
        return {
            "ds_edge": self.debug_fetch(DBG_DS_EDGE, POINT_I), "ds_surf": self.debug_fetch(DBG_DS_SURF, POINT_I),
            "edge_knn": self.debug_fetch(DBG_EDGE_KNN, np.int32).reshape(-1, 5), "surf_knn": self.debug_fetch(DBG_SURF_KNN, np.int32).reshape(-1, 5),
            "edge_d2": self.debug_fetch(DBG_EDGE_D2, np.float32).reshape(-1, 5), "surf_d2": self.debug_fetch(DBG_SURF_D2, np.float32).reshape(-1, 5),
            "edge_ok": self.debug_fetch(DBG_EDGE_OK, np.uint8), "surf_ok": self.debug_fetch(DBG_SURF_OK, np.uint8),
            "residuals": self.debug_fetch(DBG_RESIDUALS, np.float64).reshape(-1, 10),
            "lm": {"iterations": int(lm[0]), "accepted": int(lm[1]), "initial_cost": lm[2], "final_cost": lm[3], "termination": int(lm[4]),
                   "H0": lm[5:41].reshape(6, 6).copy(), "g0": lm[41:47].copy()},
            "outer_iterations": int(sc[0]), "keyframe": bool(sc[1]), "n_corr": int(sc[4]), "skip_solve": bool(sc[5]),
        }

    def launch_count(self, reset=False):
        n = C.c_int64()
        _check(lib().floam_launch_count(self.h, C.byref(n), int(reset)), "floam_launch_count")
        return n.value

    def last_frame_ms(self):
        ms = C.c_float()
        _check(lib().floam_last_frame_ms(self.h, C.byref(ms)), "floam_last_frame_ms")
        return ms.value
