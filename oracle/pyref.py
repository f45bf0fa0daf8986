"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes driver for oracle/_ref/libfloam_ref.so: the REFERENCE'S OWN class sources (laserProcessingClass, dataHandler, lidar,
lidarOptimization, odomEstimationClass, laserMappingClass .cpp, read from /root/reference at build time and compiled unmodified
against the stand-in third-party headers of oracle/stubs/; recipe `make -C oracle ref`).  It exposes the same Python surface as
oracle/pyoracle.py (the library exports the same C entry points), so a test can run one input through the restatement, through the
reference's code and through the CUDA path.

/root/reference does not exist on the GPU box: the library is built in the authoring container and travels with the snapshot.
`available()` says whether it is there; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes as C
import importlib.util
import os
import shutil
import subprocess
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libfloam_ref.so")
REFERENCE_ROOT = "/root/reference"


def build(force=False):
    """Compile the reference's sources where they lie (only possible where /root/reference exists)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        return REF_SO if os.path.exists(REF_SO) else None
    subprocess.check_call(["make", "-C", _HERE, "ref"] + (["-B"] if force else []))
    return REF_SO


def available():
    return os.path.exists(REF_SO) or build() is not None


def _configure(L):
    L.fo_replay_sequence.restype = C.c_double
    L.fo_replay_sequence_stages.restype = C.c_double
    L.fo_odom_knn_queries.restype = C.c_long
    L.fo_backend.restype = C.c_char_p
    for f in ("fo_imu_create", "fo_odom_create", "fo_mapping_create"):
        getattr(L, f).restype = C.c_void_p
    return L


_TMP = []


def fresh_lib():
    """A private copy of the library (own static state).  OdomEstimationClass::KeyFrameUpdate keeps its `first` flag in a
    function-static (src/odomEstimationClass.cpp:323): a second instance in the same image would skip the first-frame
    branch and read keyframes_.back() of an empty vector.  One OdomEstimationClass per loaded copy, like one per process."""
    if not available():
        raise RuntimeError("oracle/_ref/libfloam_ref.so is missing and /root/reference is not here to build it")
    fd, path = tempfile.mkstemp(prefix="libfloam_ref_", suffix=".so")
    os.close(fd)
    shutil.copyfile(REF_SO, path)
    L = _configure(C.CDLL(path))
    os.unlink(path)   # the mapping stays valid; nothing is left behind in /tmp
    _TMP.append(L)
    return L


# a second instance of the pyoracle module whose lib() is the reference build
_spec = importlib.util.spec_from_file_location("oracle._pyoracle_on_ref", os.path.join(_HERE, "pyoracle.py"))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
_SHARED = None


def _lib():
    global _SHARED
    if _SHARED is None:
        _SHARED = fresh_lib()
    return _SHARED


_m.lib = _lib
lib = _lib

POINT_IRT = _m.POINT_IRT
POINT_I = _m.POINT_I
feature_extract = _m.feature_extract        # src/laserProcessingClass.cpp:72-231, the reference's code
se3_plus = _m.se3_plus                      # PoseSE3Parameterization::Plus
euler2quat = _m.euler2quat                  # src/lidar.cpp:8-16
evaluate_residual = _m.evaluate_residual    # Edge/SurfNormAnalyticCostFunction::Evaluate
lm_solve = _m.lm_solve                      # problem built and solved as at src/odomEstimationClass.cpp:83-108
compensate_velocity = _m.compensate_velocity
Imu = _m.Imu                                # dmapping::ImuHandler + CenterTime + Compensate + alignment
Mapping = _m.Mapping                        # LaserMappingClass


def _with_fresh_lib(fn):
    def run(*a, **kw):   # the replay constructs an OdomEstimationClass inside the library: private copy per call (see fresh_lib)
        L = fresh_lib()
        saved = _m.lib
        _m.lib = lambda: L
        try:
            return fn(*a, **kw)
        finally:
            _m.lib = saved
    run.__doc__ = fn.__doc__
    return run


replay_sequence = _with_fresh_lib(_m.replay_sequence)
replay_sequence_stages = _with_fresh_lib(_m.replay_sequence_stages)


def _dump(fn, directory, poses, stamps, clouds, *extra):
    import numpy as np
    P = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(-1, 16)); st = np.ascontiguousarray(stamps, np.float64)
    off = np.zeros(len(clouds) + 1, np.int64); off[1:] = np.cumsum([len(c) for c in clouds])
    cat = np.ascontiguousarray(np.concatenate(clouds) if len(clouds) else np.zeros(0, POINT_I), POINT_I)
    getattr(lib(), fn)(str(directory).encode(), _m._p(P), _m._p(st), _m._p(cat), _m._p(off), len(clouds), *extra)


def save_posegraph(directory, poses, stamps, clouds):      # SavePosegraph, src/utils.cpp:3-79 (the reference's code)
    _dump("fo_save_posegraph", directory, poses, stamps, clouds)


def save_odom(directory, poses, stamps, clouds):           # SaveOdom, src/utils.cpp:82-106
    _dump("fo_save_odom", directory, poses, stamps, clouds)


def save_balm(directory, poses, stamps, clouds):           # SavePosesHomogeneousBALM, src/odomEstimationNode.cpp:93-121
    _dump("fo_save_balm", directory, poses, stamps, clouds)


def save_merged(directory, poses, stamps, clouds, downsample_size, total_order=True):   # SaveMerged, src/odomEstimationNode.cpp:66-92
    _dump("fo_save_merged", directory, poses, stamps, clouds, C.c_double(downsample_size), int(total_order))


class Odom(_m.Odom):
    """OdomEstimationClass, the reference's code; each instance gets its own copy of the library (see fresh_lib)."""

    def __init__(self, *a, **kw):
        kw["_lib"] = fresh_lib()
        super().__init__(*a, **kw)

    def debug(self):
        """Only what is observable from outside the class: downsampled clouds of keyframes, last LM summary, solve count, keyframe."""
        import numpy as np
        L = self._L
        sc = np.zeros(2, np.int32); L.fo_odom_debug(self.h, 10, _m._p(sc), 2)
        lm = np.zeros(47); L.fo_odom_debug(self.h, 9, _m._p(lm), 47)

        def cloud(what):
            n = L.fo_odom_debug(self.h, what, None, 0)
            a = np.zeros(max(n, 1), POINT_I)
            L.fo_odom_debug(self.h, what, _m._p(a), n)
            return a[:n]
        return {"ds_edge": cloud(0), "ds_surf": cloud(1), "outer_iterations": int(sc[0]), "keyframe": bool(sc[1]),
                "lm": {"iterations": int(lm[0]), "accepted": int(lm[1]), "initial_cost": lm[2], "final_cost": lm[3],
                       "termination": int(lm[4]), "H0": lm[5:41].reshape(6, 6).copy(), "g0": lm[41:47].copy()}}
