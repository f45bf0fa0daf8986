"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what include/floam_b200.h declares.
No compute call is made here (there is no GPU in the build container)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "floam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(floam_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_binding_list(capi):
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(capi):
    L = capi.lib()
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    assert not missing, missing
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.lib_path()], text=True)
    exported = set(re.findall(r" T (floam_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported


def test_library_is_built_for_sm_100a_only(capi):
    capi.lib()
    out = subprocess.check_output(["cuobjdump", "--list-elf", capi.lib_path()], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_defaults_and_loss_strings(capi):
    p = capi.default_params()
    # code defaults of src/laserProcessingNode.cpp:175-179 and src/odomEstimationNode.cpp:328-330
    assert (p.num_lines, p.scan_period, p.max_distance, p.min_distance, p.map_resolution) == (64, 0.1, 60.0, 2.0, 0.4)
    L = capi.lib()
    assert L.floam_loss_from_string(b"Huber") == capi.LOSS_HUBER          # lower-cased like src/odomEstimationClass.cpp:23
    assert L.floam_loss_from_string(b"Cauchy") == capi.LOSS_TRIVIAL       # Q1: "cauchy" leaves loss_function = nullptr
    assert L.floam_loss_from_string(b"anything") == capi.LOSS_TRIVIAL
    assert capi.status_string(capi.NO_IMU) == "no imu data"


def test_point_layouts(capi):
    # vel_point::PointXYZIRT (include/lidar.h:14-32) and pcl::PointXYZI are 32 bytes with intensity @16, ring @20, time @24
    assert capi.POINT_IRT.itemsize == 32 and capi.POINT_I.itemsize == 32
    assert capi.POINT_IRT.fields["intensity"][1] == 16 and capi.POINT_IRT.fields["ring"][1] == 20 and capi.POINT_IRT.fields["time"][1] == 24
    assert capi.POINT_I.fields["intensity"][1] == 16


def test_pointcloud2_layout_and_packing_helper(capi):
    # floam_pc2_layout is 11 32-bit words; the helper that builds msg.data for the GPU tests is checked here against a hand-built
    # Velodyne XYZIRT point (22 bytes: x y z intensity f32 @0/4/8/12, ring u16 @16, time f32 @18) and a big-endian, row-padded layout
    import ctypes
    import struct
    import numpy as np
    assert ctypes.sizeof(capi.Pc2Layout) == 44
    pts = np.zeros(4, capi.POINT_IRT)
    pts["x"] = [1.5, -2.0, 3.25, 0.0]; pts["y"] = [0.5, 8.0, -1.0, 2.0]; pts["z"] = [9.0, 0.25, 0.0, -4.0]
    pts["intensity"] = [0.1, 0.2, 0.3, 0.4]; pts["ring"] = [0, 7, 63, 15]; pts["time"] = [0.0, 0.01, 0.05, 0.099]
    L = capi.pc2_layout(4, 22)
    raw = capi.pack_pointcloud2(pts, L)
    assert raw.dtype == np.uint8 and len(raw) == 4 * 22
    for i in range(4):
        x, y, z, it, ring, t = struct.unpack_from("<ffffHf", raw.tobytes(), i * 22)
        assert (x, y, z, ring) == (pts["x"][i], pts["y"][i], pts["z"][i], pts["ring"][i])
        assert np.float32(it) == pts["intensity"][i] and np.float32(t) == pts["time"][i]
    L = capi.pc2_layout(4, 24, x=4, y=8, z=12, intensity=16, ring=0, time=20, height=2, row_step=2 * 24 + 8, bigendian=True)
    raw = capi.pack_pointcloud2(pts, L).tobytes()
    assert len(raw) == 2 * (2 * 24 + 8)
    for i in range(4):
        base = (i // 2) * L.row_step + (i % 2) * 24
        assert struct.unpack_from(">H", raw, base)[0] == pts["ring"][i] and struct.unpack_from(">f", raw, base + 12)[0] == pts["z"][i]
    assert raw[2 * 24:2 * 24 + 8] == b"\xab" * 8          # row padding is left alone
    # the library refuses a call without a context before it looks at anything else (no GPU here)
    out = np.zeros(4, capi.POINT_IRT)
    rc = capi.lib().floam_unpack_pointcloud2(None, raw, ctypes.byref(L), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == capi.ERR_ARG


def test_no_cpu_fallback_without_device(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.FloamError) as e:
        capi.Context()
    assert e.value.status == capi.ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "floam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "pyoracle" not in txt and "floam_oracle" not in txt and "oracle/" not in txt.replace("touches oracle/", ""), os.path.join(base, f)


REFERENCE = "/root/reference"


@pytest.mark.parametrize("node", ["laserProcessingNode", "odomEstimationNode", "laserMappingNode"])
def test_unmodified_reference_nodes_parse_against_the_shims(node):
    """The drop-in recipe of INTEGRATION.md section 1, proven on the reference's own node sources: floam_b200/host FIRST on the include
    path with -DFLOAM_B200_WITH_PCL, then the third-party headers (here the stand-ins of the test tree, in a catkin workspace the real
    PCL / Eigen / ROS), then the reference's include/.  The class headers the nodes include resolve to the shims; lidar.h and utils.h
    stay the reference's (PublishCloud, Dump, SavePosegraph ...).  Nothing of the reference is copied: it is read where it lies."""
    src = os.path.join(REFERENCE, "src", node + ".cpp")
    if not os.path.exists(src):
        pytest.skip("/root/reference is not present on this box")
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-DFLOAM_B200_WITH_PCL", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "floam_b200", "host"),
           "-I", os.path.join(ROOT, "oracle", "stubs"), "-I", os.path.join(REFERENCE, "include"), "-H", src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    included = r.stderr
    shim = {"laserProcessingNode": ["laserProcessingClass.h", "dataHandler.h"], "odomEstimationNode": ["odomEstimationClass.h", "dataHandler.h"],
            "laserMappingNode": ["laserMappingClass.h"]}[node]
    for h in shim:      # the class headers came from the shim directory, not from the reference
        assert re.search(r"floam_b200/host/" + re.escape(h), included), h
        assert not re.search(re.escape(REFERENCE) + r"/include/" + re.escape(h), included), h
    assert re.search(re.escape(REFERENCE) + r"/include/lidar\.h", included)       # ... while lidar.h is still the reference's (#include_next)


def test_no_prebuilt_binaries_are_tracked():
    tracked = subprocess.check_output(["git", "ls-files"], cwd=ROOT, text=True).split()
    bad = [f for f in tracked if f.endswith((".so", ".o", ".a")) or f.startswith("floam_b200/lib/") or f.startswith("oracle/_ref/")]
    for f in tracked:       # ELF executables under any name (a compiled probe was committed once)
        path = os.path.join(ROOT, f)
        if os.path.isfile(path):
            with open(path, "rb") as fh:
                if fh.read(4) == b"\x7fELF":
                    bad.append(f)
    assert not bad, bad
