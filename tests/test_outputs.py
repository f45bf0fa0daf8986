"""On-disk outputs: the product's writers against the REFERENCE'S OWN writers — SavePosegraph / SaveOdom
(src/utils.cpp:3-106) and SaveMerged / SavePosesHomogeneousBALM (src/odomEstimationNode.cpp:66-121), compiled unmodified into
oracle/_ref — byte for byte, file by file.  The text formats (iostream precision, Eigen's aligned matrix printing, Boost.Format
directives) are the reference's; pcl::io::savePCDFileBinary behind them is the stand-in of oracle/stubs (PCL 1.8 binary PCD layout).
The host writers need no GPU; SaveMerged (transform + VoxelGrid on the device) is the GPU test at the bottom."""
import filecmp
import os

import numpy as np
import pytest


def tree(root):
    out = {}
    for base, _, files in os.walk(root):
        for f in files:
            p = os.path.join(base, f)
            out[os.path.relpath(p, root)] = p
    return out


def assert_same_tree(a, b):
    ta, tb = tree(a), tree(b)
    assert sorted(ta) == sorted(tb) and len(ta) > 0, (sorted(ta), sorted(tb))
    for rel in ta:
        assert filecmp.cmp(ta[rel], tb[rel], shallow=False), rel


def scene(po, n=5, seed=0):
    rng = np.random.default_rng(seed)
    poses, stamps, clouds = [], [], []
    for i in range(n):
        ang = rng.uniform(-0.4, 0.4, 3)
        cx, cy, cz = np.cos(ang); sx, sy, sz = np.sin(ang)
        R = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = [12.3456789 * i, -0.001234 * i + 1e-7, rng.normal() * 100]
        poses.append(T)
        stamps.append(1634567890.0 + 0.1 * i + (0.123456789 if i == 2 else 0.0))
        c = np.zeros(int(rng.integers(0 if i == 3 else 50, 400)), po.POINT_I)
        for k in "xyz":
            c[k] = rng.uniform(-30, 30, len(c)).astype(np.float32)
        c["intensity"] = rng.random(len(c)).astype(np.float32); c["pad0"] = 1.0
        clouds.append(c)
    return poses, stamps, clouds


def test_pcd_binary_layout(capi, po, tmp_path):
    _, _, clouds = scene(po)
    p = tmp_path / "c.pcd"
    capi.write_pcd_binary(p, clouds[0])
    raw = p.read_bytes()
    head, body = raw.split(b"DATA binary\n", 1)
    n = len(clouds[0])
    assert head.decode().splitlines() == ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7", "FIELDS x y z intensity", "SIZE 4 4 4 4", "TYPE F F F F",
                                          "COUNT 1 1 1 1", "WIDTH %d" % n, "HEIGHT 1", "VIEWPOINT 0 0 0 1 0 0 0", "POINTS %d" % n]
    rec = np.frombuffer(body, np.float32).reshape(n, 4)
    assert np.array_equal(rec, np.stack([clouds[0][k] for k in ("x", "y", "z", "intensity")], 1))


@pytest.mark.parametrize("n", [5, 1, 0])
def test_posegraph_odom_balm_equal_the_reference_writers(capi, po, pr, tmp_path, n):
    poses, stamps, clouds = scene(po, max(n, 1), seed=n)
    poses, stamps, clouds = poses[:n], stamps[:n], clouds[:n]
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    if n > 0:      # the reference's `i < poses.size() - 1` loop (src/utils.cpp:31) wraps around for an empty pose list: not called with n = 0
        capi.save_posegraph(ours / "posegraph", poses, stamps, clouds); pr.save_posegraph(ref / "posegraph", poses, stamps, clouds)
    capi.save_odom(ours / "odom", poses, stamps, clouds); pr.save_odom(ref / "odom", poses, stamps, clouds)
    capi.save_balm(str(ours / "BALM") + "/", poses, stamps, clouds); pr.save_balm(str(ref / "BALM") + "/", poses, stamps, clouds)
    if n == 0:
        assert (ours / "BALM" / "alidarPose.csv").read_bytes() == (ref / "BALM" / "alidarPose.csv").read_bytes() == b""
        return
    assert_same_tree(ours, ref)
    g2o = (ours / "posegraph" / "graph.g2o").read_text().splitlines()
    assert g2o[0].startswith("VERTEX_SE3:QUAT 0 ") and g2o[n] == "FIX 0" and len(g2o) == 2 * n
    data = (ours / "posegraph" / "000000" / "data").read_text().splitlines()
    assert data[0] == "stamp 1634567890 0" and data[1] == "estimate" and data[6] == "odom" and data[-1] == "id 0"


@pytest.mark.gpu
def test_save_merged_equals_the_reference_writer(capi, po, pr, tmp_path):
    from conftest import SMALL
    poses, stamps, clouds = scene(po, 6, seed=3)
    ctx = capi.Context(num_lines=16, **SMALL)
    for leaf in (0.5, 1e-4):        # 1e-4: leaf too small for the extent -> VoxelGrid passes the cloud through (Q13), still saved
        ours, ref = tmp_path / ("ours%g" % leaf), tmp_path / ("ref%g" % leaf)
        ctx.save_merged(str(ours) + "/", poses, clouds, leaf)
        pr.save_merged(str(ref) + "/", poses, stamps, clouds, leaf, total_order=True)
        assert_same_tree(ours, ref)
        assert sorted(os.listdir(ours)) == ["floam_merged.pcd", "floam_merged_downsampled_leaf_%f.pcd" % leaf]
    ctx.close()
