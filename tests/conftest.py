import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs through the CUDA C-ABI library)")


def _have_gpu():
    if shutil.which("nvidia-smi") is None and not os.path.exists("/dev/nvidiactl"):
        return False
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return os.path.exists("/dev/nvidiactl")


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible in this container (GPU tests run under gpurun / on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


SMALL = dict(max_scan_points=300000, max_map_points=1 << 21, max_global_map_points=1 << 21, max_grid_cells=1 << 22)


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def pr():
    """The reference's own class sources compiled unmodified (oracle/_ref, built where /root/reference exists, shipped otherwise)."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/libfloam_ref.so is not here and /root/reference is absent")
    return pyref


@pytest.fixture(scope="session")
def synth():
    from floam_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def capi():
    from floam_b200 import capi as c
    return c


@pytest.fixture(scope="session")
def sequences(synth):
    """Lazily generated synthetic sequences, cached per (sensor, seed, frames, distort)."""
    cache = {}

    def get(sensor, frames, seed=0, distort=False, sigma=0.02, n_az=None, speed=10.0):
        key = (sensor, frames, seed, distort, sigma, n_az, speed)
        if key not in cache:
            seq = synth.Sequence(sensor, seed=seed, distort=distort, sigma=sigma, n_az=n_az, speed=speed)
            scans, off = seq.scans(0, frames)
            cache[key] = (seq, scans, off)
        return cache[key]
    return get


def xyzi(a):
    return np.stack([a["x"], a["y"], a["z"], a["intensity"]], 1)
