"""floam_b200 — B200-native replacement of dan11003/floam's per-frame odometry hot path.

csrc/     hand-written sm_100a CUDA kernels + the C ABI (include/floam_b200.h)  -> lib/libfloam_b200.so
host/     header-only C++ shims with the reference's class API on top of the C ABI
capi.py   ctypes binding used by tests/ and bench.py
synth/    deterministic synthetic LiDAR / IMU workload generator (the reference ships no data)
"""
__all__ = ["capi", "synth", "build"]
