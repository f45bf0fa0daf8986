"""In-tree native builds: the CUDA/C-ABI product library (sm_100a) and the synthetic-workload generator.

nvcc cross-compiles without a GPU; the resulting .so files are git-ignored but travel to the GPU box.
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
CUDA_LIB = os.path.join(LIB_DIR, "libfloam_b200.so")
SYNTH_LIB = os.path.join(_HERE, "synth", "libfloam_synth.so")   # the workload generator is not product code: kept out of lib/

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
              # no FMA contraction: float/double arithmetic must round like the reference's x86 build
              "-fmad=false"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_synth(force=False):
    src = os.path.join(_HERE, "synth", "synth.cpp")
    os.makedirs(LIB_DIR, exist_ok=True)
    if force or _stale(SYNTH_LIB, [src]):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", SYNTH_LIB, src])
    return SYNTH_LIB


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force=False, verbose=False):
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
        [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    os.makedirs(LIB_DIR, exist_ok=True)
    if force or _stale(CUDA_LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
            ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-shared", "-o", CUDA_LIB] + srcs
        subprocess.check_call(cmd)
    return CUDA_LIB


def build_host_shim_check():
    """Compile the C++ host shim (reference class API over the C-ABI) against the mini_pcl stand-in: syntax/link check."""
    src = os.path.join(_HERE, "host", "shim_selftest.cpp")
    if not os.path.exists(src):
        return None
    out = os.path.join(LIB_DIR, "shim_selftest")
    if _stale(out, [src] + [os.path.join(_HERE, "host", f) for f in os.listdir(os.path.join(_HERE, "host"))] + [CUDA_LIB]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(_HERE, "host"),
                               "-o", out, src, "-L", LIB_DIR, "-lfloam_b200", "-Wl,-rpath,$ORIGIN"])
    return out


def build_all(force=False):
    build_synth(force)
    build_cuda(force)
    build_host_shim_check()


if __name__ == "__main__":
    import sys
    build_all(force="-f" in sys.argv)
