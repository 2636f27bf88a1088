"""Builds hope_b200/libhope_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhope_b200.so")
SOURCES = ["hope_kernels.cu", "scene_gen.cu"]
DEPS = SOURCES + ["hope_device.cuh", "scene_gen.h", os.path.join("..", "..", "include", "hope_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # the reference rounds every product and sum separately (numpy / CPython float64)
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-pthread", "-shared",
]


def nvcc_path():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; hope_b200 needs the CUDA toolkit to build its kernels")
    return cand


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("HOPE_B200_NVCC_DEFS", "").split()  # tuning experiments, e.g. "-DHOPE_CHK_MINBLOCKS=5"
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)  # this image exports a gcc wrapper that lacks libgomp specs
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building libhope_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
