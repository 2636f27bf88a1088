"""Builds hope_b200/libhope_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhope_b200.so")
# second build of the same sources with 128 obstacle rings per scene (Dragon Lake Parking scenes, scope row f3)
VARIANTS = {16: LIB, 128: os.path.join(HERE, "libhope_b200_obs128.so")}
SOURCES = ["hope_kernels.cu", "scene_gen.cu", "policy_glue.cu", "policy_forward.cu", "img_encoder.cu", "host_wire.cpp"]
DEPS = SOURCES + ["host_wire.h", "hope_device.cuh", "render.cuh", "rs_words.cuh", "rs_walk.cuh", "rs_check.cuh", "rs_check_pooled.cuh", "div_pair.cuh", "hope_types.cuh", "observe.cuh", "observe_body.inc", "advance.cuh", "advance_body.inc", "rs_enumerate.cuh", "rs_select_body.inc", "scene_gen.h", os.path.join("..", "..", "include", "hope_b200.h")]
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


def needs_build(max_obs=16):
    lib = VARIANTS[max_obs]
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def _command(max_obs, verbose):
    extra = os.environ.get("HOPE_B200_NVCC_DEFS", "").split()  # tuning experiments, e.g. "-DHOPE_CHK_MINBLOCKS=5"
    if max_obs != 16:
        extra = extra + [f"-DHOPE_MAX_OBS={max_obs}"]
    return [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", VARIANTS[max_obs]] + \
        [os.path.join(CSRC, s) for s in SOURCES]


def build(force=False, verbose=False, variants=(16,)):
    """Build the requested capacity variants (concurrently); returns the path of the first."""
    todo = [v for v in variants if force or needs_build(v)]
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)  # this image exports a gcc wrapper that lacks libgomp specs
    procs = [(v, subprocess.Popen(_command(v, verbose), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)) for v in todo]
    for v, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed building {os.path.basename(VARIANTS[v])}")
    return VARIANTS[variants[0]]


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variants=(16, 128) if "--all" in sys.argv else (16,)))
