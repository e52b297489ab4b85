"""Build the CUDA kernel library (C ABI of include/mnv.h) in-tree for sm_100a.

    python -m minerva_b200.build [--force]

Produces minerva_b200/lib/libmnv_b200.so with plain nvcc (no torch extension machinery: the
library has no torch types in its ABI).  nvcc cross-compiles without a GPU, so this runs in the
authoring container; the .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libmnv_b200.so")
HOST_LIB = os.path.join(OUT_DIR, "libminerva_b200_host.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
          "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
# every memory-bound kernel is compiled without FMA contraction: bit-exact against the CPU oracle
SOURCES = {
    "elementwise.cu": ["--fmad=false"],
    "matrix_ops.cu": ["--fmad=false"],
    "nn_ops.cu": ["--fmad=false"],
    "gemm_conv.cu": [],
}


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "mnv.h"))
    objs = []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ccbin", "/usr/bin/g++", "-Xlinker", "-rpath=/usr/local/cuda/lib64"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
