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
TUNING_LIB = os.path.join(OUT_DIR, "libmnv_b200_tuning.so")   # -DMNV_TUNING: include/mnv_debug.h (tools, alternate-path tests)
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
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]   # incl. glibc_math.h
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
    link = ["-lcudart", "-ccbin", "/usr/bin/g++", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    if force or _stale(LIB, objs):
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + link)
    # the tuning build differs in gemm_conv.cu only (runtime-settable options + the SIMT checker kernel)
    path = os.path.join(CSRC, "gemm_conv.cu")
    tobj = os.path.join(OUT_DIR, "gemm_conv_tuning.o")
    if force or _stale(tobj, [path] + headers + [os.path.join(ROOT, "include", "mnv_debug.h")]):
        subprocess.check_call([NVCC] + ARCH + COMMON + ["-DMNV_TUNING", "-c", path, "-o", tobj])
    tobjs = [tobj if o.endswith("gemm_conv.o") else o for o in objs]
    if force or _stale(TUNING_LIB, tobjs):
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", TUNING_LIB] + tobjs + link)
    build_host(force)
    return LIB


HOST = os.path.join(CSRC, "host", "minerva")
HOST_SRCS = [os.path.join(HOST, "op", "impl", "cuda.cpp"), os.path.join(HOST, "device", "gpu_device.cpp"),
             os.path.join(HOST, "device", "stream_device.cpp")]
HOST_TEST = os.path.join(OUT_DIR, "test_host_plugin")
MNIST_APPS = os.path.join(OUT_DIR, "mnist_apps")      # apps/mnist_mlp + apps/mnist_cnn over ComputeFn::Execute on the StreamDevice


def build_host(force=False):
    """The C++ host side above the C ABI (minerva/op plug-in surface + GpuDevice) -> libminerva_b200_host.so,
    and the C++ plug-in test binary (links the CPU oracle as its checker)."""
    deps = HOST_SRCS + [os.path.join(HOST, "op", "hotpath.h"), os.path.join(HOST, "device", "gpu_device.h"),
                        os.path.join(HOST, "device", "stream_device.h"), os.path.join(HOST, "device", "task.h"), LIB]
    inc = ["-I" + HOST, "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]
    if force or _stale(HOST_LIB, deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-o", HOST_LIB] + inc + HOST_SRCS +
                              ["-L" + OUT_DIR, "-lmnv_b200", "-L/usr/local/cuda/lib64", "-lcudart",
                               "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,/usr/local/cuda/lib64"])
    test_src = os.path.join(ROOT, "tests", "cpp", "test_host_plugin.cpp")
    oracle_c = os.path.join(ROOT, "oracle", "mnv_oracle.c")
    if force or _stale(HOST_TEST, [test_src, oracle_c, HOST_LIB]):
        oracle_o = os.path.join(OUT_DIR, "mnv_oracle_test.o")
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-c", oracle_c, "-o", oracle_o])
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-o", HOST_TEST, test_src, oracle_o] + inc +
                              ["-L" + OUT_DIR, "-lminerva_b200_host", "-lmnv_b200", "-L/usr/local/cuda/lib64", "-lcudart",
                               "-fopenmp", "-lm", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,/usr/local/cuda/lib64"])
    app_src = os.path.join(CSRC, "host", "apps", "mnist_apps.cpp")
    if force or _stale(MNIST_APPS, [app_src, HOST_LIB]):
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-o", MNIST_APPS, app_src] + inc +
                              ["-L" + OUT_DIR, "-lminerva_b200_host", "-lmnv_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-lpthread",
                               "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
