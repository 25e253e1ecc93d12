"""In-tree build of the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libevab200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


HOST = os.path.join(CSRC, "host")
HDR = os.path.join(HERE, "..", "include", "evab200.h")


def _ext_path():
    import sysconfig
    return os.path.join(HERE, "_eva_b200" + sysconfig.get_config_var("EXT_SUFFIX"))


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """libevab200.so (CUDA kernels + C-ABI, nvcc sm_100a) and the pybind11 host module."""
    os.makedirs(LIBDIR, exist_ok=True)
    cuda_src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if os.path.isfile(os.path.join(CSRC, f))] + [HDR]
    if force or _newer(LIB, cuda_src):
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, "evab200.cu"), "-o", LIB]
        subprocess.check_call(cmd)
    ext = _ext_path()
    host_src = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST))] + [HDR]
    if force or _newer(ext, host_src + [LIB]):
        import pybind11
        import sysconfig
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fvisibility=hidden",
               "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(os.path.dirname(NVCC), "..", "include"),
               os.path.join(HOST, "pymodule.cpp"), "-o", ext, "-L" + LIBDIR, "-levab200", "-ldl", "-Wl,-rpath,$ORIGIN/lib"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
