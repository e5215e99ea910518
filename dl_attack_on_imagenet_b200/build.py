"""Builds libadil_b200.so (hand-written sm_100a CUDA kernels + C ABI) in-tree with nvcc.

The shared library travels with the repository snapshot to the GPU box; nothing is JIT-compiled at run time.
"""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
# ADIL_B200_LIB: load an alternative build (e.g. the -DADIL_TIMING debug build of scripts/build_timing.sh)
LIB_PATH = os.environ.get("ADIL_B200_LIB") or os.path.join(PKG_DIR, "libadil_b200.so")
SOURCES = ["adil_api.cu", "adil_fma.cu", "adil_steps.cu", "adil_tc.cu"]
HEADERS = [os.path.join(CSRC, "adil_common.cuh"), os.path.join(ROOT, "include", "adil_b200.h")]

# -cudart shared: the library binds to the CUDA runtime the process already has (PyTorch's) instead of carrying a
# private static copy of it -- one runtime instance per process, and a smaller .so
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-cudart", "shared"]


def _nvcc():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found; cannot build libadil_b200.so")
    return path


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into one shared library. Returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", LIB_PATH]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
