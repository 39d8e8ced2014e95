"""Builds the native pieces in-tree with nvcc / g++ (sm_100a only). Called by __graft_entry__.build()."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblira_b200.so")
BIN = os.path.join(ROOT, "bin")

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]


def _nvcc():
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: liblira_b200 cannot be built")
    return nv


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources():
    out = [os.path.join(ROOT, "include", "lira_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".cpp", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def build_library(force=False, verbose=False):
    """nvcc -> lira-ann-search_b200/liblira_b200.so (static cudart, no libcuda link dependency)."""
    srcs = _sources()
    if not force and not _newer(LIB, srcs):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB, os.path.join(CSRC, "lira_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB


def build_binaries(force=False):
    """compute_knn and search command-line programs (reference argv) -> bin/."""
    os.makedirs(BIN, exist_ok=True)
    built = []
    for name in ("compute_knn", "search"):
        src = os.path.join(CSRC, name + "_main.cpp")
        if not os.path.exists(src):
            continue
        out = os.path.join(BIN, name)
        if force or _newer(out, [src, LIB] + _sources()):
            subprocess.check_call([_nvcc(), "-std=c++17", "-O2", "-x", "cu", "-gencode",
                                   "arch=compute_100a,code=sm_100a", "-o", out, src,
                                   "-I" + os.path.join(ROOT, "include"), "-L" + HERE, "-llira_b200",
                                   "-Xlinker", "-rpath," + HERE, "-Xlinker", "-rpath,$ORIGIN/../lira-ann-search_b200"])
        built.append(out)
    return built
