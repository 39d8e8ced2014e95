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


def _gxx():
    # the image exports CC/CXX wrappers that lack libgomp.spec; use the distro g++ like oracle/Makefile
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def build_binaries(force=False):
    """compute_knn and search command-line programs (reference argv) -> bin/.
    compute_knn is plain C++ over the C ABI; search additionally links LibTorch (from the installed torch wheel),
    used only to read the TorchScript checkpoint written by the reference's index.py."""
    os.makedirs(BIN, exist_ok=True)
    built = []
    inc = "-I" + os.path.join(ROOT, "include")
    link = ["-L" + HERE, "-llira_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../lira-ann-search_b200"]
    src = os.path.join(CSRC, "compute_knn_main.cpp")
    out = os.path.join(BIN, "compute_knn")
    if force or _newer(out, [src, LIB]):
        subprocess.check_call([_gxx(), "-std=c++17", "-O2", "-o", out, src, inc] + link)
    built.append(out)
    src = os.path.join(CSRC, "search_main.cpp")
    out = os.path.join(BIN, "search")
    if force or _newer(out, [src, LIB]):
        import torch
        tdir = os.path.dirname(torch.__file__)
        subprocess.check_call([_gxx(), "-std=gnu++17", "-O2", "-o", out, src, inc,
                               "-I" + os.path.join(tdir, "include"), "-I" + os.path.join(tdir, "include/torch/csrc/api/include"),
                               "-L" + os.path.join(tdir, "lib"), "-Wl,-rpath," + os.path.join(tdir, "lib"),
                               "-lc10", "-ltorch", "-ltorch_cpu"] + link)
    built.append(out)
    return built
