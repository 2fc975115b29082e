"""Build libnerf_b200.so: nvcc, sm_100a only, in-tree (the .so travels to the GPU box with the repo snapshot)."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libnerf_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src):
    obj = os.path.join(OBJ, src + ".o")
    hdr_m = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    hdr_m = max(hdr_m, os.path.getmtime(os.path.join(HERE, "..", "include", "nerf_b200.h")))
    src_m = max(os.path.getmtime(os.path.join(CSRC, src)), hdr_m)
    if os.path.isfile(obj) and os.path.getmtime(obj) >= src_m:
        return obj
    cmd = ["nvcc", "-c", os.path.join(CSRC, src), "-o", obj] + NVCC_FLAGS
    subprocess.check_call(cmd)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        objs = list(ex.map(_compile, sources()))
    cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    if verbose:
        print("[build] wrote", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
