"""Build recipe for ``oracle/_ref``: the UNMODIFIED reference CUDA extensions.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this.

The two reference extensions (``/root/reference/gridencoder/src`` and
``/root/reference/raymarching/src``) are compiled *from the sources where they
lie* -- no source is copied into this repository -- and only the resulting
shared objects land in ``oracle/_ref/`` (git-ignored, but shipped to the GPU box
by ``gpurun``).  The only deviation from the reference's own build
(``gridencoder/backend.py:6-9``, ``raymarching/backend.py:6-9``) is the flag
patch SURVEY.md section 8(c) documents: ``-std=c++14`` no longer compiles
against torch >= 2.1 headers, so ``-std=c++17`` plus an explicit sm_100a
``-gencode`` is used.

On the GPU box the modules are imported by ``oracle/ref_ext.py`` to
  (a) pin the CPU restatement (``oracle/nerf_oracle.c``) against the real
      reference kernels, and
  (b) mint the golden vectors committed under ``tests/golden/``.

Usage:  python oracle/build_ref.py            (about 7 minutes of nvcc, once)
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("NERF_REFERENCE_ROOT", "/root/reference")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
    "-U__CUDA_NO_HALF2_OPERATORS__",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
]

TARGETS = {
    # module name -> (source dir, [sources])
    "_gridencoder_ref": ("gridencoder/src", ["gridencoder.cu", "bindings.cpp"]),
    "_raymarching_ref": ("raymarching/src", ["raymarching.cu", "bindings.cpp"]),
}


def _torch_paths():
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths("cuda")
    lib = ce.library_paths("cuda")
    return torch, inc, lib


def available():
    return os.path.isdir(REF)


def built(name):
    return os.path.isfile(os.path.join(OUT, name + ".so"))


def build_one(name, verbose=True):
    torch, inc, lib = _torch_paths()
    srcdir, files = TARGETS[name]
    os.makedirs(OUT, exist_ok=True)
    objdir = os.path.join(OUT, "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    pyinc = sysconfig.get_paths()["include"]
    common = ["-DTORCH_EXTENSION_NAME=" + name, "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    incs = []
    for p in inc + [pyinc]:
        incs += ["-isystem", p]
    objs = []
    for f in files:
        src = os.path.join(REF, srcdir, f)
        obj = os.path.join(objdir, f + ".o")
        objs.append(obj)
        if f.endswith(".cu"):
            cmd = ["nvcc", "-c", src, "-o", obj] + NVCC_FLAGS + common + incs
        else:
            cmd = ["g++", "-c", src, "-o", obj, "-O3", "-std=c++17", "-fPIC"] + common + incs
        if verbose:
            print("[build_ref]", " ".join(cmd[:6]), "...", flush=True)
        subprocess.check_call(cmd)
    so = os.path.join(OUT, name + ".so")
    link = ["g++", "-shared", "-o", so] + objs
    for p in lib:
        link += ["-L" + p, "-Wl,-rpath," + p]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    subprocess.check_call(link)
    shutil.rmtree(objdir, ignore_errors=True)
    return so


# The reference's own Python wrappers around its two extensions.  They are staged -- byte for byte, next to the compiled
# extensions, in the same git-ignored build directory -- so that the GPU box (which has no /root/reference) can run the
# reference's UNMODIFIED grid.py / raymarching.py over this repo's library (tests/test_gpu_reference_wrappers.py: the
# drop-in boundary proof).  Like the .so files they are build products of this recipe: never committed, never imported by
# the product.
WRAPPERS = {"gridencoder/grid.py": "pysrc/ref_gridencoder/grid.py", "raymarching/raymarching.py": "pysrc/ref_raymarching/raymarching.py"}


def stage_wrappers(verbose=True):
    for src, dst in WRAPPERS.items():
        a, b = os.path.join(REF, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        shutil.copyfile(a, b)
        init = os.path.join(os.path.dirname(b), "__init__.py")
        if not os.path.isfile(init):
            open(init, "w").close()
    if verbose:
        print("[build_ref] staged the reference's wrappers under", os.path.join(OUT, "pysrc"))


def build_all(force=False, verbose=True):
    if not available():
        if verbose:
            print("[build_ref] %s not present; keeping prebuilt oracle/_ref as is" % REF)
        return False
    for name in TARGETS:
        if force or not built(name):
            build_one(name, verbose=verbose)
    stage_wrappers(verbose=verbose)
    return True


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("[build_ref] done:", sorted(os.listdir(OUT)) if os.path.isdir(OUT) else None)
