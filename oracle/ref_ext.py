"""Loader for oracle/_ref: the unmodified reference CUDA extensions.  TEST INFRASTRUCTURE ONLY.

Importable only where a GPU is present (the modules link against libtorch_cuda and every entry
point launches kernels).  Returns None when the prebuilt .so files are missing.
"""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def _load(name):
    if name in _cache:
        return _cache[name]
    path = os.path.join(_HERE, "_ref", name + ".so")
    mod = None
    if os.path.isfile(path):
        import torch  # noqa: F401  (libtorch must be loaded first)
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def gridencoder():
    """pybind module exposing grid_encode_forward / grid_encode_backward / grad_total_variation
    (gridencoder/src/bindings.cpp:5-7)."""
    return _load("_gridencoder_ref")


def raymarching():
    """pybind module exposing the 12 raymarching entry points (raymarching/src/bindings.cpp:5-20)."""
    return _load("_raymarching_ref")


def available():
    return gridencoder() is not None and raymarching() is not None
