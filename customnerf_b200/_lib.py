"""ctypes binding of libnerf_b200.so (the C ABI declared in include/nerf_b200.h).

There is NO fallback: if the library has not been built, importing any op of this package raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnerf_b200.so")

F32, F16, F32_AS_F16 = 0, 1, 2
LAYOUT_LBC, LAYOUT_BLC = 0, 1

_lib = None


class NativeLibraryMissing(ImportError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise NativeLibraryMissing(
                "customnerf_b200: %s is missing. Build it with `python -m customnerf_b200.build` "
                "(nvcc, sm_100a). There is no CPU / PyTorch fallback for these ops." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.nb200_error_string.restype = C.c_char_p
        _lib.nb200_march_scratch_ints.restype = C.c_uint32
    return _lib


# number of this library's kernel launches so far (bench.py reports it as gpu_launches)
LAUNCHES = 0
_LAUNCHES_PER_CALL = {"march_rays_train(count)": 3, "count": 3}


def check(rc, what):
    global LAUNCHES
    LAUNCHES += _LAUNCHES_PER_CALL.get(what, 1)
    if rc != 0:
        msg = lib().nb200_error_string(C.c_int(rc)).decode()
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg, rc))


def ptr(t):
    """device pointer of a tensor (or NULL)"""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_tag(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise RuntimeError("customnerf_b200: unsupported dtype %s (float32 / float16 only)" % t.dtype)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("customnerf_b200: expected a CUDA tensor (no CPU fallback exists)")


u32, f32, i32, u64 = C.c_uint32, C.c_float, C.c_int, C.c_uint64
