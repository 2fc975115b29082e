"""`_backend` of gridencoder/grid.py served by libnerf_b200.so: the three functions of gridencoder/src/bindings.cpp:5-7 with
their pybind signatures (gridencoder/src/gridencoder.h:12-15) over the C ABI of include/nerf_b200.h.

Usable as a module (`sys.modules['_gridencoder'] = this module`: grid.py does `import _gridencoder as _backend`) or as
`from .backend_b200 import _backend` (the class below)."""
import ctypes as C
import os

import torch

_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libnerf_b200.so")
_lib = C.CDLL(_LIB)
_lib.nb200_error_string.restype = C.c_char_p
_TAG = {torch.float32: 0, torch.float16: 1}
_LAYOUT_LBC = 0          # the reference's [L, B, C] outputs / gradients (grid.py:49, :81)


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ok(rc):
    if rc:
        raise RuntimeError(_lib.nb200_error_string(C.c_int(rc)).decode())


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, Cc, L, max_level, S, H, dy_dx, gridtype, align_corners,
                        interp):
    """void grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, max_level, S, H, dy_dx, gridtype,
    align_corners, interp)   gridencoder.h:12"""
    with torch.cuda.device(inputs.device):
        _ok(_lib.nb200_grid_encode_forward(_p(inputs), _p(embeddings), _p(offsets), _p(outputs), C.c_uint32(B), C.c_uint32(D),
                                           C.c_uint32(Cc), C.c_uint32(L), C.c_uint32(max_level), C.c_float(S), C.c_uint32(H),
                                           _p(dy_dx), C.c_uint32(gridtype), C.c_int(bool(align_corners)), C.c_uint32(interp),
                                           C.c_int(_TAG[embeddings.dtype]), C.c_int(_LAYOUT_LBC), _st()))


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, Cc, L, max_level, S, H, dy_dx, grad_inputs,
                         gridtype, align_corners, interp):
    """void grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, max_level, S, H, dy_dx,
    grad_inputs, gridtype, align_corners, interp)   gridencoder.h:13.  The library accumulates the table gradient in fp32;
    a half grad_embeddings buffer (autocast, grid.py:83) receives the fp32 sums rounded once."""
    g32 = grad_embeddings if grad_embeddings.dtype == torch.float32 else torch.zeros_like(grad_embeddings, dtype=torch.float32)
    with torch.cuda.device(inputs.device):
        _ok(_lib.nb200_grid_encode_backward(_p(grad), _p(inputs), _p(offsets), _p(g32), C.c_uint32(B), C.c_uint32(D),
                                            C.c_uint32(Cc), C.c_uint32(L), C.c_uint32(max_level), C.c_float(S), C.c_uint32(H),
                                            _p(dy_dx), _p(grad_inputs), C.c_uint32(gridtype), C.c_int(bool(align_corners)),
                                            C.c_uint32(interp), C.c_int(_TAG[grad.dtype]), C.c_int(_LAYOUT_LBC), C.c_int(1), _st()))
    if g32 is not grad_embeddings:
        grad_embeddings.add_(g32.to(grad_embeddings.dtype))


def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, Cc, L, S, H, gridtype, align_corners):
    """void grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners)
    gridencoder.h:15"""
    with torch.cuda.device(inputs.device):
        _ok(_lib.nb200_grad_total_variation(_p(inputs), _p(embeddings), _p(grad), _p(offsets), C.c_float(weight), C.c_uint32(B),
                                            C.c_uint32(D), C.c_uint32(Cc), C.c_uint32(L), C.c_float(S), C.c_uint32(H),
                                            C.c_uint32(gridtype), C.c_int(bool(align_corners)), _st()))


class _backend:
    grid_encode_forward = staticmethod(grid_encode_forward)
    grid_encode_backward = staticmethod(grid_encode_backward)
    grad_total_variation = staticmethod(grad_total_variation)
