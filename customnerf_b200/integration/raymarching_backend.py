"""`_backend` of raymarching/raymarching.py served by libnerf_b200.so: the twelve functions of raymarching/src/bindings.cpp:5-20
with their pybind signatures (raymarching/src/raymarching.h:7-22) over the C ABI of include/nerf_b200.h.

Usable as a module (`sys.modules['_raymarching'] = this module`) or as `from .backend_b200 import _backend`."""
import ctypes as C
import os

import torch

_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libnerf_b200.so")
_lib = C.CDLL(_LIB)
_lib.nb200_error_string.restype = C.c_char_p
_lib.nb200_march_scratch_ints.restype = C.c_uint32
_scratch = {}


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ok(rc):
    if rc:
        raise RuntimeError(_lib.nb200_error_string(C.c_int(rc)).decode())


def _u(v):
    return C.c_uint32(int(v))


def _f(v):
    return C.c_float(float(v))


def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
    with torch.cuda.device(rays_o.device):
        _ok(_lib.nb200_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), _u(N), _f(min_near), _p(nears), _p(fars), _st()))


def sph_from_ray(rays_o, rays_d, radius, N, coords):
    with torch.cuda.device(rays_o.device):
        _ok(_lib.nb200_sph_from_ray(_p(rays_o), _p(rays_d), _f(radius), _u(N), _p(coords), _st()))


def morton3D(coords, N, indices):
    with torch.cuda.device(coords.device):
        _ok(_lib.nb200_morton3D(_p(coords.contiguous()), _u(N), _p(indices), _st()))


def morton3D_invert(indices, N, coords):
    with torch.cuda.device(indices.device):
        _ok(_lib.nb200_morton3D_invert(_p(indices.contiguous()), _u(N), _p(coords), _st()))


def packbits(grid, N, density_thresh, bitfield):
    with torch.cuda.device(grid.device):
        _ok(_lib.nb200_packbits(_p(grid), _u(N), _f(density_thresh), _p(bitfield), _st()))


def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, Cc, H, M, nears, fars, xyzs, dirs, deltas, rays, counter,
                     noises):
    """the library marches every ray once and needs a scratch buffer for the recorded samples (kept per device, grown on
    demand); offsets come out in ray-id order -- one member of the reference's atomicAdd-ordered output set"""
    dev = rays_o.device
    need = int(_lib.nb200_march_scratch_ints(_u(N)))
    buf = _scratch.get(dev)
    if buf is None or buf.numel() < need:
        buf = _scratch[dev] = torch.empty(need, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _ok(_lib.nb200_march_rays_train(_p(rays_o), _p(rays_d), _p(grid), _f(bound), _f(dt_gamma), _u(max_steps), _u(N), _u(Cc),
                                        _u(H), _u(M), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(rays), _p(counter),
                                        _p(noises), _p(buf), _st()))


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image):
    with torch.cuda.device(sigmas.device):
        _ok(_lib.nb200_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), _u(M), _u(N), _f(T_thresh),
                                                    _p(weights_sum), _p(depth), _p(image), _st()))


def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh,
                                  grad_sigmas, grad_rgbs):
    with torch.cuda.device(sigmas.device):
        _ok(_lib.nb200_composite_rays_train_backward(_p(grad_weights_sum), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas), _p(rays),
                                                     _p(weights_sum), _p(image), _u(M), _u(N), _f(T_thresh), _p(grad_sigmas),
                                                     _p(grad_rgbs), _st()))


# byte-identical math in the reference (raymarching.cu:579-657 vs :500-577, :776-857 vs :691-772): the same two entries
composite_rays_train_forward_sdf = composite_rays_train_forward
composite_rays_train_backward_sdf = composite_rays_train_backward


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, Cc, H, grid, nears, fars, xyzs,
               dirs, deltas, noises):
    with torch.cuda.device(rays_o.device):
        _ok(_lib.nb200_march_rays(_u(n_alive), _u(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), _f(bound),
                                  _f(dt_gamma), _u(max_steps), _u(Cc), _u(H), _p(grid), _p(nears), _p(fars), _p(xyzs), _p(dirs),
                                  _p(deltas), _p(noises), _st()))


def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    with torch.cuda.device(sigmas.device):
        _ok(_lib.nb200_composite_rays(_u(n_alive), _u(n_step), _f(T_thresh), _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs),
                                      _p(deltas), _p(weights_sum), _p(depth), _p(image), _st()))


class _backend:
    pass


for _name in ("near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
              "composite_rays_train_forward", "composite_rays_train_backward", "composite_rays_train_forward_sdf",
              "composite_rays_train_backward_sdf", "march_rays", "composite_rays"):
    setattr(_backend, _name, staticmethod(globals()[_name]))
