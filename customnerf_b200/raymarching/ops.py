"""Drop-in for the reference's ``raymarching`` package on top of libnerf_b200.so.

Public names, argument order, defaults and return shapes are those of ``raymarching/raymarching.py`` of the
reference (line numbers below refer to it).  Ops without a backward are plain functions here (the reference
wraps every one in an ``autograd.Function`` with only a forward); ``composite_rays_train`` keeps its custom
backward.  All ops run on the CURRENT stream of the tensors' device -- the reference launches on the legacy
default stream.

What is different underneath:
  * ``march_rays_train`` traverses every ray ONCE (count kernel that also records each sample's t, + scan), reads the
    total back (the same single D2H sync the reference has at :225), allocates ``m`` rounded up by the reference's
    rule and expands the records with one warp per ray -- instead of zero-filling N*max_steps rows (:196-209),
    marching every ray twice and slicing.  ``rays`` comes out in ray-id order with scan
    offsets (deterministic; a valid member of the reference's atomics-ordered output set).
  * no ``torch.cuda.empty_cache()`` on the hot path (:232).
"""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from .. import _lib as L

__all__ = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
           "composite_rays_train", "composite_rays_train_sdf", "march_rays", "composite_rays"]


def _f32c(t, shape=None):
    """float32, on the GPU, contiguous (custom_fwd(cast_inputs=float32) + .cuda() + .contiguous() of the reference)"""
    if not t.is_cuda:
        t = t.cuda()
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    return t if shape is None else t.view(*shape)


_scratch = {}


def _march_scratch(N, device):
    need = int(L.lib().nb200_march_scratch_ints(L.u32(N)))
    key = (device.type, device.index)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(max(need, 1024), dtype=torch.int32, device=device)
        _scratch[key] = buf
    return buf


# ----------------------------------------------------------------------------------------------- utils
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """rays_o, rays_d [N,3]; aabb [6] -> nears [N], fars [N]   (:20-50)"""
    rays_o = _f32c(rays_o, (-1, 3))
    rays_d = _f32c(rays_d, (-1, 3))
    aabb = _f32c(aabb)
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    fars = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    with torch.cuda.device(rays_o.device):
        L.check(L.lib().nb200_near_far_from_aabb(L.ptr(rays_o), L.ptr(rays_d), L.ptr(aabb), L.u32(N), L.f32(min_near),
                                                 L.ptr(nears), L.ptr(fars), L.stream()), "near_far_from_aabb")
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    """-> coords [N,2] in [-1,1]   (:53-81)"""
    rays_o = _f32c(rays_o, (-1, 3))
    rays_d = _f32c(rays_d, (-1, 3))
    N = rays_o.shape[0]
    coords = torch.empty(N, 2, dtype=torch.float32, device=rays_o.device)
    with torch.cuda.device(rays_o.device):
        L.check(L.lib().nb200_sph_from_ray(L.ptr(rays_o), L.ptr(rays_d), L.f32(radius), L.u32(N), L.ptr(coords),
                                           L.stream()), "sph_from_ray")
    return coords


def morton3D(coords):
    """coords int32 [N,3] in [0,128) -> indices int32 [N]   (:84-105)"""
    if not coords.is_cuda:
        coords = coords.cuda()
    coords = coords.int().contiguous()
    N = coords.shape[0]
    indices = torch.empty(N, dtype=torch.int32, device=coords.device)
    with torch.cuda.device(coords.device):
        L.check(L.lib().nb200_morton3D(L.ptr(coords), L.u32(N), L.ptr(indices), L.stream()), "morton3D")
    return indices


def morton3D_invert(indices):
    """indices int32 [N] -> coords int32 [N,3]   (:107-127)"""
    if not indices.is_cuda:
        indices = indices.cuda()
    indices = indices.int().contiguous()
    N = indices.shape[0]
    coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
    with torch.cuda.device(indices.device):
        L.check(L.lib().nb200_morton3D_invert(L.ptr(indices), L.u32(N), L.ptr(coords), L.stream()), "morton3D_invert")
    return coords


def packbits(grid, thresh, bitfield=None):
    """grid float [C, H^3] -> bitfield uint8 [C*H^3/8]; bit i of byte n <-> cell 8n+i, strict '>'   (:130-156)"""
    grid = _f32c(grid)
    N = grid.shape[0] * grid.shape[1] // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    with torch.cuda.device(grid.device):
        L.check(L.lib().nb200_packbits(L.ptr(grid), L.u32(N), L.f32(thresh), L.ptr(bitfield), L.stream()), "packbits")
    return bitfield


# ----------------------------------------------------------------------------------------------- training
def _align_up(m, align):
    # the reference adds a full `align` when m is already aligned (:203,227,389); reproduced for shape parity
    return m + (align - m % align) if align > 0 else m


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                     perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024, noises=None):
    """-> xyzs [M,3], dirs [M,3], deltas [M,2], rays int32 [N,3] = (ray id, offset, count)   (:162-236)

    ``noises`` (optional, [N] float) replaces the wrapper-generated ``torch.rand`` when ``perturb`` is set."""
    rays_o = _f32c(rays_o, (-1, 3))
    rays_d = _f32c(rays_d, (-1, 3))
    if not density_bitfield.is_cuda:
        density_bitfield = density_bitfield.cuda()
    density_bitfield = density_bitfield.contiguous()
    nears, fars = _f32c(nears), _f32c(fars)
    dev = rays_o.device
    N = rays_o.shape[0]
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
    if perturb:
        noises = torch.rand(N, dtype=torch.float32, device=dev) if noises is None else _f32c(noises)
    else:
        noises = None                      # NULL == all zeros
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    scratch = _march_scratch(N, dev)
    lib = L.lib()
    with torch.cuda.device(dev):
        L.check(lib.nb200_march_rays_train_count(
            L.ptr(rays_o), L.ptr(rays_d), L.ptr(density_bitfield), L.f32(bound), L.f32(dt_gamma), L.u32(max_steps),
            L.u32(N), L.u32(C), L.u32(H), L.ptr(nears), L.ptr(fars), L.ptr(noises), L.ptr(rays), L.ptr(step_counter),
            L.ptr(scratch), L.stream()), "march_rays_train(count)")
        if not force_all_rays and mean_count > 0:
            M = _align_up(mean_count, align)           # fixed budget; rays that do not fit are dropped (:201-204)
            zero_tail_from = None
        else:
            m = int(step_counter[0].item())            # the one D2H sync, as in the reference (:225)
            M = _align_up(m, align)
            zero_tail_from = m
        xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.empty(M, 2, dtype=torch.float32, device=dev)
        if zero_tail_from is None:
            xyzs.zero_(); dirs.zero_(); deltas.zero_()
        else:                                          # rows [m, M) are the zero padding the MLP still evaluates
            xyzs[zero_tail_from:].zero_(); dirs[zero_tail_from:].zero_(); deltas[zero_tail_from:].zero_()
        L.check(lib.nb200_march_rays_train_write(
            L.ptr(rays_o), L.ptr(rays_d), L.ptr(density_bitfield), L.f32(bound), L.f32(dt_gamma), L.u32(max_steps),
            L.u32(N), L.u32(C), L.u32(H), L.u32(M), L.ptr(nears), L.ptr(fars), L.ptr(noises), L.ptr(rays), L.ptr(xyzs),
            L.ptr(dirs), L.ptr(deltas), L.ptr(scratch), L.stream()), "march_rays_train(write)")
    return xyzs, dirs, deltas, rays


class _CompositeTrain(Function):
    """(:239-292); grad_depth is dropped exactly as the reference does (:276)."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        deltas = deltas.contiguous()
        rays = rays.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().nb200_composite_rays_train_forward(
                L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(rays), L.u32(M), L.u32(N), L.f32(T_thresh),
                L.ptr(weights_sum), L.ptr(depth), L.ptr(image), L.stream()), "composite_rays_train_forward")
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image)
        ctx.dims = (M, N, T_thresh)
        return weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        sigmas, rgbs, deltas, rays, weights_sum, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_weights_sum = grad_weights_sum.contiguous()
        grad_image = grad_image.contiguous()
        # rows outside every ray segment (alignment padding, dropped rays) must read as zero
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        with torch.cuda.device(sigmas.device):
            L.check(L.lib().nb200_composite_rays_train_backward(
                L.ptr(grad_weights_sum), L.ptr(grad_image), L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(rays),
                L.ptr(weights_sum), L.ptr(image), L.u32(M), L.u32(N), L.f32(T_thresh), L.ptr(grad_sigmas),
                L.ptr(grad_rgbs), L.stream()), "composite_rays_train_backward")
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _CompositeTrain.apply
# byte-identical math in the reference (raymarching.cu:579-657 vs :500-577): sdf is composited as sigma
composite_rays_train_sdf = _CompositeTrain.apply


# ----------------------------------------------------------------------------------------------- inference
def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
               align=-1, perturb=False, dt_gamma=0, max_steps=1024, noises=None):
    """-> xyzs [M,3], dirs [M,3], deltas [M,2] with M = n_alive*n_step padded by the reference's rule   (:355-405)"""
    rays_o = _f32c(rays_o, (-1, 3))
    rays_d = _f32c(rays_d, (-1, 3))
    dev = rays_o.device
    M = _align_up(n_alive * n_step, align)
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    if perturb:
        noises = torch.rand(n_alive, dtype=torch.float32, device=dev) if noises is None else _f32c(noises)
    else:
        noises = None
    with torch.cuda.device(dev):
        L.check(L.lib().nb200_march_rays(
            L.u32(n_alive), L.u32(n_step), L.ptr(rays_alive), L.ptr(rays_t), L.ptr(rays_o), L.ptr(rays_d), L.f32(bound),
            L.f32(dt_gamma), L.u32(max_steps), L.u32(C), L.u32(H), L.ptr(density_bitfield), L.ptr(near), L.ptr(far),
            L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.ptr(noises), L.stream()), "march_rays")
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """In-place accumulation into weights_sum / depth / image; marks finished rays with -1   (:408-427)"""
    sigmas = _f32c(sigmas)
    rgbs = _f32c(rgbs)
    deltas = _f32c(deltas)
    with torch.cuda.device(sigmas.device):
        L.check(L.lib().nb200_composite_rays(
            L.u32(n_alive), L.u32(n_step), L.f32(T_thresh), L.ptr(rays_alive), L.ptr(rays_t), L.ptr(sigmas), L.ptr(rgbs),
            L.ptr(deltas), L.ptr(weights_sum), L.ptr(depth), L.ptr(image), L.stream()), "composite_rays")
    return tuple()
