"""Drop-in for the ray generation of the reference's loader: ``get_rays`` (nerf/provider_utils.py:238-302), same
signature, same result dict; the index sampling (random pixels / error-map importance sampling, :263-284) is the
reference's torch code, the ray arithmetic is one CUDA kernel (csrc/raygen.cu)."""
import ctypes as C

import torch

from .. import _lib as L


def ray_kernel(poses, fx, fy, cx, cy, H, W, inds=None, offset=(0.5, 0.5), rays_o=None, rays_d=None):
    """poses [B,4,4] fp32 CUDA; inds [B,N] int64 CUDA or None (all H*W pixels in order) -> rays_o, rays_d [B,N,3]"""
    L.require_cuda(poses, inds)
    poses = poses.float().contiguous()
    B = poses.shape[0]
    N = H * W if inds is None else inds.shape[-1]
    if inds is not None:
        inds = inds.expand(B, N).contiguous().long()
    if rays_o is None:
        rays_o = torch.empty(B, N, 3, dtype=torch.float32, device=poses.device)
    if rays_d is None:
        rays_d = torch.empty(B, N, 3, dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        L.check(L.lib().nb200_get_rays(L.ptr(poses), L.f32(fx), L.f32(fy), L.f32(cx), L.f32(cy), L.u32(H), L.u32(W), L.u32(B),
                                       L.u32(N), L.ptr(inds), L.f32(offset[0]), L.f32(offset[1]), L.ptr(rays_o), L.ptr(rays_d),
                                       L.stream()), "get_rays")
    return rays_o, rays_d


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, offset=(0.5, 0.5)):
    """poses [B,4,4] cam2world, intrinsics (fx, fy, cx, cy) -> {'rays_o', 'rays_d' [B,N,3], 'inds' [B,N] when N > 0,
    'inds_coarse' with an error map}"""
    device = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    results = {}
    inds = None
    if N > 0:
        N = min(N, H * W)
        if error_map is None:
            inds = torch.randint(0, H * W, size=[N], device=device).expand([B, N])          # may duplicate (:266)
        else:
            inds_coarse = torch.multinomial(error_map.to(device), N, replacement=False)     # [B, N] in [0, 128*128)
            inds_x, inds_y = inds_coarse // 128, inds_coarse % 128
            sx, sy = H / 128, W / 128
            inds_x = (inds_x * sx + torch.rand(B, N, device=device) * sx).long().clamp(max=H - 1)
            inds_y = (inds_y * sy + torch.rand(B, N, device=device) * sy).long().clamp(max=W - 1)
            inds = inds_x * W + inds_y
            results['inds_coarse'] = inds_coarse
        results['inds'] = inds
    results['rays_o'], results['rays_d'] = ray_kernel(poses, fx, fy, cx, cy, H, W, inds, offset)
    return results
