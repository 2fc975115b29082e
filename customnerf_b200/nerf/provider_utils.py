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


def _sample_pixels(B, H, W, N, error_map, device):
    """Pixel indices [B, N] of a training batch, and the coarse cells they came from (None without an error map).
    Same draws, in the same order, as the reference (:263-284): uniform with replacement -- one index set shared by the
    batch -- or, with a 128 x 128 error map, cells drawn without replacement by error and a uniform jitter inside each."""
    if error_map is None:
        return torch.randint(0, H * W, size=[N], device=device).expand([B, N]), None
    cells = torch.multinomial(error_map.to(device), N, replacement=False)
    cell_h, cell_w = H / 128, W / 128
    rows = ((cells // 128) * cell_h + torch.rand(B, N, device=device) * cell_h).long().clamp(max=H - 1)
    cols = ((cells % 128) * cell_w + torch.rand(B, N, device=device) * cell_w).long().clamp(max=W - 1)
    return rows * W + cols, cells


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, offset=(0.5, 0.5)):
    """poses [B,4,4] cam2world, intrinsics (fx, fy, cx, cy) -> {'rays_o', 'rays_d' [B,N,3], 'inds' [B,N] when N > 0,
    'inds_coarse' with an error map}"""
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    results, inds = {}, None
    if N > 0:
        inds, cells = _sample_pixels(poses.shape[0], H, W, min(N, H * W), error_map, poses.device)
        if cells is not None:
            results['inds_coarse'] = cells      # the trainer updates the error map through these
        results['inds'] = inds
    results['rays_o'], results['rays_d'] = ray_kernel(poses, fx, fy, cx, cy, H, W, inds, offset)
    return results
