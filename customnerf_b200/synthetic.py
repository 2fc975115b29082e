"""Synthetic scene + rays used by bench.py, smoke() and the tests.

Pure torch, device-agnostic, deterministic: an analytic "bear-shaped" density (union of ellipsoidal blobs inside
[-1,1]^3 of a bound-2 box), its 128^3 x cascade occupancy grid in the reference's [cascade][morton] order, target
colours, and pin-hole camera rays on a sphere looking at the origin.
"""
import math

import torch

_BLOBS = [  # centre, radii
    ((0.0, -0.15, 0.0), (0.46, 0.40, 0.38)),     # body
    ((0.0, 0.42, 0.05), (0.27, 0.25, 0.25)),     # head
    ((-0.19, 0.66, 0.05), (0.09, 0.09, 0.07)),   # ears
    ((0.19, 0.66, 0.05), (0.09, 0.09, 0.07)),
    ((-0.32, -0.56, 0.08), (0.15, 0.18, 0.17)),  # legs
    ((0.32, -0.56, 0.08), (0.15, 0.18, 0.17)),
    ((-0.50, -0.05, 0.10), (0.12, 0.22, 0.12)),  # arms
    ((0.50, -0.05, 0.10), (0.12, 0.22, 0.12)),
]


def bear_density(p, peak=50.0, shell=0.9, width=0.03):
    """p [..., 3] -> density >= 0.  Each blob is a thin shell at ``shell`` x its ellipsoid radius (a trained NeRF
    keeps density near surfaces): with the 128^3 x 2 grid this gives ~2 % occupied cells and ~270 k samples for
    the 142 x 105 image of BASELINE.json's configs[1]."""
    out = torch.zeros(p.shape[:-1], dtype=torch.float32, device=p.device)
    for c, r in _BLOBS:
        q = (p - torch.tensor(c, device=p.device)) / torch.tensor(r, device=p.device)
        s = (q * q).sum(-1).sqrt()
        out = out + peak * torch.exp(-((s - shell) / width) ** 2)
    return out


def bear_color(p):
    """smooth target colour in [0,1]^3"""
    return 0.5 + 0.5 * torch.sin(p * torch.tensor([3.0, 5.0, 7.0], device=p.device) + torch.tensor([0.0, 1.0, 2.0], device=p.device))


def _expand_bits(v):
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def morton3d(x, y, z):
    """int64 tensors in [0, 1024) -> morton code (same bit layout as raymarching.cu:56-71)"""
    return _expand_bits(x) | (_expand_bits(y) << 1) | (_expand_bits(z) << 2)


def density_grid(bound=2, grid_size=128, device='cpu', peak=50.0):
    """[cascade, G^3] float32 in [cascade][morton] order: density at every cell centre."""
    cascade = 1 + math.ceil(math.log2(bound))
    G = grid_size
    ar = torch.arange(G, dtype=torch.int64, device=device)
    xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing='ij')
    idx = morton3d(xx.reshape(-1), yy.reshape(-1), zz.reshape(-1))
    centres = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1).float()
    centres = (centres + 0.5) / G * 2 - 1
    grid = torch.zeros(cascade, G ** 3, dtype=torch.float32, device=device)
    for cas in range(cascade):
        b = min(2 ** cas, bound)
        grid[cas, idx] = bear_density(centres * b, peak)
    return grid


def pack_bitfield_reference_order(grid, thresh):
    """bit i of byte n <-> cell 8n+i, strict '>' (host-side helper for building test inputs)"""
    bits = (grid.reshape(-1, 8) > thresh).to(torch.int32)
    w = (1 << torch.arange(8, device=grid.device, dtype=torch.int32))
    return (bits * w).sum(-1).to(torch.uint8)


def camera_rays(H, W, n_views=1, radius=1.5, fov_deg=62.0, seed=0, view=0, device='cpu'):
    """Pin-hole camera on a sphere of ``radius`` looking at the origin.  Returns rays_o, rays_d [H*W, 3] float32
    (unit-norm directions), row-major pixel order."""
    g = torch.Generator().manual_seed(seed + 7919 * view)
    theta = float(torch.rand(1, generator=g)) * 2 * math.pi
    phi = (float(torch.rand(1, generator=g)) - 0.5) * 0.8
    eye = torch.tensor([radius * math.cos(phi) * math.sin(theta), radius * math.sin(phi),
                        radius * math.cos(phi) * math.cos(theta)], dtype=torch.float32)
    fwd = -eye / eye.norm()
    up = torch.tensor([0.0, 1.0, 0.0])
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    upv = torch.linalg.cross(right, fwd)
    focal = 0.5 * W / math.tan(0.5 * math.radians(fov_deg))
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
    x = (i + 0.5 - 0.5 * W) / focal
    y = -(j + 0.5 - 0.5 * H) / focal
    d = x.reshape(-1, 1) * right + y.reshape(-1, 1) * upv + fwd
    d = d / d.norm(dim=-1, keepdim=True)
    o = eye.expand_as(d).contiguous()
    return o.to(device), d.contiguous().to(device)


def camera_pose(H, W, radius=1.5, fov_deg=62.0, seed=0, view=0):
    """The camera of ``camera_rays`` in the loader's terms: (pose [4,4] cam2world fp32, (fx, fy, cx, cy)) such that
    get_rays(pose, intrinsics, H, W) (nerf/provider_utils.py:238-302) yields the same rays."""
    g = torch.Generator().manual_seed(seed + 7919 * view)
    theta = float(torch.rand(1, generator=g)) * 2 * math.pi
    phi = (float(torch.rand(1, generator=g)) - 0.5) * 0.8
    eye = torch.tensor([radius * math.cos(phi) * math.sin(theta), radius * math.sin(phi),
                        radius * math.cos(phi) * math.cos(theta)], dtype=torch.float32)
    fwd = -eye / eye.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 1.0, 0.0]))
    right = right / right.norm()
    upv = torch.linalg.cross(right, fwd)
    pose = torch.eye(4, dtype=torch.float32)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, -upv, fwd, eye
    focal = 0.5 * W / math.tan(0.5 * math.radians(fov_deg))
    return pose, (focal, focal, 0.5 * W, 0.5 * H)


def random_rays(N, radius=1.5, seed=0, device='cpu'):
    """N rays from random points on the camera sphere towards random points of the unit ball (config C1/C5)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(N, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * radius
    tgt = (torch.rand(N, 3, generator=g) * 2 - 1) * 0.8
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True)
    return o.float().to(device), d.float().to(device)
