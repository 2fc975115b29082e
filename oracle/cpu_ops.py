"""numpy front-end of the C restatement (oracle/nerf_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Function names, argument order and
allocation rules mirror the reference's Python wrappers (raymarching/raymarching.py,
gridencoder/grid.py) so parity tests read like calls into the reference.
"""
import ctypes as C
import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


u32, f32, i32 = C.c_uint32, C.c_float, C.c_int


def _sc(scales):
    """optional per-level scale override (device-computed exp2f; see orc_locate in nerf_oracle.c)"""
    return None if scales is None else _f32(scales)


# --------------------------------------------------------------------------- grid layout
def grid_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, align_corners=False):
    """Level offset table; gridencoder/grid.py:107-108,124-134.  Returns (offsets int32[L+1], per_level_scale)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), per_level_scale


def grid_encode_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                        gridtype=0, align_corners=False, interpolation=0, max_level=None, half=False, scales=None):
    """gridencoder/grid.py:27-69 (+ kernel_grid).  Returns (outputs [B, L*C], dy_dx or None)."""
    inputs = _f32(inputs)
    emb = _f32(embeddings)
    if half:
        emb = emb.astype(np.float16).astype(np.float32)
    offsets = _i32(offsets)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    Cc = emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    max_level = L if max_level is None else min(max_level, L)
    out = np.zeros((L, B, Cc), np.float32)
    dy_dx = np.zeros((B, L * D * Cc), np.float32) if calc_grad_inputs else None
    lib().orc_grid_encode_forward(_p(inputs), _p(emb), _p(offsets), _p(out), u32(B), u32(D), u32(Cc), u32(L),
                                  u32(max_level), f32(S), u32(base_resolution), _p(dy_dx), u32(gridtype),
                                  i32(int(align_corners)), u32(interpolation), i32(int(half)), _p(_sc(scales)))
    out = np.ascontiguousarray(out.transpose(1, 0, 2)).reshape(B, L * Cc)   # grid.py:63
    return out, dy_dx


def grid_encode_backward(grad, inputs, embeddings_shape, offsets, per_level_scale, base_resolution, dy_dx=None,
                         gridtype=0, align_corners=False, interpolation=0, max_level=None, scales=None):
    """gridencoder/grid.py:74-95 (+ kernel_grid_backward / kernel_input_backward).
    grad: [B, L*C].  Returns (grad_embeddings [rows, C], grad_inputs or None)."""
    inputs = _f32(inputs)
    offsets = _i32(offsets)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    rows, Cc = embeddings_shape
    S = np.float32(np.log2(per_level_scale))
    max_level = L if max_level is None else min(max_level, L)
    g = np.ascontiguousarray(_f32(grad).reshape(B, L, Cc).transpose(1, 0, 2))   # grid.py:81
    ge = np.zeros((rows, Cc), np.float32)
    gi = np.zeros((B, D), np.float32) if dy_dx is not None else None
    lib().orc_grid_encode_backward(_p(g), _p(inputs), _p(offsets), _p(ge), u32(B), u32(D), u32(Cc), u32(L),
                                   u32(max_level), f32(S), u32(base_resolution), _p(None if dy_dx is None else _f32(dy_dx)),
                                   _p(gi), u32(gridtype), i32(int(align_corners)), u32(interpolation), _p(_sc(scales)))
    return ge, gi


def grad_total_variation(inputs, embeddings, grad, offsets, weight, per_level_scale, base_resolution,
                         gridtype=0, align_corners=False, scales=None):
    """gridencoder/grid.py:171-192 (+ kernel_grad_tv).  inputs already in [0,1].  Returns the updated grad."""
    inputs = _f32(inputs)
    emb = _f32(embeddings)
    grad = _f32(grad).copy()
    offsets = _i32(offsets)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    S = np.float32(np.log2(per_level_scale))
    lib().orc_grad_total_variation(_p(inputs), _p(emb), _p(grad), _p(offsets), f32(weight), u32(B), u32(D),
                                   u32(emb.shape[1]), u32(L), f32(S), u32(base_resolution), u32(gridtype),
                                   i32(int(align_corners)), _p(_sc(scales)))
    return grad


# --------------------------------------------------------------------------- raymarching utils
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching/raymarching.py:20-50."""
    rays_o = _f32(rays_o).reshape(-1, 3)
    rays_d = _f32(rays_d).reshape(-1, 3)
    aabb = _f32(aabb)
    N = rays_o.shape[0]
    nears = np.empty(N, np.float32)
    fars = np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), u32(N), f32(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o = _f32(rays_o).reshape(-1, 3)
    rays_d = _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().orc_sph_from_ray(_p(rays_o), _p(rays_d), f32(radius), u32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i32(coords)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().orc_morton3D(_p(coords), u32(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().orc_morton3D_invert(_p(indices), u32(N), _p(out))
    return out


def packbits(grid, thresh, bitfield=None):
    """raymarching/raymarching.py:130-156.  grid [C, H^3] float32."""
    grid = _f32(grid)
    N = grid.shape[0] * grid.shape[1] // 8
    if bitfield is None:
        bitfield = np.empty(N, np.uint8)
    lib().orc_packbits(_p(grid), u32(N), f32(np.float32(thresh)), _p(bitfield))
    return bitfield


# --------------------------------------------------------------------------- training ops
def march_rays_train(rays_o, rays_d, bound, density_bitfield, Cc, H, nears, fars, step_counter=None, mean_count=-1,
                     noises=None, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
    """raymarching/raymarching.py:162-236.  ``noises`` replaces the wrapper's torch.rand (None == perturb False)."""
    rays_o = _f32(rays_o).reshape(-1, 3)
    rays_d = _f32(rays_d).reshape(-1, 3)
    bf = np.ascontiguousarray(density_bitfield, dtype=np.uint8)
    N = rays_o.shape[0]
    M = N * max_steps
    if not force_all_rays and mean_count > 0:
        if align > 0:
            mean_count += align - mean_count % align
        M = mean_count
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    if step_counter is None:
        step_counter = np.zeros(2, np.int32)
    noises = np.zeros(N, np.float32) if noises is None else _f32(noises)
    lib().orc_march_rays_train(_p(rays_o), _p(rays_d), _p(bf), f32(bound), f32(dt_gamma), u32(max_steps), u32(N),
                               u32(Cc), u32(H), u32(M), _p(_f32(nears)), _p(_f32(fars)), _p(xyzs), _p(dirs),
                               _p(deltas), _p(rays), _p(step_counter), _p(noises))
    if force_all_rays or mean_count <= 0:
        m = int(step_counter[0])
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    return xyzs, dirs, deltas, rays


def march_rays_count(rays_o, rays_d, bound, density_bitfield, Cc, H, nears, fars, noises=None, dt_gamma=0,
                     max_steps=1024):
    rays_o = _f32(rays_o).reshape(-1, 3)
    rays_d = _f32(rays_d).reshape(-1, 3)
    bf = np.ascontiguousarray(density_bitfield, dtype=np.uint8)
    N = rays_o.shape[0]
    noises = np.zeros(N, np.float32) if noises is None else _f32(noises)
    counts = np.empty(N, np.int32)
    lib().orc_march_rays_count(_p(rays_o), _p(rays_d), _p(bf), f32(bound), f32(dt_gamma), u32(max_steps), u32(N),
                               u32(Cc), u32(H), _p(_f32(nears)), _p(_f32(fars)), _p(noises), _p(counts))
    return counts


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    """raymarching/raymarching.py:239-270."""
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws = np.zeros(N, np.float32)
    depth = np.zeros(N, np.float32)
    image = np.zeros((N, 3), np.float32)
    lib().orc_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), u32(M), u32(N), f32(T_thresh),
                                           _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image,
                                  T_thresh=1e-4):
    """raymarching/raymarching.py:272-289."""
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    gs = np.zeros_like(sigmas)
    gc = np.zeros_like(rgbs)
    lib().orc_composite_rays_train_backward(_p(_f32(grad_weights_sum)), _p(_f32(grad_image)), _p(sigmas), _p(rgbs),
                                            _p(deltas), _p(rays), _p(_f32(weights_sum)), _p(_f32(image)), u32(M),
                                            u32(N), f32(T_thresh), _p(gs), _p(gc))
    return gs, gc


# --------------------------------------------------------------------------- inference ops
def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, Cc, H, near, far,
               align=-1, noises=None, dt_gamma=0, max_steps=1024):
    """raymarching/raymarching.py:355-405."""
    rays_o = _f32(rays_o).reshape(-1, 3)
    rays_d = _f32(rays_d).reshape(-1, 3)
    bf = np.ascontiguousarray(density_bitfield, dtype=np.uint8)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    noises = np.zeros(n_alive, np.float32) if noises is None else _f32(noises)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _p(_i32(rays_alive)), _p(_f32(rays_t)), _p(rays_o), _p(rays_d),
                         f32(bound), f32(dt_gamma), u32(max_steps), u32(Cc), u32(H), _p(bf), _p(_f32(near)),
                         _p(_f32(far)), _p(xyzs), _p(dirs), _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image,
                   T_thresh=1e-2):
    """raymarching/raymarching.py:408-427.  In-place on the numpy arrays passed (must be contiguous, right dtype)."""
    for a, dt in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32), (depth, np.float32),
                  (image, np.float32)):
        assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
    lib().orc_composite_rays(u32(n_alive), u32(n_step), f32(T_thresh), _p(rays_alive), _p(rays_t), _p(_f32(sigmas)),
                             _p(_f32(rgbs)), _p(_f32(deltas)), _p(weights_sum), _p(depth), _p(image))


def get_rays(poses, intrinsics, H, W, inds=None, offset=(0.5, 0.5)):
    """numpy restatement of the direction / origin arithmetic of get_rays (nerf/provider_utils.py:238-302): pixel
    (i, j) = (p % W + off_x, p // W + off_y) (:258-260), dir = safe_normalize((i - cx) / fx, (j - cy) / fy, 1)
    (:125-126, :289-293), rays_d = dir @ R^T, rays_o = t (:294-297).  poses [B,4,4]; inds [B,N] int or None (all pixels).
    Returns rays_o, rays_d [B,N,3] float32."""
    poses = np.asarray(poses, np.float32)
    B = poses.shape[0]
    fx, fy, cx, cy = [np.float32(v) for v in intrinsics]
    if inds is None:
        inds = np.broadcast_to(np.arange(H * W, dtype=np.int64), (B, H * W))
    inds = np.asarray(inds, np.int64)
    i = (inds % W).astype(np.float32) + np.float32(offset[0])
    j = (inds // W).astype(np.float32) + np.float32(offset[1])
    xs, ys = (i - cx) / fx, (j - cy) / fy
    d = np.stack([xs, ys, np.ones_like(xs)], -1)
    d = d / np.sqrt(np.maximum((d * d).sum(-1, keepdims=True, dtype=np.float32), np.float32(1e-20)))
    rays_d = np.einsum("bnc,bkc->bnk", d, poses[:, :3, :3]).astype(np.float32)
    rays_o = np.broadcast_to(poses[:, None, :3, 3], rays_d.shape).astype(np.float32)
    return np.ascontiguousarray(rays_o), np.ascontiguousarray(rays_d)


# ---- LGIE composites on the occupancy path ---------------------------------------------------------------------------
def lgie_gate(m, variant, soft_mask, conf_thr):
    """(gate, d gate / d m) of composite variant 0 all / 1 fg / 2 bg for mask values m [M]: the edit mask of
    nerf/renderer.py:421-426 (soft: sigmoid((m - conf_thr) * 100); hard: m > 0.5) and its complement."""
    m = _f32(m)
    if variant == 0:
        return np.ones_like(m), np.zeros_like(m)
    if soft_mask:
        e = (1.0 / (1.0 + np.exp(-(m.astype(np.float64) - conf_thr) * 100.0))).astype(np.float32)
        de = np.float32(100.0) * e * (1 - e)
    else:
        e, de = (m > 0.5).astype(np.float32), np.zeros_like(m)
    return (e, de) if variant == 1 else (1 - e, -de)


def composite_lgie_forward(variant, sigmas, rgbs, masks, deltas, rays, T_thresh=1e-4, soft_mask=True, conf_thr=0.5):
    """One LGIE render (nerf/renderer.py:383-474 on the occupancy path, composed as rendering._lgie_composites does):
    the composite of raymarching.py:239-270 over sigma * gate, plus the rendered mask sum w * m (the same composite with the
    mask as colour).  Returns weights_sum, depth, image [N,3], render_mask [N]."""
    gate, _ = lgie_gate(masks, variant, soft_mask, conf_thr)
    sv = _f32(sigmas) * gate
    ws, depth, image = composite_rays_train_forward(sv, rgbs, deltas, rays, T_thresh)
    m3 = np.repeat(_f32(masks).reshape(-1, 1), 3, axis=1)
    _, _, mimg = composite_rays_train_forward(sv, m3, deltas, rays, T_thresh)
    return ws, depth, image, np.ascontiguousarray(mimg[:, 0])


def composite_lgie_backward(variant, g_ws, g_image, g_mask, sigmas, rgbs, masks, deltas, rays, T_thresh=1e-4, soft_mask=True,
                            conf_thr=0.5, detach_bg=False, detach_mask_from_field=False):
    """Gradients of one LGIE render with respect to (sigma [M], rgb [M,3], mask [M]) given the gradients of the loss with
    respect to its weights_sum / image / render_mask: the backward of raymarching.py:272-289 applied to the colour composite
    and to the mask composite, chained through the gate; detach_bg (:409-418) and detach_mask_from_field (:460-463) as in
    rendering._lgie_composites."""
    sig, m = _f32(sigmas), _f32(masks)
    gate, dgate = lgie_gate(m, variant, soft_mask, conf_thr)
    sv = sig * gate
    ws, _, image = composite_rays_train_forward(sv, rgbs, deltas, rays, T_thresh)
    gs_c, g_rgb = composite_rays_train_backward(g_ws, g_image, sv, rgbs, deltas, rays, ws, image, T_thresh)
    m3 = np.repeat(m.reshape(-1, 1), 3, axis=1)
    ws_m, _, mimg = composite_rays_train_forward(sv, m3, deltas, rays, T_thresh)
    gm3 = np.zeros_like(_f32(g_image))
    gm3[:, 0] = _f32(g_mask)
    gs_m, g_m3 = composite_rays_train_backward(np.zeros_like(_f32(g_ws)), gm3, sv, m3, deltas, rays, ws_m, mimg, T_thresh)
    gs = gs_c + (0.0 if detach_mask_from_field else 1.0) * gs_m            # gradient w.r.t. the gated density
    a = (m >= 0.5).astype(np.float32) if (variant == 0 and detach_bg) else np.ones_like(m)
    d_sigma = gs * gate * a
    d_rgb = g_rgb * a[:, None]
    d_mask = g_m3[:, 0] + gs * sig * dgate
    return d_sigma, d_rgb, d_mask
