"""Fused field network (trunk + density head + colour head + embedder + trunc_exp) on the tcgen05 kernels of
csrc/field_mlp.cu, as one autograd node.  Same math and the same flat tcnn-layout parameters as the three
``Network`` modules of ``NeRFNetwork`` (nerf/network_grid.py:98-139, :159-177)."""
import torch
from torch.autograd import Function

from .. import _lib as L


def _image_bytes():
    return int(L.lib().nb200_field_weight_image_bytes())


class _PackedWeights:
    """fp16 pre-swizzled operand images of the three parameter vectors.  Re-packed on every call (two tiny launches):
    in-place optimiser kernels such as torch's fused Adam do not bump tensor versions, so nothing cheaper is safe."""

    def __init__(self):
        self.fwd = self.bwd = None

    def get(self, trunk, density, rgb):
        n = _image_bytes()
        if self.fwd is None or self.fwd.device != trunk.device or torch.is_grad_enabled():
            # a fresh pair (one allocation from torch's caching allocator) while autograd may still hold the previous one
            # for a pending backward
            both = torch.empty(2, n, dtype=torch.uint8, device=trunk.device)
            self.fwd, self.bwd = both[0], both[1]
        L.check(L.lib().nb200_field_pack_weights(L.ptr(trunk.detach()), L.ptr(density.detach()), L.ptr(rgb.detach()),
                                                 L.ptr(self.fwd), L.ptr(self.bwd), L.stream()), "field_pack_weights")
        return self.fwd, self.bwd


_wg = {}


def _wg_scratch(dev):
    """per-device scratch for the backward kernel's per-CTA partial weight-gradient sums"""
    key = (dev.type, dev.index)
    if key not in _wg:
        L.lib().nb200_field_wgrad_scratch_bytes.restype = L.u32
        _wg[key] = torch.empty(int(L.lib().nb200_field_wgrad_scratch_bytes()) // 4, dtype=torch.float32, device=dev)
    return _wg[key]


class _FusedField(Function):
    @staticmethod
    def forward(ctx, x_en, xyz, dirs, trunk, density, rgb, packed, save):
        M = x_en.shape[0]
        dev = x_en.device
        x_en = x_en.half().contiguous()
        xyz = xyz.float().contiguous()
        dirs = dirs.float().contiguous()
        fwd_img, bwd_img = packed.get(trunk, density, rgb)
        sigma = torch.empty(M, dtype=torch.float32, device=dev)
        rgba = torch.empty(M, 4, dtype=torch.half, device=dev)
        sigma_arg = torch.empty(M, dtype=torch.float32, device=dev) if save else None
        act = torch.empty(5, M, 64, dtype=torch.half, device=dev) if save else None
        L.check(L.lib().nb200_field_forward(L.ptr(x_en), L.ptr(xyz), L.ptr(dirs), L.ptr(fwd_img), L.ptr(sigma),
                                            L.ptr(sigma_arg), L.ptr(rgba), L.ptr(act), L.u32(M), L.ptr(None),
                                            L.stream()),
                "field_forward")
        if save:
            ctx.save_for_backward(x_en, dirs, sigma_arg, rgba, act, bwd_img)
            ctx.shapes = (trunk.numel(), density.numel(), rgb.numel())
        return sigma, rgba

    @staticmethod
    def backward(ctx, d_sigma, d_rgba):
        x_en, dirs, sigma_arg, rgba, act, bwd_img = ctx.saved_tensors
        M = x_en.shape[0]
        dev = x_en.device
        d_sigma = (torch.zeros(M, device=dev) if d_sigma is None else d_sigma.float()).contiguous()
        d_rgba = (torch.zeros(M, 4, device=dev) if d_rgba is None else d_rgba.float()).contiguous()
        nt, nd, nr = ctx.shapes
        g_trunk = torch.zeros(nt, dtype=torch.float32, device=dev)
        g_density = torch.zeros(nd, dtype=torch.float32, device=dev)
        g_rgb = torch.zeros(nr, dtype=torch.float32, device=dev)
        d_x_en = torch.empty(M, 32, dtype=torch.half, device=dev)
        L.check(L.lib().nb200_field_backward(L.ptr(d_sigma), L.ptr(d_rgba), L.ptr(sigma_arg), L.ptr(rgba), L.ptr(x_en),
                                             L.ptr(dirs), L.ptr(act), L.ptr(bwd_img), L.ptr(d_x_en), L.ptr(g_trunk),
                                             L.ptr(g_density), L.ptr(g_rgb), L.u32(M), L.ptr(None), L.ptr(_wg_scratch(dev)),
                                             L.ptr(None), L.stream()),
                "field_backward")
        return d_x_en, None, None, g_trunk, g_density, g_rgb, None, None


def fused_field(x_en, xyz, dirs, trunk, density, rgb, packed):
    """-> (sigma f32 [M], rgba f16 [M, 4]).  x_en: grid encoding [M, 32]."""
    if not x_en.is_cuda:
        raise RuntimeError("customnerf_b200 fused field network runs on CUDA only (no CPU fallback)")
    save = torch.is_grad_enabled() and any(t.requires_grad for t in (x_en, trunk, density, rgb))
    return _FusedField.apply(x_en, xyz, dirs, trunk, density, rgb, packed, save)


# ---------------------------------------------------------------------------------------------------------------------
# grid encoding + field network as ONE kernel / one autograd node (csrc/field_fused.cu): the [M,32] features are gathered by
# producer warps straight into the tensor-core operand tile; in training they are written once (for the backward pass only)
def fused_encode_eligible(enc, x):
    """the fused kernel covers the product's shape under autocast: D = 3, C = 2, 16 levels, fp32 master table (entries are
    rounded to fp16 on load, as the autocast encoder does, grid.py:45-46), positions without gradient"""
    return (x.is_cuda and torch.is_autocast_enabled() and enc.input_dim == 3 and enc.level_dim == 2 and enc.num_levels == 16
            and enc.embeddings.dtype == torch.float32 and not x.requires_grad)


def _enc_args(enc):
    import numpy as np
    return (L.ptr(enc.embeddings.detach()), L.ptr(enc.offsets), L.u32(enc.num_levels), L.f32(float(np.log2(enc.per_level_scale))),
            L.u32(int(enc.base_resolution)), L.u32(enc.gridtype_id), L.i32(int(enc.align_corners)), L.u32(enc.interp_id))


class _FusedEncodeField(Function):
    @staticmethod
    def forward(ctx, xyz, dirs, embeddings, trunk, density, rgb, enc, bound, packed, save):
        M = xyz.shape[0]
        dev = xyz.device
        xyz = xyz.float().contiguous()
        dirs = dirs.float().contiguous()
        fwd_img, bwd_img = packed.get(trunk, density, rgb)
        sigma = torch.empty(M, dtype=torch.float32, device=dev)
        rgba = torch.empty(M, 4, dtype=torch.half, device=dev)
        sigma_arg = torch.empty(M, dtype=torch.float32, device=dev) if save else None
        x_en = torch.empty(M, 32, dtype=torch.half, device=dev) if save else None
        act = torch.empty(5, M, 64, dtype=torch.half, device=dev) if save else None
        L.check(L.lib().nb200_field_fused_forward(L.ptr(xyz), L.ptr(dirs), L.f32(float(bound)), *_enc_args(enc), L.ptr(fwd_img),
                                                  L.ptr(sigma), L.ptr(sigma_arg), L.ptr(rgba), L.ptr(x_en), L.ptr(act), L.u32(M),
                                                  L.ptr(None), L.stream()), "field_fused_forward")
        if save:
            ctx.save_for_backward(xyz, dirs, x_en, sigma_arg, rgba, act, bwd_img, enc.offsets)
            ctx.enc, ctx.bound = enc, float(bound)
            ctx.shapes = (tuple(embeddings.shape), trunk.numel(), density.numel(), rgb.numel())
        return sigma, rgba

    @staticmethod
    def backward(ctx, d_sigma, d_rgba):
        import numpy as np
        xyz, dirs, x_en, sigma_arg, rgba, act, bwd_img, offsets = ctx.saved_tensors
        enc = ctx.enc
        M = xyz.shape[0]
        dev = xyz.device
        d_sigma = (torch.zeros(M, device=dev) if d_sigma is None else d_sigma.float()).contiguous()
        d_rgba = (torch.zeros(M, 4, device=dev) if d_rgba is None else d_rgba.float()).contiguous()
        emb_shape, nt, nd, nr = ctx.shapes
        g_trunk = torch.zeros(nt, dtype=torch.float32, device=dev)
        g_density = torch.zeros(nd, dtype=torch.float32, device=dev)
        g_rgb = torch.zeros(nr, dtype=torch.float32, device=dev)
        g_table = torch.zeros(emb_shape, dtype=torch.float32, device=dev)
        d_x_en = torch.empty(M, 32, dtype=torch.half, device=dev)
        L.check(L.lib().nb200_field_backward(L.ptr(d_sigma), L.ptr(d_rgba), L.ptr(sigma_arg), L.ptr(rgba), L.ptr(x_en),
                                             L.ptr(dirs), L.ptr(act), L.ptr(bwd_img), L.ptr(d_x_en), L.ptr(g_trunk),
                                             L.ptr(g_density), L.ptr(g_rgb), L.u32(M), L.ptr(None), L.ptr(_wg_scratch(dev)),
                                             L.ptr(None), L.stream()),
                "field_backward")
        L.check(L.lib().nb200_fs_encode_backward(L.ptr(d_x_en), L.ptr(xyz), L.f32(ctx.bound), L.ptr(offsets), L.ptr(g_table),
                                                 L.u32(M), L.u32(enc.num_levels), L.f32(float(np.log2(enc.per_level_scale))),
                                                 L.u32(int(enc.base_resolution)), L.u32(enc.gridtype_id),
                                                 L.i32(int(enc.align_corners)), L.u32(enc.interp_id), L.ptr(None), L.stream()),
                "fs_encode_backward")
        return None, None, g_table, g_trunk, g_density, g_rgb, None, None, None, None


def fused_encode_field(xyz, dirs, enc, bound, trunk, density, rgb, packed):
    """positions [M,3] in [-bound, bound], view directions [M,3] -> (sigma f32 [M], rgba f16 [M,4]) in one launch"""
    save = torch.is_grad_enabled() and any(t.requires_grad for t in (enc.embeddings, trunk, density, rgb))
    return _FusedEncodeField.apply(xyz, dirs, enc.embeddings, trunk, density, rgb, enc, bound, packed, save)


def fused_density(xyz, enc, bound, trunk, density, rgb, packed):
    """density only (trunk + density head), no autograd: sigma f32 [M] in one launch"""
    M = xyz.shape[0]
    xyz = xyz.float().contiguous()
    fwd_img, _ = packed.get(trunk, density, rgb)
    sigma = torch.empty(M, dtype=torch.float32, device=xyz.device)
    L.check(L.lib().nb200_field_fused_forward(L.ptr(xyz), L.ptr(None), L.f32(float(bound)), *_enc_args(enc), L.ptr(fwd_img),
                                              L.ptr(sigma), L.ptr(None), L.ptr(None), L.ptr(None), L.ptr(None), L.u32(M),
                                              L.ptr(None), L.stream()), "field_fused_forward(density)")
    return sigma
