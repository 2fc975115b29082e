"""Drop-in for the reference's ``gridencoder/grid.py`` (GridEncoder / grid_encode) on top of libnerf_b200.so.

Same public surface (``GridEncoder``, ``grid_encode``, ``_grid_encode``), same argument meaning and the same
autocast behaviour; what changed underneath (reference file:line in brackets):

  * outputs are written by the kernel directly as [B, L*C]  [grid.py:49,63: [L,B,C] + permute copy];
  * the backward reads ``grad`` as [B, L*C] directly        [grid.py:81: view/permute/contiguous copy];
  * under autocast the kernel reads the fp32 master table and rounds every entry to fp16 as it loads it
    (bit-identical to gathering from a half copy)           [grid.py:45-46: full-table fp32->fp16 copy every call];
  * ``grad_embeddings`` is accumulated in fp32 by a warp-aggregated scatter and returned as fp32
                                                            [gridencoder.cu:324-337: __half2 atomics].
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from .. import _lib as L

_gridtype_to_id = {'hash': 0, 'tiled': 1}
_interp_to_id = {'linear': 0, 'smoothstep': 1}

def _fast_shape(D, C, Lv, calc_grad_inputs):
    return D == 3 and C == 2 and Lv <= 32 and not calc_grad_inputs


def _check_inputs(inputs, embeddings, offsets):
    # the reference's TORCH_CHECKs, gridencoder.cu:448-463
    for name, t in (("inputs", inputs), ("embeddings", embeddings), ("offsets", offsets)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous tensor")
    if offsets.dtype != torch.int32:
        raise RuntimeError("offsets must be an int tensor")
    if not inputs.is_floating_point():
        raise RuntimeError("inputs must be a floating tensor")
    if not embeddings.is_floating_point():
        raise RuntimeError("embeddings must be a floating tensor")


class _grid_encode(Function):
    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, interpolation=0, max_level=None):
        # inputs: [B, D] float in [0, 1]; embeddings: [sO, C]; offsets: [L + 1] int32; returns [B, L*C]
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()           # the kernels read fp32 coordinates (gridencoder.cu:89)
        B, D = inputs.shape
        Lv = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        max_level = Lv if max_level is None else min(max_level, Lv)

        master = embeddings
        tag = None
        out_dtype = embeddings.dtype
        if torch.is_autocast_enabled() and C % 2 == 0 and embeddings.dtype == torch.float32:
            # manual autocast (grid.py:43-46): half-precision table values, fp32 coordinates
            out_dtype = torch.half
            if _fast_shape(D, C, Lv, calc_grad_inputs):
                tag = L.F32_AS_F16            # the kernel rounds each fp32 entry to fp16 as it loads it: no table copy
            else:
                embeddings = embeddings.to(torch.half)
        _check_inputs(inputs, embeddings, offsets)
        if D not in (2, 3, 4, 5) or C not in (1, 2, 4, 8):
            raise RuntimeError("GridEncoding: C must be 1, 2, 4, or 8.")
        if tag is None:
            tag = L.dtype_tag(embeddings)

        if max_level < Lv:
            outputs = torch.zeros(B, Lv * C, device=inputs.device, dtype=out_dtype)
        else:
            outputs = torch.empty(B, Lv * C, device=inputs.device, dtype=out_dtype)
        if calc_grad_inputs:
            alloc = torch.zeros if max_level < Lv else torch.empty
            dy_dx = alloc(B, Lv * D * C, device=inputs.device, dtype=out_dtype)
        else:
            dy_dx = None

        L.check(L.lib().nb200_grid_encode_forward(
            L.ptr(inputs), L.ptr(embeddings), L.ptr(offsets), L.ptr(outputs), L.u32(B), L.u32(D), L.u32(C), L.u32(Lv),
            L.u32(max_level), L.f32(S), L.u32(H), L.ptr(dy_dx), L.u32(gridtype), L.i32(int(align_corners)),
            L.u32(interpolation), L.i32(tag), L.i32(L.LAYOUT_BLC), L.stream()),
            "grid_encode_forward")

        ctx.save_for_backward(inputs, offsets, dy_dx)
        ctx.dims = [B, D, C, Lv, S, H, gridtype, interpolation, max_level]
        ctx.align_corners = align_corners
        ctx.emb_shape = tuple(master.shape)
        ctx.emb_dtype = master.dtype
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        inputs, offsets, dy_dx = ctx.saved_tensors
        B, D, C, Lv, S, H, gridtype, interpolation, max_level = ctx.dims
        grad = grad.contiguous()                                   # [B, L*C], read in place
        if dy_dx is not None and grad.dtype != dy_dx.dtype:
            grad = grad.to(dy_dx.dtype)
        grad_embeddings = torch.zeros(ctx.emb_shape, dtype=torch.float32, device=grad.device)
        grad_inputs = torch.zeros(B, D, dtype=grad.dtype, device=grad.device) if dy_dx is not None else None
        L.check(L.lib().nb200_grid_encode_backward(
            L.ptr(grad), L.ptr(inputs), L.ptr(offsets), L.ptr(grad_embeddings), L.u32(B), L.u32(D), L.u32(C),
            L.u32(Lv), L.u32(max_level), L.f32(S), L.u32(H), L.ptr(dy_dx), L.ptr(grad_inputs), L.u32(gridtype),
            L.i32(int(ctx.align_corners)), L.u32(interpolation), L.i32(L.dtype_tag(grad)), L.i32(L.LAYOUT_BLC),
            L.i32(1), L.stream()), "grid_encode_backward")
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        if ctx.emb_dtype != torch.float32:
            grad_embeddings = grad_embeddings.to(ctx.emb_dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


def level_offsets(input_dim, num_levels, per_level_scale, base_resolution, max_params, align_corners=False):
    """Row offset of every level, int32 [L+1] (reference layout rule, grid.py:124-134): level i has
    min(max_params, (ceil(H * s^i) + 1)^D) rows (no +1 with align_corners), rounded up to a multiple of 8.
    Evaluated per level on Python scalars like the reference does (scalar ``**``: libm pow and arbitrary-precision integer
    powers), so that the table -- and with it checkpoint compatibility -- cannot differ by a vectorised-pow ulp or an
    int64 overflow at D = 5."""
    offsets, offset = [], 0
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        rows = min(int(max_params), (resolution if align_corners else resolution + 1) ** input_dim)
        rows = int(np.ceil(rows / 8) * 8)
        offsets.append(offset)
        offset += rows
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


class GridEncoder(nn.Module):
    """Same constructor, attributes and forward as the reference's GridEncoder (grid.py:102-168)."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 interpolation='linear'):
        super().__init__()
        # the finest resolution desired at the last level, if provided, overrides per_level_scale
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners

        self.max_params = 2 ** log2_hashmap_size
        table = level_offsets(input_dim, num_levels, per_level_scale, base_resolution, self.max_params, align_corners)
        self.register_buffer('offsets', torch.from_numpy(table))
        offset = int(table[-1])
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> "
                f"{int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} "
                f"gridtype={self.gridtype} align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1, max_level=None):
        # inputs: [..., input_dim] in [-bound, bound]; returns [..., num_levels * level_dim]
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id, max_level)
        return outputs.view(prefix_shape + [self.output_dim])

    # always run in float precision!
    @torch.amp.autocast('cuda', enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        D = self.input_dim
        C = self.embeddings.shape[1]
        Lv = self.offsets.shape[0] - 1
        S = float(np.log2(self.per_level_scale))
        H = self.base_resolution
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = (inputs + bound) / (2 * bound)
            inputs = inputs.view(-1, self.input_dim)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        inputs = inputs.float().contiguous()
        L.require_cuda(inputs, self.embeddings, self.embeddings.grad)
        L.check(L.lib().nb200_grad_total_variation(
            L.ptr(inputs), L.ptr(self.embeddings), L.ptr(self.embeddings.grad), L.ptr(self.offsets), L.f32(weight),
            L.u32(B), L.u32(D), L.u32(C), L.u32(Lv), L.f32(S), L.u32(H), L.u32(self.gridtype_id),
            L.i32(int(self.align_corners)), L.stream()), "grad_total_variation")


def level_scales(num_levels, per_level_scale, base_resolution, device='cuda'):
    """Device-evaluated per-level scales exp2f(l*S)*H - 1 (test aid, see include/nerf_b200.h)."""
    out = torch.empty(num_levels, dtype=torch.float32, device=device)
    L.check(L.lib().nb200_grid_level_scales(L.ptr(out), L.u32(num_levels), L.f32(float(np.log2(per_level_scale))),
                                            L.u32(int(base_resolution)), L.stream()), "grid_level_scales")
    return out
