"""NeRFNetwork: the reference's grid-backbone field network (nerf/network_grid.py:70-206) on the B200 ops.

Module and state-dict names follow the reference so its checkpoints map one-to-one:
    pos_en.embeddings [rows, 2] fp32, pos_en.offsets [17] int32,
    network.params, density_network.params, rgb_network.params (flat fp32 vectors).
"""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from ..gridencoder import GridEncoder
from .mlp import Network
from .fused_field import fused_field, fused_encode_field, fused_density, fused_encode_eligible, _PackedWeights
from .rendering import NeRFRenderer


class _trunc_exp(Function):
    """exp forward, gradient through exp(clamp(x, -15, 15))  (nerf/provider_utils.py:16-29)"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply


def freq_embed(d, multires=4):
    """get_embedder(4): [d, sin(2^k d), cos(2^k d)] for k < 4 -> 27 dims  (nerf/base.py:42-77)"""
    feats = [d]
    for k in range(multires):
        feats += [torch.sin(d * (2.0 ** k)), torch.cos(d * (2.0 ** k))]
    return torch.cat(feats, -1)


def get_encoder(encoding, input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                desired_resolution=2048, align_corners=False, **kwargs):
    """nerf/encoding.py:53-70 (grid encodings only)."""
    if encoding == 'None':
        return (lambda x, **kw: x), input_dim
    if encoding not in ('hashgrid', 'tiledgrid'):
        raise NotImplementedError('Unknown encoding mode, choose from [None, hashgrid, tiledgrid]')
    enc = GridEncoder(input_dim=input_dim, num_levels=num_levels, level_dim=level_dim,
                      base_resolution=base_resolution, log2_hashmap_size=log2_hashmap_size,
                      desired_resolution=desired_resolution,
                      gridtype='hash' if encoding == 'hashgrid' else 'tiled', align_corners=align_corners)
    return enc, enc.output_dim


def _mlp_cfg(out_act, n_hidden):
    return {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": out_act, "n_neurons": 64,
            "n_hidden_layers": n_hidden}


class RGB_network(nn.Module):
    """The two-head colour network of --detach_mask_from_field / --mask_no_dir (nerf/network_grid.py:13-68): a colour MLP
    [view_en | fea] -> 3 and a separate confidence (mask) MLP that must not train the field -- fed the DETACHED input, or,
    with mask_no_dir, the (detached unless mask_no_dir_nodetach) field features alone.  Output [rgb(3) | conf].  Same
    sub-module names as the reference (state-dict keys rgb_network.rgb_network.params / rgb_network.conf_network.params).
    Runs on the per-layer library path: the fused tcgen05 field kernel covers the single 3+1 head only."""

    def __init__(self, input_ch_views, opt=None):
        super().__init__()
        self.opt = opt
        self.mask_no_dir = bool(getattr(opt, 'mask_no_dir', False))
        self.rgb_network = Network(input_ch_views + 64, 3, _mlp_cfg("Sigmoid", 1), seed=1339)
        if self.mask_no_dir:
            self.conf_network = Network(64, 1, _mlp_cfg("Sigmoid", 1), seed=1340)
        else:
            ndim = 1 if getattr(opt, 'keyword2', None) is None else 2       # opt.keyword2 is undefined in main.py (Appendix B6)
            self.conf_network = Network(input_ch_views + 64, ndim, _mlp_cfg("Sigmoid", 1), seed=1340)
        self.n_output_dims = 3 + self.conf_network.n_output_dims

    def forward(self, x):
        rgb = self.rgb_network(x)
        if self.mask_no_dir:
            fea = x[..., x.shape[-1] - 64:]
            conf = self.conf_network(fea if getattr(self.opt, 'mask_no_dir_nodetach', False) else fea.detach())
        else:
            conf = self.conf_network(x.detach())
        return torch.cat([rgb, conf], dim=-1)


class NeRFNetwork(NeRFRenderer):
    def __init__(self, opt, encoding='tiledgrid', log2_hashmap_size=21, desired_resolution=8192, **unused):
        """Defaults are the reference's hard-coded grid (network_grid.py:89-96); BASELINE.json's configs use
        encoding='hashgrid', log2_hashmap_size=19, desired_resolution=2048 (nerf/encoding.py:55-58)."""
        super().__init__(opt)
        self.pos_en, self.pos_en_dim = get_encoder(encoding, input_dim=3, log2_hashmap_size=log2_hashmap_size,
                                                   desired_resolution=desired_resolution)
        self.network = Network(self.pos_en_dim, 64, _mlp_cfg("None", 2), seed=1337)
        self.density_network = Network(64, 1, _mlp_cfg("None", 1), seed=1338)
        self.input_ch_views = 27
        n_out = 3 + (1 if getattr(opt, 'train_conf', 0) else 0)
        self.two_heads = bool(getattr(opt, 'train_conf', 0) and (getattr(opt, 'detach_mask_from_field', False)
                                                                   or getattr(opt, 'mask_no_dir', False)))
        if self.two_heads:              # network_grid.py:117-118
            self.rgb_network = RGB_network(self.input_ch_views, opt=opt)
        else:
            self.rgb_network = Network(self.input_ch_views + 64, n_out, _mlp_cfg("Sigmoid", 1), seed=1339)
        self.bg_net = None
        # one fused tcgen05 kernel for trunk + heads when the shapes are the reference's (32 -> 64 ... -> 1 | 3(+1));
        # the per-layer library path (mlp.Network.forward) stays available through use_fused_field = False
        self.use_fused_field = self.pos_en_dim == 32 and not self.two_heads
        self._packed = _PackedWeights()
        # under autocast the grid gather runs inside the field kernel (no [M,32] feature tensor in HBM); False keeps the
        # encoder and the field network as two launches (tests compare the two)
        self.fuse_encoder = True
        # eval on the occupancy path: device-driven rounds in a CUDA graph (fused_infer.py) instead of the reference's
        # host-driven n_step loop; set False to run the loop of NeRFRenderer.run_cuda
        self.fast_inference = True
        self._infer = None
        # occupancy update: cells per encoder + density launch pair (0: the one-kernel form nb200_occ_density)
        self.occ_chunk_rows = 1 << 20

    def background(self, d):
        return torch.zeros(d.size(), dtype=d.dtype, device=d.device)

    def gaussian(self, x):
        """density blob at the scene centre (network_grid.py:150-156)"""
        return 5 * torch.exp(-(x ** 2).sum(-1) / (2 * 0.2 ** 2))

    def _fused(self, x, d):
        if self.fuse_encoder and x.dim() == 2 and fused_encode_eligible(self.pos_en, x):
            # grid gather + MLPs in one kernel: the [M,32] features never round-trip HBM (csrc/field_fused.cu)
            sigma, rgba = fused_encode_field(x.reshape(-1, 3), d.reshape(-1, 3), self.pos_en, self.opt.bound, self.network.params,
                                             self.density_network.params, self.rgb_network.params, self._packed)
            return sigma, rgba[:, :self.rgb_network.n_output_dims]
        x_en = self.pos_en(x, bound=self.opt.bound)
        sigma, rgba = fused_field(x_en, x, d, self.network.params, self.density_network.params,
                                  self.rgb_network.params, self._packed)
        return sigma, rgba[:, :self.rgb_network.n_output_dims]

    def forward(self, x, d, l=None, ratio=1, shading='albedo'):
        if self.use_fused_field:
            sigma, radiances = self._fused(x, d)
            return sigma, radiances, None
        x_en = self.pos_en(x, bound=self.opt.bound)
        fea = self.network(x_en)
        sigma = self.density_network(fea)
        sigma = trunc_exp(sigma.squeeze(-1) + self.gaussian(x))
        rgb_input = torch.cat([freq_embed(d).to(fea.dtype), fea], dim=-1)
        radiances = self.rgb_network(rgb_input)
        return sigma, radiances, None

    def density(self, x):
        if (self.use_fused_field and self.fuse_encoder and not torch.is_grad_enabled() and x.dim() == 2
                and fused_encode_eligible(self.pos_en, x)):
            return {'sigma': fused_density(x, self.pos_en, self.opt.bound, self.network.params, self.density_network.params,
                                           self.rgb_network.params, self._packed)}
        if self.use_fused_field:
            sigma, _ = self._fused(x, torch.zeros_like(x))
            return {'sigma': sigma}
        x_en = self.pos_en(x, bound=self.opt.bound)
        fea = self.network(x_en)
        sigma = self.density_network(fea)
        return {'sigma': trunc_exp(sigma.squeeze(-1) + self.gaussian(x))}

    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024,
                 T_thresh=1e-4, **kwargs):
        if self.training or not (self.fast_inference and self.use_fused_field and torch.is_autocast_enabled()):
            return super().run_cuda(rays_o, rays_d, dt_gamma=dt_gamma, bg_color=bg_color, perturb=perturb,
                                    force_all_rays=force_all_rays, max_steps=max_steps, T_thresh=T_thresh, **kwargs)
        from ..fused_infer import FusedInference
        prefix = rays_o.shape[:-1]
        N = rays_o.reshape(-1, 3).shape[0]
        inf = self._infer
        if (inf is None or inf.N != N or inf.T_thresh != float(T_thresh) or inf.dt_gamma != float(dt_gamma)
                or inf.max_steps != int(max_steps)):
            inf = self._infer = FusedInference(self, N, T_thresh=T_thresh, dt_gamma=dt_gamma, max_steps=max_steps)
        weights_sum, depth, image, nears, fars = inf.render(rays_o.cuda().float(), rays_d.cuda().float(), perturb=perturb)
        weights_sum, depth, image = weights_sum.clone(), depth.clone(), image.clone()
        bgc = self._flag('bg_color')
        if bgc:
            bg = torch.tensor([list(bgc)], dtype=torch.float32, device=image.device)
            image = image + (1 - weights_sum).unsqueeze(-1) * bg
        # (the ``bg_color`` argument is accepted and ignored, as in the reference: its blend is commented out at
        #  renderer.py:700, only ``opt.bg_color`` is honoured -- the editing trainer passes a random colour here,
        #  utils_init_nerf.py:358-365, which must not reach the image)
        return {'image': image.view(*prefix, 3), 'depth': depth.view(*prefix), 'weights_sum': weights_sum.reshape(*prefix),
                'mask': (nears < fars).reshape(*prefix)}

    def _occ_density_into(self, tmp_grid, decay, S):
        """Occupancy update, density query (renderer.py:1667-1698) as ONE kernel: cell -> jittered position -> hash-grid gather ->
        trunk + density head -> tmp_grid[cas, morton] (csrc/field_fused.cu: nb200_occ_density; no positions, features or
        activations in HBM).  Same jitter draws and the same arithmetic as the op-by-op path of NeRFRenderer (bit-identical
        grid under autocast, tests/test_gpu_render.py); taken when the field is the stock fused one and autocast is on (the
        fused kernel rounds table entries to fp16 as the autocast encoder does)."""
        stock = ('density' not in self.__dict__ and self.use_fused_field and torch.is_autocast_enabled()
                 and S >= self.grid_size and self.pos_en.input_dim == 3 and self.pos_en.level_dim == 2
                 and self.pos_en.num_levels == 16 and self.pos_en.embeddings.dtype == torch.float32)
        if not stock:
            return super()._occ_density_into(tmp_grid, decay, S)
        import numpy as np
        from .. import _lib as L
        dev = tmp_grid.device
        enc = self.pos_en
        _, xyzs = self._occ_cell_table(dev)
        noise = self.__dict__.get('_occ_noise')
        if noise is None or noise.device != dev or noise.shape[0] != self.cascade:
            noise = self.__dict__['_occ_noise'] = torch.empty(self.cascade, xyzs.shape[0], 3, dtype=torch.float32, device=dev)
        for cas in range(self.cascade):
            noise[cas].copy_(torch.rand_like(xyzs))         # the reference's draws, one per cascade (:1690)
        fwd_img, _ = self._packed.get(self.network.params, self.density_network.params, self.rgb_network.params)
        import os
        chunk = int(os.environ.get("NB200_OCC_CHUNK", self.occ_chunk_rows))
        if chunk > 0:
            # default: encoder + density-only field launches over chunks whose features stay in L2 (measured on B200:
            # profiles/README.md -- the standalone encoder runs at full occupancy, the one-kernel form has 8 gather warps
            # per SM); bit-identical to the one-kernel form
            chunk = (chunk + 127) // 128 * 128
            buf = self.__dict__.get('_occ_chunk_buf')
            if buf is None or buf[0].device != dev or buf[0].shape[0] != chunk:
                buf = self.__dict__['_occ_chunk_buf'] = (torch.empty(chunk, 3, dtype=torch.float32, device=dev),
                                                         torch.empty(chunk, 32, dtype=torch.float16, device=dev))
            L.check(L.lib().nb200_occ_density_chunked(L.ptr(xyzs), L.ptr(noise), L.u32(self.grid_size), L.u32(self.cascade),
                                                      L.f32(float(self.bound)), L.ptr(enc.embeddings.detach()), L.ptr(enc.offsets),
                                                      L.u32(enc.num_levels), L.f32(float(np.log2(enc.per_level_scale))),
                                                      L.u32(int(enc.base_resolution)), L.u32(enc.gridtype_id),
                                                      L.i32(int(enc.align_corners)), L.u32(enc.interp_id), L.ptr(fwd_img),
                                                      L.ptr(tmp_grid), L.ptr(buf[0]), L.ptr(buf[1]), L.u32(chunk), L.stream()),
                    "occ_density_chunked")
            L.LAUNCHES += 3 * ((self.cascade * self.grid_size ** 3 + chunk - 1) // chunk)
            return
        L.check(L.lib().nb200_occ_density(L.ptr(xyzs), L.ptr(noise), L.u32(self.grid_size), L.u32(self.cascade),
                                          L.f32(float(self.bound)), L.ptr(enc.embeddings.detach()), L.ptr(enc.offsets),
                                          L.u32(enc.num_levels), L.f32(float(np.log2(enc.per_level_scale))),
                                          L.u32(int(enc.base_resolution)), L.u32(enc.gridtype_id),
                                          L.i32(int(enc.align_corners)), L.u32(enc.interp_id), L.ptr(fwd_img),
                                          L.ptr(tmp_grid), L.stream()), "occ_density")

    def get_params(self, lr):
        return [
            {'params': self.pos_en.parameters(), 'lr': lr * 10},
            {'params': self.network.parameters(), 'lr': lr},
            {'params': self.density_network.parameters(), 'lr': lr},
            {'params': self.rgb_network.parameters(), 'lr': lr},
        ]
