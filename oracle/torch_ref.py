"""fp32 PyTorch-on-CPU restatement of the reference's Python-level hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Covers what the reference does in Python
above its two CUDA extensions, with every native op replaced by its CPU restatement:

  * grid encoder as torch ops with autograd        gridencoder/src/gridencoder.cu:50-196,270-338, grid.py:27-168
  * tcnn-shaped bias-free MLP (flat ``params``)     nerf/network_grid.py:98-139   [tcnn itself is un-vendored:
                                                    parity unpinned, SURVEY.md section 8(c)]
  * frequency embedder, trunc_exp, field network    nerf/base.py:42-77, provider_utils.py:16-29, network_grid.py:150-193
  * dense renderer run / weights_sum_i / sample_pdf nerf/renderer.py:21-55,278-474   (the "non-cuda_ray PyTorch path":
                                                    this is also the CPU baseline bench.py times)
  * occupancy renderer run_cuda (train + infer)     nerf/renderer.py:476-718 (following run_cuda2 where run_cuda is
                                                    broken as shipped)
  * update_extra_state                              nerf/renderer.py:1658-1715
"""
import math
import types

import numpy as np
import torch
import torch.nn as nn

from . import cpu_ops


# =========================================================================== grid encoder
_PRIMES = [1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737]
_M32 = 0xFFFFFFFF


def level_scale_f32(level, S, H):
    """scale = exp2f(level * S) * H - 1.0f in fp32 (gridencoder.cu:138)."""
    e = np.exp2(np.float32(np.float32(level) * np.float32(S)), dtype=np.float32)
    # e * H - 1 is exact in fp64 (24-bit x small int), so one rounding to fp32 == the device FFMA
    return np.float32(np.float64(e) * np.float64(H) - 1.0)


def _grid_index(pg, D, gridtype, align_corners, hashmap_size, resolution):
    """get_grid_index, gridencoder.cu:66-84; pg: list of D int64 tensors (uint32 values)."""
    stride, index = 1, torch.zeros_like(pg[0])
    for d in range(D):
        if stride > hashmap_size:
            break
        index = (index + pg[d] * stride) & _M32
        stride = (stride * (resolution if align_corners else resolution + 1)) & _M32
    if gridtype == 0 and stride > hashmap_size:
        index = torch.zeros_like(pg[0])
        for d in range(D):
            index = index ^ ((pg[d] * _PRIMES[d]) & _M32)
    return index % hashmap_size


def grid_encode(x01, embeddings, offsets, per_level_scale, base_resolution, gridtype=0, align_corners=False,
                interpolation=0, max_level=None, scales=None):
    """Differentiable (w.r.t. embeddings) restatement of grid_encode; returns [B, L*C] (level-major, channel-minor)."""
    B, D = x01.shape
    offs = [int(v) for v in offsets]
    L = len(offs) - 1
    Cc = embeddings.shape[1]
    S = np.float32(np.log2(per_level_scale))
    max_level = L if max_level is None else min(max_level, L)
    x01 = x01.float()
    oob = ((x01 < 0) | (x01 > 1)).any(-1)
    outs = []
    for l in range(L):
        if l >= max_level:
            outs.append(torch.zeros(B, Cc, dtype=embeddings.dtype))
            continue
        hashmap_size = offs[l + 1] - offs[l]
        scale = float(level_scale_f32(l, S, base_resolution)) if scales is None else float(scales[l])
        resolution = int(math.ceil(scale)) + 1
        # fp64 product + add, rounded once == the device FFMA (x * scale + 0.5f), gridencoder.cu:148
        pos = (x01.double() * scale + (0.0 if align_corners else 0.5)).float()
        pgf = torch.floor(pos)
        frac = pos - pgf
        if interpolation == 1:
            frac = frac * frac * (3.0 - 2.0 * frac)
        pg = [pgf[:, d].long() & _M32 for d in range(D)]
        acc = torch.zeros(B, Cc, dtype=embeddings.dtype)
        for idx in range(1 << D):
            w = torch.ones(B)
            pgl = []
            for d in range(D):
                if idx & (1 << d):
                    w = w * frac[:, d]
                    pgl.append((pg[d] + 1) & _M32)
                else:
                    w = w * (1 - frac[:, d])
                    pgl.append(pg[d])
            index = _grid_index(pgl, D, gridtype, align_corners, hashmap_size, resolution) + offs[l]
            acc = acc + w.unsqueeze(-1) * embeddings[index]
        acc = torch.where(oob.unsqueeze(-1), torch.zeros_like(acc), acc)
        outs.append(acc)
    return torch.stack(outs, 1).reshape(B, L * Cc)


class GridEncoder(nn.Module):
    """gridencoder/grid.py:102-168 on CPU."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 interpolation='linear'):
        super().__init__()
        offsets, per_level_scale = cpu_ops.grid_offsets(input_dim, num_levels, level_dim, per_level_scale,
                                                        base_resolution, log2_hashmap_size, desired_resolution,
                                                        align_corners)
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution = per_level_scale, base_resolution
        self.log2_hashmap_size = log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.gridtype, self.gridtype_id = gridtype, {'hash': 0, 'tiled': 1}[gridtype]
        self.interpolation, self.interp_id = interpolation, {'linear': 0, 'smoothstep': 1}[interpolation]
        self.align_corners = align_corners
        self.register_buffer('offsets', torch.from_numpy(offsets))
        self.n_params = int(offsets[-1]) * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def forward(self, inputs, bound=1, max_level=None):
        inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        out = grid_encode(inputs.reshape(-1, self.input_dim), self.embeddings, self.offsets.tolist(),
                          self.per_level_scale, self.base_resolution, self.gridtype_id, self.align_corners,
                          self.interp_id, max_level)
        return out.view(prefix + [self.output_dim])


# =========================================================================== tcnn-shaped MLP
def _pad16(n):
    return (n + 15) // 16 * 16


def mlp_layer_shapes(n_in, n_out, n_neurons=64, n_hidden_layers=1):
    """[(out_padded, in_padded)] per weight matrix; tcnn FullyFusedMLP layout [upstream, unverified]."""
    dims = [_pad16(n_in)] + [n_neurons] * n_hidden_layers + [_pad16(n_out)]
    return [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


def mlp_init(n_in, n_out, n_neurons=64, n_hidden_layers=1, seed=1337):
    """Xavier-uniform flat fp32 params (tcnn default initialisation, seed 1337)."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for (o, i) in mlp_layer_shapes(n_in, n_out, n_neurons, n_hidden_layers):
        s = math.sqrt(6.0 / (i + o))
        parts.append(((torch.rand(o, i, generator=g) * 2 - 1) * s).reshape(-1))
    return torch.cat(parts)


def mlp_forward(x, params, n_in, n_out, n_neurons=64, n_hidden_layers=1, output_activation='None', half=False):
    """Bias-free MLP: hidden ReLU, output None|Sigmoid.  Input lanes [n_in, pad16(n_in)) are fed 1.0
    (tcnn convention [upstream, unverified]); output sliced to n_out.  ``half`` emulates the fp16-operand /
    fp32-accumulate contract of the CUDA kernel (weights and inter-layer activations rounded to fp16)."""
    shapes = mlp_layer_shapes(n_in, n_out, n_neurons, n_hidden_layers)
    q = (lambda t: t.half().float()) if half else (lambda t: t)
    h = x.float()
    pad = shapes[0][1] - n_in
    if pad:
        h = torch.cat([h, torch.ones(h.shape[0], pad)], -1)
    h = q(h)
    off = 0
    for li, (o, i) in enumerate(shapes):
        W = q(params[off:off + o * i].view(o, i))
        off += o * i
        h = h @ W.t()
        if li < len(shapes) - 1:
            h = q(torch.relu(h))
    h = h[:, :n_out]
    if output_activation == 'Sigmoid':
        h = torch.sigmoid(h)
    return h


class Network(nn.Module):
    """Stand-in for tcnn.Network(n_in, n_out, {... FullyFusedMLP ...}) (nerf/network_grid.py:98-139)."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.n_neurons = network_config.get("n_neurons", 64)
        self.n_hidden_layers = network_config.get("n_hidden_layers", 1)
        self.output_activation = network_config.get("output_activation", "None")
        self.half = False
        self.params = nn.Parameter(mlp_init(n_input_dims, n_output_dims, self.n_neurons, self.n_hidden_layers, seed))

    def forward(self, x):
        return mlp_forward(x, self.params, self.n_input_dims, self.n_output_dims, self.n_neurons,
                           self.n_hidden_layers, self.output_activation, self.half)


# =========================================================================== field network
class _trunc_exp(torch.autograd.Function):
    """nerf/provider_utils.py:16-29."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply


def freq_embed(d, multires=4):
    """get_embedder(4): [d, sin(d*2^k), cos(d*2^k)]_{k<4} -> 27 dims (nerf/base.py:42-77)."""
    out = [d]
    for k in range(multires):
        f = 2.0 ** k
        out += [torch.sin(d * f), torch.cos(d * f)]
    return torch.cat(out, -1)


def default_opt(**kw):
    """The hot-path-relevant flags of main.py with their defaults."""
    o = dict(bound=2, min_near=0.01, density_thresh=10, max_steps=1024, num_steps=64, upsample_steps=64,
             train_conf=0.01, conf_thr=0.5, soft_mask=False, detach_bg=False, detach_mask_from_field=False,
             mask_no_dir=False, cuda_ray=False, bg_color=None, backbone='grid')
    o.update(kw)
    return types.SimpleNamespace(**o)


class NeRFNetwork(nn.Module):
    """nerf/network_grid.py:70-206 + nerf/renderer.py NeRFRenderer, on CPU in fp32."""

    def __init__(self, opt, encoder_kwargs=None, seed=1337):
        super().__init__()
        self.opt = opt
        self.bound = opt.bound
        self.cascade = 1 + math.ceil(math.log2(opt.bound))
        self.grid_size = 128
        self.cuda_ray = opt.cuda_ray
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh
        aabb = torch.tensor([-opt.bound] * 3 + [opt.bound] * 3, dtype=torch.float32)
        self.register_buffer('aabb_train', aabb)
        self.register_buffer('aabb_infer', aabb.clone())
        if self.cuda_ray:
            self.register_buffer('density_grid', torch.zeros(self.cascade, self.grid_size ** 3))
            self.register_buffer('density_bitfield', torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
            self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))
            self.mean_density, self.iter_density, self.mean_count, self.local_step = 0, 0, 0, 0
        ek = dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=21,
                  desired_resolution=8192, gridtype='tiled')           # network_grid.py:89-96
        ek.update(encoder_kwargs or {})
        self.pos_en = GridEncoder(**ek)
        self.pos_en_dim = self.pos_en.output_dim
        cfg = lambda act, nh: {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": act,
                               "n_neurons": 64, "n_hidden_layers": nh}
        self.network = Network(self.pos_en_dim, 64, cfg("None", 2), seed)
        self.density_network = Network(64, 1, cfg("None", 1), seed + 1)
        self.n_rgb_out = 3 + (1 if opt.train_conf else 0)
        self.rgb_network = Network(27 + 64, self.n_rgb_out, cfg("Sigmoid", 1), seed + 2)

    def gaussian(self, x):
        d = (x ** 2).sum(-1)
        return 5 * torch.exp(-d / (2 * 0.2 ** 2))

    def forward(self, x, d):
        x_en = self.pos_en(x, bound=self.opt.bound)
        fea = self.network(x_en)
        sigma = self.density_network(fea)
        sigma = trunc_exp(sigma.squeeze(-1) + self.gaussian(x))
        view_en = freq_embed(d)
        radiances = self.rgb_network(torch.cat([view_en, fea], dim=-1))
        return sigma, radiances, None

    def density(self, x):
        x_en = self.pos_en(x, bound=self.opt.bound)
        fea = self.network(x_en)
        sigma = self.density_network(fea)
        return {'sigma': trunc_exp(sigma.squeeze(-1) + self.gaussian(x))}

    def get_params(self, lr):
        return [{'params': self.pos_en.parameters(), 'lr': lr * 10},
                {'params': self.network.parameters(), 'lr': lr},
                {'params': self.density_network.parameters(), 'lr': lr},
                {'params': self.rgb_network.parameters(), 'lr': lr}]

    # ----------------------------------------------------------------- dense path (renderer.py:278-474)
    def run(self, rays_o, rays_d, num_steps=128, upsample_steps=128, bg_color=None, perturb=False, **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = cpu_ops.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb.numpy(), self.min_near)
        nears = torch.from_numpy(nears).unsqueeze(-1)
        fars = torch.from_numpy(fars).unsqueeze(-1)
        z_vals = torch.linspace(0.0, 1.0, num_steps).unsqueeze(0).expand((N, num_steps))
        z_vals = nears + (fars - nears) * z_vals
        sample_dist = (fars - nears) / num_steps
        if perturb:
            z_vals = z_vals + (torch.rand(z_vals.shape) - 0.5) * sample_dist
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, aabb[:3]), aabb[3:])
        density_outputs = self.density(xyzs.reshape(-1, 3))
        for k, v in density_outputs.items():
            density_outputs[k] = v.view(N, num_steps, -1)
        weights = None
        if upsample_steps > 0:
            with torch.no_grad():
                deltas = z_vals[..., 1:] - z_vals[..., :-1]
                deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
                alphas = 1 - torch.exp(-deltas * density_outputs['sigma'].squeeze(-1))
                alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
                weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
                z_vals_mid = (z_vals[..., :-1] + 0.5 * deltas[..., :-1])
                new_z_vals = sample_pdf(z_vals_mid, weights[:, 1:-1], upsample_steps, det=not self.training).detach()
                new_xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * new_z_vals.unsqueeze(-1)
                new_xyzs = torch.min(torch.max(new_xyzs, aabb[:3]), aabb[3:])
            new_density_outputs = self.density(new_xyzs.reshape(-1, 3))
            for k, v in new_density_outputs.items():
                new_density_outputs[k] = v.view(N, upsample_steps, -1)
            z_vals = torch.cat([z_vals, new_z_vals], dim=1)
            z_vals, z_index = torch.sort(z_vals, dim=1)
            xyzs = torch.cat([xyzs, new_xyzs], dim=1)
            xyzs = torch.gather(xyzs, dim=1, index=z_index.unsqueeze(-1).expand_as(xyzs))
            for k in density_outputs:
                tmp_output = torch.cat([density_outputs[k], new_density_outputs[k]], dim=1)
                density_outputs[k] = torch.gather(tmp_output, dim=1, index=z_index.unsqueeze(-1).expand_as(tmp_output))
        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        sigmas, rgbs, normals = self(xyzs.reshape(-1, 3), dirs.reshape(-1, 3))
        if rgbs.shape[-1] > 3:
            n_dim = rgbs.shape[-1] - 3
            rgbs, masks = rgbs.split([3, n_dim], dim=-1)
            masks = masks.view(N, -1, n_dim)
        else:
            masks = None
        sigmas = sigmas.view(N, -1, 1)
        rgbs = rgbs.view(N, -1, 3)
        if self.opt.train_conf:
            results = self.weights_sum_i(sample_dist, sigmas, z_vals, nears, fars, rgbs, prefix, masks=masks, is_all=True)
            if self.opt.soft_mask:
                edit_mask = torch.sigmoid((masks - self.opt.conf_thr) * 100)
                sigmas_fg = sigmas.clone() * edit_mask
                sigmas_bg = sigmas.clone() * (1 - edit_mask)
            else:
                edit_mask = masks > 0.5
                sigmas_bg = sigmas.clone()
                sigmas_bg[edit_mask] = 0
                sigmas_fg = sigmas.clone()
                sigmas_fg[~edit_mask] = 0
            results['sigma'] = sigmas
            results['rgbs'] = rgbs
            results['edit_mask'] = edit_mask
            results['fg'] = self.weights_sum_i(sample_dist, sigmas_fg, z_vals, nears, fars, rgbs, prefix, masks=masks, if_fg=True)
            results['bg'] = self.weights_sum_i(sample_dist, sigmas_bg, z_vals, nears, fars, rgbs, prefix, masks=masks)
        else:
            # the reference returns an empty dict here (renderer.py:383,405); composite anyway so the
            # restatement is usable without the mask head
            results = self.weights_sum_i(sample_dist, sigmas, z_vals, nears, fars, rgbs, prefix, masks=None, is_all=True)
        return results

    def weights_sum_i(self, sample_dist, sigmas, z_vals, nears, fars, rgbs, prefix, masks=None, bg_color=None,
                      if_fg=False, is_all=False):
        if is_all and self.opt.detach_bg and masks is not None:
            edit_points = masks.mean(-1, keepdims=True) >= 0.5
            sigmas = torch.where(edit_points, sigmas, sigmas.detach())
            rgbs = torch.where(edit_points, rgbs, rgbs.detach())
        deltas = z_vals[..., 1:] - z_vals[..., :-1]
        deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
        alphas = 1 - torch.exp(-deltas * sigmas.squeeze(-1))
        alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
        results = {}
        weights_sum = weights.sum(dim=-1)
        ori_z_vals = ((z_vals - nears) / (fars - nears)).clamp(0, 1)
        depth = torch.sum(weights * ori_z_vals, dim=-1)
        image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2)
        image = image.view(*prefix, 3)
        depth = depth.view(*prefix)
        if if_fg and bg_color is not None:
            results['black_image'] = image
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        mask = (nears < fars).reshape(*prefix)
        results['image'] = image
        if self.opt.train_conf and masks is not None:
            if self.opt.detach_mask_from_field:
                render_mask = torch.sum(weights.unsqueeze(-1).detach() * masks, dim=-2)
            else:
                render_mask = torch.sum(weights.unsqueeze(-1) * masks, dim=-2)
            results['render_mask'] = render_mask.view(*prefix, -1)
        results['depth'] = depth
        results['weights_sum'] = weights_sum
        results['weights'] = weights
        results['mask'] = mask
        return results

    # ----------------------------------------------------------------- occupancy path (renderer.py:476-718)
    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024,
                 T_thresh=1e-4, noises=None, **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = cpu_ops.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb.numpy())   # min_near default 0.2 (B5)
        bf = self.density_bitfield.numpy()
        results = {}
        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            if perturb and noises is None:
                noises = torch.rand(N).numpy()
            xyzs, dirs, deltas, rays = cpu_ops.march_rays_train(
                rays_o.numpy(), rays_d.numpy(), self.bound, bf, self.cascade, self.grid_size, nears, fars,
                counter.numpy(), self.mean_count, noises if perturb else None, 128, force_all_rays, dt_gamma, max_steps)
            xyzs, dirs, deltas, rays = map(torch.from_numpy, (xyzs, dirs, deltas, rays))
            sigmas, rgbs, _ = self(xyzs, dirs)
            results['rgba_samples'] = rgbs
            rgbs = rgbs[..., :3]                                                    # run_cuda2:510 (B3)
            weights_sum, depth, image = composite_rays_train(sigmas, rgbs.float().contiguous(), deltas, rays, T_thresh)
            results['march'] = (xyzs, dirs, deltas, rays)
        else:
            weights_sum = np.zeros(N, np.float32)
            depth = np.zeros(N, np.float32)
            image = np.zeros((N, 3), np.float32)                                    # run_cuda2:535 (B4)
            rays_alive = np.arange(N, dtype=np.int32)
            rays_t = nears.copy()
            step = 0
            while step < max_steps:
                n_alive = rays_alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                nz = None
                if perturb and step == 0:
                    nz = torch.rand(n_alive).numpy() if noises is None else noises
                xyzs, dirs, deltas = cpu_ops.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o.numpy(), rays_d.numpy(),
                                                        self.bound, bf, self.cascade, self.grid_size, nears, fars, 128,
                                                        nz, dt_gamma, max_steps)
                with torch.no_grad():
                    sigmas, rgbs, _ = self(torch.from_numpy(xyzs), torch.from_numpy(dirs))
                rgbs = rgbs[..., :3]
                cpu_ops.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas.numpy(), rgbs.contiguous().numpy(),
                                       deltas, weights_sum, depth, image, T_thresh)
                rays_alive = np.ascontiguousarray(rays_alive[rays_alive >= 0])
                step += n_step
            weights_sum, depth, image = map(torch.from_numpy, (weights_sum, depth, image))
        if bg_color is not None:
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results['image'] = image.view(*prefix, 3)
        results['depth'] = depth.view(*prefix)
        results['weights_sum'] = weights_sum.reshape(*prefix)
        results['mask'] = torch.from_numpy(nears < fars).reshape(*prefix)
        return results

    # ----------------------------------------------------------------- occupancy grid (renderer.py:1658-1715)
    @torch.no_grad()
    def update_extra_state(self, decay=0.95, noise=None):
        """noise: optional [cascade, 128^3, 3] uniform[0,1) replacing torch.rand_like (for parity tests)."""
        if not self.cuda_ray:
            return
        tmp_grid = -torch.ones_like(self.density_grid)
        G = self.grid_size
        ar = torch.arange(G, dtype=torch.int32)
        xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing='ij')
        coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
        indices = torch.from_numpy(cpu_ops.morton3D(coords.numpy())).long()
        xyzs = 2 * coords.float() / (G - 1) - 1
        for cas in range(self.cascade):
            bound = min(2 ** cas, self.bound)
            half_grid_size = bound / G
            cas_xyzs = xyzs * (bound - half_grid_size)
            r = torch.rand_like(cas_xyzs) if noise is None else noise[cas]
            cas_xyzs = cas_xyzs + (r * 2 - 1) * half_grid_size
            sigmas = self.density(cas_xyzs)['sigma'].reshape(-1).detach()
            tmp_grid[cas, indices] = sigmas.float()
        valid_mask = self.density_grid >= 0
        self.density_grid[valid_mask] = torch.maximum(self.density_grid[valid_mask] * decay, tmp_grid[valid_mask])
        self.mean_density = torch.mean(self.density_grid[valid_mask]).item()
        self.iter_density += 1
        density_thresh = min(self.mean_density, self.density_thresh)
        bf = cpu_ops.packbits(self.density_grid.numpy(), density_thresh)
        self.density_bitfield = torch.from_numpy(bf)
        total_step = min(16, self.local_step)
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0

    def render(self, rays_o, rays_d, staged=False, max_ray_batch=2048, **kwargs):
        _run = self.run_cuda if self.cuda_ray else self.run
        return _run(rays_o, rays_d, **kwargs)


def sample_pdf(bins, weights, n_samples, det=False):
    """nerf/renderer.py:21-55."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if det:
        u = torch.linspace(0. + 0.5 / n_samples, 1. - 0.5 / n_samples, steps=n_samples)
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    else:
        u = torch.rand(list(cdf.shape[:-1]) + [n_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    matched_shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(matched_shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(matched_shape), 2, inds_g)
    denom = (cdf_g[..., 1] - cdf_g[..., 0])
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])


class _composite_rays_train(torch.autograd.Function):
    """raymarching/raymarching.py:239-292 over the C restatement."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        ws, depth, image = cpu_ops.composite_rays_train_forward(sigmas.detach().numpy(), rgbs.detach().numpy(),
                                                                deltas.numpy(), rays.numpy(), T_thresh)
        ws, depth, image = map(torch.from_numpy, (ws, depth, image))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, image)
        ctx.T_thresh = T_thresh
        return ws, depth, image

    @staticmethod
    def backward(ctx, grad_ws, grad_depth, grad_image):
        sigmas, rgbs, deltas, rays, ws, image = ctx.saved_tensors
        gs, gc = cpu_ops.composite_rays_train_backward(grad_ws.contiguous().numpy(), grad_image.contiguous().numpy(),
                                                       sigmas.detach().numpy(), rgbs.detach().numpy(), deltas.numpy(),
                                                       rays.numpy(), ws.numpy(), image.numpy(), ctx.T_thresh)
        return torch.from_numpy(gs), torch.from_numpy(gc), None, None, None


composite_rays_train = _composite_rays_train.apply
