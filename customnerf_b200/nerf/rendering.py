"""NeRFRenderer: the hot-path half of the reference's ``nerf/renderer.py`` on the B200 ops.

Covered (reference line numbers):
  render()              :1719-1733   dispatch dense / occupancy path
  run()                 :278-405     dense 64+64 sampler with LGIE outputs (fg / bg / soft mask / detach_bg)
  weights_sum_i()       :407-474     torch compositing used by run()
  run_cuda()            :597-718     occupancy-grid renderer (train: march + composite; eval: n_step loop)
                                     following run_cuda2 (:476-595) where run_cuda is broken as shipped
                                     (4-channel rgbs / image)
  update_extra_state()  :1658-1715   occupancy-grid EMA update + packbits
  reset_extra_state()   :262-276

Out of scope (never reachable with --backbone grid): run_sdf, run_composite, sample_pts_*, mesh export.

The reference reads ``opt.bg_color`` which main.py never defines (Appendix B1); a missing attribute is None here.
"""
import math

import torch
import torch.nn as nn

from .. import raymarching


def sample_pdf(bins, weights, n_samples, det=False):
    """Inverse-CDF sampling of ``n_samples`` depths per ray (reference :21-55, from the original NeRF)."""
    weights = weights + 1e-5
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if det:
        u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples, device=weights.device)
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    else:
        u = torch.rand(list(cdf.shape[:-1]) + [n_samples], device=weights.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp_min(0)
    above = inds.clamp_max(cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)


class NeRFRenderer(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.bound = opt.bound
        self.cascade = 1 + math.ceil(math.log2(opt.bound))
        self.grid_size = 128
        self.cuda_ray = opt.cuda_ray
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh

        box = torch.tensor([-opt.bound] * 3 + [opt.bound] * 3, dtype=torch.float32)
        self.register_buffer('aabb_train', box)
        self.register_buffer('aabb_infer', box.clone())
        self.register_buffer('aabb_train_bg', 2 * box)
        self.register_buffer('aabb_infer_bg', 2 * box)

        if self.cuda_ray:   # state-dict keys as in the reference (:224-241)
            self.register_buffer('density_grid', torch.zeros(self.cascade, self.grid_size ** 3))
            self.register_buffer('density_bitfield',
                                 torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
            self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))
            self.mean_density = 0
            self.iter_density = 0
            self.mean_count = 0
            self.local_step = 0

    # the field network provides these
    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    def _flag(self, name, default=None):
        return getattr(self.opt, name, default)

    # ------------------------------------------------------------------------------------- dense path
    def run(self, rays_o, rays_d, num_steps=128, upsample_steps=128, bg_color=None, perturb=False, **kwargs):
        if (getattr(self, 'fused_dense', True) and rays_o.is_cuda and upsample_steps > 0 and num_steps >= 3
                and num_steps + upsample_steps <= 256):
            return self._run_dense_fused(rays_o, rays_d, num_steps, upsample_steps, perturb)
        return self._run_dense_ops(rays_o, rays_d, num_steps, upsample_steps, bg_color, perturb, **kwargs)

    def _run_dense_fused(self, rays_o, rays_d, num_steps, upsample_steps, perturb):
        """The dense renderer (renderer.py:278-405) with its sampler as two kernels (csrc/dense_sampler.cu) and everything
        after the sampler on the occupancy path's code: the merged coarse + importance samples are written as xyzs / dirs /
        deltas / rays, so the field network runs through the fused tcgen05 kernels and the three LGIE composites through the
        composite kernels (T_thresh = 0: the dense formula, no early termination).  The coarse density pass carries no
        autograd graph (nothing downstream of it is differentiated, :328-346) and the reference's second, unused density
        evaluation of the importance samples (:353) is dropped.  Same random draws as the reference, in the same order:
        torch.rand(N, num_steps) for the stratified jitter, torch.rand(N, upsample_steps) inside sample_pdf."""
        from .. import _lib as L
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        N, S, Su = rays_o.shape[0], int(num_steps), int(upsample_steps)
        T = S + Su
        dev = rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        f32 = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), torch.no_grad():
            nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)
            lin = torch.linspace(0.0, 1.0, S, device=dev)
            noise = torch.rand(N, S, device=dev) if perturb else None
            z_c, xyz_c = torch.empty(N, S, **f32), torch.empty(N * S, 3, **f32)
            L.check(L.lib().nb200_dense_coarse(L.ptr(rays_o), L.ptr(rays_d), L.ptr(nears), L.ptr(fars), L.ptr(aabb), L.ptr(lin),
                                               L.ptr(noise), L.u32(N), L.u32(S), L.ptr(z_c), L.ptr(xyz_c), L.stream()), "dense_coarse")
            sigma_c = self.density(xyz_c)['sigma'].reshape(-1).float().contiguous()
            if self.training:
                u, per_ray = torch.rand(N, Su, device=dev), 1
            else:
                u, per_ray = torch.linspace(0.5 / Su, 1.0 - 0.5 / Su, steps=Su, device=dev), 0
            z_all, xyzs, dirs = torch.empty(N, T, **f32), torch.empty(N * T, 3, **f32), torch.empty(N * T, 3, **f32)
            deltas, rays = torch.empty(N * T, 2, **f32), torch.empty(N, 3, dtype=torch.int32, device=dev)
            L.check(L.lib().nb200_dense_importance(L.ptr(rays_o), L.ptr(rays_d), L.ptr(nears), L.ptr(fars), L.ptr(aabb), L.ptr(z_c),
                                                   L.ptr(sigma_c), L.ptr(u.contiguous()), L.i32(per_ray), L.u32(N), L.u32(S), L.u32(Su),
                                                   L.ptr(z_all), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.ptr(rays), L.stream()),
                    "dense_importance")
        sigmas, rgba, _ = self(xyzs, dirs)
        rgbs = rgba[..., :3].float()
        results = {}
        if self._flag('train_conf') and rgba.shape[-1] > 3:
            results.update(self._lgie_composites(sigmas.float(), rgbs, rgba[..., 3:].float(), deltas, rays, 0.0, prefix))
            weights_sum, depth, image = results.pop('_all')
            results['sigma'], results['rgbs'] = sigmas.view(N, T, 1), rgbs.view(N, T, 3)
            results['edit_mask'] = results['edit_mask'].view(N, T, -1)
            for part in ('fg', 'bg'):
                results[part]['weights_sum'] = results[part]['weights_sum'].reshape(-1)
                results[part]['mask'] = (nears < fars).reshape(*prefix)
        else:
            weights_sum, depth, image = raymarching.composite_rays_train(sigmas.float(), rgbs, deltas, rays, 0.0)
        results['image'] = image.view(*prefix, 3)
        results['depth'] = depth.view(*prefix)
        results['weights_sum'] = weights_sum
        results['mask'] = (nears < fars).reshape(*prefix)
        results['z_vals'] = z_all
        return results

    def _run_dense_ops(self, rays_o, rays_d, num_steps=128, upsample_steps=128, bg_color=None, perturb=False, **kwargs):
        """the same renderer as the reference's op-by-op torch sequence (any sample counts; also returns 'weights')"""
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        dev = rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer

        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)
        nears = nears.unsqueeze(-1)
        fars = fars.unsqueeze(-1)

        # stratified coarse samples
        z_vals = nears + (fars - nears) * torch.linspace(0.0, 1.0, num_steps, device=dev).unsqueeze(0)
        sample_dist = (fars - nears) / num_steps
        if perturb:
            z_vals = z_vals + (torch.rand(z_vals.shape, device=dev) - 0.5) * sample_dist
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, aabb[:3]), aabb[3:])
        sigma_c = self.density(xyzs.reshape(-1, 3))['sigma'].view(N, num_steps)

        if upsample_steps > 0:   # importance samples from the coarse weights
            with torch.no_grad():
                deltas = torch.cat([z_vals[..., 1:] - z_vals[..., :-1], sample_dist], dim=-1)
                alphas = 1 - torch.exp(-deltas * sigma_c)
                trans = torch.cumprod(torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1), dim=-1)
                weights = alphas * trans[..., :-1]
                z_mid = z_vals[..., :-1] + 0.5 * deltas[..., :-1]
                new_z = sample_pdf(z_mid, weights[:, 1:-1], upsample_steps, det=not self.training).detach()
                new_xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * new_z.unsqueeze(-1)
                new_xyzs = torch.min(torch.max(new_xyzs, aabb[:3]), aabb[3:])
            # the reference also evaluates density(new_xyzs) here (:353) and merges it into a dict nothing reads;
            # no output depends on it, so it is skipped
            z_vals, order = torch.sort(torch.cat([z_vals, new_z], dim=1), dim=1)
            xyzs = torch.gather(torch.cat([xyzs, new_xyzs], dim=1), 1, order.unsqueeze(-1).expand(-1, -1, 3))

        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        sigmas, rgbs, _ = self(xyzs.reshape(-1, 3), dirs.reshape(-1, 3))
        masks = None
        if rgbs.shape[-1] > 3:
            n_dim = rgbs.shape[-1] - 3
            rgbs, masks = rgbs.split([3, n_dim], dim=-1)
            masks = masks.reshape(N, -1, n_dim)
        sigmas = sigmas.view(N, -1, 1)
        rgbs = rgbs.reshape(N, -1, 3)

        args = (sample_dist, z_vals, nears, fars, rgbs, prefix)
        if self._flag('train_conf') and masks is not None:
            results = self.weights_sum_i(sigmas, *args, masks=masks, is_all=True)
            if self._flag('soft_mask', False):
                edit_mask = torch.sigmoid((masks - self._flag('conf_thr', 0.5)) * 100)
                sigmas_fg = sigmas * edit_mask
                sigmas_bg = sigmas * (1 - edit_mask)
            else:
                edit_mask = masks > 0.5
                sigmas_fg = torch.where(edit_mask, sigmas, torch.zeros_like(sigmas))
                sigmas_bg = torch.where(edit_mask, torch.zeros_like(sigmas), sigmas)
            results['sigma'] = sigmas
            results['rgbs'] = rgbs
            results['edit_mask'] = edit_mask
            results['fg'] = self.weights_sum_i(sigmas_fg, *args, masks=masks, if_fg=True)
            results['bg'] = self.weights_sum_i(sigmas_bg, *args, masks=masks)
        else:
            results = self.weights_sum_i(sigmas, *args, masks=None, is_all=True)
        return results

    def weights_sum_i(self, sigmas, sample_dist, z_vals, nears, fars, rgbs, prefix, masks=None, bg_color=None,
                      if_fg=False, is_all=False):
        if is_all and self._flag('detach_bg', False) and masks is not None:
            # samples the mask head calls background contribute values but no gradient (:409-418)
            edit_points = masks.mean(-1, keepdim=True) >= 0.5
            sigmas = torch.where(edit_points, sigmas, sigmas.detach())
            rgbs = torch.where(edit_points, rgbs, rgbs.detach())
        deltas = torch.cat([z_vals[..., 1:] - z_vals[..., :-1], sample_dist], dim=-1)
        alphas = 1 - torch.exp(-deltas * sigmas.squeeze(-1))
        trans = torch.cumprod(torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1), dim=-1)
        weights = alphas * trans[..., :-1]
        weights_sum = weights.sum(dim=-1)
        ori_z = ((z_vals - nears) / (fars - nears)).clamp(0, 1)
        depth = torch.sum(weights * ori_z, dim=-1)
        image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2).view(*prefix, 3)
        results = {}
        if if_fg and bg_color is not None:
            results['black_image'] = image
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results['image'] = image
        if self._flag('train_conf') and masks is not None:
            w = weights.detach() if self._flag('detach_mask_from_field', False) else weights
            results['render_mask'] = torch.sum(w.unsqueeze(-1) * masks, dim=-2).view(*prefix, -1)
        results['depth'] = depth.view(*prefix)
        results['weights_sum'] = weights_sum
        results['weights'] = weights
        results['mask'] = (nears < fars).reshape(*prefix)
        return results

    # ------------------------------------------------------------------------------------- occupancy path
    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024,
                 T_thresh=1e-4, **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.cuda().contiguous().view(-1, 3)
        rays_d = rays_d.cuda().contiguous().view(-1, 3)
        N = rays_o.shape[0]
        dev = rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb)    # min_near stays 0.2 here (Appendix B5)
        results = {}
        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(
                rays_o, rays_d, self.bound, self.density_bitfield, self.cascade, self.grid_size, nears, fars, counter,
                self.mean_count, perturb, 128, force_all_rays, dt_gamma, max_steps)
            sigmas, rgba, _ = self(xyzs, dirs)
            rgbs = rgba[..., :3].float()
            if self._flag('train_conf') and rgba.shape[-1] > 3:
                results.update(self._lgie_composites(sigmas, rgbs, rgba[..., 3:].float(), deltas, rays, T_thresh, prefix))
                weights_sum, depth, image = results.pop('_all')
            else:
                weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)
        else:
            weights_sum = torch.zeros(N, dtype=torch.float32, device=dev)
            depth = torch.zeros(N, dtype=torch.float32, device=dev)
            image = torch.zeros(N, 3, dtype=torch.float32, device=dev)
            rays_alive = torch.arange(N, dtype=torch.int32, device=dev)
            rays_t = nears.clone()
            step = 0
            while step < max_steps:
                n_alive = rays_alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas = raymarching.march_rays(
                    n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound, self.density_bitfield,
                    self.cascade, self.grid_size, nears, fars, 128, perturb if step == 0 else False, dt_gamma, max_steps)
                sigmas, rgbs, _ = self(xyzs, dirs)
                rgbs = rgbs[..., :3]
                raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum,
                                           depth, image, T_thresh)
                rays_alive = rays_alive[rays_alive >= 0]
                step += n_step

        bgc = self._flag('bg_color')
        if bgc:
            bg = torch.tensor([list(bgc)], dtype=torch.float32, device=dev)
            image = image + (1 - weights_sum).unsqueeze(-1) * bg
        # (the ``bg_color`` argument is accepted and ignored, as in the reference: its blend is commented out at
        #  renderer.py:700, only ``opt.bg_color`` is honoured -- the editing trainer passes a random colour here,
        #  utils_init_nerf.py:358-365, which must not reach the image)
        results['image'] = image.view(*prefix, 3)
        results['depth'] = depth.view(*prefix)
        results['weights_sum'] = weights_sum.reshape(*prefix)
        results['mask'] = (nears < fars).reshape(*prefix)
        return results

    def _lgie_composites(self, sigmas, rgbs, masks, deltas, rays, T_thresh, prefix):
        """LGIE outputs on the occupancy path: the same quantities NeRFRenderer.run builds with weights_sum_i
        (reference :383-403, :407-474) -- all / fg / bg composites over the SAME samples, the rendered mask, soft or
        hard edit mask, detach_bg -- composed from composite_rays_train.  The reference's run_cuda never produces them
        although its trainer requires them; sample-level formulas are the dense path's."""
        comp = raymarching.composite_rays_train
        m1 = masks[..., :1]
        sig_all, rgb_all = sigmas, rgbs
        if self._flag('detach_bg', False):
            # samples the mask head calls background contribute values but no gradient to the "all" image (:409-418)
            edit_points = masks.mean(-1) >= 0.5
            sig_all = torch.where(edit_points, sigmas, sigmas.detach())
            rgb_all = torch.where(edit_points.unsqueeze(-1), rgbs, rgbs.detach())
        out = {'_all': comp(sig_all, rgb_all, deltas, rays, T_thresh)}
        if self._flag('soft_mask', False):
            edit_mask = torch.sigmoid((m1 - self._flag('conf_thr', 0.5)) * 100)
            sig_fg = sigmas * edit_mask.squeeze(-1)
            sig_bg = sigmas * (1 - edit_mask.squeeze(-1))
        else:
            edit_mask = m1 > 0.5
            sig_fg = torch.where(edit_mask.squeeze(-1), sigmas, torch.zeros_like(sigmas))
            sig_bg = torch.where(edit_mask.squeeze(-1), torch.zeros_like(sigmas), sigmas)

        def render_mask(sig):
            # sum_i w_i * mask_i: the compositing kernel with the mask as (replicated) colour; optionally with
            # detached weights (detach_mask_from_field, :460-463)
            s = sig.detach() if self._flag('detach_mask_from_field', False) else sig
            return comp(s, m1.expand(-1, 3).contiguous(), deltas, rays, T_thresh)[2][..., :1]

        def pack(sig, rgb, with_all=None):
            ws, depth, image = with_all if with_all is not None else comp(sig, rgb, deltas, rays, T_thresh)
            return {'image': image.view(*prefix, 3), 'depth': depth.view(*prefix), 'weights_sum': ws.reshape(*prefix),
                    'render_mask': render_mask(sig).view(*prefix, 1)}
        out['render_mask'] = render_mask(sig_all).view(*prefix, 1)
        out['sigma'], out['rgbs'], out['edit_mask'] = sigmas, rgbs, edit_mask
        out['fg'] = pack(sig_fg, rgbs)
        out['bg'] = pack(sig_bg, rgbs)
        return out

    # ------------------------------------------------------------------------------------- occupancy grid
    # mean_density / mean_count are produced on the device by update_extra_state (no blocking read inside the update); the
    # attributes the reference keeps as Python numbers are read back on first use
    def _occ_sync(self):
        ev = self.__dict__.get('_occ_event')
        if ev is not None:
            ev.synchronize()
            h = self._occ_state_host
            self.__dict__['_mean_density'] = float(h[0])
            if self.__dict__.get('_occ_counted'):
                self.__dict__['_mean_count'] = int(h.view(torch.int32)[2])
            self.__dict__['_occ_event'] = None

    @property
    def mean_density(self):
        self._occ_sync()
        return self.__dict__.get('_mean_density', 0)

    @mean_density.setter
    def mean_density(self, v):
        self._occ_sync()
        self.__dict__['_mean_density'] = v

    @property
    def mean_count(self):
        self._occ_sync()
        return self.__dict__.get('_mean_count', 0)

    @mean_count.setter
    def mean_count(self, v):
        self._occ_sync()
        self.__dict__['_mean_count'] = v

    def _occ_cell_table(self, dev):
        """(morton index [G^3] int64, cell centre [G^3, 3] in [-1, 1]) of every cell in the reference's x-major order.  The
        centres are evaluated on the HOST with the reference's expression (renderer.py:1678: 2 * coords.float() / (G - 1) - 1)
        -- on a CUDA tensor the division by a Python scalar becomes a multiplication by the rounded reciprocal, one ulp off
        the reference's values for some cells -- and cached: they never change."""
        G = self.grid_size
        cache = self.__dict__.get('_occ_cells')
        if cache is None or cache[0].device != dev or cache[0].numel() != G ** 3:
            axis = torch.arange(G, dtype=torch.int32)
            xx, yy, zz = torch.meshgrid(axis, axis, axis, indexing='ij')
            coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
            xyzs = (2 * coords.float() / (G - 1) - 1).to(dev)
            indices = raymarching.morton3D(coords.to(dev)).long()
            cache = self.__dict__['_occ_cells'] = (indices, xyzs)
        return cache

    def _occ_buffers(self, dev):
        from .. import _lib as L
        if self.__dict__.get('_occ_state') is None or self._occ_state.device != dev:
            L.lib().nb200_occ_scratch_bytes.restype = L.u32
            self.__dict__['_occ_state'] = torch.zeros(8, dtype=torch.float32, device=dev)
            self.__dict__['_occ_state_host'] = torch.zeros(8, dtype=torch.float32).pin_memory()
            self.__dict__['_occ_scratch'] = torch.empty(int(L.lib().nb200_occ_scratch_bytes()) // 8 + 1, dtype=torch.float64, device=dev)
            self.__dict__['_occ_tmp'] = torch.empty_like(self.density_grid)
        return self._occ_state, self._occ_scratch, self._occ_tmp

    def _occ_density_into(self, tmp_grid, decay, S):
        """density of one jittered point per cell and cascade -> tmp_grid[cas, morton] (renderer.py:1667-1698).  This is the
        reference's torch-op sequence around ``self.density`` (any callable); NeRFNetwork overrides it with the fused kernel."""
        dev = tmp_grid.device
        G = self.grid_size
        tmp_grid.fill_(-1)
        if S >= G:
            chunks = [self._occ_cell_table(dev)]
        else:                                   # smaller chunks (never used by the reference's callers): built on the fly
            axis = torch.arange(G, dtype=torch.int32)
            chunks = []
            for xs in axis.split(S):
                for ys in axis.split(S):
                    for zs in axis.split(S):
                        xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing='ij')
                        coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
                        chunks.append((raymarching.morton3D(coords.to(dev)).long(), (2 * coords.float() / (G - 1) - 1).to(dev)))
        for indices, xyzs in chunks:
            for cas in range(self.cascade):
                bound = min(2 ** cas, self.bound)
                half_grid_size = bound / G
                cas_xyzs = xyzs * (bound - half_grid_size)
                cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half_grid_size
                sigmas = self.density(cas_xyzs)['sigma'].reshape(-1).detach()
                tmp_grid[cas, indices] = sigmas.float()

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """renderer.py:1658-1715.  The density query fills tmp_grid; the EMA-max update, the mean of the valid cells, the
        threshold min(mean, density_thresh), packbits and mean_count run on the device in three launches
        (csrc/occupancy.cu: nb200_occ_finalize) -- the update never blocks on the host; mean_density / mean_count are read
        back when something asks for them."""
        if not self.cuda_ray:
            return
        from .. import _lib as L
        dev = self.density_bitfield.device
        self._occ_sync()                        # a previous update's numbers, before its state buffer is reused
        state, scratch, tmp_grid = self._occ_buffers(dev)
        with torch.cuda.device(dev):
            self._occ_density_into(tmp_grid, decay, S)
            total_step = min(16, self.local_step)
            L.check(L.lib().nb200_occ_finalize(L.ptr(self.density_grid), L.ptr(tmp_grid), L.u32(self.density_grid.numel()),
                                               L.f32(decay), L.f32(float(self.density_thresh)), L.ptr(self.step_counter),
                                               L.u32(total_step), L.ptr(self.density_bitfield), L.ptr(state), L.ptr(scratch),
                                               L.stream()), "occ_finalize")
            L.LAUNCHES += 2
            self._occ_state_host.copy_(state, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self.__dict__['_occ_counted'] = total_step > 0
        self.__dict__['_occ_event'] = ev
        self.iter_density += 1
        self.local_step = 0

    def render(self, rays_o, rays_d, staged=False, max_ray_batch=2048, **kwargs):
        if self.cuda_ray and 'neus' not in str(self._flag('backbone', 'grid')):
            return self.run_cuda(rays_o, rays_d, **kwargs)
        return self.run(rays_o, rays_d, **kwargs)
