"""Reconstruction train step on the occupancy (cuda_ray) path -- the ~100-line harness SURVEY.md section 2.1 #10 asks
for: what ``Trainer_Nerf.train_step_pretrain`` + ``train_one_epoch`` do around ``model.render`` for one batch
(nerf/utils_init_nerf.py:194-241, 599-629), with the reference's optimiser settings (main.py:182,189;
network_grid.py:196-206): Adam(betas=(0.9, 0.99), eps=1e-15), hash table at 10x LR, fp16 autocast.

The reference drives AMP with a dynamic ``GradScaler`` (one D2H sync per step for the inf check); here the loss scale
is the constant 128 tiny-cuda-nn uses internally, applied to the loss and folded back into the gradients.
"""
import torch
import torch.nn.functional as F

from . import raymarching, synthetic
from .nerf import NeRFNetwork

LOSS_SCALE = 128.0


def make_opt(**kw):
    import types
    o = dict(bound=2, min_near=0.01, density_thresh=10, max_steps=1024, num_steps=64, upsample_steps=64, train_conf=0,
             conf_thr=0.5, soft_mask=False, detach_bg=False, detach_mask_from_field=False, mask_no_dir=False,
             cuda_ray=True, bg_color=None, backbone='grid', lr=5e-4, train_rgb=1.0)
    o.update(kw)
    return types.SimpleNamespace(**o)


def build_scene_model(device, log2_hashmap_size=19, desired_resolution=2048, encoding='hashgrid', opt=None, seed=0):
    """Field network + synthetic bear occupancy grid (bitfield through the CUDA packbits)."""
    torch.manual_seed(seed)
    opt = opt or make_opt()
    model = NeRFNetwork(opt, encoding=encoding, log2_hashmap_size=log2_hashmap_size,
                        desired_resolution=desired_resolution).to(device)
    if opt.cuda_ray:
        grid = synthetic.density_grid(opt.bound, model.grid_size, device=device)
        model.density_grid.copy_(grid)
        model.mean_density = float(grid.mean())
        thr = min(model.mean_density, model.density_thresh)
        model.density_bitfield = raymarching.packbits(model.density_grid, thr, model.density_bitfield)
    return model


def pretrain_loss(outputs, rgbs, mask, opt):
    """The reconstruction loss of ``Trainer_Nerf.train_step_pretrain`` (nerf/utils_init_nerf.py:219-239) on a render result:
    train_rgb * MSE(image, rgbs) [+ train_conf * MSE(render_mask, mask) with the mask head].  Returns what the reference's
    step returns after rendering: (pred_rgb, mask_volume, loss, loss_dict); mask_volume is weights_sum clamped to
    [1e-5, 1 - 1e-5].  FusedTrainStep(mask_weight=train_conf) computes the same loss inside the compositing kernel
    (train_rgb is 1 in every shipped configuration, main.py)."""
    pred_rgb = outputs['image']
    B, N = pred_rgb.shape[:2]
    mask_volume = torch.clamp(outputs['weights_sum'].reshape(B, N), 1e-5, 1 - 1e-5)
    loss_c = getattr(opt, 'train_rgb', 1.0) * F.mse_loss(pred_rgb.reshape(-1, 3), rgbs.reshape(-1, 3).float(), reduction='mean')
    loss, loss_dict = loss_c, {'loss_c': loss_c.item()}
    if getattr(opt, 'train_conf', 0):
        loss_m = opt.train_conf * F.mse_loss(outputs['render_mask'].reshape(-1), mask.reshape(-1).float(), reduction='mean')
        loss = loss + loss_m
        loss_dict['loss_m'] = loss_m.item()
    return pred_rgb, mask_volume, loss, loss_dict


class TrainStep:
    def __init__(self, model, lr=5e-4, fp16=True, world_size=1, grad_sync=None, perturb=True):
        self.model = model
        self.fp16 = fp16
        self.perturb = perturb
        self.world_size = world_size
        self.grad_sync = grad_sync            # callable(list of params) -> None, sums gradients across ranks
        self.optimizer = torch.optim.Adam(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15, fused=True)
        self.params = [p for g in self.optimizer.param_groups for p in g['params']]
        model.train()

    def forward_backward(self, rays_o, rays_d, target_rgb, n_total=None):
        """loss = MSE(image, target) over all rays of the (global) batch; returns the detached loss tensor."""
        model = self.model
        for p in self.params:
            p.grad = None
        with torch.autocast('cuda', dtype=torch.float16, enabled=self.fp16):
            out = model.render(rays_o[None], rays_d[None], staged=False, perturb=self.perturb, force_all_rays=True,
                               **vars(model.opt))
            pred = out['image'].reshape(-1, 3)
            n_local = pred.shape[0]
            n_total = n_total or n_local * self.world_size
            # mean over the global batch: local sum / (3 * N_total)
            loss = F.mse_loss(pred, target_rgb.reshape(-1, 3), reduction='sum') / (3.0 * n_total)
        (loss * LOSS_SCALE).backward()
        return loss.detach()

    def step(self, rays_o, rays_d, target_rgb, n_total=None):
        loss = self.forward_backward(rays_o, rays_d, target_rgb, n_total)
        if self.grad_sync is not None:
            self.grad_sync(self.params)
        inv = 1.0 / LOSS_SCALE
        torch._foreach_mul_([p.grad for p in self.params if p.grad is not None], inv)
        self.optimizer.step()
        return loss
