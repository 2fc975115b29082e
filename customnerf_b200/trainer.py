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


# ------------------------------------------------------------------------------------------------ editing step (LGIE)
def editing_bg_color(opt, n, device):
    """the background colour ``train_step_editing`` hands to ``render`` (nerf/utils_init_nerf.py:357-364): one random /
    black / white colour for all ``n`` rays, or None.  (The renderer ignores the argument, as the reference's does.)"""
    if getattr(opt, 'random_bg_c', False):
        return torch.rand((1, 3), device=device).repeat(n, 1)
    if getattr(opt, 'black_bg_c', False):
        return torch.zeros((1, 3), device=device).repeat(n, 1)
    if getattr(opt, 'white_bg_c', False):
        return torch.ones((1, 3), device=device).repeat(n, 1)
    return None


class TeacherCache:
    """``Trainer_Nerf.get_pt`` (nerf/utils_init_nerf.py:243-265): the frozen pre-trained model's render of a training view --
    rendered mask, fg / bg images and the fg depth in [B, C, H, W] layout -- computed once per image path, kept on the
    host, moved back to the device (detached) on every later visit.  ``render`` is the teacher's render callable."""

    def __init__(self, render):
        self.render = render
        self.pt_dict = {}

    def get(self, rays_o, rays_d, img_path, bg_color, B, H, W, opt):
        if img_path not in self.pt_dict:
            out = self.render(rays_o, rays_d, staged=False, perturb=True, bg_color=bg_color, force_all_rays=True, **vars(opt))
            pt_mask = out['render_mask'].reshape(B, H, W, -1).contiguous()
            pt_rgb_bg = out['bg']['image'].reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
            pt_rgb_fg = out['fg']['image'].reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
            pt_depth_fg = out['fg']['depth'].reshape(B, H, W, 1).permute(0, 3, 1, 2).contiguous()
            self.pt_dict[img_path] = (pt_rgb_bg.cpu().detach(), pt_rgb_fg.cpu().detach(), pt_mask.cpu().detach(),
                                      pt_depth_fg.cpu().detach(), None)
        else:
            dev = rays_o.device
            pt_rgb_bg, pt_rgb_fg, pt_mask, pt_depth_fg = [t.detach().to(dev) for t in self.pt_dict[img_path][:-1]]
        return pt_rgb_fg, pt_rgb_bg, pt_mask, pt_depth_fg, None


def editing_loss(outputs, rgbs, teacher, opt, B, H, W, guidance_loss=None):
    """The loss of ``Trainer_Nerf.train_step_editing`` (nerf/utils_init_nerf.py:366-392) on a render result:
        loss = guidance_loss(pred_rgb, outputs)            (the Stable-Diffusion SDS term, out of scope: any callable
                                                            returning (loss, dict); the reference's train_step_sd)
             + keep_bg * L1(teacher background, rendered background)
    with, under ``ori_bg``, the teacher background replaced by the ground-truth pixels wherever neither the teacher's nor the
    current rendered mask marks the pixel as edited ((pt_mask + pred_mask) < 0.5).  ``teacher`` = (pt_rgb_fg, pt_rgb_bg,
    pt_mask, pt_depth_fg, _) as TeacherCache.get returns it.  Returns (pred_rgb, pred_ws, loss, loss_dict)."""
    pred_rgb = outputs['image'].reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
    pred_ws = outputs['weights_sum'].reshape(B, H, W)
    pred_mask = outputs['render_mask'].reshape(B, H, W, -1)
    pred_rgb_bg = outputs['bg']['image'].reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
    pt_rgb_fg, pt_rgb_bg, pt_mask, pt_depth_fg, _ = teacher
    if getattr(opt, 'ori_bg', False):
        # the reference multiplies [B,3,H,W] by this [B,H,W,1] mask as it is (:375-377), which only broadcasts for H == 3;
        # the mask is brought to [B,1,H,W] here
        non_edit = ((pt_mask + pred_mask) < 0.5)[..., :1].permute(0, 3, 1, 2)
        pt_rgb_bg = rgbs.reshape(B, H, W, 3).permute(0, 3, 1, 2) * non_edit + (~non_edit) * pt_rgb_bg
    loss, loss_dict = None, {}
    if getattr(opt, 'lambda_sd', 0) and guidance_loss is not None:
        loss, d = guidance_loss(pred_rgb, outputs)
        loss_dict.update(d)
    if getattr(opt, 'keep_bg', 0):
        loss_bg = opt.keep_bg * F.l1_loss(pt_rgb_bg, pred_rgb_bg)
        loss = loss_bg if loss is None else loss + loss_bg
        loss_dict.update(dict(loss_bg=loss_bg.item()))
    return pred_rgb, pred_ws, loss, loss_dict
