"""The editing train step's host logic (customnerf_b200/trainer.py: editing_loss, TeacherCache, editing_bg_color) against the
reference's own Trainer_Nerf.train_step_editing / get_pt run on the CPU (tests/golden/ref_trainer.npz, minted by
tests/golden/make_golden_trainer.py from /root/reference/nerf/utils_init_nerf.py:243-265, 353-394)."""
import os
import types

import numpy as np
import torch

from customnerf_b200 import trainer

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_trainer.npz"))
VARIANTS = {"keep": dict(keep_bg=0.5, ori_bg=False, lambda_sd=1.0, random_bg_c=False, black_bg_c=True, white_bg_c=False),
            "rand": dict(keep_bg=2.0, ori_bg=False, lambda_sd=1.0, random_bg_c=True, black_bg_c=False, white_bg_c=False),
            "nobg": dict(keep_bg=0, ori_bg=False, lambda_sd=1.0, random_bg_c=False, black_bg_c=False, white_bg_c=True)}


def _tree(prefix, grad):
    def t(name):
        x = torch.from_numpy(G[prefix + name].copy())
        return x.requires_grad_() if grad else x
    out = {k: t(k) for k in ("image", "depth", "weights_sum", "render_mask")}
    for p in ("fg", "bg"):
        out[p] = {k: t("%s_%s" % (p, k)) for k in ("image", "depth", "weights_sum", "render_mask")}
    return out


def test_editing_step_loss_teacher_cache_and_background_colour_match_the_reference():
    B, H, W = [int(v) for v in G["meta_BHW"]]
    rgbs = torch.from_numpy(G["rgbs"])
    rays = torch.zeros(B, H * W, 3)
    for tag, v in VARIANTS.items():
        opt = types.SimpleNamespace(**v)
        out, teacher_out = _tree("out_", True), _tree("teacher_", False)
        calls = []
        cache = trainer.TeacherCache(lambda ro, rd, **kw: (calls.append(1), teacher_out)[1])
        torch.manual_seed(11)
        for visit in range(2):
            bg = trainer.editing_bg_color(opt, B * H * W, rays.device)          # same generator draw as the reference's (:358)
            teacher = cache.get(rays, rays, "img0", bg, B, H, W, opt)
            pred_rgb, pred_ws, loss, ld = trainer.editing_loss(
                out, rgbs, teacher, opt, B, H, W, guidance_loss=lambda pr, o: (0.25 * pr.mean(), {"loss_sd": float(0.25 * pr.mean())}))
        loss.backward()
        assert len(calls) == int(G["teacher_calls_%s" % tag]) == 1               # rendered once, served from the cache after
        want_bg = G["bg_color_%s" % tag]
        assert (bg is None and want_bg.shape[0] == 0) or np.array_equal(bg.numpy(), want_bg), tag
        np.testing.assert_allclose(loss.item(), float(G["loss_%s" % tag]), rtol=1e-6)
        np.testing.assert_allclose(ld.get("loss_bg", -1.0), float(G["loss_bg_%s" % tag]), rtol=1e-6)
        np.testing.assert_array_equal(pred_rgb.detach().numpy(), G["pred_rgb_%s" % tag])
        np.testing.assert_array_equal(pred_ws.detach().numpy(), G["pred_ws_%s" % tag])
        np.testing.assert_allclose(out["image"].grad.numpy(), G["grad_image_%s" % tag], rtol=1e-6, atol=1e-9)
        gb = out["bg"]["image"].grad
        np.testing.assert_allclose(gb.numpy() if gb is not None else np.zeros_like(G["grad_bg_image_%s" % tag]),
                                   G["grad_bg_image_%s" % tag], rtol=1e-6, atol=1e-9)


def test_ori_bg_branch_works_where_the_reference_cannot_broadcast():
    """--ori_bg multiplies a [B,3,H,W] image by a [B,H,W,1] mask in the reference (utils_init_nerf.py:375-377): a shape error for
    every H != 3 (recorded by the golden script).  The product applies the evident intent: pixels neither the teacher nor the
    current render marks as edited keep the ground-truth colour as background target."""
    assert int(G["ori_bg_runs_in_reference"]) == 0
    B, H, W = [int(v) for v in G["meta_BHW"]]
    rgbs = torch.from_numpy(G["rgbs"])
    out, teacher_out = _tree("out_", True), _tree("teacher_", False)
    opt = types.SimpleNamespace(keep_bg=1.0, ori_bg=True, lambda_sd=0)
    teacher = trainer.TeacherCache(lambda ro, rd, **kw: teacher_out).get(torch.zeros(B, H * W, 3), torch.zeros(B, H * W, 3), "x", None, B, H, W, opt)
    _, _, loss, ld = trainer.editing_loss(out, rgbs, teacher, opt, B, H, W)
    pt_mask = teacher_out["render_mask"].reshape(B, H, W, 1)
    non_edit = ((pt_mask + out["render_mask"].detach().reshape(B, H, W, 1)) < 0.5).permute(0, 3, 1, 2)
    tgt = torch.where(non_edit, rgbs.reshape(B, H, W, 3).permute(0, 3, 1, 2), teacher_out["bg"]["image"].reshape(B, H, W, 3).permute(0, 3, 1, 2))
    want = (tgt - out["bg"]["image"].detach().reshape(B, H, W, 3).permute(0, 3, 1, 2)).abs().mean()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=1e-6)
