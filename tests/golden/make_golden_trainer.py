"""Mint golden values from the reference's editing train step, imported UNMODIFIED on the CPU:

  * Trainer_Nerf.train_step_editing   /root/reference/nerf/utils_init_nerf.py:353-394  (background colour choice, the
        reshapes of the LGIE render outputs, the teacher cache, the ori_bg blend, loss = SDS term + keep_bg * L1)
  * Trainer_Nerf.get_pt               /root/reference/nerf/utils_init_nerf.py:243-265  (teacher render cached per image path)

on canned render results (the renderer itself is pinned elsewhere).  The Stable-Diffusion term (train_step_sd, out of scope
by north_star) is replaced by a stand-in that returns 0.25 * mean(pred_rgb) so that the sum and its gradients are checked.
Run in the build container:  python tests/golden/make_golden_trainer.py  ->  tests/golden/ref_trainer.npz
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden_python import stub_modules  # noqa: E402

B, H, W = 1, 6, 8
VARIANTS = [dict(tag="keep", keep_bg=0.5, ori_bg=False, lambda_sd=1.0, random_bg_c=False, black_bg_c=True, white_bg_c=False),
            dict(tag="rand", keep_bg=2.0, ori_bg=False, lambda_sd=1.0, random_bg_c=True, black_bg_c=False, white_bg_c=False),
            dict(tag="nobg", keep_bg=0, ori_bg=False, lambda_sd=1.0, random_bg_c=False, black_bg_c=False, white_bg_c=True)]


def canned(rng, grad):
    N = H * W

    def t(*shape):
        x = torch.from_numpy(rng.uniform(0, 1, shape).astype(np.float32))
        return x.requires_grad_() if grad else x
    part = lambda: {"image": t(B, N, 3), "depth": t(B, N), "weights_sum": t(N), "render_mask": t(B, N, 1)}  # noqa: E731
    out = part()
    out["fg"], out["bg"] = part(), part()
    return out


def main():
    stub_modules()
    for name in ("imageio", "tensorboardX", "clip"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sd_stub, clip_stub = types.ModuleType("nerf.sd"), types.ModuleType("nerf.clip")
    sd_stub.StableDiffusion, clip_stub.CLIP = object, object
    sys.modules["nerf.sd"], sys.modules["nerf.clip"] = sd_stub, clip_stub
    utils = importlib.import_module("nerf.utils_init_nerf")
    rng = np.random.RandomState(3)
    G = {"meta_BHW": np.array([B, H, W])}
    out = canned(rng, True)
    teacher = canned(rng, False)
    rgbs = torch.from_numpy(rng.uniform(0, 1, (B, H * W, 3)).astype(np.float32))
    for k in ("image", "depth", "weights_sum", "render_mask"):
        G["out_" + k] = out[k].detach().numpy()
        for p in ("fg", "bg"):
            G["out_%s_%s" % (p, k)] = out[p][k].detach().numpy()
            G["teacher_%s_%s" % (p, k)] = teacher[p][k].numpy()
        G["teacher_" + k] = teacher[k].numpy()
    G["rgbs"] = rgbs.numpy()
    rays = torch.zeros(B, H * W, 3)
    for v in VARIANTS:
        opt = types.SimpleNamespace(clip_view=False, **{k: val for k, val in v.items() if k != "tag"})
        calls = {"student": [], "teacher": 0}

        def render(ro, rd, **kw):
            calls["student"].append(kw.get("bg_color"))
            return out

        def teacher_render(ro, rd, **kw):
            calls["teacher"] += 1
            return teacher
        me = types.SimpleNamespace(opt=opt, model=types.SimpleNamespace(render=render),
                                   model_pretrained=types.SimpleNamespace(render=teacher_render), pt_dict={})
        me.get_pt = types.MethodType(utils.Trainer_Nerf.get_pt, me)
        me.train_step_sd = lambda pred_rgb, outputs, B, H, W, img_path, match_probs=None, pose=None, tuning_cls=False: \
            (0.25 * pred_rgb.mean(), {"loss_sd": float(0.25 * pred_rgb.mean())})
        for leaf in (out["image"], out["bg"]["image"], out["render_mask"]):
            leaf.grad = None
        torch.manual_seed(11)
        data = (rgbs, torch.zeros(B, H * W, 1), rays, rays, H, W, "img0")
        keep_cuda = torch.Tensor.cuda               # get_pt moves the cached teacher back with .cuda() (:262): identity here
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            for visit in range(2):                  # the second visit reads the teacher from the cache
                pred_rgb, pred_ws, loss, ld = utils.Trainer_Nerf.train_step_editing(me, data)
        finally:
            torch.Tensor.cuda = keep_cuda
        loss.backward()
        tag = v["tag"]
        G["loss_%s" % tag] = np.float64(loss.item())
        G["loss_bg_%s" % tag] = np.float64(ld.get("loss_bg", -1.0))
        G["pred_rgb_%s" % tag], G["pred_ws_%s" % tag] = pred_rgb.detach().numpy(), pred_ws.detach().numpy()
        G["grad_image_%s" % tag] = out["image"].grad.numpy().copy()
        G["grad_bg_image_%s" % tag] = (out["bg"]["image"].grad if out["bg"]["image"].grad is not None
                                       else torch.zeros_like(out["bg"]["image"])).numpy().copy()
        G["teacher_calls_%s" % tag] = np.int64(calls["teacher"])
        bgc = calls["student"][-1]
        G["bg_color_%s" % tag] = bgc.numpy() if bgc is not None else np.zeros((0, 3), np.float32)
    # --ori_bg (:375-377) multiplies a [B,3,H,W] image by a [B,H,W,1] mask: it only broadcasts for H == 3 (a reference defect,
    # DESIGN.md section 8b B15); recorded here so that the test documents the deviation of the product's (working) branch
    opt = types.SimpleNamespace(clip_view=False, keep_bg=1.0, ori_bg=True, lambda_sd=1.0, random_bg_c=False, black_bg_c=False,
                                white_bg_c=False)
    me = types.SimpleNamespace(opt=opt, model=types.SimpleNamespace(render=lambda *a, **k: out),
                               model_pretrained=types.SimpleNamespace(render=lambda *a, **k: teacher), pt_dict={})
    me.get_pt = types.MethodType(utils.Trainer_Nerf.get_pt, me)
    me.train_step_sd = lambda *a, **k: (torch.zeros(()), {})
    try:
        utils.Trainer_Nerf.train_step_editing(me, (rgbs, None, rays, rays, H, W, "img1"))
        G["ori_bg_runs_in_reference"] = np.int64(1)
    except RuntimeError:
        G["ori_bg_runs_in_reference"] = np.int64(0)
    # ---- RGB_network (nerf/network_grid.py:13-68): separate colour and confidence heads, the confidence head fed DETACHED
    #      inputs (--detach_mask_from_field) or the detached field features only (--mask_no_dir); tcnn.Network served by the
    #      oracle MLP (what is pinned is the wiring: which inputs, which detach, output order, parameter names)
    from oracle import torch_ref
    tc = types.ModuleType("tinycudann")
    tc.Network = torch_ref.Network
    sys.modules["tinycudann"] = tc
    ng = importlib.import_module("nerf.network_grid")
    xin = torch.from_numpy(rng.uniform(-1, 1, (40, 91)).astype(np.float32))
    G["rgbnet_x"] = xin.numpy()
    for tag, o in (("detach", dict(mask_no_dir=False, keyword2=None, mask_no_dir_nodetach=False)),
                   ("nodir", dict(mask_no_dir=True, keyword2=None, mask_no_dir_nodetach=False)),
                   ("nodir_nodetach", dict(mask_no_dir=True, keyword2=None, mask_no_dir_nodetach=True))):
        net = ng.RGB_network(27, opt=types.SimpleNamespace(**o))
        for name in ("rgb_network", "conf_network"):
            m = getattr(net, name)
            w = (rng.uniform(-1, 1, m.params.numel()) * 0.4).astype(np.float32)
            m.params.data.copy_(torch.from_numpy(w))
            G["rgbnet_%s_%s" % (tag, name)] = w
        x = xin.clone().requires_grad_()
        y = net(x)
        G["rgbnet_%s_out" % tag] = y.detach().numpy()
        y[:, 3:].sum().backward()
        G["rgbnet_%s_grad_x_from_conf" % tag] = x.grad.numpy().copy()
        G["rgbnet_%s_keys" % tag] = np.array(sorted(net.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, "ref_trainer.npz"), **G)
    print("wrote ref_trainer.npz", {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in G.items() if "loss" in k or "calls" in k})


if __name__ == "__main__":
    main()
