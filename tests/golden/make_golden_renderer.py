"""Mint golden vectors for the pure-PyTorch compositing of the reference's renderer from its UNMODIFIED code:
``sample_pdf`` (/root/reference/nerf/renderer.py:21-55) and ``NeRFRenderer.weights_sum_i`` (:407-474: the dense path's
compositing with the LGIE switches detach_bg / train_conf / detach_mask_from_field / bg_color), run here on the CPU.

The module is imported as ``nerf.renderer`` under a stub package (so that nerf/__init__.py is not executed) with stubs
for its absent top-level imports (trimesh, plyfile, skimage, torchtyping, and the CUDA-only ``raymarching`` extension,
none of which these two functions touch).  ``weights_sum_i`` only reads ``self.opt``: it is called unbound on a namespace.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_renderer.py
    -> tests/golden/ref_renderer.npz   (inputs stored next to outputs and gradients)
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
FLAG_SETS = [dict(train_conf=0.01, detach_bg=True, detach_mask_from_field=False),
             dict(train_conf=0.01, detach_bg=False, detach_mask_from_field=True),
             dict(train_conf=0, detach_bg=False, detach_mask_from_field=False)]
CALLS = [dict(is_all=True, if_fg=False, bg=False), dict(is_all=False, if_fg=True, bg=True), dict(is_all=False, if_fg=False, bg=False)]


def load_reference():
    for name in ("trimesh", "plyfile", "skimage", "skimage.measure", "raymarching"):
        sys.modules.setdefault(name, types.ModuleType(name))
    tt = types.ModuleType("torchtyping")
    tt.TensorType = type("TensorType", (), {"__class_getitem__": classmethod(lambda cls, item: cls)})
    sys.modules.setdefault("torchtyping", tt)
    pkg = types.ModuleType("nerf")
    pkg.__path__ = ["/root/reference/nerf"]
    sys.modules["nerf"] = pkg
    return importlib.import_module("nerf.renderer")


def inputs(rng, N=6, T=12):
    z = np.sort(rng.uniform(0.3, 2.5, (N, T)).astype(np.float32), axis=-1)
    nears, fars = z[:, :1] - 0.05, z[:, -1:] + 0.1
    nears[1], fars[1] = 1.0, 0.5                                   # a ray that misses the box: mask = False
    return dict(z_vals=z, nears=nears.astype(np.float32), fars=fars.astype(np.float32),
                sample_dist=((fars - nears) / T).astype(np.float32),
                sigmas=rng.uniform(0, 40, (N, T, 1)).astype(np.float32), rgbs=rng.uniform(0, 1, (N, T, 3)).astype(np.float32),
                masks=np.clip(rng.normal(0.5, 0.2, (N, T, 1)), 0, 1).astype(np.float32),
                bg_color=rng.uniform(0, 1, (N, 3)).astype(np.float32),
                g_image=rng.randn(N, 3).astype(np.float32), g_mask=rng.randn(N, 1).astype(np.float32),
                g_ws=rng.randn(N).astype(np.float32), g_depth=rng.randn(N).astype(np.float32))


def main():
    ref = load_reference()
    rng = np.random.RandomState(20261017)
    G = {}
    # ---- sample_pdf, deterministic branch (the random branch draws torch.rand with the same call, :37)
    bins = np.sort(rng.uniform(0.2, 3.0, (5, 9)).astype(np.float32), axis=-1)
    w = rng.uniform(0, 1, (5, 8)).astype(np.float32)
    w[2] = 0.0                                                      # an empty ray: uniform pdf from the 1e-5 floor
    w[3, :6] = 0.0                                                  # mass in the last two bins only
    G["pdf_bins"], G["pdf_weights"] = bins, w
    G["pdf_samples_det16"] = ref.sample_pdf(torch.from_numpy(bins), torch.from_numpy(w), 16, det=True).numpy()
    torch.manual_seed(7)
    G["pdf_samples_rand16_seed7"] = ref.sample_pdf(torch.from_numpy(bins), torch.from_numpy(w), 16, det=False).numpy()
    # ---- weights_sum_i
    inp = inputs(rng)
    for k, v in inp.items():
        G["ws_" + k] = v
    for fi, flags in enumerate(FLAG_SETS):
        opt = types.SimpleNamespace(**flags)
        for ci, call in enumerate(CALLS):
            t = {k: torch.from_numpy(v.copy()) for k, v in inp.items()}
            for k in ("sigmas", "rgbs", "masks"):
                t[k].requires_grad_()
            res = ref.NeRFRenderer.weights_sum_i(types.SimpleNamespace(opt=opt), t["sample_dist"], t["sigmas"], None, None, None,
                                                 t["z_vals"], t["nears"], t["fars"], t["rgbs"], (t["z_vals"].shape[0],),
                                                 masks=t["masks"], bg_color=t["bg_color"] if call["bg"] else None,
                                                 if_fg=call["if_fg"], is_all=call["is_all"])
            loss = (res["image"] * t["g_image"]).sum() + (res["weights_sum"] * t["g_ws"]).sum() + (res["depth"] * t["g_depth"]).sum()
            if "render_mask" in res:
                loss = loss + (res["render_mask"] * t["g_mask"]).sum()
            loss.backward()
            tag = "ws_f%d_c%d_" % (fi, ci)
            for key in ("image", "depth", "weights_sum", "weights", "mask", "render_mask", "black_image"):
                if key in res:
                    G[tag + key] = res[key].detach().numpy()
            for key in ("sigmas", "rgbs", "masks"):
                g = t[key].grad
                G[tag + "grad_" + key] = (g if g is not None else torch.zeros_like(t[key])).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_renderer.npz"), **G)
    print("wrote ref_renderer.npz with", len(G), "arrays")


if __name__ == "__main__":
    main()
