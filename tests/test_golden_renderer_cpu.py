"""The dense renderer's compositing -- oracle (oracle/torch_ref.py) AND product (customnerf_b200/nerf/rendering.py, plain
torch ops, so it runs on the CPU) -- against golden vectors minted from the reference's own ``sample_pdf`` and
``NeRFRenderer.weights_sum_i`` (tests/golden/make_golden_renderer.py imports /root/reference/nerf/renderer.py unmodified).

Tolerance: fp32 rel 1e-5 / abs 1e-6 on values and gradients (same torch ops in a slightly different arrangement:
torch.where instead of masked assignment for detach_bg)."""
import os
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_renderer.npz"))
FLAG_SETS = [dict(train_conf=0.01, detach_bg=True, detach_mask_from_field=False),
             dict(train_conf=0.01, detach_bg=False, detach_mask_from_field=True),
             dict(train_conf=0, detach_bg=False, detach_mask_from_field=False)]
CALLS = [dict(is_all=True, if_fg=False, bg=False), dict(is_all=False, if_fg=True, bg=True), dict(is_all=False, if_fg=False, bg=False)]


def _close(got, want, what):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.abs(got - want)
    assert (err <= 1e-6 + 1e-5 * np.abs(want)).all(), (what, err.max())


def _impls():
    from oracle import torch_ref
    from customnerf_b200.nerf import rendering

    def oracle_ws(opt, t, **kw):
        return torch_ref.NeRFNetwork.weights_sum_i(types.SimpleNamespace(opt=opt), t["sample_dist"], t["sigmas"], t["z_vals"],
                                                    t["nears"], t["fars"], t["rgbs"], (t["z_vals"].shape[0],), **kw)

    def product_ws(opt, t, **kw):
        me = types.SimpleNamespace(opt=opt, _flag=lambda name, default=None: getattr(opt, name, default))
        return rendering.NeRFRenderer.weights_sum_i(me, t["sigmas"], t["sample_dist"], t["z_vals"], t["nears"], t["fars"],
                                                    t["rgbs"], (t["z_vals"].shape[0],), **kw)
    return (("oracle", torch_ref.sample_pdf, oracle_ws), ("product", rendering.sample_pdf, product_ws))


def test_sample_pdf_matches_the_reference():
    bins, w = torch.from_numpy(G["pdf_bins"]), torch.from_numpy(G["pdf_weights"])
    for name, sample_pdf, _ in _impls():
        _close(sample_pdf(bins, w, 16, det=True).numpy(), G["pdf_samples_det16"], name + " det")
        torch.manual_seed(7)                       # the random branch draws the same torch.rand call (renderer.py:37)
        _close(sample_pdf(bins, w, 16, det=False).numpy(), G["pdf_samples_rand16_seed7"], name + " rand")


def test_weights_sum_i_values_and_gradients_match_the_reference():
    names = ("z_vals", "nears", "fars", "sample_dist", "sigmas", "rgbs", "masks", "bg_color", "g_image", "g_mask", "g_ws", "g_depth")
    for impl, _, ws_i in _impls():
        for fi, flags in enumerate(FLAG_SETS):
            opt = types.SimpleNamespace(**flags)
            for ci, call in enumerate(CALLS):
                t = {k: torch.from_numpy(G["ws_" + k].copy()) for k in names}
                for k in ("sigmas", "rgbs", "masks"):
                    t[k].requires_grad_()
                res = ws_i(opt, t, masks=t["masks"], bg_color=t["bg_color"] if call["bg"] else None, if_fg=call["if_fg"],
                           is_all=call["is_all"])
                tag = "ws_f%d_c%d_" % (fi, ci)
                loss = (res["image"] * t["g_image"]).sum() + (res["weights_sum"] * t["g_ws"]).sum() + (res["depth"] * t["g_depth"]).sum()
                if (tag + "render_mask") in G:
                    loss = loss + (res["render_mask"] * t["g_mask"]).sum()
                else:
                    assert "render_mask" not in res
                loss.backward()
                for key in ("image", "depth", "weights_sum", "weights", "render_mask", "black_image"):
                    if (tag + key) in G:
                        _close(res[key].detach().numpy(), G[tag + key], (impl, tag, key))
                assert np.array_equal(res["mask"].numpy(), G[tag + "mask"])
                for key in ("sigmas", "rgbs", "masks"):
                    g = t[key].grad
                    _close((g if g is not None else torch.zeros_like(t[key])).numpy(), G[tag + "grad_" + key], (impl, tag, "grad", key))
