"""-m gpu: the fused step actually TRAINS a scene.  A few hundred graph-replayed steps over rotating synthetic views
(analytic bear colours as targets), with the occupancy grid refreshed from the learned density every 50 steps exactly as
the reference's train loop does (nerf/utils_init_nerf.py:602-607), must cut the held-out image error by a large factor."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_training_converges_with_occupancy_updates():
    from customnerf_b200 import trainer, fused_trainer, synthetic as syn
    dev = torch.device("cuda")
    model = trainer.build_scene_model(dev, log2_hashmap_size=17, desired_resolution=1024, seed=0)
    H, W = 48, 64
    views = []
    for v in range(9):
        o, d = syn.camera_rays(H, W, view=v)
        views.append((o.to(dev), d.to(dev), syn.bear_color(o + d * 1.5).to(dev)))
    held = views.pop()                                   # never trained on
    fs = fused_trainer.FusedTrainStep(model, H * W, lr=5e-3, lr_decay_base=0.1, lr_decay_iters=400)

    def held_out_mse():
        model.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            img = model.render(held[0][None], held[1][None], perturb=False)["image"].reshape(-1, 3)
        model.train()
        return float(((img - held[2]) ** 2).mean())

    before = held_out_mse()
    first = last = None
    for it in range(400):
        o, d, t = views[it % len(views)]
        fs.step(o, d, t)
        if it % 50 == 49:
            loss, samples, used = fs.last_stats()        # also grows the sample buffers if a view overflowed them
            assert math.isfinite(loss) and used <= fs.m_cap
            first = loss if first is None else first
            last = loss
            with torch.autocast("cuda", dtype=torch.float16):
                model.update_extra_state()               # occupancy grid from the LEARNED density
            # (the captured graph stays valid: the bitfield is updated in place and the sample count is read on the device)
    after = held_out_mse()
    assert int(fs.step_count) == 400
    assert last < 0.5 * first, (first, last)
    assert after < 0.35 * before, (before, after)
    assert int(model.density_bitfield.count_nonzero()) > 0
