"""-m gpu: checkpoints in the reference's format (nerf/utils_init_nerf.py:779-901) load into NeRFNetwork + FusedTrainStep and
training continues; a checkpoint written here loads into the reference's own objects (torch.optim.Adam over get_params,
LambdaLR, GradScaler)."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N_RAYS = 2048


def _model(seed=3):
    from customnerf_b200 import trainer
    m = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=15, desired_resolution=512, seed=seed,
                                  opt=trainer.make_opt(train_conf=0.01))
    with torch.no_grad():
        m.pos_en.embeddings.uniform_(-0.5, 0.5)
    return m


def _batch():
    from customnerf_b200 import synthetic as syn
    o, d = syn.camera_rays(105, 142)
    sel = torch.arange(5000, 5000 + N_RAYS)
    o, d = o[sel].contiguous(), d[sel].contiguous()
    return o.cuda(), d.cuda(), syn.bear_color(o + d * 1.5).cuda()


def test_reference_format_checkpoint_loads_and_training_continues():
    """a checkpoint assembled exactly as Trainer_Nerf.save_checkpoint(full=True) assembles it -- torch's own Adam / LambdaLR /
    GradScaler state dicts next to the model's state dict and the two occupancy numbers -- after three reference-style steps"""
    from customnerf_b200 import checkpoint, fused_trainer, trainer
    ref = _model()
    o, d, tgt = _batch()
    lr = 5e-4
    opt = torch.optim.Adam(ref.get_params(lr), betas=(0.9, 0.99), eps=1e-15)                       # main.py:182
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** min(it / 1000, 1))            # main.py:189
    scaler = torch.amp.GradScaler("cuda", init_scale=128.0)                                        # utils_init_nerf.py:100
    ref.train()
    for _ in range(3):                                                                             # train_one_epoch, :612-629
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            out = ref.render(o[None], d[None], staged=False, perturb=False, force_all_rays=True, **vars(ref.opt))
            loss = ((out["image"].reshape(-1, 3) - tgt) ** 2).mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        sched.step()
    ref.mean_count, ref.mean_density = 4321, 0.125
    state = {"epoch": 7, "global_step": 3, "stats": {"loss": [0.5], "valid_loss": [], "results": [], "checkpoints": [], "best_result": None},
             "mean_count": ref.mean_count, "mean_density": ref.mean_density, "optimizer": opt.state_dict(),
             "lr_scheduler": sched.state_dict(), "scaler": scaler.state_dict(), "model": ref.state_dict()}
    buf = io.BytesIO()
    torch.save(state, buf)
    buf.seek(0)
    ckpt = torch.load(buf, map_location="cuda", weights_only=False)

    new = _model(seed=9)                                     # different initial weights
    fs = fused_trainer.FusedTrainStep(new, N_RAYS, perturb=False, lr=lr, lr_decay_base=0.1, lr_decay_iters=1000)
    info = checkpoint.load_checkpoint(ckpt, new, fs, log=print)
    assert info["missing_keys"] == [] and info["unexpected_keys"] == [] and (info["epoch"], info["global_step"]) == (7, 3)
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), new.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2), n1
    assert torch.equal(new.density_bitfield, ref.density_bitfield) and torch.equal(new.density_grid, ref.density_grid)
    assert (new.mean_count, new.mean_density) == (4321, 0.125)
    assert int(fs.step_count) == 3 and int(fs.scaler[2]) == 3                    # Adam's step and the LambdaLR epoch
    sd = opt.state_dict()["state"]
    off, n = fs.layout[0][1], fs.layout[0][2]
    assert torch.equal(fs.exp_avg[off:off + n], sd[0]["exp_avg"].reshape(-1))
    assert torch.equal(fs.exp_avg_sq[off:off + n], sd[0]["exp_avg_sq"].reshape(-1))
    scale, skipped, steps = fs.scaler_state()
    assert scale == scaler.get_scale() and steps == 3
    # training continues from there: the loss keeps falling and the step / epoch counters advance
    losses = []
    for _ in range(4):
        fs.step(o, d, tgt)
        losses.append(fs.last_stats()[0])
    assert np.isfinite(losses).all() and int(fs.step_count) == 7 and int(fs.scaler[2]) == 7
    # the learning rate the device computed for the last step is LambdaLR's for epoch 6
    want_lr = lr * 0.1 ** (6 / 1000)
    got = fs.hyper.cpu()
    np.testing.assert_allclose([float(got[0]), float(got[8])], [10 * want_lr, want_lr], rtol=1e-6)


def test_checkpoint_written_here_loads_into_the_reference_objects():
    from customnerf_b200 import checkpoint, fused_trainer
    m = _model()
    o, d, tgt = _batch()
    fs = fused_trainer.FusedTrainStep(m, N_RAYS, perturb=False, lr_decay_base=0.1, lr_decay_iters=1000)
    for _ in range(3):
        fs.step(o, d, tgt)
    fs.last_stats()
    state = checkpoint.checkpoint_state(m, fs, epoch=2, global_step=3, full=True)
    assert set(state) == {"epoch", "global_step", "stats", "mean_count", "mean_density", "optimizer", "lr_scheduler", "scaler", "model"}
    other = _model(seed=5)
    missing, unexpected = other.load_state_dict(state["model"], strict=False)
    assert not missing and not unexpected
    opt = torch.optim.Adam(other.get_params(5e-4), betas=(0.9, 0.99), eps=1e-15)
    opt.load_state_dict(state["optimizer"])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** min(it / 1000, 1))
    sched.load_state_dict(state["lr_scheduler"])
    assert sched.last_epoch == 3
    scaler = torch.amp.GradScaler("cuda")
    scaler.load_state_dict(state["scaler"])
    assert scaler.get_scale() == 128.0 and scaler.get_growth_interval() == 2000
    assert int(float(opt.state_dict()["state"][0]["step"])) == 3
