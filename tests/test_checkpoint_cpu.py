"""CPU: the state-dict mapping of customnerf_b200/checkpoint.py (no GPU: stand-in objects carry the tensors the fused step
keeps on the device).  Reference: Trainer_Nerf.save_checkpoint / load_checkpoint, nerf/utils_init_nerf.py:779-901."""
import types

import torch

from customnerf_b200 import checkpoint


def _fake_step(scale=256.0, tracker=17, iteration=1234, interval=2000, lr=5e-4, decay_base=0.1, decay_iters=10000.0):
    scaler = torch.zeros(8, dtype=torch.int32)
    scaler[0:1] = torch.tensor([scale]).view(torch.int32)
    scaler[1], scaler[2], scaler[6] = tracker, iteration, interval
    sched = torch.tensor([lr * 10, lr, 0.9, 0.99, 1e-15, 1 / 128.0, decay_base, decay_iters])
    return types.SimpleNamespace(scaler=scaler, sched=sched, lr=lr, step_count=torch.tensor([iteration]))


def test_scaler_and_scheduler_state_dicts_have_torch_s_keys():
    fs = _fake_step()
    sd = checkpoint._scaler_state_dict(fs)
    ref = torch.amp.GradScaler("cpu", enabled=True).state_dict() if hasattr(torch, "amp") else None
    assert sd == {"scale": 256.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 17}
    if ref:                                  # same key set as torch's own scaler writes
        assert set(sd) == set(ref)
    ls = checkpoint._lr_scheduler_state_dict(fs)
    assert ls["last_epoch"] == 1234 and ls["_step_count"] == 1235
    assert ls["base_lrs"] == [5e-3, 5e-4, 5e-4, 5e-4]
    decay = 0.1 ** (1234 / 10000.0)          # main.py:189: 0.1 ** min(iter / iters, 1)
    for got, base in zip(ls["_last_lr"], ls["base_lrs"]):
        assert abs(got - base * decay) <= 1e-6 * base      # (the schedule words are fp32 on the device)
    # what torch's LambdaLR writes for a plain-function lambda
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([{"params": [p], "lr": 5e-3}], lr=5e-4)
    ref_ls = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** min(it / 10000.0, 1)).state_dict()
    assert set(ref_ls) <= set(ls) | {"_is_initial"}, (set(ref_ls) - set(ls))


def test_constant_scale_step_writes_no_scaler_state():
    fs = _fake_step()
    fs.scaler = None
    assert checkpoint._scaler_state_dict(fs) == {}
    assert checkpoint._lr_scheduler_state_dict(fs)["last_epoch"] == 1234      # falls back to the Adam step count
