#!/usr/bin/env python
"""Timings of the BASELINE.json configs that are NOT the bench line (configs[1] is bench.py): run on the GPU box,
prints one JSON object per config.  usage: python profiles/bench_configs.py > gpurun_out/configs.jsonl"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from customnerf_b200 import trainer, fused_trainer, synthetic as syn, raymarching  # noqa: E402

dev = torch.device("cuda")


def timeit(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def c0_dense_path():
    """configs[0] on the GPU: 4096 rays through the dense (non-cuda_ray) renderer, forward + backward + Adam"""
    opt = trainer.make_opt(cuda_ray=False)
    model = trainer.build_scene_model(dev, opt=opt)
    ts = trainer.TrainStep(model)
    o, d = syn.random_rays(4096, seed=1)
    o, d = o.to(dev), d.to(dev)
    tgt = syn.bear_color(o + d * 1.5)
    orig = model.render
    model.render = lambda ro, rd, **kw: orig(ro, rd, **{**kw, "num_steps": 64, "upsample_steps": 64})
    ms = timeit(lambda: ts.step(o, d, tgt), 10)
    # the same step as one CUDA-graph replay: FusedTrainStep(dense=(64, 64)) -- sampler kernels, density-only fused launch,
    # then the occupancy path's encode .. encode^T and the fused Adam
    model2 = trainer.build_scene_model(dev, opt=opt)
    fs = fused_trainer.FusedTrainStep(model2, 4096, dense=(64, 64))
    fs.set_batch(o, d, tgt)
    ms_fused = timeit(lambda: fs.step(), 20, warm=5)
    loss = fs.last_stats()[0]
    return {"config": "configs[0] dense renderer, 4096 rays, 64+64 samples, fwd+bwd+Adam",
            "ms_per_step": ms_fused, "rays_per_s": 4096 / ms_fused * 1e3, "how": "FusedTrainStep(dense=(64, 64)), one CUDA-graph replay per step",
            "ms_per_step_autograd_over_the_drop_in_ops": ms, "final_loss": loss}


def c2_inference():
    """configs[2]: full 248x184 image through march_rays / composite_rays (n_step loop) + one occupancy-grid update"""
    model = trainer.build_scene_model(dev)
    model.eval()
    o, d = syn.camera_rays(184, 248)
    o, d = o.to(dev), d.to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ms = timeit(lambda: model.render(o[None], d[None], perturb=False), 5, warm=1)
        model.fast_inference = False
        ms_loop = timeit(lambda: model.render(o[None], d[None], perturb=False), 3, warm=1)
        model.fast_inference = True
        model.train()
        upd = timeit(lambda: model.update_extra_state(), 5, warm=2)                  # default: encoder + density-only field over 8 chunks
        model.occ_chunk_rows = 0
        upd_one = timeit(lambda: model.update_extra_state(), 5, warm=2)              # the one-kernel density query
        model.occ_chunk_rows = 1 << 20
        model.density = model.density                                               # instance attribute -> the op-by-op path
        upd_ops = timeit(lambda: model.update_extra_state(), 3, warm=1)
    return {"config": "configs[2] inference 248x184 (45632 rays), device-driven rounds; occupancy update of 2x128^3 cells",
            "ms_per_frame": ms, "rays_per_s": o.shape[0] / ms * 1e3, "ms_per_frame_host_loop": ms_loop, "update_extra_state_ms": upd,
            "update_extra_state_one_kernel_ms": upd_one, "update_extra_state_op_by_op_ms": upd_ops}


def c3_lgie():
    """configs[3] on one GPU: the LGIE editing render -- foreground-masked local render + full-image global render with the
    soft edit mask and detach_bg, all / fg / bg composites over the same samples (rendering._lgie_composites) -- forward +
    backward + Adam through the autograd composition of the drop-in ops.  The SDS guidance that produces the editing
    loss is out of scope (north_star); a per-pixel stand-in loss touches every rendered output instead."""
    import torch.nn.functional as F
    opt = trainer.make_opt(train_conf=0.01, soft_mask=True, detach_bg=True)
    model = trainer.build_scene_model(dev, opt=opt)
    ts = trainer.TrainStep(model)
    o, d = syn.camera_rays(105, 142)
    tgt = syn.bear_color(o + d * 1.5).to(dev)
    o, d = o.to(dev), d.to(dev)
    gt_mask = (tgt.mean(-1, keepdim=True) > 0.5).float()

    def step():
        for p in ts.params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(o[None], d[None], staged=False, perturb=True, force_all_rays=True, **vars(model.opt))
            loss = (F.mse_loss(out["image"].reshape(-1, 3), tgt) + F.mse_loss(out["fg"]["image"].reshape(-1, 3), tgt * gt_mask) +
                    F.mse_loss(out["bg"]["image"].reshape(-1, 3), tgt * (1 - gt_mask)) +
                    0.01 * F.mse_loss(out["render_mask"].reshape(-1, 1), gt_mask))
        (loss * trainer.LOSS_SCALE).backward()
        torch._foreach_mul_([p.grad for p in ts.params if p.grad is not None], 1.0 / trainer.LOSS_SCALE)
        ts.optimizer.step()
        return loss
    ms = timeit(step, 20, warm=3)
    samples = int(model.step_counter[(model.local_step - 1) % 16, 0])
    # the same step fused: one CUDA-graph replay (customnerf_b200/fused_edit.py)
    from customnerf_b200 import fused_edit

    def loss_fn(out):
        return (F.mse_loss(out["image"].reshape(-1, 3), tgt) + F.mse_loss(out["fg"]["image"].reshape(-1, 3), tgt * gt_mask) +
                F.mse_loss(out["bg"]["image"].reshape(-1, 3), tgt * (1 - gt_mask)) +
                0.01 * F.mse_loss(out["render_mask"].reshape(-1, 1), gt_mask))
    model2 = trainer.build_scene_model(dev, opt=opt)
    fe = fused_edit.FusedEditStep(model2, o.shape[0], loss_fn)
    fe.step(o, d, tgt)
    fe.last_stats()
    ms_fused = timeit(lambda: fe.step(), 50, warm=5)
    _, samples_f, used_f = fe.last_stats()
    return {"config": "configs[3] at 1 GPU: LGIE editing render (all / fg / bg composites, soft mask, detach_bg), 14910 rays, "
                      "fwd+bwd+Adam (autograd over the drop-in ops, stand-in per-pixel loss)",
            "ms_per_step_autograd_path": ms, "rays_per_s_autograd_path": o.shape[0] / ms * 1e3, "samples_per_step": samples,
            "ms_per_step": ms_fused, "rays_per_s": o.shape[0] / ms_fused * 1e3, "samples_per_step_fused": samples_f,
            "fused": "FusedEditStep: one CUDA-graph replay (3 gated composites fwd/bwd, tcgen05 field, loss via autograd on the per-ray outputs)"}


def c4_scale(n_rays=1 << 20, log2_T=22):
    """configs[4] on one GPU: 2^22 table, 1 M rays per step (fused graph step)"""
    model = trainer.build_scene_model(dev, log2_hashmap_size=log2_T)
    o, d = syn.random_rays(n_rays, seed=2)
    tgt = syn.bear_color(o + d * 1.5)
    fs = fused_trainer.FusedTrainStep(model, n_rays)
    fs.step(o.to(dev), d.to(dev), tgt.to(dev))
    loss, samples, used = fs.last_stats()
    ms = timeit(lambda: fs.step(), 5, warm=2)
    loss, samples, used = fs.last_stats()
    fs.use_graph = False
    stages = fs.profile_stages(3)
    return {"config": "configs[4] at 1 GPU: 2^%d table, %d rays/step" % (log2_T, n_rays), "ms_per_step": ms,
            "rays_per_s": n_rays / ms * 1e3, "samples_per_step": samples, "rows_used": used,
            "stage_us": {k: round(v, 1) for k, v in stages.items()}}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c0", "c2", "c3", "c4"]
    for name, fn in (("c0", c0_dense_path), ("c2", c2_inference), ("c3", c3_lgie), ("c4", c4_scale)):
        if name in which:
            try:
                print(json.dumps(fn()), flush=True)
            except Exception as e:                       # keep going: one config failing must not hide the others
                print(json.dumps({"config": name, "error": repr(e)}), flush=True)
            torch.cuda.empty_cache()
