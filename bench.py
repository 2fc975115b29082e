#!/usr/bin/env python
"""bench.py -- train rays/s (forward + backward + optimiser) of the NeRF render/train hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): single-B200 reconstruction step at train_resolution_level 7 -- one 142 x 105
image (14 910 rays) of the synthetic bear scene, bound 2, 128^3 x 2 occupancy grid, ~270 k samples per step,
16-level hash grid (2^19, F=2) + 64-wide MLPs, occupancy (cuda_ray) path, fp16 autocast, Adam.
With N GPUs every rank renders its own view of the scene (weak scaling) and the gradients of the hash table and
the MLPs are all-reduced once per step (NCCL).

A "step" = one full train step over one image: near/far -> march -> encode -> MLP -> composite -> loss ->
backward (composite, MLP, encode scatter) -> Adam.
  value : rays/s with the ray batch already resident in HBM (device-timed with CUDA events, L2 flushed between
          steps outside the event pairs, max over ranks)
  e2e   : rays/s through the public API with HOST (pinned) ray / target buffers: H2D copies and the D2H read of
          the loss inside the timed region
The reference arm (--impl reference) times the CPU restatement of the reference's PyTorch render path
(oracle/torch_ref.py, dense 64+64 sampler, fp32) on a bounded sample of the same image's rays.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train_rays_per_sec_fwd_bwd"
UNIT = "rays/s"
IMG_H, IMG_W = 105, 142
WORKLOAD = "configs[1]: 142x105 image (14910 rays), bear scene, bound 2, 128^3x2 occupancy grid, hash 2^19 L16 F2, " \
           "64-wide MLPs, cuda_ray path, fp16 autocast, Adam"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(steps, warmup, n_rays):
    """The reference's non-cuda_ray PyTorch render path (NeRFRenderer.run, nerf/renderer.py:278-474) restated in
    fp32 PyTorch on the host cores, train step = render + MSE + backward + Adam, on ``n_rays`` rays of the image."""
    import torch
    from oracle import torch_ref
    from customnerf_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    opt = torch_ref.default_opt(cuda_ray=False, train_conf=0)
    net = torch_ref.NeRFNetwork(opt, encoder_kwargs=dict(log2_hashmap_size=19, desired_resolution=2048, gridtype="hash"))
    net.train()
    optim = torch.optim.Adam(net.get_params(5e-4), betas=(0.9, 0.99), eps=1e-15)
    o, d = syn.camera_rays(IMG_H, IMG_W)
    sel = torch.linspace(0, o.shape[0] - 1, n_rays).long()
    o, d = o[sel].contiguous(), d[sel].contiguous()
    target = syn.bear_color(o + d * 1.5)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        optim.zero_grad(set_to_none=True)
        out = net.render(o[None], d[None], num_steps=64, upsample_steps=64, perturb=True)
        loss = ((out["image"].reshape(-1, 3) - target) ** 2).mean()
        loss.backward()
        optim.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": n_rays / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of the image's 14910 rays per step, dense 64+64 sampler (393k field evaluations incl. "
                      "density-only passes), fp32, render+MSE+backward+Adam, %d timed steps, %.2f s/step"
                      % (n_rays, len(times), sec), "ms_per_step": sec * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rays = 2048
    res = cpu_reference_run(max(1, min(args.steps, 5)), max(1, min(args.warmup, 1)), n_rays)
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference's dense PyTorch render path on "
                       "a %d-ray sample of the image; rank 0 only" % n_rays},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def op_breakdown(model, rays_o, rays_d, reps=20):
    """Device time of every native kernel of one step, each timed alone with CUDA events (L2 flushed in between)."""
    import numpy as np
    import torch
    from customnerf_b200 import raymarching as rm, _lib as L
    dev = rays_o.device
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn):
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    out = {}
    N = rays_o.shape[0]
    nears, fars = rm.near_far_from_aabb(rays_o, rays_d, model.aabb_train)
    out["near_far_from_aabb"] = timeit(lambda: rm.near_far_from_aabb(rays_o, rays_d, model.aabb_train))
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    noises = torch.rand(N, device=dev)
    xyzs, dirs, deltas, rays = rm.march_rays_train(rays_o, rays_d, model.bound, model.density_bitfield, model.cascade,
                                                   model.grid_size, nears, fars, counter, -1, True, 128, True, 0, 1024,
                                                   noises=noises)
    M = xyzs.shape[0]
    scratch = torch.empty(int(L.lib().nb200_march_scratch_ints(L.u32(N))), dtype=torch.int32, device=dev)
    lib = L.lib()

    def march_count():
        counter.zero_()
        L.check(lib.nb200_march_rays_train_count(L.ptr(rays_o), L.ptr(rays_d), L.ptr(model.density_bitfield),
                L.f32(model.bound), L.f32(0), L.u32(1024), L.u32(N), L.u32(model.cascade), L.u32(model.grid_size),
                L.ptr(nears), L.ptr(fars), L.ptr(noises), L.ptr(rays), L.ptr(counter), L.ptr(scratch), L.stream()), "count")

    def march_write():
        L.check(lib.nb200_march_rays_train_write(L.ptr(rays_o), L.ptr(rays_d), L.ptr(model.density_bitfield),
                L.f32(model.bound), L.f32(0), L.u32(1024), L.u32(N), L.u32(model.cascade), L.u32(model.grid_size), L.u32(M),
                L.ptr(nears), L.ptr(fars), L.ptr(noises), L.ptr(rays), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.stream()), "write")
    out["march_count(+scan)"] = timeit(march_count)
    out["march_write"] = timeit(march_write)

    enc = model.pos_en
    emb16 = enc.embeddings.detach().half()
    x01 = ((xyzs + model.bound) / (2 * model.bound)).contiguous()
    feat = torch.empty(M, 32, dtype=torch.half, device=dev)
    S = float(np.log2(enc.per_level_scale))

    def enc_fwd():
        L.check(lib.nb200_grid_encode_forward(L.ptr(x01), L.ptr(emb16), L.ptr(enc.offsets), L.ptr(feat), L.u32(M), L.u32(3),
                L.u32(2), L.u32(16), L.u32(16), L.f32(S), L.u32(16), L.ptr(None), L.u32(enc.gridtype_id), L.i32(0), L.u32(0),
                L.i32(L.F16), L.i32(L.LAYOUT_BLC), L.stream()), "fwd")
    gfeat = torch.randn(M, 32, device=dev).half()
    gemb = torch.zeros_like(enc.embeddings)

    def enc_bwd(agg):
        def f():
            L.check(lib.nb200_grid_encode_backward(L.ptr(gfeat), L.ptr(x01), L.ptr(enc.offsets), L.ptr(gemb), L.u32(M),
                    L.u32(3), L.u32(2), L.u32(16), L.u32(16), L.f32(S), L.u32(16), L.ptr(None), L.ptr(None),
                    L.u32(enc.gridtype_id), L.i32(0), L.u32(0), L.i32(L.F16), L.i32(L.LAYOUT_BLC), L.i32(agg), L.stream()), "bwd")
        return f
    out["grid_encode_forward_f16"] = timeit(enc_fwd)
    out["grid_encode_backward_f16_agg"] = timeit(enc_bwd(1))
    out["grid_encode_backward_f16_noagg"] = timeit(enc_bwd(0))

    sig = (torch.rand(M, device=dev) * 50).requires_grad_()
    rgb = torch.rand(M, 3, device=dev).requires_grad_()
    ws = torch.empty(N, device=dev); dp = torch.empty(N, device=dev); im = torch.empty(N, 3, device=dev)
    gs = torch.empty(M, device=dev); gc = torch.empty(M, 3, device=dev)
    gws = torch.randn(N, device=dev); gim = torch.randn(N, 3, device=dev)
    out["composite_train_forward"] = timeit(lambda: L.check(lib.nb200_composite_rays_train_forward(
        L.ptr(sig), L.ptr(rgb), L.ptr(deltas), L.ptr(rays), L.u32(M), L.u32(N), L.f32(1e-4), L.ptr(ws), L.ptr(dp), L.ptr(im),
        L.stream()), "cf"))
    out["composite_train_backward"] = timeit(lambda: L.check(lib.nb200_composite_rays_train_backward(
        L.ptr(gws), L.ptr(gim), L.ptr(sig), L.ptr(rgb), L.ptr(deltas), L.ptr(rays), L.ptr(ws), L.ptr(im), L.u32(M), L.u32(N),
        L.f32(1e-4), L.ptr(gs), L.ptr(gc), L.stream()), "cb"))
    # fused tcgen05 field network (trunk + heads), forward with activations saved, and backward
    from customnerf_b200.nerf.fused_field import _image_bytes
    nb = _image_bytes()
    fimg = torch.empty(nb, dtype=torch.uint8, device=dev); bimg = torch.empty(nb, dtype=torch.uint8, device=dev)
    tp, dp_, rp = model.network.params.detach(), model.density_network.params.detach(), model.rgb_network.params.detach()
    L.check(lib.nb200_field_pack_weights(L.ptr(tp), L.ptr(dp_), L.ptr(rp), L.ptr(fimg), L.ptr(bimg), L.stream()), "pack")
    sigma = torch.empty(M, device=dev); sarg = torch.empty(M, device=dev)
    rgba = torch.empty(M, 4, dtype=torch.half, device=dev); act = torch.empty(5, M, 64, dtype=torch.half, device=dev)
    enc_fwd()
    out["field_forward_tcgen05"] = timeit(lambda: L.check(lib.nb200_field_forward(
        L.ptr(feat), L.ptr(xyzs), L.ptr(dirs), L.ptr(fimg), L.ptr(sigma), L.ptr(sarg), L.ptr(rgba), L.ptr(act), L.u32(M),
        L.stream()), "ff"))
    out["field_forward_tcgen05_nosave"] = timeit(lambda: L.check(lib.nb200_field_forward(
        L.ptr(feat), L.ptr(xyzs), L.ptr(dirs), L.ptr(fimg), L.ptr(sigma), L.ptr(None), L.ptr(rgba), L.ptr(None), L.u32(M),
        L.stream()), "ff"))
    dsig = torch.randn(M, device=dev) * 0.01; drgba = torch.randn(M, 4, device=dev)
    dx = torch.empty(M, 32, dtype=torch.half, device=dev)
    gt, gd_, gr = torch.zeros_like(tp), torch.zeros_like(dp_), torch.zeros_like(rp)
    out["field_backward_tcgen05"] = timeit(lambda: L.check(lib.nb200_field_backward(
        L.ptr(dsig), L.ptr(drgba), L.ptr(sarg), L.ptr(rgba), L.ptr(feat), L.ptr(dirs), L.ptr(act), L.ptr(bimg), L.ptr(dx),
        L.ptr(gt), L.ptr(gd_), L.ptr(gr), L.u32(M), L.stream()), "fb"))
    model.use_fused_field = False
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        out["field_forward(torch/cuBLAS per-layer path)"] = timeit(lambda: model(xyzs, dirs))
    model.use_fused_field = True
    samples = int(counter[0])
    return out, samples, M


def roofline_from(breakdown, samples, peaks):
    """Dominant native kernel of the step and its achieved algorithmic bandwidth (SURVEY.md 8(d) per-unit bytes)."""
    per_unit = {  # bytes per sample
        "grid_encode_forward_f16": 588.0,      # 12 B coords + 16*8 corners * 4 B + 64 B out
        "grid_encode_backward_f16_agg": 1100.0 + 1024.0,   # 12 + 64 B grad + fp32 atomics RMW 2*8*16*8 B
        "composite_train_forward": 24.0,
        "composite_train_backward": 40.0,
        "march_write": 32.0,
        # field network: HBM bytes per point (x_en 64 + xyz/dirs 24 + outputs 12 + saved activations 644)
        "field_forward_tcgen05": 744.0,
        # reads activations 640 + x_en 64 + dirs 12 + grads/outputs 36, writes d_x_en 64
        "field_backward_tcgen05": 816.0,
    }
    mine = {k: v for k, v in breakdown.items() if k in per_unit}
    top = max(mine, key=mine.get)
    us = mine[top]
    achieved = per_unit[top] * samples / (us * 1e-6) / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "us_per_launch": us, "bytes_per_unit": per_unit[top], "units_per_launch": samples,
            "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from customnerf_b200 import parallel, synthetic as syn, trainer, _lib as L

    rank, local_rank, world = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    L.lib()   # fail loudly if the native library is missing

    model = trainer.build_scene_model(dev)
    sync = parallel.FlatGradSync([p for g in model.get_params(5e-4) for p in g["params"]]) if world > 1 else None
    ts = trainer.TrainStep(model, lr=5e-4, fp16=True, world_size=world, grad_sync=sync)
    o, d = syn.camera_rays(IMG_H, IMG_W, view=rank)           # weak scaling: one image per rank
    target = syn.bear_color(o + d * 1.5)
    n_rays = o.shape[0]
    o_h, d_h, t_h = o.pin_memory(), d.pin_memory(), target.pin_memory()
    o_d, d_d, t_d = o.to(dev), d.to(dev), target.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, host_inputs):
        evs = []
        for _ in range(n_steps):
            flush.zero_()                                      # L2 flush, outside the per-step event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if host_inputs:
                ro, rd, tg = o_h.to(dev, non_blocking=True), d_h.to(dev, non_blocking=True), t_h.to(dev, non_blocking=True)
                loss = ts.step(ro, rd, tg, n_total=n_rays * world)
                _ = loss.item()                                # D2H read of the step's result
            else:
                ts.step(o_d, d_d, t_d, n_total=n_rays * world)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / 1e3    # seconds of device time

    for _ in range(max(args.warmup, 3)):
        ts.step(o_d, d_d, t_d, n_total=n_rays * world)
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    L.LAUNCHES = 0
    sec = timed(args.steps, host_inputs=False)
    launches = L.LAUNCHES
    barrier()
    sec_e2e = timed(args.steps, host_inputs=True)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([sec, sec_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec, sec_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        total_rays = n_rays * world * args.steps
        line = {"metric": METRIC, "value": total_rays / sec, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "rays_per_gpu_per_step": n_rays, "parallelism": "ray-sharded dp%d" % world,
                           "l2": "256 MB memset between steps, outside the per-step CUDA-event pairs",
                           "timing": "sum of per-step CUDA-event intervals, max over ranks"},
                "e2e": {"value": total_rays / sec_e2e, "unit": UNIT,
                        "h2d_bytes_per_step": int(3 * n_rays * 3 * 4), "d2h_bytes_per_step": 4,
                        "ms_per_step": sec_e2e / args.steps * 1e3},
                "gpu_launches": int(launches), "clocks": clk}
        if world == 1 and not args.no_breakdown:
            bd, samples, M = op_breakdown(model, o_d, d_d)
            line["kernel_us"] = {k: round(v, 2) for k, v in bd.items()}
            line["samples_per_step"] = samples
            line["roofline"] = roofline_from(bd, samples, peaks)
            n_cpu = 2048
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_run(3, 1, n_cpu).items() if k != "ms_per_step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-breakdown", action="store_true",
                    help="skip the per-kernel breakdown and the CPU baseline (profiling runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
