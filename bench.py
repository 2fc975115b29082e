#!/usr/bin/env python
"""bench.py -- train rays/s (forward + backward + optimiser) of the NeRF render/train hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): single-B200 reconstruction step at train_resolution_level 7 -- one 142 x 105
image (14 910 rays) of the synthetic bear scene, bound 2, 128^3 x 2 occupancy grid, ~270 k samples per step,
16-level hash grid (2^19, F=2) + 64-wide MLPs, occupancy (cuda_ray) path, fp16 autocast, Adam.
With N GPUs every rank renders its own view of the scene (weak scaling); the gradients of the hash table and the MLPs
are summed and the optimiser step taken by ONE kernel over NVLink peer memory (csrc/peer_update.cu; --update nccl: NCCL
all-reduce + Adam instead).  --config 4 runs BASELINE.json configs[4] (2^22 table, 1 M rays per step) the same way.

A "step" = one full train step over one image: near/far -> march -> encode -> MLP -> composite -> loss ->
backward (composite, MLP, encode scatter) -> Adam, replayed as one CUDA graph (customnerf_b200/fused_trainer.py).
  value : rays/s with the ray batch already resident in HBM (device-timed with CUDA events, L2 flushed between
          steps outside the event pairs, max over ranks)
  e2e   : rays/s through the public API with HOST (pinned) ray / target buffers: H2D copies and the D2H read of
          every step's loss inside the timed region, in the asynchronous loop a trainer runs (the result of step k - 1 is
          read while step k executes); e2e_sync is the same with the host waiting for every step before issuing the next
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

TRAIN_CONF = 0.01          # main.py:132 default: 4-output colour head, loss += train_conf * MSE(render_mask, gt_mask)


def silhouette(o, d, n=48):
    """synthetic ground-truth mask [N]: 1 where the ray meets the analytic bear (48 density probes along the ray)"""
    import torch
    from customnerf_b200 import synthetic as syn
    t = torch.linspace(0.4, 2.8, n)
    p = o[:, None, :] + d[:, None, :] * t[None, :, None]
    return (syn.bear_density(p).max(dim=1).values > 1.0).float()


METRIC = "train_rays_per_sec_fwd_bwd"
UNIT = "rays/s"
IMG_H, IMG_W = 105, 142
WORKLOAD = "configs[1]: 142x105 image (14910 rays), bear scene, bound 2, 128^3x2 occupancy grid, hash 2^19 L16 F2, " \
           "64-wide MLPs (rgb + mask head), cuda_ray path, fp16 autocast, loss MSE(rgb) + 0.01 MSE(mask), Adam"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a thread (the timed
    region of this bench is tens of milliseconds, shorter than nvidia-smi's fastest loop), nvidia-smi as a fallback."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.thread, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.reasons, self.nvml = [], None, set(), None

    def _poll(self):
        n, h = self.nvml, self.handle
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # handle and max clock are fetched here (blocking, before the timed region) so the thread only polls
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            try:
                self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            except Exception:
                self.mx = None
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
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
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml, 2 ms poll during the timed region"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


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
    opt = torch_ref.default_opt(cuda_ray=False, train_conf=TRAIN_CONF)      # the reference's default: mask head + mask loss
    net = torch_ref.NeRFNetwork(opt, encoder_kwargs=dict(log2_hashmap_size=19, desired_resolution=2048, gridtype="hash"))
    net.train()
    optim = torch.optim.Adam(net.get_params(5e-4), betas=(0.9, 0.99), eps=1e-15)
    o, d = syn.camera_rays(IMG_H, IMG_W)
    sel = torch.linspace(0, o.shape[0] - 1, n_rays).long()
    o, d = o[sel].contiguous(), d[sel].contiguous()
    target = syn.bear_color(o + d * 1.5)
    gt_mask = silhouette(o, d)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        optim.zero_grad(set_to_none=True)
        out = net.render(o[None], d[None], num_steps=64, upsample_steps=64, perturb=True)
        loss = ((out["image"].reshape(-1, 3) - target) ** 2).mean() \
            + TRAIN_CONF * ((out["render_mask"].reshape(-1) - gt_mask) ** 2).mean()      # utils_init_nerf.py:224-234
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
    """--impl reference: the CPU restatement, EXACTLY args.steps timed steps after args.warmup warm-up steps (what the line
    prints is what ran).  A step is a bounded sample of the workload: ``n_rays`` of the image's rays, sized from the step
    count so that the whole run stays within a few minutes on the box's host cores (~0.7 s per 2048-ray step on 16 cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget_rays = 2048 * 60                        # ~40 s of CPU work in total at the measured ~3 k rays/s
    n_rays = int(min(2048, max(128, budget_rays // (steps + warmup))))
    res = cpu_reference_run(steps, warmup, n_rays)
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference's dense PyTorch render path on "
                       "a %d-ray sample of the image per step; %d timed + %d warm-up steps really run; rank 0 only"
                       % (n_rays, steps, warmup)},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
# Algorithmic bytes per sample of every stage (DESIGN.md "Kernels"): what the stage must move if every
# byte were touched exactly once.  fp16 features / activations, fp32 table master and gradients.
BYTES_PER_SAMPLE = {
    "march_write": 32.0,                    # xyz 12 + dir 12 + deltas 8 written
    "grid_encode_forward": 12.0 + 16 * 8 * 8 + 64.0,       # coords + 128 corner float2 gathers (fp32 master) + 64 B features
    "field_forward": 64.0 + 24.0 + 4 + 4 + 8 + 640.0,      # x_en + xyz/dirs + sigma + sigma_arg + rgba + 5 saved activations
    "composite_forward": 4.0 + 8 + 8,                      # sigma + rgba(f16x4) + deltas
    "composite_backward": 4.0 + 8 + 8 + 4 + 16,            # same reads + d_sigma + d_rgba(float4)
    "composite_forward_backward": 2 * (4.0 + 8 + 8) + 4 + 16,      # one launch: the reads twice (second time from L1 / L2), the gradients out
    "field_backward": 640.0 + 64 + 12 + 4 + 16 + 4 + 8 + 64,   # activations, x_en, dirs, d_sigma, d_rgba, sigma_arg, rgba; d_x_en out
    "grid_encode_backward": 12.0 + 64 + 2 * 16 * 8 * 8,    # coords + feature grads + 128 float2 atomic read-modify-writes
}
BYTES_PER_PARAM = {"adam": 32.0}            # p, g, m, v read; p, m, v, g written (16 + 16)


def l2_probe(dev):
    """L2 read bandwidth of this GPU measured in place (MEASURED_PEAKS.json has no L2 figure): a coalesced
    stream over a 32 MB buffer, and random 8-byte gathers over a 4 MB (2^19 x 8 B: one hashed level) and a 32 MB region."""
    import ctypes as C
    import torch
    from customnerf_b200 import _lib as L
    lib = L.lib()
    buf = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
    sink = torch.zeros(1, dtype=torch.float32, device=dev)

    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e-3
    reps = 20
    s_stream = t(lambda: L.check(lib.nb200_l2_stream_probe(L.ptr(buf), C.c_uint64(32 << 20), L.u32(reps), L.ptr(sink), L.stream()), "l2s"))
    out = {"stream_32MB_gbs": round((32 << 20) * reps / s_stream / 1e9, 1)}
    nthr, per = 148 * 2048, 64
    for name, words in (("gather_4MB", 1 << 19), ("gather_32MB", 1 << 22)):
        s_g = t(lambda: L.check(lib.nb200_l2_gather_probe(L.ptr(buf), L.u32(words), L.u32(nthr), L.u32(per), L.ptr(sink), L.stream()), "l2g"))
        out[name + "_gsectors_per_s"] = round(nthr * per / s_g / 1e9, 2)
        out[name + "_sector_gbs"] = round(nthr * per * 32 / s_g / 1e9, 1)
    out["how"] = "ld.global.cg (L1 bypassed); stream: float4 per thread; gather: 8-byte words at LCG-hashed positions, 32-byte sectors"
    return out


# what bounds each stage: the encoder's 23-47 MB table and its gradient stay in the 126 MB L2 (ncu: the
# scatter touches DRAM for 8 % of its algorithmic bytes), so the encode kernels are measured against the L2 rates probed in
# this run; the field kernels stream activations from / to HBM and run the only dense contraction (tensor pipe); the
# composites and Adam stream HBM; the march is issue / latency bound (samples/s reported).
STAGE_BOUND = {"grid_encode_forward": "l2", "grid_encode_backward": "l2", "composite_forward": "l2", "composite_backward": "l2", "composite_forward_backward": "l2",
               "field_forward": "hbm", "field_backward": "hbm", "adam": "hbm", "march_write": "hbm"}
FLOPS_PER_SAMPLE = {"field_forward": 2.0 * 20480, "field_backward": 4.0 * 20480}     # useful MACs x 2; backward = dgrad + wgrad


def rooflines(stage_us, samples, n_params, peaks, l2=None, fused_composite=False):
    """achieved algorithmic GB/s of every stage against the roofline that bounds it; the dominant stage is the headline"""
    hbm = peaks.get("hbm_gbs", 6650.0)
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1370.0))
    l2_stream = (l2 or {}).get("stream_32MB_gbs")
    l2_sector = (l2 or {}).get("gather_32MB_sector_gbs")
    per = {}
    stage_us = dict(stage_us)
    # compositing forward + MSE + backward as ONE launch (FusedTrainStep.fused_composite): the second stage is an empty pair of
    # events -- report the launch as one stage with the bytes of both directions
    if fused_composite and "composite_forward" in stage_us and "composite_backward" in stage_us:
        stage_us["composite_forward_backward"] = stage_us.pop("composite_forward") + stage_us.pop("composite_backward")
    for k, us in stage_us.items():
        if k in BYTES_PER_SAMPLE:
            units, bpu = samples, BYTES_PER_SAMPLE[k]
        elif k in BYTES_PER_PARAM:
            units, bpu = n_params, BYTES_PER_PARAM[k]
        else:
            continue
        ach = bpu * units / (us * 1e-6) / 1e9
        bound = STAGE_BOUND.get(k, "hbm")
        if bound == "l2" and not l2_stream:
            bound = "hbm"
        peak = l2_stream if bound == "l2" else hbm
        row = {"us": round(us, 2), "bytes_per_unit": bpu, "units": units, "bound": bound, "achieved_gbs": round(ach, 1),
               "peak_gbs": peak, "frac": round(ach / peak, 4), "frac_of_hbm_peak": round(ach / hbm, 4)}
        if bound == "l2":
            row["frac_of_l2_random_sector_rate"] = round(ach / l2_sector, 4) if l2_sector else None
        if k in FLOPS_PER_SAMPLE:
            tfs = FLOPS_PER_SAMPLE[k] * units / (us * 1e-6) / 1e12
            row["tensor"] = {"achieved_tflops": round(tfs, 1), "peak_tflops": tf, "frac": round(tfs / tf, 4),
                             "flops_per_unit": FLOPS_PER_SAMPLE[k]}
        per[k] = row
    if "march_count" in stage_us:
        us = stage_us["march_count"] + stage_us.get("march_write", 0.0)
        per["march"] = {"us": round(us, 2), "bound": "issue/latency", "samples_per_s": round(samples / (us * 1e-6), 1),
                        "units": samples}
    timed = {k: v for k, v in per.items() if "achieved_gbs" in v}
    top = max(timed, key=lambda k: timed[k]["us"])
    t = per[top]
    # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture, when one exists
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top)
    except Exception:
        pass
    head = {"kernel": top, "bound": t["bound"], "achieved": t["achieved_gbs"], "peak": t["peak_gbs"], "unit": "GB/s",
            "frac": t["achieved_gbs"] / t["peak_gbs"], "traffic": traffic, "us_per_launch": t["us"],
            "bytes_per_unit": t["bytes_per_unit"], "units_per_launch": t["units"],
            "peak_source": ("L2 read stream over a 32 MB resident buffer measured in this run (l2_probe.stream_32MB_gbs; "
                            "MEASURED_PEAKS.json has no L2 figure)" if t["bound"] == "l2" else src),
            "frac_of_hbm_peak": t["frac_of_hbm_peak"],
            "how": "achieved = bytes_per_unit x units_per_launch / us_per_launch; CUDA events recorded between the stages of "
                   "real train steps on the launching stream (L2 flushed between steps), mean over the profiled steps"}
    if t["bound"] == "l2":
        head["frac_of_l2_random_sector_rate"] = t.get("frac_of_l2_random_sector_rate")
    # the field network against the tensor pipe (the only dense contraction): forward + backward together
    if "field_forward" in per and "field_backward" in per:
        us = per["field_forward"]["us"] + per["field_backward"]["us"]
        fl = (FLOPS_PER_SAMPLE["field_forward"] + FLOPS_PER_SAMPLE["field_backward"]) * samples
        head["mlp_tensor"] = {"bound": "tensor", "achieved": round(fl / (us * 1e-6) / 1e12, 1), "peak": tf, "unit": "TFLOP/s",
                              "frac": round(fl / (us * 1e-6) / 1e12 / tf, 4), "us": round(us, 2),
                              "flops": fl, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a step)"}
    return head, per


def run_b200(args):
    import torch
    import torch.distributed as dist
    from customnerf_b200 import parallel, synthetic as syn, trainer, fused_trainer, _lib as L

    rank, local_rank, world = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    L.lib()   # fail loudly if the native library is missing

    workload, scaling, raygen, cam_pose = WORKLOAD, "weak", None, None
    if args.config == 3:
        model = trainer.build_scene_model(dev, opt=trainer.make_opt(train_conf=TRAIN_CONF, soft_mask=True, detach_bg=True))
        o, d = syn.camera_rays(IMG_H, IMG_W, view=rank)
        target = syn.bear_color(o + d * 1.5)
        gt_mask = silhouette(o, d)
        workload = "configs[3]: LGIE editing step (foreground-masked local render + background render + full-image global " \
                   "render over the same samples, soft edit mask conf_thr 0.5, detach_bg), 142x105 image (14910 rays) per rank, " \
                   "bear scene, hash 2^19 L16 F2, 64-wide MLPs (rgb + mask head), cuda_ray path, fp16 autocast, stand-in " \
                   "per-pixel loss on all three renders + rendered mask (the SDS guidance is out of scope), Adam"
    elif args.config == 4:
        # BASELINE.json configs[4] (not the driver's line): 2^22 table, 1 M random rays per step -- sharded over the ranks
        # (strong scaling, as the config words it) or, with --weak, 1 M rays on EVERY rank
        model = trainer.build_scene_model(dev, log2_hashmap_size=22, opt=trainer.make_opt(train_conf=TRAIN_CONF))
        total = 1 << 20
        if args.weak:
            o, d = syn.random_rays(total, seed=2 + rank)
        else:
            o, d = syn.random_rays(total, seed=2)
            idx = parallel.shard_rays(total, rank, world)
            o, d, scaling = o[idx].contiguous(), d[idx].contiguous(), "strong"
        target = syn.bear_color(o + d * 1.5)
        gt_mask = torch.cat([silhouette(o[i:i + 65536], d[i:i + 65536]) for i in range(0, o.shape[0], 65536)])
        workload = "configs[4]: 2^22 hash table, %d random rays per %s, bear scene, bound 2, 128^3x2 occupancy grid, L16 F2, " \
                   "64-wide MLPs (rgb + mask head), cuda_ray path, fp16 autocast, loss MSE(rgb) + 0.01 MSE(mask), Adam" \
                   % (total, "rank and step" if args.weak else "step, sharded over the ranks")
    else:
        model = trainer.build_scene_model(dev, opt=trainer.make_opt(train_conf=TRAIN_CONF))
        o, d = syn.camera_rays(IMG_H, IMG_W, view=rank)           # weak scaling: one image per rank
        target = syn.bear_color(o + d * 1.5)
        gt_mask = silhouette(o, d)
        cam_pose, cam_intr = syn.camera_pose(IMG_H, IMG_W, view=rank)      # the same camera in the loader's terms
        raygen = dict(H=IMG_H, W=IMG_W, intrinsics=cam_intr)
    n_rays = o.shape[0]
    # N > 1, default: the update is ONE kernel over NVLink peer memory (csrc/peer_update.cu: every rank reduces + Adam-
    # updates the slice it owns out of the peers' gradients and stores the new parameters into every replica) -- no NCCL
    # call inside the step.  --update nccl: one NCCL all-reduce of the flat gradient + a full Adam sweep per rank (measured
    # at N = 2: 0.67 ms/step; in 4 pieces pipelined with Adam -- allreduce_chunks=4 -- 0.77 ms).
    peer, sync = None, None
    if args.update == "auto":
        # measured (profiles/r02t_*): the multicast form moves n (1 + 1/W) words per link direction against 2 n (W - 1) / W
        # for P2P loads / stores -- fewer from W = 4 on (0.523 vs 0.560 ms/step at W = 8), more at W = 2
        args.update = "nvls" if world >= 4 else "peer"
        if args.update == "nvls":
            try:
                peer = parallel.PeerMemory(fused_trainer.flat_parameter_count(model), dev, multicast=True)
                ok = 1
            except Exception as e:                      # no multicast object on this fabric / driver: P2P form, on every rank
                sys.stderr.write("[bench] NVSwitch multicast unavailable (%s); P2P peer update\n" % (e,))
                peer, ok = None, 0
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 0:
                peer, args.update = None, "peer"
    if peer is not None:
        pass
    elif (world > 1 or args.peer_at_1) and args.update in ("peer", "nvls"):
        peer = parallel.PeerMemory(fused_trainer.flat_parameter_count(model), dev, multicast=args.update == "nvls")
    elif world > 1:
        sync = (lambda flat: dist.all_reduce(flat, op=dist.ReduceOp.SUM))
    if args.config == 3:
        # BASELINE.json configs[3] (not the driver's line): the LGIE editing step -- fg / bg / all renders with the soft edit
        # mask and detach_bg -- on the same image, ray-sharded like configs[1].  The Stable-Diffusion guidance that gives the
        # reference its editing loss is out of scope (north_star): a per-pixel stand-in loss touches all three renders and
        # the rendered mask, normalised by the global ray count like the reconstruction loss.
        from customnerf_b200 import fused_edit
        tgt_d, msk_d = target.to(dev), gt_mask.to(dev)[:, None]
        inv = 1.0 / (n_rays * world)

        def edit_loss(out):
            return inv * (((out["image"].reshape(-1, 3) - tgt_d) ** 2).sum() / 3 +
                          ((out["fg"]["image"].reshape(-1, 3) - tgt_d * msk_d) ** 2).sum() / 3 +
                          ((out["bg"]["image"].reshape(-1, 3) - tgt_d * (1 - msk_d)) ** 2).sum() / 3 +
                          TRAIN_CONF * ((out["render_mask"].reshape(-1, 1) - msk_d) ** 2).sum())
        fs = fused_edit.FusedEditStep(model, n_rays, edit_loss, lr=5e-4, world_size=world, grad_sync=sync,
                                      use_graph=not args.no_graph, peer=peer)
        raygen = None
        args.no_breakdown = True
    else:
        fs = fused_trainer.FusedTrainStep(model, n_rays, lr=5e-4, world_size=world, grad_sync=sync, use_graph=not args.no_graph,
                                          pipeline_update=not args.no_pipeline, mask_weight=TRAIN_CONF, peer=peer, raygen=raygen,
                                          fused_forward=os.environ.get("NB200_FUSED_FORWARD", "0") == "1")
    fs.target_mask.copy_(gt_mask)     # [N] ground-truth mask: resident (59 KB; not part of the per-step H2D count)
    # the batch is handed over the way a loader would: written into one of the trainer's two pinned staging slots, from where
    # step() copies it to the device (one H2D copy of 537 KB per step, inside the timed region)
    o_h, d_h, t_h = fs.pinned_batch()
    o_h.copy_(o); d_h.copy_(d); t_h.copy_(target)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the second staging slot (a loader alternates the two when it issues steps without waiting for the one before)
    slots = [(o_h, d_h, t_h), fs.pinned_batch(1)]
    for dst, src in zip(slots[1], (o, d, target)):
        dst.copy_(src)
    pose_slots = None
    if raygen is not None:
        pose_slots = [fs.pinned_pose_batch(0), fs.pinned_pose_batch(1)]
        for p_h, _ in pose_slots:
            p_h.copy_(cam_pose)

    import ctypes as C_
    lib_ = L.lib()

    def timed(n_steps, host_inputs, lagged=False):
        """host_inputs: False (batch resident in HBM) | True (host rays + target) | "pose" (host pose + target).
        lagged: the result of step k - 1 is read while step k runs (previous_stats) instead of waiting for step k
        (last_stats); every step's result is still read back exactly once, the last one after the loop."""
        evs = []
        for k in range(n_steps):
            flush.zero_()                                      # L2 flush, outside the per-step event pair
            if peer is not None and world > 1:
                # N > 1: line the ranks up again after the flush (a one-CTA flag barrier over the peer memory), so that one
                # rank's 256 MB memset is not billed to the step of a peer that waits for it inside the update kernel
                L.check(lib_.nb200_peer_rank_barrier(C_.byref(fs.peer_plan), L.stream()), "peer_rank_barrier")
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if host_inputs == "pose":
                p_h, tt_h = pose_slots[k & 1]
                fs.step(pose=p_h, target=tt_h)                 # H2D of pose + target pixels, rays generated in the step
            elif host_inputs:
                fs.step(*slots[k & 1])                         # H2D of the batch (pinned) + the step
            else:
                fs.step()                                      # batch already resident in HBM
            if host_inputs:
                fs.previous_stats() if lagged else fs.last_stats()   # D2H read of loss / sample count (32 B)
            b.record()
            evs.append((a, b))
        if host_inputs and lagged:
            fs.last_stats()                                    # the last step's result
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / 1e3    # seconds of device time

    fs.step(o_h, d_h, t_h)
    for _ in range(max(args.warmup, 3)):
        fs.step()
    _, samples, used = fs.last_stats()
    assert used == samples, "sample buffers overflowed during warm-up (%d > %d)" % (samples, fs.m_cap)
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    L.LAUNCHES = 0
    sec = timed(args.steps, host_inputs=False)
    launches = L.LAUNCHES
    barrier()
    for sl in slots:                                           # captures the graph of either staging slot (outside the timing)
        fs.step(*sl)
    fs.last_stats()
    barrier()
    sec_e2e_sync = timed(args.steps, host_inputs=True)
    barrier()
    sec_e2e = timed(args.steps, host_inputs=True, lagged=True)
    barrier()
    sec_pose = None
    if raygen is not None:
        for p_h, tt_h in pose_slots:                           # captures the pose-driven graphs
            fs.step(pose=p_h, target=tt_h)
        fs.last_stats()
        barrier()
        sec_pose = timed(args.steps, host_inputs="pose", lagged=True)
        barrier()
    clk = clocks.stop() if rank == 0 else None
    loss, samples, used = fs.last_stats()
    update_us = None
    if world > 1:
        # the update alone (all ranks in lock step; the gradient is zero by now, which changes nothing about the traffic)
        import ctypes as C
        fs.flush()
        torch.cuda.synchronize()
        dist.barrier()
        reps = 20
        with torch.cuda.device(dev):
            fs._update(L.stream())
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            for _ in range(reps):
                fs._update(L.stream())
            eb.record()
            torch.cuda.synchronize()
        update_us = ea.elapsed_time(eb) / reps * 1e3
        t = torch.tensor([sec, sec_e2e, update_us, sec_pose or 0.0, sec_e2e_sync], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec, sec_e2e, update_us, sec_e2e_sync = float(t[0]), float(t[1]), float(t[2]), float(t[4])
        sec_pose = float(t[3]) if sec_pose is not None else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        total_rays = n_rays * world * args.steps
        line = {"metric": METRIC, "value": total_rays / sec, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": workload, "rays_per_gpu_per_step": n_rays, "parallelism": "ray-sharded dp%d" % world,
                           "step": ("LGIE editing step -- 3 mask-gated composites forward / backward, stand-in loss through "
                                    "autograd on the per-ray outputs; " if args.config == 3 else "") +
                                   "one CUDA-graph replay: near/far, march, encode, field MLP, composite, MSE, backward, "
                                   "fused Adam" + (", NCCL all-reduce of the flat gradient" if sync is not None else "") +
                                   ("; update = one reduce + Adam + broadcast kernel over NVLink peer memory%s (no NCCL call "
                                    "in the step)" % (" through the NVSwitch multicast mapping" if args.update == "nvls" else "")
                                    if peer is not None else "") +
                                   ("; the update of step k runs on a second stream next to the march of step k+1 (every "
                                    "replay contains exactly one update and one forward/backward)" if not args.no_pipeline else ""),
                           "sample_rows_capacity": fs.m_cap,
                           "l2_persist": ({k: v for k, v in fs.l2.items() if k != "base"} if getattr(fs, "l2", None) else None),
                           "l2": "256 MB memset between steps, outside the per-step CUDA-event pairs",
                           "timing": "sum of per-step CUDA-event intervals, max over ranks" +
                                     ("; the ranks are re-aligned by a one-CTA flag barrier after each L2 flush, outside the "
                                      "event pairs" if peer is not None and world > 1 else "")},
                "e2e": {"value": total_rays / sec_e2e, "unit": UNIT,
                        "h2d_bytes_per_step": int(3 * n_rays * 3 * 4), "d2h_bytes_per_step": 32,
                        "ms_per_step": sec_e2e / args.steps * 1e3,
                        "api": "the asynchronous training loop: FusedTrainStep.step(*fs.pinned_batch(k % 2)) (batch in one of two "
                               "pinned staging slots; one H2D copy per step on a copy stream into the matching device slot, which "
                               "the step's graph waits for), then previous_stats(): "
                               "the 32-byte result of step k - 1 is read while step k runs; every step's result is read once, "
                               "the last one after the loop"},
                "e2e_sync": {"value": total_rays / sec_e2e_sync, "unit": UNIT, "h2d_bytes_per_step": int(3 * n_rays * 3 * 4),
                             "d2h_bytes_per_step": 32, "ms_per_step": sec_e2e_sync / args.steps * 1e3,
                             "api": "the same with last_stats() after every step: the host waits for step k before it issues "
                                    "step k + 1 (launch latency, copies and the wake-up are exposed each step)"},
                "gpu_launches": int(launches), "clocks": clk, "samples_per_step": samples, "final_loss": loss}
        if sec_pose is not None:
            # the same end-to-end step fed the way the reference's loader feeds it (provider.py:344-470 hands the trainer a
            # pose; get_rays runs on the device): host inputs = 64 B of pose + the target pixels, rays generated by the
            # first kernel of the step's graph (csrc/raygen.cu)
            line["e2e_from_pose"] = {"value": total_rays / sec_pose, "unit": UNIT, "h2d_bytes_per_step": int(64 + n_rays * 3 * 4),
                                     "d2h_bytes_per_step": 32, "ms_per_step": sec_pose / args.steps * 1e3,
                                     "api": "FusedTrainStep.step(pose=, target=) with pinned_pose_batch(k % 2) + previous_stats() (asynchronous loop as e2e)"}
        if update_us is not None:
            nb = fs.params_flat.numel() * 4
            # per rank and link direction: IN = the peers' gradient slices it reads (read responses) + the parameters the
            # peers store into it; OUT = the same two streams the other way round: 2 x 4 B x n x (W - 1) / W each way
            # multicast form: OUT = its gradient to the switch (n words, reduced there) + its parameter slice once (n / W);
            # IN = the reduced slice (n / W) + everybody's parameters (n)
            wire = nb * (1 + 1.0 / world) if args.update == "nvls" else 2 * nb * (world - 1) / world
            line["update"] = {"kind": args.update, "us": round(update_us, 1),
                              "what": "adam_hyper + " + ("k_peer_reduce_adam_bcast" if peer is not None else
                                                        "ncclAllReduce(grads_flat) + k_fused_adam") + " + weight re-pack, "
                                      "back to back on all ranks, max over ranks",
                              "nvlink_bytes_each_way_per_rank": int(wire),
                              "nvlink_gbs_each_way_per_rank": round(wire / (update_us * 1e-6) / 1e9, 1)}
        if world == 1 and not args.no_breakdown:
            fs.flush()
            fs.use_graph, fs.pipeline_update = False, False      # stage times: one stream, one kernel at a time
            stage_us = fs.profile_stages(10, flush=flush.zero_)
            fs.use_graph, fs.pipeline_update = not args.no_graph, not args.no_pipeline
            line["kernel_us"] = {k: round(v, 2) for k, v in stage_us.items()}
            l2 = None
            try:
                l2 = l2_probe(dev)
                # the encoder against the L2: 128 corner gathers per sample (forward), 128 reductions (backward)
                l2["grid_encode_forward_gsectors_per_s"] = round(samples * 128 / (stage_us["grid_encode_forward"] * 1e-6) / 1e9, 2)
                l2["grid_encode_forward_frac_of_gather_32MB"] = round(l2["grid_encode_forward_gsectors_per_s"] / l2["gather_32MB_gsectors_per_s"], 3)
                # 128 ALGORITHMIC corner reductions per sample; the atomics actually issued after the warp aggregation are
                # fewer (ncu lts__t_requests_srcunit_tex_op_red per launch: profiles/r02_encode_bwd_atomics.txt)
                l2["grid_encode_backward_algorithmic_greductions_per_s"] = round(samples * 128 / (stage_us["grid_encode_backward"] * 1e-6) / 1e9, 2)
                line["l2_probe"] = l2
            except Exception as e:
                line["l2_probe"] = {"error": repr(e)}
                l2 = None
            line["roofline"], line["stage_rooflines"] = rooflines(stage_us, samples, fs.params_flat.numel(), peaks, l2, fused_composite=getattr(fs, "fused_composite", False))
            if not args.no_cpu:
                n_cpu = 2048
                line["cpu_baseline"] = {k: v for k, v in cpu_reference_run(3, 1, n_cpu).items() if k != "ms_per_step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured step holds NCCL kernels: tearing the communicator down under a live graph can dead-lock, so drop
        # the graph first, and leave without the (optional) communicator teardown
        fs.graph = None
        torch.cuda.synchronize()
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize()
        time.sleep(0.2 if rank == 0 else 1.0)      # rank 0 (the one that printed) leaves first
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels directly instead of replaying a CUDA graph")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="run each step's optimiser update before the next step starts instead of next to its ray march")
    ap.add_argument("--update", default="auto", choices=["auto", "peer", "nvls", "nccl"],
                    help="N > 1: optimiser update as one NVLink peer-memory kernel (peer: P2P loads / stores; nvls: through "
                         "the NVSwitch multicast mapping, reduced in the switch; auto = nvls from 4 GPUs on, else peer) or "
                         "NCCL all-reduce + Adam")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3, 4],
                    help="BASELINE.json configs[] index: 1 = the bench line (default); 3 = LGIE editing step; "
                         "4 = 2^22 table, 1 M rays per step")
    ap.add_argument("--weak", action="store_true", help="--config 4: 1 M rays on every rank instead of 1 M sharded over the ranks")
    ap.add_argument("--peer-at-1", action="store_true", help="N = 1: run the update through the peer-memory kernel too (tuning)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (kernel tuning runs)")
    ap.add_argument("--no-breakdown", action="store_true",
                    help="skip the per-kernel breakdown and the CPU baseline (profiling runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
