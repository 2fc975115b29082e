"""FusedTrainStep: the reconstruction train step of the occupancy (cuda_ray) path as ONE replayable CUDA graph.

Same computation as ``trainer.TrainStep`` (the autograd composition of the drop-in ops, i.e. what
``Trainer_Nerf.train_step_pretrain`` + ``train_one_epoch`` do around ``model.render``,
nerf/utils_init_nerf.py:194-241, 599-629), restructured for the B200:

  * no host synchronisation inside the step: the step's sample count stays on the device (the reference blocks on
    ``counter[0].item()`` every step, raymarching/raymarching.py:225); buffers have a fixed capacity ``m_cap`` rows
    and every kernel bounds its work by the device-side count;
  * no autograd graph, no allocator traffic: fixed buffers, the backward kernels are called directly
    (csrc/fused_step.cu: nb200_train_forward_backward / nb200_train_update);
  * one flat fp32 vector holds [hash table | trunk | density head | colour head]; the module's Parameters are views
    into it, so state-dict names are unchanged and the gradient exchange of the ray-sharded
    multi-GPU step works on one buffer, ``grads_flat``: by default ONE kernel over NVLink peer memory that also takes the
    optimiser step (``peer=``, csrc/peer_update.cu), or a single NCCL all-reduce (``grad_sync=``);
  * zero-grad + unscale + Adam are one pass over that vector (csrc/optim.cu) with the reference's hyper-parameters
    (Adam betas (0.9, 0.99), eps 1e-15, table at 10x LR: main.py:182, network_grid.py:196-206);
  * the whole sequence is captured once and replayed; per-step scalars (step count, learning rate, bias corrections)
    are computed on the device, so a replay reads no host memory.

Capacity: rays whose samples do not fit in ``m_cap`` rows are dropped for that step exactly as the reference drops
them when its ``mean_count`` budget overflows (raymarching.cu:415-416).  ``last_stats()`` returns the step's true
sample count; ``step()`` grows the buffers (and re-captures) when a step overflowed, so at most one step per growth
sees dropped rays.  The default capacity is 1.25x the count measured on the first batch.
"""
import contextlib
import ctypes as C
import math

import numpy as np
import weakref

import torch

from . import _lib as L

LOSS_SCALE = 128.0          # initial loss scale (tiny-cuda-nn's own backward loss scale); dynamic from there on, see below
GROWTH_INTERVAL = 2000      # torch.cuda.amp.GradScaler's default: the scale doubles after this many steps without overflow
# this library's kernels in one step (the ncu launch list profiles/r03_ncu_summary.txt shows the same): march count
# (+ near/far) + scan + finalize + expand, encode, field, composite (+ MSE), composite^T, field^T + weight-gradient reduce,
# encode^T, adam hyper, adam, weight pack (+ scaler commit)
KERNELS_PER_STEP = 13
STAGES = ["march_count", "march_write", "grid_encode_forward", "field_forward", "composite_forward",
          "composite_backward", "field_backward", "grid_encode_backward", "adam", "pack_weights"]


def _release_status_word(lib, stats, dev):
    try:
        with torch.cuda.device(dev):
            lib.nb200_release_kernel_status_word(C.c_void_p(stats[7:8].data_ptr()))
    except Exception:           # interpreter shutdown
        pass


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s (code %d)" % (what, L.lib().nb200_error_string(C.c_int(rc)).decode(), rc))


class TrainPlan(C.Structure):
    """mirror of nb200_train_plan (include/nerf_b200.h)"""
    _u32 = ["N", "M_cap", "C", "H", "L", "base_res", "gridtype", "max_steps"]
    _f32 = ["bound", "dt_gamma", "S", "T_thresh", "min_near", "loss_scale", "inv_n_total"]
    _flags = ["flags"]
    _u64 = ["n_params", "n_table_params"]
    _ptr0 = ["rays_o", "rays_d", "target", "aabb", "noises", "target_mask", "render_mask", "g_render_mask"]
    _f32b = ["mask_weight", "pad1"]
    _ptr = ["bitfield",
            "params_flat", "grads_flat", "exp_avg", "exp_avg_sq", "hyper", "sched", "step",
            "table", "trunk", "density", "rgb", "offsets",
            "g_table", "g_trunk", "g_density", "g_rgb", "w_fwd", "w_bwd",
            "nears", "fars", "weights_sum", "depth", "image", "g_weights_sum", "g_image", "loss",
            "rays", "counter", "m_eff", "scratch",
            "xyzs", "dirs", "deltas", "sigma", "sigma_arg", "d_sigma", "d_rgba",
            "x_en", "rgba", "act", "d_x_en", "wg_scratch", "timer", "scaler"]
    _tail = ["adam_grid", "adam_threads", "adam_unroll", "split_level"]
    _fields_ = ([(n, C.c_uint32) for n in _u32] + [(n, C.c_float) for n in _f32] + [(n, C.c_uint32) for n in _flags] +
                [(n, C.c_uint64) for n in _u64] +
                [(n, C.c_void_p) for n in _ptr0] + [(n, C.c_float) for n in _f32b] + [(n, C.c_void_p) for n in _ptr] +
                [(n, C.c_uint32) for n in _tail] + [("split_elem", C.c_uint64)])


def flat_parameter_count(model):
    return (model.pos_en.embeddings.numel() + model.network.params.numel() + model.density_network.params.numel() +
            model.rgb_network.params.numel())


def flatten_parameters(model, flat=None):
    """Move the four parameter tensors of ``model`` (NeRFNetwork) into one flat fp32 vector
    [table | trunk | density | rgb] (``flat``: use this buffer, e.g. peer-visible memory); the Parameters become views.
    Returns (flat, [(name, offset, numel)])."""
    named = [("pos_en.embeddings", model.pos_en.embeddings), ("network.params", model.network.params),
             ("density_network.params", model.density_network.params), ("rgb_network.params", model.rgb_network.params)]
    total = sum(p.numel() for _, p in named)
    dev = named[0][1].device
    if flat is None:
        flat = torch.empty(total, dtype=torch.float32, device=dev)
    elif flat.numel() != total or flat.dtype != torch.float32 or flat.device != dev:
        raise RuntimeError("flatten_parameters: the buffer must be fp32 [%d] on %s" % (total, dev))
    layout, off = [], 0
    for name, p in named:
        n = p.numel()
        if n % 4:
            raise RuntimeError("parameter %s: numel %d is not a multiple of 4" % (name, n))
        flat[off:off + n].copy_(p.detach().reshape(-1))
        p.data = flat[off:off + n].view(p.shape)
        layout.append((name, off, n))
        off += n
    return flat, layout


class FusedTrainStep:
    kernels_per_step = KERNELS_PER_STEP     # this library's launches in one replay (bench.py: gpu_launches)

    def __init__(self, model, n_rays, lr=5e-4, m_cap=None, world_size=1, grad_sync=None, use_graph=True, perturb=True,
                 betas=(0.9, 0.99), eps=1e-15, T_thresh=1e-4, dt_gamma=0.0, max_steps=1024, lr_decay_base=1.0,
                 lr_decay_iters=0, allreduce_chunks=0, process_group=None, pipeline_update=False, mask_weight=0.0,
                 peer=None, raygen=None, fused_forward=False, dynamic_loss_scale=True, update_shape=(64, 512, 4), split_level=None,
                 rgb_weight=1.0, dense=None):
        # dense = (num_steps, upsample_steps): the DENSE renderer (nerf/renderer.py:278-405, the path the published -O2 commands
        # run) instead of the occupancy march -- per step: near/far, coarse samples (csrc/dense_sampler.cu), their densities
        # by the fused encode + trunk + density-head launch (no autograd, no saves), inverse-cdf importance samples merged with
        # the coarse ones and written as xyzs / dirs / deltas / rays -- then exactly the occupancy path's encode .. encode^T
        # (T_thresh = 0: the dense compositing formula) and the update.  The sample count is the constant N (S + Su).
        self.dense = None if dense is None else (int(dense[0]), int(dense[1]))
        if self.dense is None and not model.cuda_ray:
            raise RuntimeError("FusedTrainStep drives the occupancy (cuda_ray) path; pass dense=(num_steps, upsample_steps) "
                               "for the dense renderer")
        if self.dense is not None:
            if self.dense[0] < 2 or self.dense[1] < 1 or model.pos_en.num_levels != 16:
                raise RuntimeError("FusedTrainStep(dense=(S, Su)): S >= 2, Su >= 1 and the 16-level encoder are required")
            if pipeline_update or peer is not None or raygen is not None:
                raise RuntimeError("FusedTrainStep(dense=...): single stream, single GPU, rays from the caller")
            T_thresh, m_cap = 0.0, int(n_rays) * (self.dense[0] + self.dense[1])
        if model.pos_en.input_dim != 3 or model.pos_en.level_dim != 2 or model.pos_en_dim != 32:
            raise RuntimeError("FusedTrainStep needs the reference field shape (D=3, 16 levels x 2 features)")
        if getattr(model, "two_heads", False):
            raise RuntimeError("FusedTrainStep covers the single colour + mask head; the two-head RGB_network "
                               "(--detach_mask_from_field / --mask_no_dir) trains through trainer.TrainStep")
        self.model = model
        self.lib = L.lib()
        self.dev = model.pos_en.embeddings.device
        self.N = int(n_rays)
        self.lr = float(lr)
        self.betas, self.eps = betas, float(eps)
        self.world_size = world_size
        self.grad_sync = grad_sync          # callable(flat fp32 grad tensor) -> None (sums across ranks in place)
        # allreduce_chunks > 1: the flat gradient is all-reduced (torch.distributed, NCCL) in that many pieces and the
        # Adam sweep of piece i runs while piece i+1 is still on the wire (replaces grad_sync)
        self.allreduce_chunks = int(allreduce_chunks)
        self.process_group = process_group
        # peer (parallel.PeerMemory): parameters and gradient live in NVLink peer-visible memory and the update of the
        # sharded step is ONE kernel (csrc/peer_update.cu): every rank reduces and Adam-updates the slice it owns straight
        # out of the peers' gradients and stores the new parameters into every replica.  Replaces grad_sync (no NCCL call
        # in the step); the moments of slices a rank does not own stay untouched on that rank.
        self.peer = peer
        # raygen = dict(H=, W=, intrinsics=(fx, fy, cx, cy)) with H * W == n_rays: step(pose=, target=) takes the camera
        # pose instead of rays -- the rays are generated by the first kernel of the step (csrc/raygen.cu, get_rays of
        # nerf/provider_utils.py:238-302), so a step's host inputs are 64 B of pose + the target pixels
        self.raygen = raygen
        if raygen is not None and int(raygen["H"]) * int(raygen["W"]) != int(n_rays):
            raise RuntimeError("FusedTrainStep: raygen H x W must equal n_rays")
        if allreduce_chunks > 1:
            dynamic_loss_scale = False          # the chunked all-reduce experiment keeps the constant scale
        if peer is not None and (grad_sync is not None or allreduce_chunks > 1):
            raise RuntimeError("FusedTrainStep: peer replaces grad_sync / allreduce_chunks")
        # pipeline_update: the optimiser update of step k (all-reduce, Adam, weight re-pack) runs on a second stream
        # CONCURRENTLY with the ray march of step k + 1, which reads nothing the update writes -- a memory-bound sweep next
        # to an issue-bound traversal, and at N > 1 the all-reduce hides behind the march.  The parameters then lag one
        # update behind the last step() until flush() (call it before reading the model: eval, update_extra_state,
        # checkpoints).  Same arithmetic, same order per tensor: results equal the unpipelined step's.
        # mask_weight > 0: the reference's reconstruction loss with train_conf (utils_init_nerf.py:224-234):
        # MSE(image, target) + mask_weight * MSE(render_mask, target_mask); render_mask composites the 4th field output
        self.mask_weight = float(mask_weight)
        # rgb_weight = opt.train_rgb (utils_init_nerf.py:224: loss = train_rgb * MSE(rgb) [+ train_conf * MSE(mask)]; 1 in every
        # shipped configuration).  Folded into the kernel's two constants: inv_n carries it, the mask weight is divided by it.
        self.rgb_weight = float(rgb_weight)
        if not self.rgb_weight > 0.0:
            raise NotImplementedError("FusedTrainStep: rgb_weight (opt.train_rgb) must be positive")
        # fused_forward: grid gather + field network forward as ONE kernel (csrc/field_fused.cu: the features are gathered by
        # producer warps straight into the tensor-core operand tile); False (default): two launches (encode, then field).
        # Measured on B200 at configs[1] (profiles/r02c_fused_forward_ncu.txt, r02e_bench_*.json): the fused kernel is bit-identical but SLOWER in the
        # training step (166 us against 55 + 47 us): the MLP's 2 x 112 KB of shared memory leave the gathers ~4 KB of L1 and
        # 8 warps per SM, where the standalone encoder has ~220 KB and ~31
        self.fused_forward = bool(fused_forward) and model.pos_en.num_levels == 16
        # compositing forward + MSE + compositing backward as one launch (csrc/raymarching.cu: k_composite_train_fused); the
        # two-launch form stays selectable (NB200_SPLIT_COMPOSITE=1) for comparison
        import os as _os
        self.fused_composite = _os.environ.get("NB200_SPLIT_COMPOSITE", "0") != "1"
        self.kernels_per_step = KERNELS_PER_STEP - (1 if self.fused_forward else 0) - (1 if self.fused_composite else 0)
        if self.dense is not None:
            pass                                # near/far, coarse, density, importance instead of the march's four launches
        self.pipeline_update = bool(pipeline_update)
        self.update_shape = tuple(int(v) for v in update_shape) if update_shape else None   # (CTAs, threads, unroll) of the pipelined sweep
        # split update (ray-sharded, peer-memory update, pipelined; OFF by default, split_level=10 or NB200_SPLIT_LEVEL=10 turns it
        # on): the table gradient is scattered in two launches -- levels [split_level, 16) first -- and the NVLink update of those
        # levels (+ the MLPs; ~half of the bytes) runs beside the second scatter launch AND the next step's ray march
        # (_launch_split).  Correct (tests/peer_worker.py) but measured slower at 2 GPUs: 0.470 vs 0.449 ms/step
        # (profiles/r03k_split_2gpu.txt) -- the second scatter launch and the L2 traffic the update adds to it cost more than the
        # ~35 us of exposed update they hide.  Within each part rank r owns the r-th 1/world of the part's element range
        # (update_ranges).
        # split_level=None: on (level 8) for tables of 2^25 parameters and more, where the kernels are milliseconds long and the
        # split gains (configs[4], 2^22 rows per level: 12.20 vs 12.62 ms/step at 2 GPUs, 3.48 vs 3.54 at 8 --
        # profiles/r03o_c4_split_2gpu.txt, r03p_c4_strong_8gpu_split*.json); off below that.
        import os
        if "NB200_SPLIT_LEVEL" in os.environ:
            self.split_level = int(os.environ["NB200_SPLIT_LEVEL"])
        elif split_level is None:
            self.split_level = 8 if model.pos_en.embeddings.numel() >= (1 << 25) else 0
        else:
            self.split_level = int(split_level)
        if peer is None or peer.world < 2 or not pipeline_update or not (0 < self.split_level < model.pos_en.num_levels):
            self.split_level = 0
        if self.split_level:
            self.kernels_per_step += 2      # a second scatter launch and a second peer-update launch
        self._pending_update = False
        self._side = None
        self._scattered = None
        self.use_graph = use_graph
        self.perturb = perturb
        self.T_thresh, self.dt_gamma, self.max_steps = float(T_thresh), float(dt_gamma), int(max_steps)
        self.graphs = {}                    # captured steps by input mode: False | True | 'pose' (slot 0), 'rays1' | 'pose1' (slot 1)
        self.overflows = 0
        model.train()

        dev = self.dev
        self.params_flat, self.layout = flatten_parameters(model, None if peer is None else peer.params)
        self.grads_flat = torch.zeros_like(self.params_flat) if peer is None else peer.grads.zero_()
        self.exp_avg = torch.zeros_like(self.params_flat)
        self.exp_avg_sq = torch.zeros_like(self.params_flat)
        self.hyper = torch.zeros(16, dtype=torch.float32, device=dev)
        # {lr0 table (10x, network_grid.py:199), lr0 MLPs, beta1, beta2, eps, 1/loss_scale, LambdaLR base, iters (main.py:189)}
        self.sched = torch.tensor([self.lr * 10.0, self.lr, betas[0], betas[1], self.eps, 1.0 / LOSS_SCALE,
                                   float(lr_decay_base), float(lr_decay_iters)], dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        # dynamic loss scaling with skipped steps, all on the device (torch.cuda.amp.GradScaler as the reference trains with it,
        # utils_init_nerf.py:100,612-629; csrc/adam.cuh): 8 words {scale, growth tracker, iteration, skipped steps, two found-inf
        # flags, growth interval, -}.  The backward kernels raise the flag when a gradient is not finite; the update then leaves
        # p, m, v and the step count alone and halves the scale -- no host synchronisation, captured in the step's graph.  With
        # the peer-memory update the words live in peer-visible memory: any rank's overflow skips the step on every rank.
        self.dynamic_loss_scale = bool(dynamic_loss_scale)
        self.scaler = None
        if self.dynamic_loss_scale:
            self.scaler = peer.scaler if peer is not None else torch.zeros(8, dtype=torch.int32, device=dev)
            self.scaler.zero_()
            self.scaler[0:1].view(torch.float32).fill_(LOSS_SCALE)
            self.scaler[6] = GROWTH_INTERVAL
        nb = int(self.lib.nb200_field_weight_image_bytes())
        self.w_fwd = torch.empty(nb, dtype=torch.uint8, device=dev)
        self.w_bwd = torch.empty(nb, dtype=torch.uint8, device=dev)
        self.lib.nb200_field_wgrad_scratch_bytes.restype = C.c_uint32
        self.wg_scratch = torch.empty(int(self.lib.nb200_field_wgrad_scratch_bytes()) // 4, dtype=torch.float32, device=dev)

        N = self.N
        f32 = dict(dtype=torch.float32, device=dev)
        # a batch lives in ONE device buffer [rays_o | rays_d | target] mirrored by a pinned host staging buffer: a loader
        # that writes into pinned_batch() hands a step its inputs with a single H2D copy (no per-tensor copy calls)
        self.batch_dev = torch.zeros(2, 3, N, 3, **f32)
        # TWO staging slots, host (pinned) and device: the H2D copy of step k + 1's batch runs on a copy stream while step k
        # computes, and the loader fills the other host slot meanwhile
        self.batch_host = torch.zeros(2, 3, N, 3, dtype=torch.float32).pin_memory()
        self.rays_o, self.rays_d, self.target = self.batch_dev[0][0], self.batch_dev[0][1], self.batch_dev[0][2]
        self._copy_stream = None
        self._slot_ready = [torch.cuda.Event(), torch.cuda.Event()]     # H2D into device slot s has landed
        self._slot_free = [torch.cuda.Event(), torch.cuda.Event()]      # the last step that read device slot s has finished
        self.pose_dev = torch.zeros(4, 4, **f32)
        self.pose_host = torch.zeros(2, 4, 4, dtype=torch.float32).pin_memory()
        self.target_mask = torch.zeros(N, **f32)
        self.render_mask, self.g_render_mask = torch.zeros(N, **f32), torch.zeros(N, **f32)
        self.noises = torch.zeros(N, **f32)
        self.nears, self.fars = torch.empty(N, **f32), torch.empty(N, **f32)
        self.weights_sum, self.depth = torch.empty(N, **f32), torch.empty(N, **f32)
        self.image, self.g_image = torch.empty(N, 3, **f32), torch.empty(N, 3, **f32)
        self.g_weights_sum = torch.zeros(N, **f32)
        self.rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        self.scratch = torch.empty(int(self.lib.nb200_march_scratch_ints(L.u32(N))), dtype=torch.int32, device=dev)
        if self.dense is not None:
            S_, Su_ = self.dense
            self.lin = torch.linspace(0.0, 1.0, S_, **f32)
            self.noise_c, self.u_fine = torch.zeros(N, S_, **f32), torch.zeros(N, Su_, **f32)
            self.z_c, self.xyz_c, self.sigma_c = torch.empty(N, S_, **f32), torch.empty(N * S_, 3, **f32), torch.empty(N * S_, **f32)
            self.z_all = torch.empty(N, S_ + Su_, **f32)
        # [counter0, counter1, m_eff, loss bits, peer-update status, max samples over ranks, parked sample count, kernel
        # status (NB200_STATUS_*: a field kernel's bounded mbarrier wait timed out)]: one 32-byte D2H returns everything the
        # host wants to know
        self.stats = torch.zeros(8, dtype=torch.int32, device=dev)
        _check(self.lib.nb200_set_kernel_status_word(C.c_void_p(self.stats[7:8].data_ptr())), "set_kernel_status_word")
        weakref.finalize(self, _release_status_word, self.lib, self.stats, self.dev)    # keeps `stats` alive until then
        self.stats_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        # results of the last two steps, copied out after each step and fenced by an event each: previous_stats() reads step
        # k - 1 while step k runs (no pipeline bubble between steps; last_stats() is the synchronous form)
        self.stats_ring = torch.zeros(2, 8, dtype=torch.int32).pin_memory()
        self._ring_events = [torch.cuda.Event(), torch.cuda.Event()]
        self._steps_launched = 0
        self.peer_plan = None if peer is None else peer.plan(self.layout[0][2], self.exp_avg, self.exp_avg_sq, self.hyper,
                                                             status=self.stats[4:5], use_scaler=self.dynamic_loss_scale)
        self.n_total = N * world_size
        self.l2 = self._setup_l2_persist()
        self._pack()
        self.m_cap = 0
        self._alloc_samples(int(m_cap) if m_cap else 0)

    # ------------------------------------------------------------------------------------------ setup
    def _setup_l2_persist(self):
        """L2 residency of the hash table (north_star kernel 1; csrc/l2_residency.cu): a carve-out of the L2 for persisting
        lines and an access-policy window over the table on the step's stream(s), so that the ~350 MB of activations a step
        streams through the L2 do not evict the table between two encode passes.  NB200_L2_PERSIST=1 turns it on,
        NB200_L2_PERSIST=grad puts the window on the table's gradient instead (the scatter's read-modify-write target).
        Measured on B200 at configs[1] (2^19 table, carve-out limit 47.4 MiB): window on the table: encode forward 51.9 ->
        50.5 us but field forward + 3.6 us and Adam + 3.9 us (the carve-out is taken from everybody's L2): 0.502 vs 0.488
        ms/step; window on the gradient: scatter 130 -> 119 us, encode / field forward + 2.5 / + 3.8 us: 0.4905 ms/step."""
        import os
        mode = os.environ.get("NB200_L2_PERSIST", "0")       # measured (profiles/r02m_l2_persist.txt): no net gain at 2^19 -> off by default
        if mode == "0":
            return None
        n_table = self.layout[0][2]
        base = (self.grads_flat if mode == "grad" else self.params_flat).data_ptr()
        granted, maxwin = C.c_uint64(0), C.c_uint64(0)
        with torch.cuda.device(self.dev):
            rc = self.lib.nb200_l2_persist_limit(C.c_uint64(n_table * 4), C.byref(granted), C.byref(maxwin))
        if rc != 0 or granted.value == 0:
            return None
        win = min(n_table * 4, int(maxwin.value))
        ratio = min(1.0, granted.value / float(win))
        info = {"mode": "table" if mode != "grad" else "grad", "carve_out_bytes": int(granted.value), "window_bytes": int(win),
                "hit_ratio": round(ratio, 3), "base": base}
        self._apply_l2_window(torch.cuda.current_stream(self.dev), info)
        return info

    def _apply_l2_window(self, stream, info=None):
        info = info or self.l2
        if info is None:
            return
        with torch.cuda.device(self.dev):
            self.lib.nb200_stream_access_window(C.c_void_p(stream.cuda_stream), C.c_void_p(info["base"]),
                                                C.c_uint64(info["window_bytes"]), C.c_float(info["hit_ratio"]))

    def _pack(self):
        m = self.model
        with torch.cuda.device(self.dev):
            L.check(self.lib.nb200_field_pack_weights(L.ptr(m.network.params.detach()), L.ptr(m.density_network.params.detach()),
                                                      L.ptr(m.rgb_network.params.detach()), L.ptr(self.w_fwd),
                                                      L.ptr(self.w_bwd), L.stream()), "field_pack_weights")

    def _alloc_samples(self, m_cap):
        dev = self.dev
        if self.split_level and self._pending_update:
            self.flush()                    # the pending coarse scatter reads the buffers that are about to be replaced
        self.graphs = {}
        self.m_cap = m_cap
        if m_cap == 0:
            return
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = dict(dtype=torch.float16, device=dev)
        M = m_cap
        self.xyzs, self.dirs, self.deltas = torch.zeros(M, 3, **f32), torch.zeros(M, 3, **f32), torch.zeros(M, 2, **f32)
        self.sigma, self.sigma_arg, self.d_sigma = torch.zeros(M, **f32), torch.zeros(M, **f32), torch.zeros(M, **f32)
        self.d_rgba = torch.zeros(M, 4, **f32)
        self.x_en, self.d_x_en = torch.zeros(M, 32, **f16), torch.zeros(M, 32, **f16)
        self.rgba = torch.zeros(M, 4, **f16)
        self.act = torch.zeros(5, M, 64, **f16)
        if self.dense is not None:          # the sample count is a constant: counter[0] = m_eff = N (S + Su)
            self.stats[0], self.stats[2] = M, M
        self._fill_plan()

    def _fill_plan(self):
        m, enc = self.model, self.model.pos_en
        p = TrainPlan()
        p.N, p.M_cap, p.C, p.H = self.N, self.m_cap, int(getattr(m, 'cascade', 1)), int(getattr(m, 'grid_size', 128))
        p.L, p.base_res, p.gridtype, p.max_steps = enc.num_levels, int(enc.base_resolution), enc.gridtype_id, self.max_steps
        p.bound, p.dt_gamma = float(m.bound), self.dt_gamma
        p.S = float(np.log2(enc.per_level_scale))
        p.T_thresh, p.min_near = self.T_thresh, 0.2          # run_cuda leaves min_near at its default (Appendix B5)
        p.loss_scale, p.inv_n_total = LOSS_SCALE, self.rgb_weight / (3.0 * self.n_total)
        p.flags = (1 if self.fused_forward else 0) | (0 if self.fused_composite else 4)
        p.n_params, p.n_table_params = self.params_flat.numel(), self.layout[0][2]

        def a(t):
            return None if t is None else t.data_ptr()
        p.rays_o, p.rays_d, p.target, p.aabb = a(self.rays_o), a(self.rays_d), a(self.target), a(m.aabb_train)
        p.noises = a(self.noises) if self.perturb else None
        use_mask = self.mask_weight > 0.0
        p.target_mask = a(self.target_mask) if use_mask else None
        p.render_mask, p.g_render_mask, p.mask_weight = a(self.render_mask), a(self.g_render_mask), self.mask_weight / self.rgb_weight
        p.bitfield = a(getattr(m, 'density_bitfield', None))
        p.params_flat, p.grads_flat, p.exp_avg, p.exp_avg_sq = a(self.params_flat), a(self.grads_flat), a(self.exp_avg), a(self.exp_avg_sq)
        p.hyper, p.sched, p.step = a(self.hyper), a(self.sched), a(self.step_count)
        es = 4
        base, gbase = self.params_flat.data_ptr(), self.grads_flat.data_ptr()
        offs = {name: off for name, off, _ in self.layout}
        p.table, p.g_table = base + es * offs["pos_en.embeddings"], gbase + es * offs["pos_en.embeddings"]
        p.trunk, p.g_trunk = base + es * offs["network.params"], gbase + es * offs["network.params"]
        p.density, p.g_density = base + es * offs["density_network.params"], gbase + es * offs["density_network.params"]
        p.rgb, p.g_rgb = base + es * offs["rgb_network.params"], gbase + es * offs["rgb_network.params"]
        p.offsets = a(enc.offsets)
        p.w_fwd, p.w_bwd = a(self.w_fwd), a(self.w_bwd)
        p.nears, p.fars, p.weights_sum, p.depth, p.image = a(self.nears), a(self.fars), a(self.weights_sum), a(self.depth), a(self.image)
        p.g_weights_sum, p.g_image = a(self.g_weights_sum), a(self.g_image)
        sp = self.stats.data_ptr()
        p.counter, p.m_eff, p.loss = sp, sp + 8, sp + 12
        p.rays, p.scratch = a(self.rays), a(self.scratch)
        p.xyzs, p.dirs, p.deltas = a(self.xyzs), a(self.dirs), a(self.deltas)
        p.sigma, p.sigma_arg, p.d_sigma, p.d_rgba = a(self.sigma), a(self.sigma_arg), a(self.d_sigma), a(self.d_rgba)
        p.x_en, p.rgba, p.act, p.d_x_en = a(self.x_en), a(self.rgba), a(self.act), a(self.d_x_en)
        p.wg_scratch = a(self.wg_scratch)
        p.scaler = a(self.scaler)
        # pipelined update: the Adam sweep runs beside the next step's ray march, as a NARROW grid -- 64 CTAs of 512 threads with
        # 4 float4 groups of every vector in flight pull 3.9 TB/s from 64 SMs and leave the other 84 (and half of those 64)
        # to the march.  Measured on B200 (profiles/r02n_adam_narrow_sweep.txt): 0.464 ms/step against 0.491 with the wide
        # sweep (which the march cannot share an SM with); more than ~72 CTAs and the overlap collapses again.
        if self.pipeline_update and self.update_shape is not None:
            p.adam_grid, p.adam_threads, p.adam_unroll = self.update_shape
        n = self.params_flat.numel()
        self.update_ranges = [(0, n)]
        if self.split_level:
            p.split_level = self.split_level
            p.split_elem = 2 * int(enc.offsets[self.split_level])
            self.update_ranges = [(int(p.split_elem), n), (0, int(p.split_elem))]
        assert C.sizeof(p) == int(self.lib.nb200_train_plan_bytes()), "nb200_train_plan layout mismatch"
        self.plan = p

    def measure_samples(self, rays_o, rays_d):
        """Samples the occupancy grid yields for this ray batch (one synchronous count pass; used to size m_cap)."""
        from . import raymarching as rm
        m = self.model
        o, d = rays_o.to(self.dev).float().contiguous(), rays_d.to(self.dev).float().contiguous()
        nears, fars = rm.near_far_from_aabb(o, d, m.aabb_train)
        counter = torch.zeros(2, dtype=torch.int32, device=self.dev)
        rays = torch.empty(o.shape[0], 3, dtype=torch.int32, device=self.dev)
        L.check(self.lib.nb200_march_rays_train_count(
            L.ptr(o), L.ptr(d), L.ptr(m.density_bitfield), L.f32(m.bound), L.f32(self.dt_gamma), L.u32(self.max_steps),
            L.u32(o.shape[0]), L.u32(m.cascade), L.u32(m.grid_size), L.ptr(nears), L.ptr(fars), L.ptr(None), L.ptr(rays),
            L.ptr(counter), L.ptr(self.scratch), L.stream()), "count")
        return int(counter[0].item())

    def _max_over_ranks(self, n):
        """the sample-buffer capacity is the same on every rank of a ray-sharded job: growth (which drops the captured graph
        and re-runs warm-up steps that contain the collective / peer update) must happen on all ranks at the same step"""
        if self.world_size > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                t = torch.tensor([int(n)], dtype=torch.int64, device=self.dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.process_group if self.peer is None else self.peer.group)
                return int(t[0])
        return int(n)

    @staticmethod
    def _round_cap(samples):
        return max(4096, int(math.ceil(samples * 1.25 / 4096.0)) * 4096)

    # ------------------------------------------------------------------------------------------ the step
    # the captured graphs by their historical names (tests, bench teardown)
    graph = property(lambda self: self.graphs.get(False), lambda self, v: self.graphs.clear() if v is None else self.graphs.__setitem__(False, v))
    graph_staged = property(lambda self: self.graphs.get("rays1"))
    graph_pose = property(lambda self: self.graphs.get("pose"))

    def pinned_batch(self, slot=0):
        """(rays_o, rays_d, target): [N,3] views of pinned host staging slot ``slot`` (0 or 1).  Fill them in place and call
        ``step(*pinned_batch(slot))``: ONE H2D copy of the whole batch is issued on a copy stream into the matching device
        slot (it overlaps whatever the device is still computing) and the step's graph waits for it.  Alternate the slots
        when steps are issued without waiting for the one before (``previous_stats()``)."""
        b = self.batch_host[slot]
        return b[0], b[1], b[2]

    def pinned_pose_batch(self, slot=0):
        """(pose [4,4], target [N,3]): views of pinned host staging slot ``slot`` for ``step(pose=, target=)``"""
        return self.pose_host[slot], self.batch_host[slot][2]

    def _generate_rays(self):
        r = self.raygen
        fx, fy, cx, cy = [float(v) for v in r["intrinsics"]]
        L.check(self.lib.nb200_get_rays(L.ptr(self.pose_dev), L.f32(fx), L.f32(fy), L.f32(cx), L.f32(cy), L.u32(int(r["H"])),
                                        L.u32(int(r["W"])), L.u32(1), L.u32(self.N), L.ptr(None), L.f32(0.5), L.f32(0.5),
                                        L.ptr(self.rays_o), L.ptr(self.rays_d), L.stream()), "get_rays")

    def _stage(self, staged):
        """host -> device copies (and ray generation) that head a staged step; staged: False | True (rays) | 'pose' """
        if staged in ("pose", "pose1"):     # ray batches (True / 'rays1') are copied ahead of the step by _prefetch
            slot = 1 if staged == "pose1" else 0
            self.pose_dev.copy_(self.pose_host[slot], non_blocking=True)
            self.target.copy_(self.batch_host[slot][2], non_blocking=True)
            self._generate_rays()

    def _prefetch(self, slot):
        """H2D copy of pinned host slot -> device slot on the copy stream; the current stream waits for it"""
        main = torch.cuda.current_stream(self.dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        cs = self._copy_stream
        cs.wait_event(self._slot_free[slot])
        with torch.cuda.stream(cs):
            self.batch_dev[slot].copy_(self.batch_host[slot], non_blocking=True)
            self._slot_ready[slot].record(cs)
        main.wait_event(self._slot_ready[slot])

    @contextlib.contextmanager
    def _plan_slot(self, slot):
        """the plan with its ray / target pointers on device slot ``slot`` (what a launch or a capture inside sees)"""
        if slot == 0:
            yield
            return
        p, b = self.plan, self.batch_dev[1]
        keep = (p.rays_o, p.rays_d, p.target)
        p.rays_o, p.rays_d, p.target = b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr()
        try:
            yield
        finally:
            p.rays_o, p.rays_d, p.target = keep

    def _update(self, st):
        if self.allreduce_chunks > 1:
            self._pipelined_allreduce_update(st)
        else:
            if self.peer_plan is not None:
                if self.split_level:
                    self._update_part(1, st)
                    self._update_part(2, st)
                    return
                _check(self.lib.nb200_train_update_peer(C.byref(self.plan), C.byref(self.peer_plan), st), "train_update_peer")
                return
            if self.grad_sync is not None:
                self.grad_sync(self.grads_flat)
                if self.scaler is not None and self.world_size > 1:
                    # an overflow on any rank skips the step on every rank: the found-inf flags travel with the gradient
                    import torch.distributed as dist
                    dist.all_reduce(self.scaler[4:6], op=dist.ReduceOp.MAX, group=self.process_group)
            _check(self.lib.nb200_train_update(C.byref(self.plan), st), "train_update")

    def _update_part(self, part, st):
        _check(self.lib.nb200_train_update_peer_part(C.byref(self.plan), C.byref(self.peer_plan), C.c_int(part), st),
               "train_update_peer_part")

    def _launch_split(self, staged, st):
        """one step of the split, pipelined, ray-sharded form.  The table gradient is scattered in two launches -- levels
        [split_level, L) with the rest of the backward (REST_A), the coarse levels afterwards (REST_B) -- and the NVLink update
        of the fine levels + MLPs (part 1, ~all of the bytes) starts as soon as REST_A and the hyper kernel are through.  A
        replay is cut AFTER the hyper kernel, so that part 1 of step k runs beside BOTH the coarse scatter of step k and the
        ray march of step k + 1:
              side:  update part 1 (k)                          -> [after REST_B] update part 2 (k) + re-pack + commit
              main:  REST_B (k: coarse scatter) -> march (k + 1)                                   -> join
                     REST_A (k + 1: encode .. field^T, fine scatter) -> hyper kernel (k + 1)
        REST_B stays on the main stream in front of the march (it reads the sample buffers the march overwrites)."""
        main = torch.cuda.current_stream(self.dev)
        if self._side is None:
            import os
            prio = -1 if os.environ.get("NB200_SIDE_PRIORITY", "1") == "1" else 0
            self._side = torch.cuda.Stream(device=self.dev, priority=prio)
            self._apply_l2_window(self._side)
        if self._scattered is None:
            self._scattered = torch.cuda.Event()
        side = self._side
        side.wait_stream(main)
        self.plan.flags |= 2                # the hyper kernel of the pending step ran at the end of its own launch
        with torch.cuda.stream(side):
            self._update_part(1, L.stream())
        _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(8), st), "train_phase(rest B)")
        self._scattered.record(main)
        _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(1), st), "train_phase(march)")
        with torch.cuda.stream(side):
            side.wait_event(self._scattered)
            self._update_part(2, L.stream())
        self.plan.flags &= ~2
        main.wait_stream(side)
        _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(4), st), "train_phase(rest A)")
        _check(self.lib.nb200_train_update_hyper(C.byref(self.plan), C.c_int(1), st), "train_update_hyper")

    def _launch(self, staged=False):
        """every device-side action of one step, on the current stream (this is what the graph captures)"""
        st = L.stream()
        self._stage(staged)
        if self.perturb:
            self.noises.uniform_()
        if self.dense is not None:
            self._launch_dense_sampler(st)
            _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(2), st), "train_phase(rest)")
            self._update(st)
        elif not self.pipeline_update:
            _check(self.lib.nb200_train_forward_backward(C.byref(self.plan), st), "train_forward_backward")
            self._update(st)
        elif self.split_level:
            self._launch_split(staged, st)
        else:
            # [update of the previous step] on the side stream  ||  [march of this step] here, then join
            main = torch.cuda.current_stream(self.dev)
            if self._side is None:
                # high priority: when both branches are runnable the update's (thin) grid is placed first and the march fills
                # the rest of every SM -- launched the other way round, the march's single resident wave leaves the update no
                # room until it drains and the two run back to back
                import os
                prio = -1 if os.environ.get("NB200_SIDE_PRIORITY", "1") == "1" else 0
                self._side = torch.cuda.Stream(device=self.dev, priority=prio)
                self._apply_l2_window(self._side)
            # the update's first (one-thread) kernel runs here, BEFORE the fork: the sweep and the march then become runnable
            # at the same moment and the high-priority side stream gets its (thin) grid resident first -- forked earlier, the
            # march's single wave takes every SM while that little kernel runs and the sweep waits for the wave to drain
            hyper_first = self.allreduce_chunks <= 1 and self.grad_sync is None
            if hyper_first:
                _check(self.lib.nb200_train_update_hyper(C.byref(self.plan), C.c_int(1 if self.peer_plan is not None else 0), st),
                       "train_update_hyper")
                self.plan.flags |= 2
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self._update(L.stream())
            self.plan.flags &= ~2
            _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(1), st), "train_phase(march)")
            main.wait_stream(self._side)
            _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(2), st), "train_phase(rest)")
        self.stats_host.copy_(self.stats, non_blocking=True)

    def _launch_dense_sampler(self, st):
        """the dense renderer's sampler (nerf/renderer.py:297-367) in five launches, writing the rows the rest of the step reads;
        random draws in the reference's order: rand(N, S) for the stratified jitter, rand(N, Su) inside sample_pdf"""
        p, m, lib = self.plan, self.model, self.lib
        S_, Su_ = self.dense
        V, N = C.c_void_p, C.c_uint32(self.N)
        if self.perturb:
            self.noise_c.uniform_()
        self.u_fine.uniform_()              # training mode: sample_pdf draws u per ray (det=False)
        L.LAUNCHES += 2 if self.perturb else 1
        _check(lib.nb200_near_far_from_aabb(V(p.rays_o), V(p.rays_d), V(p.aabb), N, C.c_float(float(m.min_near)), V(p.nears),
                                            V(p.fars), st), "near_far_from_aabb")
        _check(lib.nb200_dense_coarse(V(p.rays_o), V(p.rays_d), V(p.nears), V(p.fars), V(p.aabb), V(self.lin.data_ptr()),
                                      V(self.noise_c.data_ptr()) if self.perturb else None, N, C.c_uint32(S_),
                                      V(self.z_c.data_ptr()), V(self.xyz_c.data_ptr()), st), "dense_coarse")
        _check(lib.nb200_field_fused_forward(V(self.xyz_c.data_ptr()), None, C.c_float(p.bound), V(p.table), V(p.offsets),
                                             C.c_uint32(p.L), C.c_float(p.S), C.c_uint32(p.base_res), C.c_uint32(p.gridtype),
                                             C.c_int(0), C.c_uint32(0), V(p.w_fwd), V(self.sigma_c.data_ptr()), None, None, None,
                                             None, C.c_uint32(self.N * S_), None, st), "field_fused_forward(density)")
        _check(lib.nb200_dense_importance(V(p.rays_o), V(p.rays_d), V(p.nears), V(p.fars), V(p.aabb), V(self.z_c.data_ptr()),
                                          V(self.sigma_c.data_ptr()), V(self.u_fine.data_ptr()), C.c_int(1), N, C.c_uint32(S_),
                                          C.c_uint32(Su_), V(self.z_all.data_ptr()), V(p.xyzs), V(p.dirs), V(p.deltas), V(p.rays),
                                          st), "dense_importance")

    def scaler_state(self):
        """(loss scale, skipped steps, optimiser steps taken) -- synchronises with the device"""
        if self.scaler is None:
            return LOSS_SCALE, 0, int(self.step_count)
        h = self.scaler.cpu()
        return float(h[0:1].view(torch.float32)[0]), int(h[3]), int(self.step_count)

    def flush(self):
        """pipeline_update: apply the update of the last step() now, so that the parameters are current"""
        if self.pipeline_update and self._pending_update:
            with torch.cuda.device(self.dev):
                if self.split_level and self._pending_update == "split":
                    # the step ended after its hyper kernel: coarse scatter, then both halves of the update
                    _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(8), L.stream()), "train_phase(rest B)")
                    self.plan.flags |= 2
                    self._update_part(1, L.stream())
                    self._update_part(2, L.stream())
                    self.plan.flags &= ~2
                else:
                    self._update(L.stream())
            self._pending_update = False

    def _pipelined_allreduce_update(self, st):
        """all-reduce(sum) of grads_flat in pieces on NCCL's stream, each piece's Adam sweep as soon as it has arrived"""
        import torch.distributed as dist
        n, n_table = self.params_flat.numel(), self.layout[0][2]
        per = ((n + self.allreduce_chunks - 1) // self.allreduce_chunks + 1023) // 1024 * 1024
        ranges = [(a, min(n, a + per)) for a in range(0, n, per)]
        works = [dist.all_reduce(self.grads_flat[a:b], op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
                 for a, b in ranges]
        p = self.plan
        _check(self.lib.nb200_adam_hyper(C.c_void_p(p.step), C.c_void_p(p.sched), C.c_void_p(p.hyper), st), "adam_hyper")
        for (a, b), wk in zip(ranges, works):
            wk.wait()                       # the current stream waits for this piece only
            split = min(max(n_table - a, 0), b - a)
            _check(self.lib.nb200_fused_adam(C.c_void_p(p.params_flat + 4 * a), C.c_void_p(p.grads_flat + 4 * a),
                                             C.c_void_p(p.exp_avg + 4 * a), C.c_void_p(p.exp_avg_sq + 4 * a),
                                             C.c_uint64(b - a), C.c_uint64(split), C.c_void_p(p.hyper), C.c_int(1), st),
                   "fused_adam")
        _check(self.lib.nb200_field_pack_weights(C.c_void_p(p.trunk), C.c_void_p(p.density), C.c_void_p(p.rgb),
                                                 C.c_void_p(p.w_fwd), C.c_void_p(p.w_bwd), st), "field_pack_weights")

    def _capture(self, staged=False):
        """warm up on a side stream (first-call cudaFuncSetAttribute, allocator, RNG registration), capture one step,
        then restore the optimiser state the warm-up steps advanced"""
        # (hyper too: in the split form the second half of an update uses the hyper-parameters its first half computed one launch earlier)
        state = (self.params_flat, self.exp_avg, self.exp_avg_sq, self.step_count, self.hyper) + (() if self.scaler is None else (self.scaler,))
        if self.pipeline_update:            # the gradient of the step before is still waiting for its update: keep it
            state = state + (self.grads_flat,)
        keep = [t.clone() for t in state]
        try:
            s = torch.cuda.Stream(device=self.dev)
            self._apply_l2_window(s)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._launch(staged)
            torch.cuda.current_stream(self.dev).wait_stream(s)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            # thread_local: NCCL's watchdog thread may touch the CUDA API while the all-reduce of a sharded step is captured
            # the capture stream carries the L2 access-policy window: kernel nodes inherit it from the stream they are captured on
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                self._launch(staged)
            self.graphs[staged] = g
        finally:
            torch.cuda.synchronize(self.dev)
            for t, k in zip(state, keep):
                t.copy_(k)
            if not self.pipeline_update:
                self.grads_flat.zero_()
            self._pack()

    def forward_backward(self):
        """forward + backward only, not captured (tests): gradients accumulate into ``grads_flat``"""
        with torch.cuda.device(self.dev):
            if self.dense is not None:
                self._launch_dense_sampler(L.stream())
                _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(2), L.stream()), "train_phase(rest)")
            else:
                if self.perturb:
                    self.noises.uniform_()
                _check(self.lib.nb200_train_forward_backward(C.byref(self.plan), L.stream()), "train_forward_backward")
            self.stats_host.copy_(self.stats, non_blocking=True)

    def profile_stages(self, n_steps=10, flush=None):
        """Device time of every stage of ``n_steps`` real (non-captured) train steps, measured with CUDA events recorded
        between the stages on the launching stream.  Returns {stage: mean microseconds}.  ``flush``: optional callable
        run before each step (e.g. an L2-flushing memset)."""
        timer = C.c_void_p()
        _check(self.lib.nb200_stage_timer_create(C.byref(timer)), "stage_timer_create")
        out = (C.c_float * len(STAGES))()
        acc = np.zeros(len(STAGES))
        try:
            self.plan.timer = timer.value
            with torch.cuda.device(self.dev):
                for _ in range(n_steps):
                    if flush is not None:
                        flush()
                    self._launch()
                    L.LAUNCHES += self.kernels_per_step
                    _check(self.lib.nb200_stage_timer_read(timer, out), "stage_timer_read")
                    acc += np.array(list(out))
        finally:
            self.plan.timer = None
            self.lib.nb200_stage_timer_destroy(timer)
        return {k: float(v / n_steps) for k, v in zip(STAGES, acc)}

    def set_batch(self, rays_o, rays_d, target, target_mask=None):
        """copy one ray batch (host -- ideally pinned -- or device tensors) into the step's static input buffers"""
        self.rays_o.copy_(rays_o.reshape(-1, 3), non_blocking=True)
        self.rays_d.copy_(rays_d.reshape(-1, 3), non_blocking=True)
        self.target.copy_(target.reshape(-1, 3), non_blocking=True)
        if target_mask is not None:
            self.target_mask.copy_(target_mask.reshape(-1), non_blocking=True)

    def step(self, rays_o=None, rays_d=None, target=None, target_mask=None, pose=None):
        """One train step on the current stream.  Never synchronises; ``last_stats()`` reads the result back.
        ``pose`` (needs ``raygen``): camera-to-world [4,4] instead of rays; with the tensors of ``pinned_pose_batch()`` the
        copies and the ray generation are nodes of the step's graph."""
        with torch.cuda.device(self.dev):
            pose_staged = False
            if pose is not None:
                if self.raygen is None:
                    raise RuntimeError("FusedTrainStep.step(pose=...): construct the step with raygen=dict(H, W, intrinsics)")
                pose_staged = False
                for slot, name in ((0, "pose"), (1, "pose1")):
                    if (not pose.is_cuda and pose.data_ptr() == self.pose_host[slot].data_ptr() and target is not None
                            and target.data_ptr() == self.batch_host[slot][2].data_ptr()):
                        pose_staged = name
                if not pose_staged:
                    self.pose_dev.copy_(pose.reshape(4, 4), non_blocking=True)
                    if target is not None:
                        self.target.copy_(target.reshape(-1, 3), non_blocking=True)
                    self._generate_rays()
                    L.LAUNCHES += 1
                elif self.m_cap == 0:
                    self._stage(pose_staged)  # the capacity measurement below needs this camera's rays
                rays_o = rays_d = target = None
            if self.m_cap == 0:
                n0 = self.measure_samples(rays_o if rays_o is not None else self.rays_o,
                                          rays_d if rays_d is not None else self.rays_d)
                self._alloc_samples(self._round_cap(self._max_over_ranks(n0)))
            # a batch handed over in the pinned staging buffer is copied by the graph itself (one H2D node)
            staged = False
            if rays_o is not None and not rays_o.is_cuda:
                for slot, name in ((0, True), (1, "rays1")):
                    b = self.batch_host[slot]
                    if (rays_o.data_ptr() == b[0].data_ptr() and rays_d.data_ptr() == b[1].data_ptr()
                            and target.data_ptr() == b[2].data_ptr()):
                        staged = name
            if pose_staged:
                staged = pose_staged
            if rays_o is not None and not staged:
                self.set_batch(rays_o, rays_d, target, target_mask)
            elif target_mask is not None:
                self.target_mask.copy_(target_mask.reshape(-1), non_blocking=True)
            slot = -1
            if staged is True or staged == "rays1":
                slot = 1 if staged == "rays1" else 0
                self._prefetch(slot)
                staged = "rays1" if slot else False       # slot 0 is where the resident batch lives: same graph
            with self._plan_slot(max(slot, 0)):
                self._step_staged(staged)
            if slot >= 0:
                self._slot_free[slot].record()
            self._publish_stats()

    def _try_capture(self, staged):
        try:
            self._capture(staged)
        except Exception as e:      # e.g. a collective that cannot be captured on this NCCL build
            import warnings
            warnings.warn("FusedTrainStep: CUDA-graph capture failed (%s); launching the step's kernels "
                          "directly instead" % (e,))
            self.use_graph = False
            self.graphs.clear()
            torch.cuda.synchronize(self.dev)

    def _step_staged(self, staged):
        need_graph = self.use_graph and self.graphs.get(staged) is None
        if self.split_level and need_graph and self._pending_update:
            # the split form's graph starts with the second half of the step before: capture only from a clean state (its
            # warm-up launches would otherwise scatter the pending coarse levels from their own samples)
            self.flush()
        if self.pipeline_update and not self._pending_update:
            if self.split_level and need_graph:
                self._try_capture(staged)   # warm-ups run on a zero gradient; every piece of state they touch is restored
            # first step of a pipelined run: nothing to update yet -- forward + backward only, launched directly
            self._stage(staged)
            if self.dense is not None:
                self._launch_dense_sampler(L.stream())
                _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(2), L.stream()), "train_phase(rest)")
            elif self.split_level:          # the split form ends a step after REST_A + the hyper kernel (see _launch_split)
                if self.perturb:
                    self.noises.uniform_()
                _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(1), L.stream()), "train_phase(march)")
                _check(self.lib.nb200_train_phase(C.byref(self.plan), C.c_int(4), L.stream()), "train_phase(rest A)")
                _check(self.lib.nb200_train_update_hyper(C.byref(self.plan), C.c_int(1), L.stream()), "train_update_hyper")
            else:
                if self.perturb:
                    self.noises.uniform_()
                _check(self.lib.nb200_train_forward_backward(C.byref(self.plan), L.stream()), "train_forward_backward")
            self.stats_host.copy_(self.stats, non_blocking=True)
            self._pending_update = "split" if self.split_level else True
            L.LAUNCHES += self.kernels_per_step - 4
            return
        if self.use_graph and self.graphs.get(staged) is None:
            self._try_capture(staged)
        if self.use_graph:
            self.graphs[staged].replay()
        else:
            self._launch(staged)
        L.LAUNCHES += self.kernels_per_step + (1 if staged in ("pose", "pose1") else 0)

    def _publish_stats(self):
        i = self._steps_launched & 1
        self.stats_ring[i].copy_(self.stats, non_blocking=True)
        self._ring_events[i].record()
        self._steps_launched += 1

    def _parse_stats(self, s):
        if int(s[4]):
            raise RuntimeError("FusedTrainStep: the peer-memory update timed out waiting for another rank (status %d): "
                               "a rank left the job or launched fewer steps" % int(s[4]))
        if int(s[7]):
            raise RuntimeError("FusedTrainStep: a field kernel gave up waiting for its tensor-core work (status 0x%x: 1 forward, "
                               "2 backward, 4 fused forward) -- the results of this step are invalid" % int(s[7]))
        return float(s[3:4].view(torch.float32)[0]), int(s[0]), int(s[2])

    def _needed_rows(self, s, samples):
        """rows the sample buffers must hold.  One rank: this step's count.  Ray-sharded: ONLY the maximum over all ranks that
        the update published in the stats block (the same number on every rank, csrc/adam.cuh status words), so that every
        rank grows -- and re-captures, warm-up updates included -- at the same step."""
        if self.world_size > 1 and self.scaler is not None:
            return int(s[5])
        return samples

    def previous_stats(self):
        """(loss, samples, rows_used) of the step BEFORE the most recent ``step()`` -- waits for that step only, so the
        device never idles between steps (issue step k + 1, then read step k).  None before the second step.  A step that
        overflowed ``m_cap`` is noticed one step late: the buffers grow before the next ``step()``."""
        if self._steps_launched < 2:
            return None
        i = (self._steps_launched - 2) & 1
        self._ring_events[i].synchronize()
        out = self._parse_stats(self.stats_ring[i])
        need = self._needed_rows(self.stats_ring[i], out[1])
        if need > self.m_cap:
            torch.cuda.current_stream(self.dev).synchronize()
            self.overflows += 1
            self._alloc_samples(self._round_cap(need))
        return out

    # ------------------------------------------------------------------------------------------ checkpoints
    def optimizer_state_dict(self):
        """The optimiser state in ``torch.optim.Adam.state_dict()`` form for ``Adam(model.get_params(lr), betas, eps)`` --
        the object the reference trains with and stores under ``state['optimizer']`` (main.py:182,
        utils_init_nerf.py:795) -- so that its checkpoints carry over: four parameters in get_params order
        (network_grid.py:196-206), 'step' / 'exp_avg' / 'exp_avg_sq' each.  Applies a pending pipelined update first; with
        the peer-memory update (moments kept only where owned) the slices are gathered from the ranks (a collective call:
        every rank must make it)."""
        self.flush()
        m, v = self.exp_avg, self.exp_avg_sq
        if self.peer is not None and self.peer.world > 1:
            from . import parallel
            m, v = m.clone(), v.clone()
            for a, b in self.update_ranges:          # rank r owns the r-th 1/world of every update range
                parallel.gather_owned_slices(m[a:b], self.peer.world, self.peer.rank, self.peer.group)
                parallel.gather_owned_slices(v[a:b], self.peer.world, self.peer.rank, self.peer.group)
        t = float(int(self.step_count))
        params = dict(self.model.named_parameters())
        state, groups = {}, []
        for i, (name, off, n) in enumerate(self.layout):
            shape = params[name].shape
            state[i] = {"step": torch.tensor(t), "exp_avg": m[off:off + n].view(shape).clone(),
                        "exp_avg_sq": v[off:off + n].view(shape).clone()}
            groups.append({"lr": self.lr * (10.0 if i == 0 else 1.0), "betas": tuple(self.betas), "eps": self.eps,
                           "weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                           "differentiable": False, "fused": None, "decoupled_weight_decay": False, "params": [i]})
        return {"state": state, "param_groups": groups}

    def load_optimizer_state_dict(self, sd):
        """inverse of ``optimizer_state_dict`` (also takes the state dict of the reference's own Adam)"""
        self.flush()
        steps = set()
        for i, (name, off, n) in enumerate(self.layout):
            st = sd["state"].get(i)
            if st is None:                       # a parameter that never received a gradient: no state yet
                self.exp_avg[off:off + n].zero_(); self.exp_avg_sq[off:off + n].zero_()
                continue
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise RuntimeError("load_optimizer_state_dict: the parameters disagree on the step count: %r" % sorted(steps))
        self.step_count.fill_(steps.pop() if steps else 0)

    def last_stats(self):
        """(loss, samples, rows_used) of the most recent step -- synchronises with the device.  Grows the sample
        buffers (and drops the captured graph) when that step overflowed ``m_cap``."""
        torch.cuda.current_stream(self.dev).synchronize()
        loss, samples, used = self._parse_stats(self.stats_host)
        need = self._needed_rows(self.stats_host, samples)
        if need > self.m_cap:
            self.overflows += 1
            self._alloc_samples(self._round_cap(need))
        return loss, samples, used
