"""FusedInference: full-image rendering on the occupancy path without a host round trip per round.

The reference's eval loop (``NeRFRenderer.run_cuda``, nerf/renderer.py:651-688; same structure in
``customnerf_b200/nerf/rendering.py``) reads ``n_alive`` back on every round -- to choose ``n_step`` and to compact
``rays_alive`` with a boolean mask -- so the GPU idles behind the host for ~60 rounds per frame.  Here the round state
(n_alive, n_step, step) lives on the device, every kernel of a round is launched for the worst case and bounded by that
state, one round is captured in a CUDA graph, and the host looks at ``n_alive`` only every ``rounds_per_check`` rounds.
Per ray, the samples and their compositing order are exactly those of the reference loop, so the image is the same
(tests/test_gpu_render.py compares the two paths bit for bit under autocast).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s (code %d)" % (what, L.lib().nb200_error_string(C.c_int(rc)).decode(), rc))


class FusedInference:
    def __init__(self, model, n_rays, T_thresh=1e-4, dt_gamma=0.0, max_steps=1024, rounds_per_check=8, use_graph=True,
                 recorded_march=True):
        if not model.cuda_ray:
            raise RuntimeError("FusedInference drives the occupancy (cuda_ray) path")
        if model.pos_en.input_dim != 3 or model.pos_en.level_dim != 2 or model.pos_en_dim != 32:
            raise RuntimeError("FusedInference needs the reference field shape (D=3, 16 levels x 2 features)")
        self.model, self.lib = model, L.lib()
        self.dev = model.pos_en.embeddings.device
        self.N = N = int(n_rays)
        self.T_thresh, self.dt_gamma, self.max_steps = float(T_thresh), float(dt_gamma), int(max_steps)
        self.rounds_per_check, self.use_graph = int(rounds_per_check), use_graph
        dev = self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = dict(dtype=torch.float16, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.rays_o, self.rays_d = torch.zeros(N, 3, **f32), torch.zeros(N, 3, **f32)
        self.nears, self.fars, self.rays_t, self.noises = (torch.zeros(N, **f32) for _ in range(4))
        self.rays_alive, self.tmp = torch.zeros(N, **i32), torch.zeros(N, **i32)
        M = N + 128
        self.xyzs, self.dirs, self.deltas = torch.zeros(M, 3, **f32), torch.zeros(M, 3, **f32), torch.zeros(M, 2, **f32)
        self.rgba, self.sigma = torch.zeros(M, 4, **f16), torch.zeros(M, **f32)
        self.weights_sum, self.depth, self.image = torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(N, 3, **f32)
        self.state = torch.zeros(8, **i32)
        self.state_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        nb = int(self.lib.nb200_field_weight_image_bytes())
        self.w_fwd = torch.empty(nb, dtype=torch.uint8, device=dev)
        self.w_bwd = torch.empty(nb, dtype=torch.uint8, device=dev)
        self.M_cap = M
        # recorded_march: ONE whole-ray traversal per frame (the training march's speculative-segment kernel) records every
        # ray's sample parameters; a round then looks its next n_step samples up instead of walking the occupancy grid from
        # rays_t again (csrc/raymarching.cu: k_march_rays_rec; same samples bit for bit, rays that leave the recorded chain
        # fall back to the walk -- state[5] counts them)
        self.recorded_march = bool(recorded_march)
        if self.recorded_march:
            self.lib.nb200_march_scratch_ints.restype = C.c_uint32
            self.rec_rays = torch.zeros(N, 3, **i32)
            self.rec_counter = torch.zeros(2, **i32)
            self.rec_scratch = torch.empty(int(self.lib.nb200_march_scratch_ints(L.u32(N))), **i32)
            self.consumed = torch.zeros(N, **i32)
        self.graph = None
        self.use_noise = False
        self.rounds = 0

    def _round(self):
        """one round on the current stream: plan -> march -> encode + field (one kernel) -> composite -> compact"""
        m, enc, lib, st = self.model, self.model.pos_en, self.lib, L.stream()
        p = L.ptr
        cnt = C.c_void_p(self.state.data_ptr() + 12)
        _check(lib.nb200_infer_plan(p(self.state), L.u32(self.N), L.u32(self.max_steps), st), "infer_plan")
        if self.recorded_march:
            _check(lib.nb200_march_rays_rec(p(self.state), L.u32(self.N), p(self.rays_alive), p(self.rays_t), p(self.rays_o),
                                            p(self.rays_d), L.f32(m.bound), L.f32(self.dt_gamma), L.u32(self.max_steps),
                                            L.u32(m.cascade), L.u32(m.grid_size), p(m.density_bitfield), p(self.fars),
                                            p(self.xyzs), p(self.dirs), p(self.deltas),
                                            p(self.noises if self.use_noise else None), p(self.rec_rays), p(self.rec_scratch),
                                            p(self.consumed), st), "march_rays_rec")
        else:
            _check(lib.nb200_march_rays_dev(p(self.state), L.u32(self.N), p(self.rays_alive), p(self.rays_t), p(self.rays_o),
                                            p(self.rays_d), L.f32(m.bound), L.f32(self.dt_gamma), L.u32(self.max_steps),
                                            L.u32(m.cascade), L.u32(m.grid_size), p(m.density_bitfield), p(self.fars),
                                            p(self.xyzs), p(self.dirs), p(self.deltas),
                                            p(self.noises if self.use_noise else None), st), "march_rays_dev")
        # grid gather + field network in one launch: the features never exist in HBM (csrc/field_fused.cu)
        _check(lib.nb200_field_fused_forward(p(self.xyzs), p(self.dirs), L.f32(m.bound), p(enc.embeddings.detach()), p(enc.offsets),
                                             L.u32(enc.num_levels), L.f32(float(np.log2(enc.per_level_scale))),
                                             L.u32(int(enc.base_resolution)), L.u32(enc.gridtype_id), L.i32(int(enc.align_corners)),
                                             L.u32(enc.interp_id), p(self.w_fwd), p(self.sigma), p(None), p(self.rgba), p(None), p(None),
                                             L.u32(self.M_cap), cnt, st), "field_fused_forward")
        _check(lib.nb200_composite_rays_dev(p(self.state), L.u32(self.N), L.f32(self.T_thresh), p(self.rays_alive),
                                            p(self.rays_t), p(self.sigma), p(self.rgba), p(self.deltas), p(self.weights_sum),
                                            p(self.depth), p(self.image), st), "composite_rays_dev")
        _check(lib.nb200_compact_alive(p(self.state), L.u32(self.N), p(self.rays_alive), p(self.tmp), st), "compact_alive")
        L.LAUNCHES += 7

    def _capture(self):
        live = [self.state, self.rays_alive, self.rays_t, self.weights_sum, self.depth, self.image]
        if self.recorded_march:
            live.append(self.consumed)
        keep = [t.clone() for t in live]
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self._round()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._round()
        for t, k in zip(live, keep):
            t.copy_(k)
        self.graph = g

    @torch.no_grad()
    def render(self, rays_o, rays_d, perturb=False, noises=None):
        """-> (weights_sum [N], depth [N], image [N,3], nears, fars); tensors are views of internal buffers"""
        from . import raymarching as rm
        m = self.model
        with torch.cuda.device(self.dev):
            self.rays_o.copy_(rays_o.reshape(-1, 3), non_blocking=True)
            self.rays_d.copy_(rays_d.reshape(-1, 3), non_blocking=True)
            aabb = m.aabb_train if m.training else m.aabb_infer
            _check(self.lib.nb200_near_far_from_aabb(L.ptr(self.rays_o), L.ptr(self.rays_d), L.ptr(aabb), L.u32(self.N),
                                                     L.f32(0.2), L.ptr(self.nears), L.ptr(self.fars), L.stream()), "near_far")
            _check(self.lib.nb200_field_pack_weights(L.ptr(m.network.params.detach()), L.ptr(m.density_network.params.detach()),
                                                     L.ptr(m.rgb_network.params.detach()), L.ptr(self.w_fwd), L.ptr(self.w_bwd),
                                                     L.stream()), "field_pack_weights")
            self.weights_sum.zero_(); self.depth.zero_(); self.image.zero_()
            self.rays_alive.copy_(torch.arange(self.N, dtype=torch.int32, device=self.dev))
            self.rays_t.copy_(self.nears)
            self.state.copy_(torch.tensor([self.N, 0, 0, 0, 0, 0, 0, 0], dtype=torch.int32), non_blocking=True)
            if perturb != self.use_noise:
                self.use_noise, self.graph = bool(perturb), None
            if perturb:
                self.noises.copy_(noises) if noises is not None else self.noises.uniform_()
            if self.recorded_march:         # the frame's one traversal: every ray's sample parameters up to its far point
                self.rec_counter.zero_()
                _check(self.lib.nb200_march_rays_train_count(
                    L.ptr(self.rays_o), L.ptr(self.rays_d), L.ptr(m.density_bitfield), L.f32(m.bound), L.f32(self.dt_gamma),
                    L.u32(self.max_steps), L.u32(self.N), L.u32(m.cascade), L.u32(m.grid_size), L.ptr(self.nears), L.ptr(self.fars),
                    L.ptr(self.noises if self.use_noise else None), L.ptr(self.rec_rays), L.ptr(self.rec_counter),
                    L.ptr(self.rec_scratch), L.stream()), "march_rays_train_count")
                self.consumed.zero_()
                L.LAUNCHES += 3
            if self.use_graph and self.graph is None:
                self._capture()
            self.rounds = 0
            while True:
                for _ in range(self.rounds_per_check):
                    if self.use_graph:
                        self.graph.replay()
                        L.LAUNCHES += 7
                    else:
                        self._round()
                self.rounds += self.rounds_per_check
                self.state_host.copy_(self.state, non_blocking=True)
                torch.cuda.current_stream(self.dev).synchronize()
                if int(self.state_host[0]) <= 0 or self.rounds > self.max_steps + self.rounds_per_check:
                    break
            self.march_fallbacks = int(self.state_host[5])      # rays that left the recorded chain and walked the grid instead
        return self.weights_sum, self.depth, self.image, self.nears, self.fars
