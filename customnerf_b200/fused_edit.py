"""FusedEditStep: the LGIE editing step of the occupancy (cuda_ray) path as one replayable CUDA graph
(BASELINE.json configs[3]).

What ``Trainer_Nerf.train_step_editing`` renders (nerf/utils_init_nerf.py:243-265 around ``model.render``): the
foreground-masked local render, the background render and the full-image global render of NeRFRenderer.run
(nerf/renderer.py:383-474) -- here over the occupancy-grid samples, as ``rendering._lgie_composites`` composes them from
the drop-in ops -- with the soft / hard edit mask from the mask head, ``detach_bg`` and ``detach_mask_from_field``.

    forward  : near/far -> march -> encode -> field (tcgen05) -> 3 gated composites   (nb200_train_lgie_forward)
    loss     : the caller's ``loss_fn(outputs) -> scalar`` on the rendered per-ray outputs (the reference's is the
               Stable-Diffusion guidance, out of scope); its gradients w.r.t. the outputs come from torch.autograd on that
               tiny per-ray graph (an SDS ``SpecifyGradient`` node works unchanged)
    backward : 3 composites^T summed per sample -> field^T -> encode^T                 (nb200_train_lgie_backward)
    update   : fused Adam (or the NVLink peer-memory update with ``peer=``)

``outputs`` has the keys and shapes of ``model.render``'s result on this path ([1, N, ...]): image, depth, weights_sum,
render_mask and the same four under 'fg' and 'bg'.
"""
import ctypes as C

import torch

from . import _lib as L
from .fused_trainer import FusedTrainStep, LOSS_SCALE, _check

VARIANTS = ("all", "fg", "bg")


class LgiePlan(C.Structure):
    """mirror of nb200_lgie_plan (include/nerf_b200.h)"""
    _fields_ = ([("conf_thr", C.c_float), ("soft_mask", C.c_int32), ("detach_bg", C.c_int32),
                 ("detach_mask_from_field", C.c_int32)] +
                [(n, C.c_void_p) for n in ("weights_sum", "depth", "image", "render_mask", "g_weights_sum", "g_image",
                                           "g_render_mask")])


class FusedEditStep(FusedTrainStep):
    kernels_per_step = FusedTrainStep.kernels_per_step + 4      # 3 + 3 gated composites instead of 1 + 1

    def __init__(self, model, n_rays, loss_fn, **kw):
        if kw.pop("pipeline_update", False):
            raise RuntimeError("FusedEditStep: pipeline_update is not supported (the loss sits between forward and backward)")
        if getattr(model, "two_heads", False):
            raise RuntimeError("FusedEditStep covers the single colour + mask head; the two-head RGB_network "
                               "(--detach_mask_from_field / --mask_no_dir) renders through NeRFNetwork.render")
        if model.rgb_network.n_output_dims < 4:
            raise RuntimeError("FusedEditStep needs the mask head (opt.train_conf > 0: 4-output colour network)")
        super().__init__(model, n_rays, **kw)
        self.loss_fn = loss_fn
        N, dev = self.N, self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        self.out_ws, self.out_depth = torch.zeros(3, N, **f32), torch.zeros(3, N, **f32)
        self.out_image, self.out_mask = torch.zeros(3, N, 3, **f32), torch.zeros(3, N, **f32)
        self.g_ws, self.g_image, self.g_mask = torch.zeros(3, N, **f32), torch.zeros(3, N, 3, **f32), torch.zeros(3, N, **f32)
        flag = model._flag
        g = LgiePlan()
        g.conf_thr = float(flag("conf_thr", 0.5))
        g.soft_mask, g.detach_bg = int(bool(flag("soft_mask", False))), int(bool(flag("detach_bg", False)))
        g.detach_mask_from_field = int(bool(flag("detach_mask_from_field", False)))
        g.weights_sum, g.depth, g.image, g.render_mask = (t.data_ptr() for t in (self.out_ws, self.out_depth, self.out_image, self.out_mask))
        g.g_weights_sum, g.g_image, g.g_render_mask = (t.data_ptr() for t in (self.g_ws, self.g_image, self.g_mask))
        self.lib.nb200_lgie_plan_bytes.restype = C.c_uint32
        assert C.sizeof(g) == int(self.lib.nb200_lgie_plan_bytes()), "nb200_lgie_plan layout mismatch"
        self.lgie = g

    def outputs(self, leaves=False):
        """the rendered per-ray outputs in ``model.render``'s layout; leaves=True: detached copies that require grad"""
        def pick(v):
            d = {"image": self.out_image[v].view(1, self.N, 3), "depth": self.out_depth[v].view(1, self.N),
                 "weights_sum": self.out_ws[v].view(1, self.N), "render_mask": self.out_mask[v].view(1, self.N, 1)}
            if leaves:
                d = {k: t.detach().clone().requires_grad_(k != "depth") for k, t in d.items()}
            return d
        out = pick(0)
        out["fg"], out["bg"] = pick(1), pick(2)
        return out

    def _loss_and_output_grads(self):
        """loss_fn on leaf copies of the outputs; d(loss * LOSS_SCALE)/d(outputs) into the static gradient buffers"""
        out = self.outputs(leaves=True)
        with torch.enable_grad():
            loss = self.loss_fn(out)
            flat, slots = [], []
            for v, d in enumerate((out, out["fg"], out["bg"])):
                for key, buf in (("weights_sum", self.g_ws), ("image", self.g_image), ("render_mask", self.g_mask)):
                    flat.append(d[key]); slots.append(buf[v])
            # the loss scale is the device-side scaler's current value (a tensor: no host read), or the constant
            scale = LOSS_SCALE if self.scaler is None else self.scaler[0:1].view(torch.float32)[0]
            grads = torch.autograd.grad(loss * scale, flat, allow_unused=True)
        for gr, slot in zip(grads, slots):
            if gr is None:
                slot.zero_()
            else:
                slot.copy_(gr.reshape(slot.shape))
        self.stats[3:4].view(torch.float32).copy_(loss.detach().reshape(1).float())

    def _launch(self, staged=False):
        st = L.stream()
        self._stage(staged)
        if self.perturb:
            self.noises.uniform_()
        _check(self.lib.nb200_train_lgie_forward(C.byref(self.plan), C.byref(self.lgie), st), "train_lgie_forward")
        self._loss_and_output_grads()
        _check(self.lib.nb200_train_lgie_backward(C.byref(self.plan), C.byref(self.lgie), st), "train_lgie_backward")
        self._update(st)
        self.stats_host.copy_(self.stats, non_blocking=True)

    def forward_backward(self):
        """forward + loss + backward only, not captured (tests): gradients accumulate into ``grads_flat``"""
        with torch.cuda.device(self.dev):
            if self.perturb:
                self.noises.uniform_()
            st = L.stream()
            _check(self.lib.nb200_train_lgie_forward(C.byref(self.plan), C.byref(self.lgie), st), "train_lgie_forward")
            self._loss_and_output_grads()
            _check(self.lib.nb200_train_lgie_backward(C.byref(self.plan), C.byref(self.lgie), st), "train_lgie_backward")
            self.stats_host.copy_(self.stats, non_blocking=True)
