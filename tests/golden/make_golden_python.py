"""Mint golden values from the reference's remaining pure-Python pieces of the hot path, imported UNMODIFIED on the CPU:

  * get_embedder(4)            /root/reference/nerf/base.py:42-77        (view-direction frequency encoding, 27 dims)
  * trunc_exp                  /root/reference/nerf/provider_utils.py:16-29   (forward and its clamped backward)
  * GridEncoder.__init__       /root/reference/gridencoder/grid.py:103-146    (level table: offsets / n_params / output_dim)
  * GridEncoder.forward + _grid_encode   /root/reference/gridencoder/grid.py:24-99,151-168  (the autograd wrapper: [-bound, bound]
                                       -> [0, 1] mapping, prefix shapes, [L,B,C] <-> [B,L*C] permutes, max_level, input gradients)
                                       with the two native entry points it calls served by the C oracle
  * NeRFNetwork.__init__ / forward / density   /root/reference/nerf/network_grid.py:70-193 (+ encoding.py:51-69): the field
                                       network's composition -- tiled 2^21 -> 8192 encoder, trunk, density head + gaussian blob +
                                       trunc_exp, [view embedding | features] -> colour + mask head -- with ``tinycudann.Network``
                                       (absent, un-vendored, unpinned) served by the oracle's MLP restatement and the
                                       encoder's native entry points by the C oracle: what is pinned is the WIRING
  * NeRFRenderer.run           /root/reference/nerf/renderer.py:278-405   (the dense 'non-cuda_ray' renderer -- the path the
                                       CPU baseline of bench.py restates: stratified + importance sampling, sort / gather, LGIE
                                       all / fg / bg composites) on an analytic field, eval and training mode
  * NeRFRenderer.run_cuda, training branch   /root/reference/nerf/renderer.py:597-640,688-716  (the occupancy renderer's glue:
                                       near/far with the default min_near, the step_counter ring, march -> field -> composite,
                                       result dict) with its native ops served by the C oracle and ``Tensor.cuda()`` made the
                                       identity for the duration of the call (the method moves its inputs to the GPU, :603-604)
  * Trainer_Nerf.train_step_pretrain   /root/reference/nerf/utils_init_nerf.py:194-241  (the reconstruction loss: train_rgb * MSE
                                       + train_conf * MSE of the rendered mask, the clamped mask volume) on a canned render result,
                                       with the guidance modules it imports (nerf.sd, nerf.clip, clip, tensorboardX, imageio) stubbed
  * NeRFRenderer.update_extra_state   /root/reference/nerf/renderer.py:1658-1715  (occupancy-grid EMA update, thresholding,
                                       mean_count) with the two native ops it calls -- raymarching.morton3D / packbits,
                                       CUDA-only in the reference -- served by the CPU oracle (oracle/cpu_ops.py), so what
                                       is pinned here is the reference's own torch logic around them.

Absent top-level imports (trimesh, plyfile, skimage, torchtyping, the compiled extension modules) are stubbed; none is
touched by the code exercised.  Run in the build container:  python tests/golden/make_golden_python.py
    -> tests/golden/ref_python.npz
"""
import hashlib
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import cpu_ops  # noqa: E402
from customnerf_b200 import synthetic as syn  # noqa: E402

GRID_CONFIGS = [dict(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048, gridtype="hash"),
                dict(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=21, desired_resolution=8192, gridtype="tiled"),
                dict(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=22, desired_resolution=2048, gridtype="hash"),
                dict(num_levels=8, level_dim=4, base_resolution=8, log2_hashmap_size=15, desired_resolution=512, gridtype="hash",
                     align_corners=True),
                dict(input_dim=2, num_levels=4, level_dim=1, base_resolution=4, log2_hashmap_size=10, per_level_scale=2, gridtype="hash")]


def stub_modules():
    for name in ("trimesh", "plyfile", "skimage", "skimage.measure", "_gridencoder", "_raymarching"):
        sys.modules.setdefault(name, types.ModuleType(name))
    tt = types.ModuleType("torchtyping")
    tt.TensorType = type("TensorType", (), {"__class_getitem__": classmethod(lambda cls, item: cls)})
    sys.modules.setdefault("torchtyping", tt)
    rm = types.ModuleType("raymarching")          # the two native ops update_extra_state calls, served by the CPU oracle
    rm.morton3D = lambda coords: torch.from_numpy(cpu_ops.morton3D(coords.numpy()))
    rm.packbits = lambda grid, thresh, bitfield=None: torch.from_numpy(cpu_ops.packbits(grid.numpy(), thresh))

    def near_far(rays_o, rays_d, aabb, min_near=0.2):
        n, f = cpu_ops.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb.numpy(), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)
    rm.near_far_from_aabb = near_far

    def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                         perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        assert not perturb                   # the wrapper would draw torch.rand(N) on the GPU (raymarching.py:214-217)
        cnt = step_counter.numpy()
        out = cpu_ops.march_rays_train(rays_o.numpy(), rays_d.numpy(), bound, density_bitfield.numpy(), C, H, nears.numpy(),
                                       fars.numpy(), cnt, mean_count, None, align, force_all_rays, dt_gamma, max_steps)
        return tuple(torch.from_numpy(a) for a in out)
    rm.march_rays_train = march_rays_train
    from oracle import torch_ref as _tr
    rm.composite_rays_train = _tr.composite_rays_train
    sys.modules["raymarching"] = rm
    pkg = types.ModuleType("nerf")
    pkg.__path__ = ["/root/reference/nerf"]
    sys.modules["nerf"] = pkg
    gpk = types.ModuleType("gridencoder")
    gpk.__path__ = ["/root/reference/gridencoder"]
    sys.modules["gridencoder"] = gpk


def scene_density(x):
    return {"sigma": syn.bear_density(x)}


def field_forward(x, d):
    """analytic stand-in for NeRFNetwork.forward: sigma [M], (rgb | mask) [M,4], no normals"""
    mask = torch.sigmoid(4.0 * x[:, :1] + d[:, 1:2])
    return syn.bear_density(x), torch.cat([syn.bear_color(x), mask], -1), None


RUN_OPT = dict(bound=2, cuda_ray=False, min_near=0.01, density_thresh=10, train_conf=0.01, soft_mask=True, conf_thr=0.5,
               detach_bg=True, detach_mask_from_field=False)
RUN_KEYS = ("image", "depth", "weights_sum", "render_mask")


def run_rays():
    o, d = syn.camera_rays(12, 16)
    sel = torch.arange(40, 40 + 48)
    return o[sel].contiguous()[None], d[sel].contiguous()[None]


ENC_CFG = dict(input_dim=3, num_levels=8, level_dim=2, base_resolution=8, log2_hashmap_size=12, desired_resolution=128,
               gridtype="hash")


def install_grid_backend(grid):
    """grid._backend.grid_encode_forward / _backward (bindings.cpp:5-7) on CPU tensors through oracle/cpu_ops.py"""
    def fwd(inputs, embeddings, offsets, outputs, B, D, C, L, max_level, S, H, dy_dx, gridtype, align_corners, interp):
        out, dd = cpu_ops.grid_encode_forward(inputs.detach().numpy(), embeddings.detach().numpy(), offsets.numpy(), 2.0 ** S, H,
                                              dy_dx is not None, gridtype, align_corners, interp, max_level)
        outputs.copy_(torch.from_numpy(np.ascontiguousarray(out.reshape(B, L, C).transpose(1, 0, 2))))
        if dy_dx is not None:
            dy_dx.copy_(torch.from_numpy(dd))

    def bwd(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, max_level, S, H, dy_dx, grad_inputs, gridtype,
            align_corners, interp):
        g = np.ascontiguousarray(grad.numpy().transpose(1, 0, 2)).reshape(B, L * C)
        ge, gi = cpu_ops.grid_encode_backward(g, inputs.detach().numpy(), tuple(embeddings.shape), offsets.numpy(), 2.0 ** S, H,
                                              None if dy_dx is None else dy_dx.numpy(), gridtype, align_corners, interp, max_level)
        grad_embeddings.copy_(torch.from_numpy(ge))
        if grad_inputs is not None:
            grad_inputs.copy_(torch.from_numpy(gi))
    def tv(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
        out = cpu_ops.grad_total_variation(inputs.detach().numpy(), embeddings.detach().numpy(), grad.numpy(), offsets.numpy(), weight,
                                           2.0 ** S, H, gridtype, align_corners)
        grad.copy_(torch.from_numpy(out))
    grid._backend.grid_encode_forward, grid._backend.grid_encode_backward = fwd, bwd
    grid._backend.grad_total_variation = tv


def table_fill(rows, C):
    """deterministic table values in [-1, 1) from the entry index (the 2^21 table is too large to store)"""
    i = np.arange(rows * C, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    return (((h >> np.uint64(8)).astype(np.float64) / float(1 << 23)) - 1.0).astype(np.float32).reshape(rows, C)


FIELD_OPT = dict(bound=2, cuda_ray=False, min_near=0.01, density_thresh=10, train_conf=0.01, detach_mask_from_field=False,
                 mask_no_dir=False)


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def main():
    stub_modules()
    base = importlib.import_module("nerf.base")
    pu = importlib.import_module("nerf.provider_utils")
    renderer = importlib.import_module("nerf.renderer")
    grid = importlib.import_module("gridencoder.grid")
    rng = np.random.RandomState(20261017)
    G = {}
    # ---- frequency embedding of the view direction
    d = rng.randn(64, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    emb, out_dim = base.get_embedder(4)
    G["embed_in"], G["embed_out"], G["embed_dim"] = d, emb(torch.from_numpy(d)).numpy(), np.int64(out_dim)
    # ---- trunc_exp
    x = np.concatenate([rng.uniform(-20, 20, 61), [-15.0, 15.0, 16.5]]).astype(np.float32)
    tx = torch.from_numpy(x).requires_grad_()
    y = pu.trunc_exp(tx)
    g = rng.randn(64).astype(np.float32)
    y.backward(torch.from_numpy(g))
    G["texp_in"], G["texp_out"], G["texp_gout"], G["texp_gin"] = x, y.detach().numpy(), g, tx.grad.numpy()
    # ---- level tables
    for i, cfg in enumerate(GRID_CONFIGS):
        enc = grid.GridEncoder(**cfg)
        G["grid%d_offsets" % i] = enc.offsets.numpy().astype(np.int64)
        G["grid%d_meta" % i] = np.array([enc.n_params, enc.output_dim, enc.embeddings.shape[0], enc.embeddings.shape[1]], np.int64)
        G["grid%d_scale" % i] = np.float64(enc.per_level_scale)
        G["grid%d_init_absmax" % i] = np.float32(enc.embeddings.detach().abs().max())
    # ---- the encoder's autograd wrapper (native entry points served by the C oracle)
    install_grid_backend(grid)
    enc = grid.GridEncoder(**ENC_CFG)
    emb = rng.uniform(-1, 1, tuple(enc.embeddings.shape)).astype(np.float32)
    enc.embeddings.data.copy_(torch.from_numpy(emb))
    xin = rng.uniform(-2, 2, (2, 5, 3)).astype(np.float32)
    gout = rng.randn(2, 5, enc.output_dim).astype(np.float32)
    G["enc_embeddings"], G["enc_inputs"], G["enc_gout"] = emb, xin, gout
    tx = torch.from_numpy(xin.copy()).requires_grad_()
    out = enc(tx, bound=2)
    out.backward(torch.from_numpy(gout))
    G["enc_out"], G["enc_grad_inputs"], G["enc_grad_embeddings"] = out.detach().numpy(), tx.grad.numpy(), enc.embeddings.grad.numpy().copy()
    enc.embeddings.grad = None
    out5 = enc(torch.from_numpy(xin.copy()), bound=2, max_level=5)             # levels >= 5 stay zero, no input gradient
    out5.backward(torch.from_numpy(gout))
    G["enc_out_max5"], G["enc_grad_embeddings_max5"] = out5.detach().numpy(), enc.embeddings.grad.numpy().copy()
    # ---- total-variation gradient wrapper (grid.py:171-192): explicit locations in [-bound, bound], then random ones
    enc.embeddings.grad = torch.zeros_like(enc.embeddings)
    tv_x = rng.uniform(-2, 2, (300, 3)).astype(np.float32)
    enc.grad_total_variation(weight=1e-2, inputs=torch.from_numpy(tv_x), bound=2)
    G["tv_inputs"], G["tv_grad_explicit"] = tv_x, enc.embeddings.grad.numpy().copy()
    torch.manual_seed(11)
    enc.grad_total_variation(weight=1e-2, B=500)                  # accumulates on top; draws torch.rand(B, 3)
    G["tv_grad_then_random"] = enc.embeddings.grad.numpy().copy()
    try:
        enc.embeddings.grad = None
        enc.grad_total_variation()
        G["tv_raises_without_grad"] = np.int64(0)
    except ValueError:
        G["tv_raises_without_grad"] = np.int64(1)
    # ---- the field network's wiring (tcnn.Network served by the oracle MLP)
    from oracle import torch_ref
    tc = types.ModuleType("tinycudann")
    tc.Network = torch_ref.Network
    sys.modules["tinycudann"] = tc
    sys.modules["gridencoder"].GridEncoder = grid.GridEncoder          # what gridencoder/__init__.py:1 exports
    ng = importlib.import_module("nerf.network_grid")
    net = ng.NeRFNetwork(types.SimpleNamespace(**FIELD_OPT))
    assert (net.pos_en.gridtype, net.pos_en_dim, tuple(net.pos_en.embeddings.shape)) == ("tiled", 32, (23967296, 2))
    net.pos_en.embeddings.data.copy_(torch.from_numpy(table_fill(*net.pos_en.embeddings.shape)))
    for name in ("network", "density_network", "rgb_network"):
        m = getattr(net, name)
        w = (rng.uniform(-1, 1, m.params.numel()) * 0.35).astype(np.float32)
        m.params.data.copy_(torch.from_numpy(w))
        G["field_params_" + name] = w
    fx = rng.uniform(-1.6, 1.6, (200, 3)).astype(np.float32)
    fx[:20] *= 0.1                                                  # points inside the gaussian blob at the centre
    fd = rng.randn(200, 3).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=-1, keepdims=True)
    with torch.no_grad():
        sig, rad, _ = net(torch.from_numpy(fx), torch.from_numpy(fd))
        dens = net.density(torch.from_numpy(fx))["sigma"]
    G["field_x"], G["field_d"], G["field_sigma"], G["field_radiances"], G["field_density"] = fx, fd, sig.numpy(), rad.numpy(), dens.numpy()
    G["field_param_group_lrs"] = np.array([g["lr"] for g in net.get_params(5e-4)], np.float64)
    del net
    # ---- the dense renderer on an analytic field: eval (deterministic importance sampling) and training (perturbed, random)
    rr = renderer.NeRFRenderer(types.SimpleNamespace(**RUN_OPT))
    rr.density = scene_density
    rr.forward = field_forward
    o, d = run_rays()
    for mode in ("eval", "train"):
        rr.train(mode == "train")
        torch.manual_seed(5)
        res = rr.run(o, d, num_steps=16, upsample_steps=16, perturb=(mode == "train"))
        for key in RUN_KEYS:
            G["run_%s_%s" % (mode, key)] = res[key].detach().numpy()
            G["run_%s_fg_%s" % (mode, key)] = res["fg"][key].detach().numpy()
            G["run_%s_bg_%s" % (mode, key)] = res["bg"][key].detach().numpy()
        G["run_%s_edit_mask" % mode] = res["edit_mask"].detach().numpy()
    # ---- the occupancy renderer's training branch on the bear bit field
    grid_d = syn.density_grid(2, 128)
    thr = min(float(grid_d.mean()), 10.0)
    rc = renderer.NeRFRenderer(types.SimpleNamespace(bound=2, cuda_ray=True, min_near=0.01, density_thresh=10, bg_color=None,
                                                     if_smooth=False))
    rc.density_bitfield = torch.from_numpy(cpu_ops.packbits(grid_d.numpy(), thr))
    rc.forward = lambda x, d, *a, **k: (syn.bear_density(x), syn.bear_color(x) * (0.5 + 0.5 * d[:, :1].abs()), None)
    rc.train()
    o, d = run_rays()
    keep_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for it in range(2):                       # two calls: the step_counter ring advances
            res = rc.run_cuda(o, d, perturb=False, force_all_rays=True)
            for key in ("image", "depth", "weights_sum", "mask"):
                G["runcuda%d_%s" % (it, key)] = res[key].detach().numpy()
        G["runcuda_step_counter"], G["runcuda_local_step"] = rc.step_counter.numpy().copy(), np.int64(rc.local_step)
    finally:
        torch.Tensor.cuda = keep_cuda
    # ---- the reconstruction loss of the trainer
    for name in ("imageio", "tensorboardX", "clip"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sd_stub, clip_stub = types.ModuleType("nerf.sd"), types.ModuleType("nerf.clip")
    sd_stub.StableDiffusion, clip_stub.CLIP = object, object
    sys.modules["nerf.sd"], sys.modules["nerf.clip"] = sd_stub, clip_stub
    utils = importlib.import_module("nerf.utils_init_nerf")
    Npx = 96
    canned = {"image": torch.from_numpy(rng.uniform(0, 1, (1, Npx, 3)).astype(np.float32)).requires_grad_(),
              "weights_sum": torch.from_numpy(np.concatenate([[0.0, 1.0], rng.uniform(0, 1, Npx - 2)]).astype(np.float32)),
              "render_mask": torch.from_numpy(rng.uniform(0, 1, (1, Npx, 1)).astype(np.float32)).requires_grad_()}
    gt_rgb, gt_mask = rng.uniform(0, 1, (Npx, 3)).astype(np.float32), (rng.uniform(0, 1, (Npx, 1)) > 0.5).astype(np.float32)
    for tag, tconf in (("conf", 0.01), ("noconf", 0)):
        topt = types.SimpleNamespace(batch_rays=0, train_rgb=1.0, train_conf=tconf)
        me = types.SimpleNamespace(device="cpu", opt=topt, model=types.SimpleNamespace(render=lambda *a, **k: canned))
        for t in (canned["image"], canned["render_mask"]):
            t.grad = None
        data = (torch.from_numpy(gt_rgb), torch.from_numpy(gt_mask), torch.zeros(Npx, 3), torch.zeros(Npx, 3), 8, 12, "img")
        pred, mvol, loss, ld = utils.Trainer_Nerf.train_step_pretrain(me, data)
        loss.backward()
        G["loss_%s_value" % tag], G["loss_%s_mask_volume" % tag] = np.float64(loss.item()), mvol.detach().numpy()
        G["loss_%s_dict" % tag] = np.array([ld["loss_c"], ld.get("loss_m", -1.0)], np.float64)
        G["loss_%s_grad_image" % tag] = canned["image"].grad.numpy().copy()
        G["loss_%s_grad_mask" % tag] = (canned["render_mask"].grad if canned["render_mask"].grad is not None
                                        else torch.zeros_like(canned["render_mask"])).numpy().copy()
    for k in ("image", "weights_sum", "render_mask"):
        G["loss_in_" + k] = canned[k].detach().numpy()
    G["loss_in_rgb"], G["loss_in_gtmask"] = gt_rgb, gt_mask
    # ---- occupancy-grid update (two consecutive updates: fresh grid, then the EMA-max path), bound 2 -> 2 cascades
    opt = types.SimpleNamespace(bound=2, cuda_ray=True, min_near=0.01, density_thresh=10)
    r = renderer.NeRFRenderer(opt)
    r.density = scene_density
    torch.manual_seed(123)
    r.local_step = 3
    r.step_counter[:3, 0] = torch.tensor([100, 200, 301], dtype=torch.int32)
    for k in range(2):
        r.update_extra_state()
        g_ = r.density_grid.numpy()
        G["occ%d_grid_sha256" % k] = digest(g_)
        G["occ%d_grid_sample" % k] = g_.reshape(-1)[::1009].copy()
        G["occ%d_bitfield_sha256" % k] = digest(r.density_bitfield.numpy())
        G["occ%d_bits_set" % k] = np.int64(np.unpackbits(r.density_bitfield.numpy()).sum())
        G["occ%d_mean_density" % k] = np.float64(r.mean_density)
        G["occ%d_mean_count" % k] = np.int64(r.mean_count)
        G["occ%d_iter_local" % k] = np.array([r.iter_density, r.local_step], np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_python.npz"), **G)
    print("wrote ref_python.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in G.items() if "occ" in k or "meta" in k})


if __name__ == "__main__":
    main()
