"""Oracle (oracle/) and the product's host-side Python (customnerf_b200/) against golden values minted from the reference's
own Python code, imported unmodified by tests/golden/make_golden_python.py: the view-direction frequency encoding
(nerf/base.py:42-77), trunc_exp forward / backward (nerf/provider_utils.py:16-29), the GridEncoder level table
(gridencoder/grid.py:103-146) and the occupancy-grid update (nerf/renderer.py:1658-1715).

Tolerances: integer tables, bit fields and counters exact; fp32 values rel 1e-6 (same torch ops on the same CPU)."""
import hashlib
import os
import types

import numpy as np
import torch

from oracle import cpu_ops, torch_ref

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_python.npz"))
from golden.make_golden_python import GRID_CONFIGS  # noqa: E402  (the configurations the vectors were minted for)


def test_frequency_embedding_matches_the_reference():
    from customnerf_b200.nerf import freq_embed
    d = torch.from_numpy(G["embed_in"])
    assert int(G["embed_dim"]) == 27
    for name, fn in (("oracle", torch_ref.freq_embed), ("product", freq_embed)):
        out = fn(d).numpy()
        assert out.shape == G["embed_out"].shape, name
        np.testing.assert_allclose(out, G["embed_out"], rtol=1e-6, atol=1e-7, err_msg=name)


def test_trunc_exp_forward_and_clamped_backward_match_the_reference():
    from customnerf_b200.nerf import trunc_exp
    for name, fn in (("oracle", torch_ref._trunc_exp.apply), ("product", trunc_exp)):
        x = torch.from_numpy(G["texp_in"].copy()).requires_grad_()
        y = fn(x)
        y.backward(torch.from_numpy(G["texp_gout"]))
        np.testing.assert_allclose(y.detach().numpy(), G["texp_out"], rtol=1e-6, err_msg=name)
        np.testing.assert_allclose(x.grad.numpy(), G["texp_gin"], rtol=1e-6, err_msg=name)      # exp(clamp(x, -15, 15)) * g


def test_level_tables_match_the_reference_encoder():
    from customnerf_b200.gridencoder import GridEncoder
    for i, cfg in enumerate(GRID_CONFIGS):
        want = G["grid%d_offsets" % i]
        n_params, out_dim, rows, C = [int(v) for v in G["grid%d_meta" % i]]
        okw = {k: v for k, v in cfg.items() if k != "gridtype"}
        offs, scale = cpu_ops.grid_offsets(**okw)
        assert np.array_equal(offs.astype(np.int64), want), ("oracle", i)
        assert abs(float(scale) - float(G["grid%d_scale" % i])) < 1e-12
        enc = GridEncoder(**cfg)                                   # host-side construction: no kernel is launched
        assert np.array_equal(enc.offsets.numpy().astype(np.int64), want), ("product", i)
        assert (enc.n_params, enc.output_dim, tuple(enc.embeddings.shape)) == (n_params, out_dim, (rows, C)), i
        assert abs(float(enc.per_level_scale) - float(G["grid%d_scale" % i])) < 1e-12
        assert float(enc.embeddings.detach().abs().max()) <= 1e-4 and float(G["grid%d_init_absmax" % i]) <= 1e-4   # U(-1e-4, 1e-4)


def _digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_occupancy_update_matches_the_reference_renderer():
    """two consecutive update_extra_state calls (fresh grid, then the EMA-max path) with the generator seeded as the
    reference run was: grid, bit field, mean density, mean_count"""
    from customnerf_b200 import synthetic as syn
    net = torch_ref.NeRFNetwork(torch_ref.default_opt(cuda_ray=True), encoder_kwargs=dict(log2_hashmap_size=12, desired_resolution=64,
                                                                                          gridtype="hash"))
    net.density = lambda x: {"sigma": syn.bear_density(x)}
    torch.manual_seed(123)
    net.local_step = 3
    net.step_counter[:3, 0] = torch.tensor([100, 200, 301], dtype=torch.int32)
    for k in range(2):
        net.update_extra_state()
        grid, bits = net.density_grid.numpy(), net.density_bitfield.numpy()
        np.testing.assert_allclose(grid.reshape(-1)[::1009], G["occ%d_grid_sample" % k], rtol=1e-6, atol=1e-7)
        assert int(np.unpackbits(bits).sum()) == int(G["occ%d_bits_set" % k])
        assert np.array_equal(_digest(bits), G["occ%d_bitfield_sha256" % k])
        assert np.array_equal(_digest(grid), G["occ%d_grid_sha256" % k])
        assert abs(net.mean_density - float(G["occ%d_mean_density" % k])) <= 1e-6 * float(G["occ%d_mean_density" % k])
        assert int(net.mean_count) == int(G["occ%d_mean_count" % k])
        assert [net.iter_density, net.local_step] == list(G["occ%d_iter_local" % k])


def test_dense_renderer_run_matches_the_reference_renderer():
    """oracle/torch_ref.py NeRFNetwork.run -- the restatement bench.py times as the CPU baseline -- against the reference's
    NeRFRenderer.run (renderer.py:278-405) on the same analytic field: eval mode (deterministic importance sampling) and
    training mode (perturbed strata, random importance samples; the reference draws torch.randn(3) for its unused light
    direction first, :303, so the generator is advanced by the same draw here)."""
    import types
    from golden.make_golden_python import RUN_OPT, RUN_KEYS, field_forward, scene_density, run_rays
    opt = torch_ref.default_opt(**RUN_OPT)
    net = torch_ref.NeRFNetwork(opt, encoder_kwargs=dict(log2_hashmap_size=12, desired_resolution=64, gridtype="hash"))
    net.density = scene_density
    net.forward = field_forward
    o, d = run_rays()
    for mode in ("eval", "train"):
        net.train(mode == "train")
        torch.manual_seed(5)
        torch.randn(3)                                    # the reference's light_d draw (renderer.py:303)
        res = net.run(o, d, num_steps=16, upsample_steps=16, perturb=(mode == "train"))
        for key in RUN_KEYS:
            for sub, r in (("", res), ("fg_", res["fg"]), ("bg_", res["bg"])):
                want = G["run_%s_%s%s" % (mode, sub, key)]
                np.testing.assert_allclose(r[key].detach().numpy().reshape(want.shape), want, rtol=1e-5, atol=1e-6,
                                           err_msg="%s %s%s" % (mode, sub, key))
        np.testing.assert_allclose(res["edit_mask"].detach().numpy(), G["run_%s_edit_mask" % mode], rtol=1e-5, atol=1e-6)


def test_encoder_wrapper_matches_the_reference_autograd_function():
    """oracle/torch_ref.py GridEncoder (a pure-torch encoder differentiated by autograd) against the reference's own
    GridEncoder.forward + _grid_encode Function (gridencoder/grid.py:24-99, 151-168) whose two native entry points were
    served by the C oracle when the vectors were minted: world -> [0,1] mapping, prefix shapes, level-major permutes,
    max_level, gradients w.r.t. embeddings and inputs.  Two independent fp32 implementations of the kernel: rel 1e-4 with an
    absolute floor of 3e-5 on O(1) features (a position is x * scale + 0.5 with scale up to 128, so one ulp of x moves the
    interpolation weights by ~1e-5)."""
    from golden.make_golden_python import ENC_CFG
    enc = torch_ref.GridEncoder(**ENC_CFG)
    enc.embeddings.data.copy_(torch.from_numpy(G["enc_embeddings"]))
    x = torch.from_numpy(G["enc_inputs"].copy()).requires_grad_()
    out = enc(x, bound=2)
    assert tuple(out.shape) == G["enc_out"].shape
    out.backward(torch.from_numpy(G["enc_gout"]))
    np.testing.assert_allclose(out.detach().numpy(), G["enc_out"], rtol=1e-4, atol=3e-5)
    np.testing.assert_allclose(enc.embeddings.grad.numpy(), G["enc_grad_embeddings"], rtol=1e-4, atol=3e-5)
    gi, want = x.grad.numpy(), G["enc_grad_inputs"]
    # d/d(world x) = d/d(x01) / (2 bound): the reference's kernel differentiates w.r.t. the [0,1] input and its wrapper
    # returns that gradient for the [0,1] tensor autograd then chains through (inputs + bound) / (2 bound)
    np.testing.assert_allclose(gi, want, rtol=1e-3, atol=1e-4 * np.abs(want).max())
    enc.embeddings.grad = None
    out5 = enc(torch.from_numpy(G["enc_inputs"].copy()), bound=2, max_level=5)
    out5.backward(torch.from_numpy(G["enc_gout"]))
    np.testing.assert_allclose(out5.detach().numpy(), G["enc_out_max5"], rtol=1e-4, atol=3e-5)
    assert float(np.abs(out5.detach().numpy().reshape(-1, 8, 2)[:, 5:]).max()) == 0.0
    np.testing.assert_allclose(enc.embeddings.grad.numpy(), G["enc_grad_embeddings_max5"], rtol=1e-4, atol=3e-5)


def test_field_network_wiring_matches_the_reference_network_grid():
    """oracle/torch_ref.py NeRFNetwork.forward / density against the reference's NeRFNetwork (nerf/network_grid.py:70-193)
    run with its real tiled 2^21 -> 8192 encoder; tinycudann.Network (absent, unpinned) was served by the oracle's MLP
    restatement when the vectors were minted, so what this pins is the wiring: which features feed which network, the
    gaussian blob + trunc_exp on the density head, [view embedding | trunk features] into the colour + mask head, and
    the parameter groups / learning rates of get_params.  Two independent fp32 encoders meet here at resolution 8192 (one ulp
    of a [0,1] coordinate is 5e-4 of a finest-level cell), followed by three layers: outputs abs 2e-4, the density's exponent
    abs 1e-3 (0.1 % of sigma) -- a wrong wiring moves these by O(1)."""
    from golden.make_golden_python import FIELD_OPT, table_fill
    opt = torch_ref.default_opt(**FIELD_OPT)
    net = torch_ref.NeRFNetwork(opt)                       # default encoder = network_grid.py:89-96
    assert tuple(net.pos_en.embeddings.shape) == (23967296, 2) and net.pos_en.gridtype == "tiled"
    with torch.no_grad():
        net.pos_en.embeddings.copy_(torch.from_numpy(table_fill(*net.pos_en.embeddings.shape)))
        for name in ("network", "density_network", "rgb_network"):
            getattr(net, name).params.copy_(torch.from_numpy(G["field_params_" + name]))
        x, d = torch.from_numpy(G["field_x"]), torch.from_numpy(G["field_d"])
        sigma, rad, _ = net(x, d)
        dens = net.density(x)["sigma"]
    assert tuple(rad.shape) == (200, 4)
    np.testing.assert_allclose(rad.numpy(), G["field_radiances"], rtol=1e-4, atol=2e-4)
    # sigma = exp(head + blob): a relative tolerance on the exponent's argument (values reach 360)
    np.testing.assert_allclose(np.log(sigma.numpy()), np.log(G["field_sigma"]), rtol=0, atol=1e-3)
    np.testing.assert_allclose(np.log(dens.numpy()), np.log(G["field_density"]), rtol=0, atol=1e-3)
    assert [g["lr"] for g in net.get_params(5e-4)] == list(G["field_param_group_lrs"])
    # the product's module has the same parameter groups and state-dict names (constructed on the CPU: no kernel runs)
    from customnerf_b200.nerf import NeRFNetwork
    from customnerf_b200 import trainer
    prod = NeRFNetwork(trainer.make_opt(cuda_ray=False, train_conf=0.01), encoding="tiledgrid", log2_hashmap_size=12,
                       desired_resolution=64)
    assert [g["lr"] for g in prod.get_params(5e-4)] == list(G["field_param_group_lrs"])
    assert {k for k in prod.state_dict() if "params" in k or "embeddings" in k} == {
        "pos_en.embeddings", "network.params", "density_network.params", "rgb_network.params"}
    for name in ("network", "density_network", "rgb_network"):
        assert getattr(prod, name).params.numel() == G["field_params_" + name].size, name


def test_occupancy_renderer_training_branch_matches_the_reference_run_cuda():
    """oracle/torch_ref.py NeRFNetwork.run_cuda (training branch) against the reference's NeRFRenderer.run_cuda
    (renderer.py:597-640, 688-716) run on the CPU with its native ops served by the C oracle: near/far with the default
    min_near, the step_counter ring over two calls, march -> field -> composite, the result dict."""
    from customnerf_b200 import synthetic as syn
    from golden.make_golden_python import run_rays
    net = torch_ref.NeRFNetwork(torch_ref.default_opt(cuda_ray=True), encoder_kwargs=dict(log2_hashmap_size=12, desired_resolution=64,
                                                                                          gridtype="hash"))
    grid = syn.density_grid(2, 128)
    net.density_bitfield = torch.from_numpy(cpu_ops.packbits(grid.numpy(), min(float(grid.mean()), 10.0)))
    net.forward = lambda x, d: (syn.bear_density(x), syn.bear_color(x) * (0.5 + 0.5 * d[:, :1].abs()), None)
    net.train()
    o, d = run_rays()
    for it in range(2):
        res = net.run_cuda(o, d, perturb=False, force_all_rays=True)
        for key in ("image", "depth", "weights_sum"):
            np.testing.assert_allclose(res[key].detach().numpy(), G["runcuda%d_%s" % (it, key)], rtol=1e-5, atol=1e-6, err_msg=key)
        assert np.array_equal(res["mask"].numpy(), G["runcuda%d_mask" % it])
    assert np.array_equal(net.step_counter.numpy(), G["runcuda_step_counter"]) and net.local_step == int(G["runcuda_local_step"])


def test_total_variation_wrapper_matches_the_reference():
    """oracle grad_total_variation against the reference's GridEncoder.grad_total_variation (grid.py:171-192) run on the CPU with
    its native entry point served by the C oracle: the [-bound, bound] -> [0,1] mapping of explicit locations, accumulation into
    .grad, the torch.rand(B, D) draw of the default path, the ValueError without a gradient.  Exact (same C arithmetic)."""
    from golden.make_golden_python import ENC_CFG
    assert int(G["tv_raises_without_grad"]) == 1
    offs, scale = cpu_ops.grid_offsets(**{k: v for k, v in ENC_CFG.items() if k != "gridtype"})
    emb = G["enc_embeddings"]
    grad = np.zeros_like(emb)
    x01 = (G["tv_inputs"] + np.float32(2)) / np.float32(4)
    grad = cpu_ops.grad_total_variation(x01, emb, grad, offs, 1e-2, scale, ENC_CFG["base_resolution"])
    assert np.array_equal(grad, G["tv_grad_explicit"]) and float(np.abs(grad).max()) > 0
    torch.manual_seed(11)
    grad = cpu_ops.grad_total_variation(torch.rand(500, 3).numpy(), emb, grad, offs, 1e-2, scale, ENC_CFG["base_resolution"])
    assert np.array_equal(grad, G["tv_grad_then_random"])


def test_reconstruction_loss_matches_the_reference_trainer():
    """customnerf_b200/trainer.py: pretrain_loss against the reference's Trainer_Nerf.train_step_pretrain
    (utils_init_nerf.py:194-241) run on a canned render result: loss, its parts, the clamped mask volume, and the gradients
    with respect to the rendered image and mask (what FusedTrainStep's compositing kernel produces as g_image /
    g_render_mask).  Same torch ops: rel 1e-6."""
    import types
    from customnerf_b200 import trainer
    for tag, tconf in (("conf", 0.01), ("noconf", 0)):
        out = {"image": torch.from_numpy(G["loss_in_image"].copy()).requires_grad_(),
               "weights_sum": torch.from_numpy(G["loss_in_weights_sum"]),
               "render_mask": torch.from_numpy(G["loss_in_render_mask"].copy()).requires_grad_()}
        opt = types.SimpleNamespace(train_rgb=1.0, train_conf=tconf)
        pred, mvol, loss, ld = trainer.pretrain_loss(out, torch.from_numpy(G["loss_in_rgb"]), torch.from_numpy(G["loss_in_gtmask"]), opt)
        loss.backward()
        assert abs(loss.item() - float(G["loss_%s_value" % tag])) <= 1e-6 * float(G["loss_%s_value" % tag])
        want_c, want_m = G["loss_%s_dict" % tag]
        assert abs(ld["loss_c"] - want_c) <= 1e-6 * want_c and (("loss_m" in ld) == (want_m >= 0))
        if want_m >= 0:
            assert abs(ld["loss_m"] - want_m) <= 1e-6 * want_m
        np.testing.assert_allclose(mvol.numpy(), G["loss_%s_mask_volume" % tag], rtol=1e-6)
        assert mvol.min() >= 1e-5 and mvol.max() <= 1 - 1e-5
        np.testing.assert_allclose(out["image"].grad.numpy(), G["loss_%s_grad_image" % tag], rtol=1e-6, atol=1e-9)
        gm = out["render_mask"].grad
        np.testing.assert_allclose((gm if gm is not None else torch.zeros_like(out["render_mask"])).numpy(),
                                   G["loss_%s_grad_mask" % tag], rtol=1e-6, atol=1e-9)
