"""-m gpu: grid encoding + field network as ONE kernel (csrc/field_fused.cu: producer warps gather the hash-grid features
straight into the tensor-core operand tile) against the two-kernel path it replaces (nb200_grid_encode_forward +
nb200_field_forward, themselves pinned to the reference encoder kernel and the fp32 oracle by test_gpu_ref_ext.py /
test_gpu_field_mlp.py) and against the oracle directly.

Contract: the features (x_en), the trunk / density-head activations and sigma are BIT-IDENTICAL to the two-kernel path (same
gather function, same MMA sequence); the colour head accumulates its two K blocks in the other order (view part first), so
hr / rgba agree to fp16 rounding (abs 2e-3)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import torch_ref

pytestmark = pytest.mark.gpu


def _net(train_conf=0.01, seed=0, log2=14, res=256):
    from customnerf_b200.nerf import NeRFNetwork
    torch.manual_seed(seed)
    opt = torch_ref.default_opt(train_conf=train_conf, cuda_ray=True)
    net = NeRFNetwork(opt, encoding="hashgrid", log2_hashmap_size=log2, desired_resolution=res).cuda()
    with torch.no_grad():
        net.pos_en.embeddings.uniform_(-1, 1)
    return net, opt


def _inputs(M, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(M, 3, generator=g) * 2 - 1) * 1.9
    x[: M // 8] *= 0.05                      # some points inside the gaussian density blob
    if M > 16:
        x[-3:] = torch.tensor([[2.0, -2.0, 0.5], [2.5, 0.0, 0.0], [0.0, 0.0, -2.0]])   # box faces and one point outside
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return x.cuda(), d.cuda()


@pytest.mark.parametrize("M", [1, 127, 128, 129, 5000, 100003])
def test_fused_inference_forward_matches_two_kernel_path(M):
    net, opt = _net()
    x, d = _inputs(M)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        net.fuse_encoder = True
        s1, c1, _ = net(x, d)
        dens = net.density(x)["sigma"]
        net.fuse_encoder = False
        s0, c0, _ = net(x, d)
    assert s1.shape == (M,) and c1.shape == (M, 4) and torch.isfinite(s1).all()
    assert torch.equal(s1, s0), float((s1 - s0).abs().max())
    assert torch.equal(dens, s0), "density-only kernel == full kernel's sigma"
    assert_close(c1.float().cpu().numpy(), c0.float().cpu().numpy(), 0, 2e-3, "rgba fused vs two kernels")


def test_fused_forward_matches_oracle():
    net, opt = _net()
    M = 20000
    x, d = _inputs(M, seed=5)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        s1, c1, _ = net(x, d)
    ref = torch_ref.NeRFNetwork(opt, encoder_kwargs=dict(log2_hashmap_size=14, desired_resolution=256, gridtype="hash"))
    ref.pos_en.embeddings.data.copy_(net.pos_en.embeddings.detach().half().float().cpu())    # the autocast table (grid.py:45-46)
    for name in ("network", "density_network", "rgb_network"):
        getattr(ref, name).params.data.copy_(getattr(net, name).params.detach().cpu())
        getattr(ref, name).half = True
    with torch.no_grad():
        sr, cr, _ = ref(x.cpu(), d.cpu())
    assert_close(c1.float().cpu().numpy(), cr.numpy(), 1e-2, 2e-3, "rgba fused vs oracle")
    assert_close(s1.cpu().numpy(), sr.numpy(), 2e-2, 1e-3, "sigma fused vs oracle")


@pytest.mark.parametrize("M", [129, 30000])
def test_fused_training_forward_saves_what_the_backward_needs(M):
    """training forward: x_en, sigma_arg and the five activation planes against the two-kernel path, then the gradients of the
    one-node autograd function (fused forward, field backward, encode scatter) against the two-node composition"""
    from customnerf_b200 import _lib as L
    from customnerf_b200.nerf import fused_field as ff
    net, opt = _net()
    x, d = _inputs(M, seed=9)
    enc = net.pos_en
    lib = L.lib()
    fwd_img, _ = net._packed.get(net.network.params, net.density_network.params, net.rgb_network.params)
    f16, f32 = dict(dtype=torch.half, device="cuda"), dict(dtype=torch.float32, device="cuda")
    out = {}
    for fused in (True, False):
        sigma, sarg, rgba = torch.zeros(M, **f32), torch.zeros(M, **f32), torch.zeros(M, 4, **f16)
        x_en, act = torch.zeros(M, 32, **f16), torch.zeros(5, M, 64, **f16)
        if fused:
            L.check(lib.nb200_field_fused_forward(L.ptr(x), L.ptr(d), L.f32(2.0), *ff._enc_args(enc), L.ptr(fwd_img), L.ptr(sigma),
                                                  L.ptr(sarg), L.ptr(rgba), L.ptr(x_en), L.ptr(act), L.u32(M), L.ptr(None),
                                                  L.stream()), "fused")
        else:
            L.check(lib.nb200_fs_encode_forward(L.ptr(x), L.f32(2.0), L.ptr(enc.embeddings.detach()), L.ptr(enc.offsets), L.ptr(x_en),
                                                L.u32(M), L.u32(16), L.f32(float(np.log2(enc.per_level_scale))), L.u32(16),
                                                L.u32(enc.gridtype_id), L.i32(0), L.u32(0), L.ptr(None), L.stream()), "enc")
            L.check(lib.nb200_field_forward(L.ptr(x_en), L.ptr(x), L.ptr(d), L.ptr(fwd_img), L.ptr(sigma), L.ptr(sarg), L.ptr(rgba),
                                            L.ptr(act), L.u32(M), L.ptr(None), L.stream()), "field")
        torch.cuda.synchronize()
        out[fused] = (sigma, sarg, rgba, x_en, act)
    a, b = out[True], out[False]
    assert torch.equal(a[3], b[3]), "x_en"
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), "sigma / sigma_arg"
    for pl, name in enumerate(("h1", "h2", "fea", "hd")):
        assert torch.equal(a[4][pl], b[4][pl]), name
    assert_close(a[4][4].float().cpu().numpy(), b[4][4].float().cpu().numpy(), 0, 4e-3, "hr")
    assert_close(a[2].float().cpu().numpy(), b[2].float().cpu().numpy(), 0, 2e-3, "rgba")

    g = torch.Generator(device="cuda").manual_seed(5)
    gs = torch.randn(M, device="cuda", generator=g) * 0.1
    gc = torch.randn(M, 4, device="cuda", generator=g)
    grads = {}
    for fused in (True, False):
        net.fuse_encoder = fused
        net.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            s, c, _ = net(x, d)
        torch.autograd.backward([s, c], [gs, gc.to(c.dtype)])
        grads[fused] = {n: p.grad.detach().float().cpu().numpy().copy() for n, p in net.named_parameters()}
    for name in ("rgb_network.params", "density_network.params", "network.params", "pos_en.embeddings"):
        u, v = grads[True][name], grads[False][name]
        assert np.isfinite(u).all() and np.abs(v).max() > 0, name
        assert_close(u, v, 1e-2, 1e-2 * np.abs(v).max(), name)


def test_fused_forward_honours_the_device_side_row_count():
    from customnerf_b200 import _lib as L
    from customnerf_b200.nerf import fused_field as ff
    net, opt = _net()
    M, live = 1000, 333
    x, d = _inputs(M, seed=2)
    fwd_img, _ = net._packed.get(net.network.params, net.density_network.params, net.rgb_network.params)
    sigma = torch.full((M,), -7.0, device="cuda")
    rgba = torch.full((M, 4), -7.0, dtype=torch.half, device="cuda")
    cnt = torch.tensor([live], dtype=torch.int32, device="cuda")
    L.check(L.lib().nb200_field_fused_forward(L.ptr(x), L.ptr(d), L.f32(2.0), *ff._enc_args(net.pos_en), L.ptr(fwd_img),
                                              L.ptr(sigma), L.ptr(None), L.ptr(rgba), L.ptr(None), L.ptr(None), L.u32(M),
                                              L.ptr(cnt), L.stream()), "fused")
    torch.cuda.synchronize()
    assert (sigma[live:] == -7.0).all() and (rgba[live:] == -7.0).all() and (sigma[:live] > 0).all()


def test_fused_occupancy_update_equals_the_op_by_op_update():
    """update_extra_state three ways -- encoder + density-only field launches over chunks (nb200_occ_density_chunked, the
    default; a chunk size that does not divide the grid), the one-kernel density query (nb200_occ_density), and the
    reference-shaped op-by-op path (cell table -> jitter -> self.density per cascade -> scatter by Morton index) -- from the
    same generator state: bit-identical density grid and bit field, same mean density and mean_count."""
    grids = {}
    for mode in ("chunked", "one_kernel", "ops"):
        net, opt = _net(train_conf=0, seed=4)
        net.fuse_encoder = True
        net.local_step = 2
        net.step_counter[:2, 0] = torch.tensor([111, 224], dtype=torch.int32, device="cuda")
        if mode == "ops":
            net.density = net.density               # an instance attribute: the stock-field test fails -> op-by-op path
        net.occ_chunk_rows = {"chunked": 300000, "one_kernel": 0, "ops": 0}[mode]
        torch.manual_seed(77)
        with torch.autocast("cuda", dtype=torch.float16):
            net.update_extra_state()
            net.update_extra_state()
        grids[mode] = (net.density_grid.clone(), net.density_bitfield.clone(), net.mean_density, net.mean_count, net.iter_density)
    b = grids["ops"]
    for mode in ("chunked", "one_kernel"):
        a = grids[mode]
        assert torch.equal(a[0], b[0]), (mode, float((a[0] - b[0]).abs().max()))
        assert torch.equal(a[1], b[1]), mode
        assert a[2] == b[2] and a[2] > 0 and a[3] == b[3] == 167 and a[4] == b[4] == 2, mode
    assert int(b[1].count_nonzero()) > 0


def test_fused_train_step_matches_the_two_kernel_step():
    """FusedTrainStep with the fused encode + field forward against the same step with two launches: same loss (the
    colour head's K blocks are summed in the other order: rel 1e-4), gradients within the fp16 contract"""
    from customnerf_b200 import fused_trainer, synthetic as syn, trainer
    res = {}
    for fused in (True, False):
        model = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=15, desired_resolution=512, seed=3,
                                          opt=trainer.make_opt(train_conf=0.01))
        with torch.no_grad():
            g = torch.Generator(device="cuda").manual_seed(11)
            model.pos_en.embeddings.copy_(torch.rand(model.pos_en.embeddings.shape, device="cuda", generator=g) - 0.5)
        o, d = syn.camera_rays(105, 142)
        sel = torch.arange(5000, 5000 + 2048)
        o, d = o[sel].contiguous().cuda(), d[sel].contiguous().cuda()
        tgt = syn.bear_color(o.cpu() + d.cpu() * 1.5).cuda()
        fs = fused_trainer.FusedTrainStep(model, o.shape[0], perturb=False, use_graph=False, fused_forward=fused, mask_weight=0.01)
        fs.set_batch(o, d, tgt)
        fs.m_cap == 0 and fs._alloc_samples(fs._round_cap(fs.measure_samples(o, d)))
        fs.forward_backward()
        loss, samples, used = fs.last_stats()
        res[fused] = (loss, samples, fs.grads_flat.clone(), fs.image.clone())
    (l1, n1, g1, i1), (l0, n0, g0, i0) = res[True], res[False]
    assert n1 == n0 and n1 > 1000
    assert abs(l1 - l0) <= 1e-4 * abs(l0), (l1, l0)
    assert_close(i1.cpu().numpy(), i0.cpu().numpy(), 0, 2e-3, "image")
    assert_close(g1.cpu().numpy(), g0.cpu().numpy(), 1e-2, 1e-2 * float(g0.abs().max()), "flat gradient")
