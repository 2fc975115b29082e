"""-m gpu: the fused tcgen05 field kernels (csrc/field_mlp.cu) against (a) the per-layer library path of the same
module and (b) the fp32 oracle with fp16 operand rounding emulated.

tiny-cuda-nn is un-vendored and un-pinned (parity unpinned): the contract checked here is this repo's --
fp16 operands, fp32 accumulate: outputs abs 2e-3 / rel 1e-2, gradients rel 2e-2 of the largest entry.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import torch_ref

pytestmark = pytest.mark.gpu


def _net(train_conf=0.01, seed=0):
    from customnerf_b200.nerf import NeRFNetwork
    torch.manual_seed(seed)
    opt = torch_ref.default_opt(train_conf=train_conf)
    net = NeRFNetwork(opt, encoding="hashgrid", log2_hashmap_size=14, desired_resolution=256).cuda()
    with torch.no_grad():
        net.pos_en.embeddings.uniform_(-1, 1)
    return net, opt


def _inputs(M, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(M, 3, generator=g) * 2 - 1) * 1.5
    x[: M // 8] *= 0.05                      # some points inside the gaussian density blob
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return x.cuda(), d.cuda()


@pytest.mark.parametrize("M", [1, 127, 128, 129, 5000, 40000, 100003])
def test_fused_forward_matches_library_path_and_oracle(M):
    net, opt = _net()
    x, d = _inputs(M)
    with torch.no_grad():
        net.use_fused_field = True
        s1, c1, _ = net(x, d)
        net.use_fused_field = False
        s0, c0, _ = net(x, d)
    assert torch.isfinite(s1).all() and s1.shape == (M,) and c1.shape == (M, 4)
    assert_close(c1.float().cpu().numpy(), c0.float().cpu().numpy(), 1e-2, 2e-3, "rgba fused vs library")
    assert_close(s1.cpu().numpy(), s0.cpu().numpy(), 2e-2, 1e-3, "sigma fused vs library")
    ref = torch_ref.NeRFNetwork(opt, encoder_kwargs=dict(log2_hashmap_size=14, desired_resolution=256, gridtype="hash"))
    ref.pos_en.embeddings.data.copy_(net.pos_en.embeddings.detach().cpu())
    for name in ("network", "density_network", "rgb_network"):
        getattr(ref, name).params.data.copy_(getattr(net, name).params.detach().cpu())
        getattr(ref, name).half = True
    with torch.no_grad():
        sr, cr, _ = ref(x.cpu(), d.cpu())
    assert_close(c1.float().cpu().numpy(), cr.numpy(), 1e-2, 2e-3, "rgba fused vs oracle")
    assert_close(s1.cpu().numpy(), sr.numpy(), 2e-2, 1e-3, "sigma fused vs oracle")


@pytest.mark.parametrize("M", [300, 20000, 70001])
def test_fused_backward_matches_library_path(M):
    net, opt = _net()
    x, d = _inputs(M, seed=3)
    g = torch.Generator(device="cuda").manual_seed(5)
    gs = torch.randn(M, device="cuda", generator=g) * 0.1
    gc = torch.randn(M, 4, device="cuda", generator=g)
    grads = {}
    for fused in (True, False):
        net.use_fused_field = fused
        net.zero_grad(set_to_none=True)
        s, c, _ = net(x, d)
        torch.autograd.backward([s, c], [gs, gc.to(c.dtype)])
        grads[fused] = {n: p.grad.detach().float().cpu().numpy().copy() for n, p in net.named_parameters()}
    for name in ("rgb_network.params", "density_network.params", "network.params", "pos_en.embeddings"):
        a, b = grads[True][name], grads[False][name]
        assert np.isfinite(a).all(), name
        assert_close(a, b, 3e-2, 2e-2 * np.abs(b).max(), name)
    # padded output rows of the heads receive no gradient (tcnn layout: 16 padded outputs, 1 / 4 used)
    gd = grads[True]["density_network.params"][64 * 64:].reshape(16, 64)
    assert not gd[1:].any()
    gr = grads[True]["rgb_network.params"][64 * 96:].reshape(16, 64)
    assert not gr[4:].any() and np.abs(gr[:4]).sum() > 0


def test_fused_density_only_and_grad_free_mode():
    net, opt = _net(train_conf=0)
    x, d = _inputs(1000, seed=7)
    with torch.no_grad():
        s = net.density(x)["sigma"]
        s2, c2, _ = net(x, d)
    assert c2.shape == (1000, 3)
    assert_close(s.cpu().numpy(), s2.cpu().numpy(), 1e-6, 0, "density() == forward() sigma")
    assert net(torch.zeros(0, 3).cuda(), torch.zeros(0, 3).cuda())[0].shape == (0,)


def test_weight_repack_follows_parameter_updates():
    net, opt = _net()
    x, d = _inputs(512, seed=9)
    with torch.no_grad():
        a = net(x, d)[1].float().clone()
        net.rgb_network.params.mul_(0.5)
        b = net(x, d)[1].float()
        net.use_fused_field = False
        b0 = net(x, d)[1].float()
    assert (a - b).abs().max() > 1e-3
    assert_close(b.cpu().numpy(), b0.cpu().numpy(), 1e-2, 2e-3, "after in-place parameter update")
