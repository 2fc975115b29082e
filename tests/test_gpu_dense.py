"""-m gpu: the dense (non-cuda_ray) renderer with its sampler as two kernels (csrc/dense_sampler.cu) and its tail on the
occupancy path's kernels, against (a) the torch restatement of the reference's sampler (renderer.py:297-367, :21-55; itself
pinned to the reference's own run on the CPU by tests/test_golden_python_cpu.py), (b) the op-by-op dense renderer
(NeRFRenderer._run_dense_ops, the reference's torch sequence) on the real field network, values and gradients.
The golden vectors of the reference's own NeRFRenderer.run are checked in tests/test_gpu_reference_python.py."""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import torch_ref

pytestmark = pytest.mark.gpu


def _rays(n=512, seed=0):
    from customnerf_b200 import synthetic as syn
    o, d = syn.camera_rays(105, 142)
    g = torch.Generator().manual_seed(seed)
    sel = torch.randperm(o.shape[0], generator=g)[:n]
    return o[sel].contiguous().cuda(), d[sel].contiguous().cuda()


@pytest.mark.parametrize("S,Su,det", [(64, 64, True), (64, 64, False), (16, 16, True), (33, 20, False), (128, 128, True)])
def test_sampler_kernels_match_the_torch_restatement(S, Su, det):
    from customnerf_b200 import _lib as L, raymarching as rm, synthetic as syn
    from customnerf_b200.nerf.rendering import sample_pdf
    o, d = _rays(300)
    N = o.shape[0]
    aabb = torch.tensor([-2, -2, -2, 2, 2, 2], dtype=torch.float32, device="cuda")
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.01)
    lin = torch.linspace(0, 1, S, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    noise = None if det else torch.rand(N, S, device="cuda", generator=g)
    z_c, xyz_c = torch.empty(N, S, device="cuda"), torch.empty(N * S, 3, device="cuda")
    lib = L.lib()
    L.check(lib.nb200_dense_coarse(L.ptr(o), L.ptr(d), L.ptr(nears), L.ptr(fars), L.ptr(aabb), L.ptr(lin), L.ptr(noise), L.u32(N),
                                   L.u32(S), L.ptr(z_c), L.ptr(xyz_c), L.stream()), "coarse")
    # reference expressions (renderer.py:306-317)
    z_ref = nears[:, None] + (fars - nears)[:, None] * lin[None]
    sd = (fars - nears)[:, None] / S
    if noise is not None:
        z_ref = z_ref + (noise - 0.5) * sd
    xyz_ref = torch.min(torch.max(o[:, None] + d[:, None] * z_ref[..., None], aabb[:3]), aabb[3:])
    assert torch.equal(z_c, z_ref)
    assert torch.equal(xyz_c.view(N, S, 3), xyz_ref)
    # a haze of 0.3 everywhere keeps every bin's probability mass well above sample_pdf's 1e-5 threshold (:50-51) -- on exactly
    # empty bins the reference's own result flips between "denominator" and "1" with the last bit of a cumulative sum
    sigma = (syn.bear_density(xyz_c.cpu()) * 3.0 + 0.3).cuda().contiguous()
    u = torch.linspace(0.5 / Su, 1 - 0.5 / Su, Su, device="cuda") if det else torch.rand(N, Su, device="cuda", generator=g)
    T = S + Su
    z_all, xyzs, dirs = torch.empty(N, T, device="cuda"), torch.empty(N * T, 3, device="cuda"), torch.empty(N * T, 3, device="cuda")
    deltas, rays = torch.empty(N * T, 2, device="cuda"), torch.empty(N, 3, dtype=torch.int32, device="cuda")
    L.check(lib.nb200_dense_importance(L.ptr(o), L.ptr(d), L.ptr(nears), L.ptr(fars), L.ptr(aabb), L.ptr(z_c), L.ptr(sigma),
                                       L.ptr(u), L.i32(0 if det else 1), L.u32(N), L.u32(S), L.u32(Su), L.ptr(z_all), L.ptr(xyzs),
                                       L.ptr(dirs), L.ptr(deltas), L.ptr(rays), L.stream()), "importance")
    # torch restatement on the CPU in fp32 (renderer.py:328-361)
    zc, sg, sdc = z_c.cpu(), sigma.view(N, S).cpu(), sd.cpu()
    dl = torch.cat([zc[:, 1:] - zc[:, :-1], sdc], -1)
    al = 1 - torch.exp(-dl * sg)
    w = al * torch.cumprod(torch.cat([torch.ones_like(al[:, :1]), 1 - al + 1e-15], -1), -1)[:, :-1]
    mid = zc[:, :-1] + 0.5 * dl[:, :-1]
    # sample_pdf with the same u
    ww = w[:, 1:-1] + 1e-5
    pdf = ww / ww.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros(N, 1), torch.cumsum(pdf, -1)], -1)
    uu = (u.cpu()[None].expand(N, Su) if det else u.cpu()).contiguous()
    inds = torch.searchsorted(cdf, uu, right=True)
    below, above = (inds - 1).clamp_min(0), inds.clamp_max(cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(mid, 1, below), torch.gather(mid, 1, above)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    new_z = b0 + (uu - c0) / den * (b1 - b0)
    z_want, _ = torch.sort(torch.cat([zc, new_z], 1), 1)
    span = float((fars - nears).max())
    # The cumulative sums differ from torch's by summation order only (~1e-7), so the depths agree to 2e-5 of the ray span --
    # except where u sits within that rounding of a cdf entry AND the neighbouring bin is (nearly) empty: sample_pdf replaces a
    # denominator below 1e-5 by 1 (:50-51), a discontinuity of up to one bin width.  Allow that for at most 0.1 % of the samples.
    err = (z_all.cpu() - z_want).abs()
    off = err > 2e-5 * span
    assert float(off.float().mean()) <= 1e-3, float(off.float().mean())
    assert float(err.max()) <= 1.01 * span / (S - 1), float(err.max())
    got = z_all.cpu()
    assert (got[:, 1:] >= got[:, :-1]).all(), "sorted"
    want_d0 = torch.cat([got[:, 1:] - got[:, :-1], sdc], -1)
    assert_close(deltas[:, 0].view(N, T).cpu().numpy(), want_d0.numpy(), 0, 1e-6 * span, "deltas")
    oz = ((got - nears.cpu()[:, None]) / (fars - nears).cpu()[:, None]).clamp(0, 1)
    assert_close(torch.cumsum(deltas[:, 1].view(N, T).cpu(), 1).numpy(), oz.numpy(), 0, 2e-6, "depth coordinate")
    want_xyz = torch.min(torch.max(o.cpu()[:, None] + d.cpu()[:, None] * got[..., None], aabb.cpu()[:3]), aabb.cpu()[3:])
    assert torch.equal(xyzs.view(N, T, 3).cpu(), want_xyz)
    assert torch.equal(dirs.view(N, T, 3).cpu(), d.cpu()[:, None].expand(N, T, 3))
    assert torch.equal(rays.cpu(), torch.stack([torch.arange(N), torch.arange(N) * T, torch.full((N,), T)], 1).int())


@pytest.mark.parametrize("soft_mask,detach_bg", [(True, True), (False, False)])
def test_fused_dense_renderer_matches_the_op_by_op_renderer(soft_mask, detach_bg):
    """eval mode (deterministic importance samples): the fused dense renderer against the reference-shaped op sequence on the
    real field network under autocast: all / fg / bg images, rendered masks, weights_sum, depth; then the gradients of a
    loss over all of them (training-mode network, but the same deterministic samples through perturb=False and a fixed u)."""
    from customnerf_b200 import trainer
    from customnerf_b200.nerf import NeRFNetwork
    torch.manual_seed(0)
    opt = trainer.make_opt(cuda_ray=False, train_conf=0.01, soft_mask=soft_mask, detach_bg=detach_bg)
    net = NeRFNetwork(opt, encoding="hashgrid", log2_hashmap_size=15, desired_resolution=512).cuda()
    with torch.no_grad():
        net.pos_en.embeddings.uniform_(-0.5, 0.5)
        net.rgb_network.params[-16 * 64:].view(16, 64)[3] *= 12.0        # a mask head that splits fg / bg
    o, d = _rays(512, seed=1)
    net.eval()
    outs = {}
    for fused in (True, False):
        net.fused_dense = fused
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            outs[fused] = net.run(o[None], d[None], num_steps=64, upsample_steps=64, perturb=False)
    a, b = outs[True], outs[False]
    for key in ("image", "depth", "weights_sum", "render_mask"):
        assert_close(a[key].float().cpu().numpy().reshape(-1), b[key].float().cpu().numpy().reshape(-1), 2e-3, 2e-3, key)
    for part in ("fg", "bg"):
        x, y = a[part]["image"].float().cpu().numpy(), b[part]["image"].float().cpu().numpy()
        if soft_mask:
            assert_close(x.reshape(-1), y.reshape(-1), 2e-3, 2e-3, part + ".image")
        else:       # hard mask: a sample whose mask value sits within fp16 rounding of 0.5 switches sides -- allow a few rays
            bad = np.abs(x - y) > 2e-3 + 2e-3 * np.abs(y)
            assert bad.mean() < 0.03, (part, bad.mean())
    assert tuple(a["image"].shape) == tuple(b["image"].shape) and tuple(a["depth"].shape) == tuple(b["depth"].shape)
    assert torch.equal(a["mask"], b["mask"])

    # gradients (eval-mode sampling is deterministic: both paths differentiate through the same sample set)
    grads = {}
    for fused in (True, False):
        net.fused_dense = fused
        net.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            out = net.run(o[None], d[None], num_steps=64, upsample_steps=64, perturb=False)
        loss = (out["image"].float() ** 2).mean() + out["render_mask"].float().mean() + (out["fg"]["image"].float() ** 2).mean()
        loss = loss + out["bg"]["weights_sum"].float().mean()
        (loss * 128.0).backward()
        grads[fused] = {n: p.grad.detach().float().cpu().numpy().copy() for n, p in net.named_parameters() if p.grad is not None}
    for name in ("rgb_network.params", "density_network.params", "network.params", "pos_en.embeddings"):
        u, v = grads[True][name], grads[False][name]
        assert np.isfinite(u).all() and np.abs(v).max() > 0, name
        if soft_mask:
            assert_close(u, v, 2e-2, 2e-2 * np.abs(v).max(), name)
        else:
            assert np.abs(u - v).max() <= 0.2 * np.abs(v).max(), name


def test_fused_dense_train_step_matches_autograd_over_the_dense_renderer():
    """FusedTrainStep(dense=(64, 64)): sampler kernels + density-only fused launch + the occupancy path's encode .. encode^T,
    against autograd over NeRFNetwork.run (training mode, the same rand(N, Su) draw through the same seed, no jitter):
    loss rel 1e-4, every parameter gradient within 2e-2 of its largest entry; then the step trains (graph replay)."""
    import torch.nn.functional as F
    from customnerf_b200 import trainer, fused_trainer
    from customnerf_b200.nerf import NeRFNetwork
    w, N = 0.5, 1024
    opt = trainer.make_opt(cuda_ray=False, train_conf=w)
    nets = []
    for _ in range(3):
        torch.manual_seed(0)
        net = NeRFNetwork(opt, encoding="hashgrid", log2_hashmap_size=15, desired_resolution=512).cuda()
        with torch.no_grad():
            net.pos_en.embeddings.uniform_(-0.5, 0.5)
        nets.append(net.train())
    ma, mb, mc = nets
    o, d = _rays(N, seed=2)
    g = torch.Generator().manual_seed(3)
    tgt, gt_mask = torch.rand(N, 3, generator=g).cuda(), (torch.rand(N, generator=g) > 0.5).float().cuda()
    torch.manual_seed(11)
    with torch.autocast("cuda", dtype=torch.float16):
        out = ma.run(o[None], d[None], num_steps=64, upsample_steps=64, perturb=False)
        loss_ref = (F.mse_loss(out["image"].reshape(-1, 3).float(), tgt, reduction="sum") / (3.0 * N)
                    + w * F.mse_loss(out["render_mask"].reshape(-1).float(), gt_mask, reduction="sum") / N)
    (loss_ref * fused_trainer.LOSS_SCALE).backward()
    fs = fused_trainer.FusedTrainStep(mb, N, perturb=False, use_graph=False, mask_weight=w, dense=(64, 64), dynamic_loss_scale=False)
    fs.set_batch(o, d, tgt, gt_mask)
    torch.manual_seed(11)
    fs.forward_backward()
    loss, samples, used = fs.last_stats()
    assert samples == used == N * 128
    ref = float(loss_ref.detach())
    assert abs(loss - ref) <= 1e-4 * abs(ref) + 1e-7, (loss, ref)
    np.testing.assert_allclose(fs.z_all.cpu().numpy(), out["z_vals"].cpu().numpy(), rtol=0, atol=1e-6)
    for name, off, n in fs.layout:
        mod, attr = name.split(".")
        g_ref = getattr(getattr(ma, mod), attr).grad.reshape(-1).float().cpu().numpy()
        got = fs.grads_flat[off:off + n].cpu().numpy()
        scale = np.abs(g_ref).max()
        assert scale > 0 and np.abs(got - g_ref).max() <= 2e-2 * scale, (name, np.abs(got - g_ref).max(), scale)

    # and it trains: captured graph, jittered samples
    fs2 = fused_trainer.FusedTrainStep(mc, N, perturb=True, use_graph=True, mask_weight=w, dense=(64, 64))
    fs2.set_batch(o, d, tgt, gt_mask)
    losses = []
    for _ in range(40):
        fs2.step()
        losses.append(fs2.last_stats()[0])
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.9 * np.mean(losses[:5]), losses   # (the targets are noise)
