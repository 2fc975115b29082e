"""-m gpu: the whole hot path (march -> encode -> MLP -> composite, forward + backward) and the renderers built on it
against the CPU oracle restatement (oracle/torch_ref.py) with identical weights.

MLP numerics contract (tiny-cuda-nn is un-vendored: parity unpinned): fp16 operands, fp32 accumulation; the oracle
emulates the operand rounding (half=True), so the remaining differences are accumulation order: abs 2e-3 / rel 1e-2.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import torch_ref

pytestmark = pytest.mark.gpu
ENC = dict(log2_hashmap_size=15, desired_resolution=512)


def _pair(opt, scene=None):
    from customnerf_b200.nerf import NeRFNetwork
    torch.manual_seed(0)
    net = NeRFNetwork(opt, encoding="hashgrid", **ENC).cuda()
    with torch.no_grad():
        net.pos_en.embeddings.uniform_(-0.5, 0.5)
    ref = torch_ref.NeRFNetwork(opt, encoder_kwargs=dict(gridtype="hash", **ENC))
    ref.pos_en.embeddings.data.copy_(net.pos_en.embeddings.detach().cpu())
    for name in ("network", "density_network", "rgb_network"):
        getattr(ref, name).params.data.copy_(getattr(net, name).params.detach().cpu())
        getattr(ref, name).half = True
    if scene is not None and opt.cuda_ray:
        bf = torch.from_numpy(scene["bitfield"])
        net.density_bitfield.copy_(bf.cuda())
        ref.density_bitfield = bf
    return net, ref


def test_field_network_forward_and_density():
    opt = torch_ref.default_opt(train_conf=0.01)
    net, ref = _pair(opt)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(4096, 3, generator=g) * 2 - 1) * 1.9
    d = torch.randn(4096, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    sigma, rgba, _ = net(x.cuda(), d.cuda())
    s0, c0, _ = ref(x, d)
    assert rgba.shape == (4096, 4) and sigma.dtype == torch.float32
    assert_close(rgba.float().detach().cpu().numpy(), c0.detach().numpy(), 1e-2, 2e-3, "radiances")
    assert_close(sigma.detach().cpu().numpy(), s0.detach().numpy(), 2e-2, 1e-3, "sigma")
    assert_close(net.density(x.cuda())["sigma"].detach().cpu().numpy(), s0.detach().numpy(), 2e-2, 1e-3, "density()")


def test_cuda_ray_train_step_matches_oracle(scene):
    opt = torch_ref.default_opt(cuda_ray=True, train_conf=0)
    net, ref = _pair(opt, scene)
    sel = slice(7000, 7512)
    o, d = torch.from_numpy(scene["rays_o"][sel]), torch.from_numpy(scene["rays_d"][sel])
    target = torch.rand(512, 3, generator=torch.Generator().manual_seed(2))
    net.train(); ref.train()
    out = net.render(o[None].cuda(), d[None].cuda(), perturb=False, force_all_rays=True)
    out0 = ref.render(o[None], d[None], perturb=False, force_all_rays=True)
    assert int(net.step_counter[0, 0]) == int(ref.step_counter[0, 0]) > 1000
    assert_close(out["image"].detach().cpu().numpy(), out0["image"].detach().numpy(), 1e-2, 2e-3, "image")
    assert_close(out["weights_sum"].detach().cpu().numpy(), out0["weights_sum"].detach().numpy(), 1e-2, 2e-3, "weights_sum")
    assert_close(out["depth"].detach().cpu().numpy(), out0["depth"].detach().numpy(), 1e-2, 2e-3, "depth")
    ((out["image"].reshape(-1, 3) - target.cuda()) ** 2).mean().backward()
    ((out0["image"].reshape(-1, 3) - target) ** 2).mean().backward()
    for name in ("network", "density_network", "rgb_network"):
        g, g0 = getattr(net, name).params.grad.cpu().numpy(), getattr(ref, name).params.grad.numpy()
        assert_close(g, g0, 5e-2, 2e-2 * np.abs(g0).max(), name + ".params.grad")
    g, g0 = net.pos_en.embeddings.grad.cpu().numpy(), ref.pos_en.embeddings.grad.numpy()
    assert_close(g, g0, 5e-2, 2e-2 * np.abs(g0).max(), "embeddings.grad")


def test_cuda_ray_eval_matches_oracle(scene):
    opt = torch_ref.default_opt(cuda_ray=True, train_conf=0)
    net, ref = _pair(opt, scene)
    sel = slice(7100, 7356)
    o, d = torch.from_numpy(scene["rays_o"][sel]), torch.from_numpy(scene["rays_d"][sel])
    net.eval(); ref.eval()
    with torch.no_grad():
        out = net.render(o[None].cuda(), d[None].cuda(), perturb=False)
        out0 = ref.render(o[None], d[None], perturb=False)
    assert_close(out["image"].cpu().numpy(), out0["image"].numpy(), 1e-2, 3e-3, "image")
    assert_close(out["weights_sum"].cpu().numpy(), out0["weights_sum"].numpy(), 1e-2, 3e-3, "weights_sum")


@pytest.mark.parametrize("soft_mask,detach_bg", [(False, False), (True, True)])
def test_dense_lgie_render_matches_oracle(soft_mask, detach_bg):
    """NeRFRenderer.run with the LGIE outputs (fg / bg / render_mask; renderer.py:383-403) at deterministic samples"""
    from customnerf_b200 import synthetic as syn
    opt = torch_ref.default_opt(cuda_ray=False, train_conf=0.01, soft_mask=soft_mask, detach_bg=detach_bg)
    net, ref = _pair(opt)
    o, d = syn.random_rays(256, seed=4)
    net.eval(); ref.eval()           # eval => deterministic importance sampling (det=True), no perturbation
    out = net.render(o[None].cuda(), d[None].cuda(), num_steps=32, upsample_steps=32, perturb=False)
    out0 = ref.render(o[None], d[None], num_steps=32, upsample_steps=32, perturb=False)
    for key in ("image", "render_mask", "weights_sum", "depth"):
        assert_close(out[key].detach().cpu().numpy(), out0[key].detach().numpy(), 2e-2, 5e-3, key)
    for part in ("fg", "bg"):
        a, b = out[part]["image"].detach().cpu().numpy(), out0[part]["image"].detach().numpy()
        if soft_mask:
            assert_close(a, b, 2e-2, 5e-3, part + ".image")
        else:
            # hard mask (masks > 0.5, renderer.py:391): a sample whose mask value sits within fp16 rounding of 0.5
            # switches sides, which moves a whole sample between fg and bg -- allow a few such rays
            bad = np.abs(a - b) > 5e-3 + 2e-2 * np.abs(b)
            assert bad.mean() < 0.03, (part, bad.mean())
    loss = out["image"].mean() + out["render_mask"].mean() + out["fg"]["image"].mean()
    loss.backward()
    assert net.pos_en.embeddings.grad.abs().sum() > 0 and torch.isfinite(net.rgb_network.params.grad).all()


def test_update_extra_state_builds_the_same_bitfield(scene):
    """occupancy-grid update (renderer.py:1658-1715): EMA + mean + packbits on the GPU vs the oracle, same densities"""
    from customnerf_b200 import raymarching as rm
    opt = torch_ref.default_opt(cuda_ray=True, train_conf=0)
    net, ref = _pair(opt, scene)
    net.update_extra_state()
    assert net.density_bitfield.dtype == torch.uint8 and net.iter_density == 1
    grid = net.density_grid.cpu().numpy()
    thr = min(net.mean_density, net.density_thresh)
    from oracle import cpu_ops
    assert np.array_equal(net.density_bitfield.cpu().numpy(), cpu_ops.packbits(grid, thr))
    assert abs(net.mean_density - float(grid.mean())) < 1e-4 * max(1.0, abs(float(grid.mean())))
    assert (grid >= 0).all()


def test_update_extra_state_reproduces_the_reference_bitfield(monkeypatch):
    """NeRFRenderer.update_extra_state on the GPU against the reference's own update_extra_state (renderer.py:1658-1715, run
    on the CPU by tests/golden/make_golden_python.py): same analytic density (evaluated on the host so that its values are the
    reference run's bit for bit), same jitter (the reference's torch.rand_like draws, replayed from the same CPU generator
    state), two consecutive updates (fresh grid, then the EMA-max path).  The grid must be bit-identical (SHA-256); the bit
    field too, except for cells that sit within one part in 10^6 of the threshold (the mean is summed in a different order on
    the device) -- none on this scene."""
    import hashlib
    import os
    from customnerf_b200 import synthetic as syn, trainer
    from customnerf_b200.nerf import NeRFNetwork
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python.npz"))
    net = NeRFNetwork(trainer.make_opt(bound=2, cuda_ray=True, min_near=0.01, density_thresh=10), encoding="hashgrid",
                      log2_hashmap_size=12, desired_resolution=64).cuda()
    net.density = lambda x: {"sigma": syn.bear_density(x.cpu()).cuda()}
    monkeypatch.setattr(torch, "rand_like", lambda t, **kw: torch.rand(t.shape, dtype=t.dtype).to(t.device))
    torch.manual_seed(123)
    net.local_step = 3
    net.step_counter[:3, 0] = torch.tensor([100, 200, 301], dtype=torch.int32, device="cuda")

    def digest(a):
        return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    for k in range(2):
        net.update_extra_state()
        grid, bits = net.density_grid.cpu().numpy(), net.density_bitfield.cpu().numpy()
        np.testing.assert_allclose(grid.reshape(-1)[::1009], G["occ%d_grid_sample" % k], rtol=0, atol=0)
        assert np.array_equal(digest(grid), G["occ%d_grid_sha256" % k]), "density grid differs from the reference's"
        md = float(G["occ%d_mean_density" % k])
        assert abs(float(net.mean_density) - md) <= 1e-6 * md
        thr = min(md, 10.0)
        borderline = int((np.abs(grid - thr) <= 1e-6 * thr).sum())
        if borderline == 0:
            assert np.array_equal(digest(bits), G["occ%d_bitfield_sha256" % k]), "bit field differs from the reference's"
        assert abs(int(np.unpackbits(bits).sum()) - int(G["occ%d_bits_set" % k])) <= borderline
        assert int(net.mean_count) == int(G["occ%d_mean_count" % k])
        assert [net.iter_density, net.local_step] == list(G["occ%d_iter_local" % k])


def test_train_step_harness_runs_and_reduces_loss(scene):
    """30 Adam steps on rays that hit the bear: the loss over those pixels goes down (evaluated without perturbation)"""
    from customnerf_b200 import trainer, synthetic as syn
    model = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=15, desired_resolution=512)
    ts = trainer.TrainStep(model, lr=1e-2)
    o, d = syn.camera_rays(105, 142)
    hit = torch.from_numpy(scene["rays_o"][:, 0] * 0).bool()
    from oracle import cpu_ops
    counts = cpu_ops.march_rays_count(scene["rays_o"], scene["rays_d"], 2.0, scene["bitfield"], 2, 128, scene["nears"],
                                      scene["fars"])
    sel = torch.from_numpy(np.nonzero(counts > 10)[0][:3000])
    o, d = o[sel].cuda(), d[sel].cuda()
    target = syn.bear_color(o + d * 1.5)

    def eval_loss():
        model.train()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out = model.render(o[None], d[None], perturb=False, force_all_rays=True)
        return float(((out["image"].reshape(-1, 3) - target) ** 2).mean())

    before = eval_loss()
    losses = [float(ts.step(o, d, target)) for _ in range(40)]
    after = eval_loss()
    assert np.isfinite(losses).all()
    assert after < 0.9 * before, (before, after)


@pytest.mark.parametrize("soft", [False, True])
def test_lgie_composites_on_cuda_path_match_dense_formulas(soft):
    """all / fg / bg composites, rendered mask, soft | hard edit mask and detach_bg on the occupancy path against the
    dense path's sample-level formulas (nerf/renderer.py:383-474) evaluated in fp64 on the same samples.
    Tolerance abs 2e-4 (T_thresh = 1e-4 early-out + __expf)."""
    from customnerf_b200.nerf import NeRFNetwork
    opt = torch_ref.default_opt(cuda_ray=True, train_conf=0.01, soft_mask=soft, detach_bg=True, conf_thr=0.5)
    net = NeRFNetwork(opt, encoding="hashgrid", **ENC).cuda()
    g = torch.Generator().manual_seed(11)
    counts = torch.randint(0, 70, (300,), generator=g)
    counts[::17] = 0
    offs = torch.cumsum(counts, 0) - counts
    M = int(counts.sum())
    rays = torch.stack([torch.arange(300), offs, counts], -1).int().cuda()
    sig = (torch.rand(M, generator=g) * 40).cuda().requires_grad_()
    rgb = torch.rand(M, 3, generator=g).cuda().requires_grad_()
    msk = torch.rand(M, 1, generator=g).cuda()
    deltas = torch.stack([torch.full((M,), 0.0034), torch.rand(M, generator=g) * 0.02 + 0.0034], -1).cuda()
    out = net._lgie_composites(sig, rgb, msk, deltas, rays, 1e-4, (300,))
    ws_all, _, img_all = out['_all']

    s64, c64, m64, d64 = (t.detach().double().cpu() for t in (sig, rgb, msk, deltas))
    em = torch.sigmoid((m64 - 0.5) * 100) if soft else (m64 > 0.5).double()

    def dense(sg):
        img, ws, rm = torch.zeros(300, 3, dtype=torch.float64), torch.zeros(300, dtype=torch.float64), torch.zeros(300, dtype=torch.float64)
        for r in range(300):
            o, n = int(offs[r]), int(counts[r])
            a = 1 - torch.exp(-sg[o:o + n] * d64[o:o + n, 0])
            T = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64), 1 - a]), 0)[:-1]
            w = a * T
            img[r], ws[r], rm[r] = (w[:, None] * c64[o:o + n]).sum(0), w.sum(), (w * m64[o:o + n, 0]).sum()
        return img, ws, rm
    for name, sg, got in (("all", s64, (img_all, ws_all, out['render_mask'])),
                          ("fg", s64 * em[:, 0], (out['fg']['image'], out['fg']['weights_sum'], out['fg']['render_mask'])),
                          ("bg", s64 * (1 - em[:, 0]), (out['bg']['image'], out['bg']['weights_sum'], out['bg']['render_mask']))):
        img, ws, rm = dense(sg)
        assert_close(got[0].detach().cpu().numpy(), img.numpy(), 0, 2e-4, name + " image")
        assert_close(got[1].detach().cpu().numpy(), ws.numpy(), 0, 2e-4, name + " weights_sum")
        assert_close(got[2].detach().cpu().numpy().reshape(-1), rm.numpy(), 0, 2e-4, name + " render_mask")
    # detach_bg: background samples (mask < 0.5) get no gradient from the "all" image, foreground samples do
    img_all.sum().backward()
    bgs = (msk[:, 0] < 0.5)
    assert float(sig.grad[bgs].abs().max()) == 0.0 and float(rgb.grad[bgs].abs().max()) == 0.0
    assert float(sig.grad[~bgs].abs().max()) > 0.0


def test_device_driven_inference_matches_the_host_loop(scene):
    """FusedInference (device-side round state, CUDA-graph rounds) against the reference-shaped host loop of
    NeRFRenderer.run_cuda: per ray the same samples are composited in the same order, so the frames must be identical."""
    opt = torch_ref.default_opt(cuda_ray=True, train_conf=0)
    net, _ = _pair(opt, scene)
    net.eval()
    sel = slice(3000, 3000 + 4096)
    o, d = torch.from_numpy(scene["rays_o"][sel]).cuda(), torch.from_numpy(scene["rays_d"][sel]).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        net.fast_inference = False
        slow = net.render(o[None], d[None], perturb=False)
        net.fast_inference = True
        fast = net.render(o[None], d[None], perturb=False)
        again = net.render(o[None], d[None], perturb=False)           # graph replay on a re-used instance
        assert net._infer.recorded_march                               # default: rounds served from one traversal's records
        fallbacks = net._infer.march_fallbacks
        net._infer.recorded_march, net._infer.graph = False, None      # every round walks the grid from rays_t
        walk = net.render(o[None], d[None], perturb=False)
        net._infer.recorded_march, net._infer.graph = True, None
        noise = torch.rand(4096, device="cuda")
        pert = {}
        for rec in (True, False):                                      # perturbed first round, both forms, same draws
            net._infer.recorded_march, net._infer.graph = rec, None
            w, dp, im, _, _ = net._infer.render(o, d, perturb=True, noises=noise)
            pert[rec] = (w.clone(), dp.clone(), im.clone())
        net._infer.recorded_march, net._infer.graph = True, None
    assert net._infer is not None and net._infer.rounds >= 8
    assert float(slow["weights_sum"].sum()) > 100
    for k in ("image", "depth", "weights_sum"):
        assert torch.equal(fast[k], slow[k]), k
        assert torch.equal(again[k], slow[k]), k
        assert torch.equal(walk[k], slow[k]), k
    assert torch.equal(fast["mask"], slow["mask"])
    for a, b in zip(pert[True], pert[False]):
        assert torch.equal(a, b)
    assert 0 <= fallbacks <= 0.05 * 4096, fallbacks                    # the recorded chain serves (nearly) every ray
