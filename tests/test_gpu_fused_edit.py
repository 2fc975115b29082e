"""-m gpu: the fused LGIE editing step (customnerf_b200/fused_edit.py, nb200_train_lgie_forward/backward, the gated
composite kernels) against the autograd composition of the drop-in ops (NeRFRenderer.run_cuda + _lgie_composites, the
occupancy-path form of nerf/renderer.py:383-474).

Tolerances: rendered outputs rel 1e-4 (same kernels, fp32 compositing); loss rel 1e-4; parameter gradients within 1e-2 of
the largest entry (fp16 activations, fp32 atomics in another order), as tests/test_gpu_fused_step.py.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
N_RAYS = 2048


def _split_mask_head(model, spread=12.0):
    """A freshly initialised mask head sits on one side of 0.5 everywhere (empty foreground or empty background).  Shift and
    widen its output row -- w3 <- spread * (w3 - beta * 1), beta found by bisection so that the median mask over points of
    the scene is 0.5 (sum_j h_j > 0 after the ReLU, so the median falls monotonically with beta) -- to get a real split."""
    g = torch.Generator(device="cuda").manual_seed(5)
    o, d = _batch()[:2]                      # points along the test's own rays: the head sees these view directions
    o, d = o.repeat(4, 1), d.repeat(4, 1)
    x = o + d * (0.7 + 1.6 * torch.rand(o.shape[0], 1, device="cuda", generator=g))
    row = model.rgb_network.params[-16 * 64:].view(16, 64)[3]
    w3 = row.detach().clone()

    def median_mask(beta):
        with torch.no_grad():
            row.copy_(spread * (w3 - beta))
            with torch.autocast("cuda", dtype=torch.float16):
                out = model(x, d)
            return float(out[1][:, 3].float().median())
    lo, hi = -4.0, 4.0
    assert median_mask(lo) > 0.5 > median_mask(hi)
    for _ in range(30):
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if median_mask(mid) > 0.5 else (lo, mid)
    median_mask(0.5 * (lo + hi))
    return row.detach().clone()


def _models(**flags):
    """two identical single-head models.  detach_mask_from_field is switched on AFTER construction: at construction time the
    flag selects the two-head RGB_network (network_grid.py:117-118), which the fused kernels do not cover (separate test);
    here it exercises the other half of its meaning, the mask composited with detached weights (renderer.py:437-441)"""
    from customnerf_b200 import trainer
    late = {k: flags.pop(k) for k in ("detach_mask_from_field",) if k in flags}
    opt = dict(train_conf=0.01, **flags)
    a = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=15, desired_resolution=512, seed=3, opt=trainer.make_opt(**opt))
    b = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=15, desired_resolution=512, seed=3, opt=trainer.make_opt(**opt))
    for m in (a, b):
        for k, v in late.items():
            setattr(m.opt, k, v)
    with torch.no_grad():
        a.pos_en.embeddings.uniform_(-0.5, 0.5)
        b.pos_en.embeddings.copy_(a.pos_en.embeddings)
        row = _split_mask_head(a)
        b.rgb_network.params[-16 * 64:].view(16, 64)[3].copy_(row)
    return a, b


def _batch():
    from customnerf_b200 import synthetic as syn
    o, d = syn.camera_rays(105, 142)
    sel = torch.arange(5000, 5000 + N_RAYS)
    o, d = o[sel].contiguous(), d[sel].contiguous()
    tgt = syn.bear_color(o + d * 1.5).cuda()
    return o.cuda(), d.cuda(), tgt, (tgt.mean(-1, keepdim=True) > 0.5).float()


def _make_loss(tgt, gt_mask):
    def loss_fn(out):        # touches every rendered output that carries a gradient
        return (F.mse_loss(out["image"].reshape(-1, 3).float(), tgt) +
                F.mse_loss(out["fg"]["image"].reshape(-1, 3).float(), tgt * gt_mask) +
                F.mse_loss(out["bg"]["image"].reshape(-1, 3).float(), tgt * (1 - gt_mask)) +
                0.5 * F.mse_loss(out["render_mask"].reshape(-1, 1).float(), gt_mask) +
                0.25 * F.mse_loss(out["fg"]["render_mask"].reshape(-1, 1).float(), gt_mask) +
                0.1 * (out["bg"]["weights_sum"].reshape(-1).float() ** 2).mean() +
                0.1 * (out["weights_sum"].reshape(-1).float() - 1).abs().mean())
    return loss_fn


@pytest.mark.parametrize("flags", [dict(soft_mask=True, detach_bg=True), dict(soft_mask=False, detach_bg=False),
                                   dict(soft_mask=True, detach_bg=False, detach_mask_from_field=True)])
def test_edit_step_matches_autograd_composition(flags):
    from customnerf_b200 import trainer, fused_edit
    ma, mb = _models(**dict(flags))
    o, d, tgt, gt_mask = _batch()
    loss_fn = _make_loss(tgt, gt_mask)
    # reference-shaped path: autograd over the drop-in ops
    ma.train()
    with torch.autocast("cuda", dtype=torch.float16):
        out = ma.render(o[None], d[None], staged=False, perturb=False, force_all_rays=True, **vars(ma.opt))
        loss_ref = loss_fn(out)
    (loss_ref * trainer.LOSS_SCALE).backward()
    # fused path
    fs = fused_edit.FusedEditStep(mb, N_RAYS, loss_fn, perturb=False, use_graph=False)
    fs._alloc_samples(fs._round_cap(fs.measure_samples(o, d)))
    fs.set_batch(o, d, tgt)
    fs.forward_backward()
    loss, samples, used = fs.last_stats()
    assert samples == used > 1000
    got = fs.outputs()
    for name, ref_d, got_d in (("all", out, got), ("fg", out["fg"], got["fg"]), ("bg", out["bg"], got["bg"])):
        for key in ("image", "weights_sum", "depth", "render_mask"):
            r, g = ref_d[key].detach().float().cpu().numpy().reshape(-1), got_d[key].cpu().numpy().reshape(-1)
            assert np.abs(g - r).max() <= 1e-4 * max(1.0, np.abs(r).max()), (name, key, np.abs(g - r).max())
    assert float(got["fg"]["weights_sum"].sum()) > 1.0 and float(got["bg"]["weights_sum"].sum()) > 1.0   # both non-trivial
    ref = float(loss_ref.detach())
    assert abs(loss - ref) <= 1e-4 * abs(ref) + 1e-7
    for name, off, n in fs.layout:
        mod, attr = name.split(".")
        g_ref = getattr(getattr(ma, mod), attr).grad.reshape(-1).float().cpu().numpy()
        g = fs.grads_flat[off:off + n].cpu().numpy()
        scale = np.abs(g_ref).max()
        assert scale > 0
        assert np.abs(g - g_ref).max() <= 1e-2 * scale, (flags, name, np.abs(g - g_ref).max(), scale)


def test_edit_step_graph_replay_trains():
    from customnerf_b200 import fused_edit
    _, mb = _models(soft_mask=True, detach_bg=True)
    _, mc = _models(soft_mask=True, detach_bg=True)
    o, d, tgt, gt_mask = _batch()
    loss_fn = _make_loss(tgt, gt_mask)
    runs = {}
    for graph, model in ((True, mb), (False, mc)):
        fs = fused_edit.FusedEditStep(model, N_RAYS, loss_fn, perturb=False, use_graph=graph)
        ls = []
        for it in range(8):
            fs.step(o, d, tgt)
            ls.append(fs.last_stats()[0])
        runs[graph] = ls
    np.testing.assert_allclose(runs[True], runs[False], rtol=2e-3)     # replayed graph == direct launches
    assert runs[True][-1] < runs[True][0]


def test_edit_step_needs_the_mask_head():
    from customnerf_b200 import trainer, fused_edit
    model = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=15, desired_resolution=512)
    with pytest.raises(RuntimeError):
        fused_edit.FusedEditStep(model, 128, lambda out: out["image"].sum())


def test_gated_composite_kernels_match_the_cpu_oracle():
    """nb200_fs_composite_lgie_forward / _backward (C ABI) against oracle.cpu_ops.composite_lgie_* on ragged rays
    (empty rays, a ray longer than one warp trip, early termination at T_thresh): fp32 rel 1e-4; the three variants'
    backward launches accumulate into one set of per-sample rows."""
    import ctypes as C
    from customnerf_b200 import _lib as L
    from oracle import cpu_ops
    lib = L.lib()
    rng = np.random.RandomState(11)
    counts = np.array([40, 0, 3, 97, 1, 64, 0, 33], np.int32)
    N = counts.size
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    order = rng.permutation(N).astype(np.int32)                         # rays rows in any order: column 0 is the ray id
    rays = np.stack([np.arange(N, dtype=np.int32), offs, counts], 1)[order]
    M = int(counts.sum())
    sig = rng.uniform(0, 60, M).astype(np.float32)
    sig[offs[3]:offs[3] + 30] = 400.0                                    # ray 3 saturates: T < 1e-4 inside the segment
    rgba_h = np.concatenate([rng.uniform(0, 1, (M, 3)), np.clip(rng.normal(0.5, 0.02, (M, 1)), 0, 1)], 1).astype(np.float16)
    rgb, msk = rgba_h[:, :3].astype(np.float32), rgba_h[:, 3].astype(np.float32)
    dl = np.stack([rng.uniform(0.002, 0.01, M), rng.uniform(0.002, 0.01, M)], 1).astype(np.float32)
    g_ws, g_img, g_m = (rng.randn(3, N).astype(np.float32), rng.randn(3, N, 3).astype(np.float32), rng.randn(3, N).astype(np.float32))
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_sig, d_rgba_t, d_dl, d_rays = cu(sig), cu(rgba_h), cu(dl), cu(rays)
    for soft, dbg, dmf in ((1, 1, 0), (0, 0, 0), (1, 0, 1)):
        ws, dp, img, rm = (torch.zeros(3, N, device="cuda"), torch.zeros(3, N, device="cuda"),
                           torch.zeros(3, N, 3, device="cuda"), torch.zeros(3, N, device="cuda"))
        for v in range(3):
            L.check(lib.nb200_fs_composite_lgie_forward(C.c_int(v), L.ptr(d_sig), L.ptr(d_rgba_t), L.ptr(d_dl), L.ptr(d_rays), L.u32(M),
                                                        L.u32(N), L.f32(1e-4), L.f32(0.5), C.c_int(soft), L.ptr(ws[v]), L.ptr(dp[v]),
                                                        L.ptr(img[v]), L.ptr(rm[v]), L.stream()), "lgie_fwd")
        gs, grgba = torch.full((M,), 7.0, device="cuda"), torch.full((M, 4), 7.0, device="cuda")    # variant 0 must overwrite
        t_gws, t_gimg, t_gm = cu(g_ws), cu(g_img), cu(g_m)
        for v in range(3):
            L.check(lib.nb200_fs_composite_lgie_backward(C.c_int(v), L.ptr(t_gws[v]), L.ptr(t_gimg[v]), L.ptr(t_gm[v]), L.ptr(d_sig),
                                                         L.ptr(d_rgba_t), L.ptr(d_dl), L.ptr(d_rays), L.ptr(ws[v]), L.ptr(img[v]),
                                                         L.ptr(rm[v]), L.u32(M), L.u32(N), L.f32(1e-4), L.f32(0.5), C.c_int(soft),
                                                         C.c_int(dbg), C.c_int(dmf), L.ptr(gs), L.ptr(grgba), L.stream()), "lgie_bwd")
        want_s, want_c, want_m = np.zeros(M, np.float32), np.zeros((M, 3), np.float32), np.zeros(M, np.float32)
        for v in range(3):
            o_ws, o_dp, o_img, o_rm = cpu_ops.composite_lgie_forward(v, sig, rgb, msk, dl, rays, 1e-4, bool(soft), 0.5)
            for got, want, what in ((ws[v], o_ws, "weights_sum"), (dp[v], o_dp, "depth"), (img[v], o_img, "image"), (rm[v], o_rm, "render_mask")):
                got = got.cpu().numpy()
                assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), (v, soft, what, np.abs(got - want).max())
            a, b, c = cpu_ops.composite_lgie_backward(v, g_ws[v], g_img[v], g_m[v], sig, rgb, msk, dl, rays, 1e-4, bool(soft), 0.5,
                                                      bool(dbg), bool(dmf))
            want_s += a; want_c += b; want_m += c
        got_s, got_q = gs.cpu().numpy(), grgba.cpu().numpy()
        for got, want, what in ((got_s, want_s, "d_sigma"), (got_q[:, :3], want_c, "d_rgb"), (got_q[:, 3], want_m, "d_mask")):
            assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), (soft, dbg, dmf, what, np.abs(got - want).max())


def test_edit_step_refuses_the_two_head_network():
    """--detach_mask_from_field builds RGB_network (two MLPs); the fused steps cover the single 3 + 1 head and say so"""
    from customnerf_b200 import trainer, fused_edit, fused_trainer
    opt = trainer.make_opt(train_conf=0.01, detach_mask_from_field=True)
    m = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=14, desired_resolution=256, opt=opt)
    assert m.two_heads
    with pytest.raises(RuntimeError, match="two-head"):
        fused_edit.FusedEditStep(m, 256, lambda out: out["image"].sum())
    with pytest.raises(RuntimeError, match="two-head"):
        fused_trainer.FusedTrainStep(m, 256)
