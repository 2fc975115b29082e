"""-m gpu: the fused, graph-captured train step (customnerf_b200/fused_trainer.py, csrc/fused_step.cu, csrc/optim.cu)
against the autograd composition of the drop-in ops (trainer.TrainStep) and against torch.optim.Adam.

Tolerances: gradients come out of the same kernels in both paths (fp16 activations, fp32 atomics whose order differs
run to run), so loss rel 1e-5, gradients rel 1e-2 of the largest entry (SURVEY.md Appendix D, fp16 grid grads).
The Adam kernel is compared with torch.optim.Adam on identical gradients: abs 1e-6 on O(1) parameters (2-4 ulp).
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N_RAYS = 2048


def _models():
    from customnerf_b200 import trainer
    a = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=15, desired_resolution=512, seed=3)
    b = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=15, desired_resolution=512, seed=3)
    with torch.no_grad():
        a.pos_en.embeddings.uniform_(-0.5, 0.5)
        b.pos_en.embeddings.copy_(a.pos_en.embeddings)
    return a, b


def _batch():
    from customnerf_b200 import synthetic as syn
    o, d = syn.camera_rays(105, 142)
    sel = torch.arange(5000, 5000 + N_RAYS)
    o, d = o[sel].contiguous(), d[sel].contiguous()
    return o.cuda(), d.cuda(), syn.bear_color(o + d * 1.5).cuda()


def test_forward_backward_matches_autograd_path():
    from customnerf_b200 import trainer, fused_trainer
    ma, mb = _models()
    o, d, tgt = _batch()
    ref = trainer.TrainStep(ma, perturb=False)
    loss_ref = ref.forward_backward(o, d, tgt)
    fs = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False, use_graph=False)
    fs.m_cap or fs._alloc_samples(fs._round_cap(fs.measure_samples(o, d)))
    fs.set_batch(o, d, tgt)
    fs.forward_backward()
    loss, samples, used = fs.last_stats()
    assert samples == used == int(ma.step_counter[0, 0]) > 1000
    assert abs(loss - float(loss_ref)) <= 1e-5 * abs(float(loss_ref)) + 1e-7
    for name, off, n in fs.layout:
        mod, attr = name.split(".")
        g_ref = getattr(getattr(ma, mod), attr).grad.reshape(-1).float().cpu().numpy()
        g = fs.grads_flat[off:off + n].cpu().numpy()
        scale = np.abs(g_ref).max()
        assert scale > 0
        assert np.abs(g - g_ref).max() <= 1e-2 * scale, (name, np.abs(g - g_ref).max(), scale)


def test_capacity_overflow_drops_whole_rays_and_recovers():
    from customnerf_b200 import fused_trainer
    _, mb = _models()
    o, d, tgt = _batch()
    fs = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False, use_graph=False, m_cap=4096)
    fs.set_batch(o, d, tgt)
    fs.forward_backward()
    torch.cuda.synchronize()
    rays = fs.rays.cpu().numpy()
    loss, samples, used = fs.last_stats()          # grows the buffers
    assert samples > 4096 and used <= 4096 and np.isfinite(loss)
    fits = rays[:, 1] + rays[:, 2] <= 4096
    assert used == (rays[fits, 1] + rays[fits, 2]).max()      # complete segments only
    assert fs.m_cap >= samples and fs.overflows == 1
    fs.grads_flat.zero_()
    fs.set_batch(o, d, tgt)
    fs.forward_backward()
    _, samples2, used2 = fs.last_stats()
    assert samples2 == used2 == samples


def test_fused_adam_matches_torch_adam():
    from customnerf_b200 import _lib as L
    lib = L.lib()
    torch.manual_seed(0)
    n, split = 40000 + 8, 30000
    p0 = torch.randn(n, device="cuda")
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    ref_a = torch.nn.Parameter(p0[:split].clone())
    ref_b = torch.nn.Parameter(p0[split:].clone())
    opt = torch.optim.Adam([{"params": [ref_a], "lr": 5e-3}, {"params": [ref_b], "lr": 5e-4}], betas=(0.9, 0.99), eps=1e-15)
    sched = torch.tensor([5e-3, 5e-4, 0.9, 0.99, 1e-15, 1.0 / 128.0, 0.1, 100.0], device="cuda")
    lam = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** min(it / 100.0, 1))
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    hyper = torch.zeros(16, device="cuda")
    for it in range(5):
        g = torch.randn(n, device="cuda") * (10.0 ** (-it))
        g[::7] = 0.0
        ref_a.grad, ref_b.grad = g[:split].clone(), g[split:].clone()
        opt.step(); lam.step()
        gs = (g * 128.0).contiguous()
        L.check(lib.nb200_adam_hyper(L.ptr(step), L.ptr(sched), L.ptr(hyper), L.stream()), "hyper")
        L.check(lib.nb200_fused_adam(L.ptr(p), L.ptr(gs), L.ptr(m), L.ptr(v), C.c_uint64(n), C.c_uint64(split),
                                     L.ptr(hyper), C.c_int(1), L.stream()), "adam")
        assert float(gs.abs().max()) == 0.0
        ref = torch.cat([ref_a.detach(), ref_b.detach()])
        err = (p - ref).abs().max().item()
        assert err <= 1e-6, (it, err)      # parameters are O(1): one fp32 ulp is 2.4e-7 .. 4.8e-7
    assert int(step) == 5


def test_graph_replay_trains_and_matches_eager_launches():
    from customnerf_b200 import fused_trainer
    ma, mb = _models()
    o, d, tgt = _batch()
    eager = fused_trainer.FusedTrainStep(ma, N_RAYS, perturb=False, use_graph=False)
    graph = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False, use_graph=True)
    mc, _ = _models()
    staged = fused_trainer.FusedTrainStep(mc, N_RAYS, perturb=False, use_graph=True)
    ho, hd, ht = staged.pinned_batch()                 # batch handed over in the pinned staging buffer:
    ho.copy_(o.cpu()); hd.copy_(d.cpu()); ht.copy_(tgt.cpu())     # one H2D copy on the copy stream ahead of the step's graph
    losses_e, losses_g, losses_s = [], [], []
    for _ in range(6):
        eager.step(o, d, tgt); losses_e.append(eager.last_stats()[0])
        graph.step(o, d, tgt); losses_g.append(graph.last_stats()[0])
        staged.step(ho, hd, ht); losses_s.append(staged.last_stats()[0])
    assert set(staged.graphs) == {False}               # slot 0 is where the resident batch lives: one graph serves both
    assert losses_e[-1] < losses_e[0]                  # it trains
    assert int(graph.step_count) == int(eager.step_count) == int(staged.step_count) == 6
    np.testing.assert_allclose(losses_g, losses_e, rtol=2e-2)
    np.testing.assert_allclose(losses_s, losses_e, rtol=2e-2)
    # parameters stay views of the flat vector (state-dict names unchanged)
    assert mb.pos_en.embeddings.data_ptr() == graph.params_flat.data_ptr()
    assert set(k for k in mb.state_dict() if "params" in k or "embeddings" in k) == {
        "pos_en.embeddings", "network.params", "density_network.params", "rgb_network.params"}


def test_async_loop_two_staging_slots_and_lagged_stats():
    """the asynchronous loop: step k + 1 is issued from the other pinned staging slot before step k's result is read
    (previous_stats): same losses, in the same order, as the synchronous loop"""
    from customnerf_b200 import fused_trainer
    ma, mb = _models()
    o, d, tgt = _batch()
    sync = fused_trainer.FusedTrainStep(ma, N_RAYS, perturb=False)
    asyn = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False)
    slots = [asyn.pinned_batch(0), asyn.pinned_batch(1)]
    assert slots[0][0].data_ptr() != slots[1][0].data_ptr()
    for ho, hd, ht in slots:
        ho.copy_(o.cpu()); hd.copy_(d.cpu()); ht.copy_(tgt.cpu())
    want, got = [], []
    assert asyn.previous_stats() is None
    for k in range(7):
        sync.step(o, d, tgt); want.append(sync.last_stats()[0])
        asyn.step(*slots[k & 1])
        prev = asyn.previous_stats()
        assert (prev is None) == (k == 0)
        if prev is not None:
            got.append(prev[0])
            assert prev[1] == prev[2] > 1000
    got.append(asyn.last_stats()[0])
    assert set(asyn.graphs) == {False, "rays1"}                    # one captured graph per staging slot
    np.testing.assert_allclose(got, want, rtol=2e-2)
    assert got[-1] < got[0]


def test_optimizer_state_round_trip_and_torch_adam_format():
    """optimizer_state_dict() loads into the reference's optimiser object (torch.optim.Adam over get_params) and back:
    a run resumed from it continues exactly like the uninterrupted one"""
    from customnerf_b200 import fused_trainer
    ma, mb = _models()
    o, d, tgt = _batch()
    a = fused_trainer.FusedTrainStep(ma, N_RAYS, perturb=False)
    for _ in range(3):
        a.step(o, d, tgt)
    a.last_stats()
    sd = a.optimizer_state_dict()
    # the reference's optimiser accepts it (main.py:182) and hands the same state back
    opt = torch.optim.Adam(ma.get_params(5e-4), betas=(0.9, 0.99), eps=1e-15)
    opt.load_state_dict(sd)
    sd2 = opt.state_dict()
    assert [g["lr"] for g in sd2["param_groups"]] == [5e-3, 5e-4, 5e-4, 5e-4]
    assert torch.equal(sd2["state"][0]["exp_avg"].reshape(-1), a.exp_avg[:a.layout[0][2]])
    assert int(float(sd2["state"][3]["step"])) == 3
    # resume in a fresh trainer: parameters + optimiser state -> identical continuation
    with torch.no_grad():
        for (_, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            pb.copy_(pa)
    b = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False)
    b.load_optimizer_state_dict(sd2)
    la, lb = [], []
    for _ in range(3):
        a.step(o, d, tgt); la.append(a.last_stats()[0])
        b.step(o, d, tgt); lb.append(b.last_stats()[0])
    assert int(b.step_count) == 6
    np.testing.assert_allclose(lb, la, rtol=1e-3)       # fp32 atomics order differs run to run; no drift beyond it


def test_stage_profile_reports_every_stage():
    from customnerf_b200 import fused_trainer
    _, mb = _models()
    o, d, tgt = _batch()
    fs = fused_trainer.FusedTrainStep(mb, N_RAYS, use_graph=False)
    fs.step(o, d, tgt)
    prof = fs.profile_stages(3)
    assert list(prof) == fused_trainer.STAGES
    assert all(v > 0 for v in prof.values())


def test_graph_capture_includes_nccl_allreduce():
    """The sharded step's only collective (all-reduce of the flat gradient) is captured inside the step's CUDA graph.
    A one-rank NCCL group exercises exactly that capture path on a single GPU; with one rank the sum is the identity,
    so the losses must follow the un-synchronised run."""
    import socket
    import torch.distributed as dist
    from customnerf_b200 import fused_trainer
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)
    try:
        ma, mb = _models()
        o, d, tgt = _batch()
        plain = fused_trainer.FusedTrainStep(ma, N_RAYS, perturb=False, use_graph=True)
        synced = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False, use_graph=True,
                                              grad_sync=lambda flat: dist.all_reduce(flat, op=dist.ReduceOp.SUM))
        mc, _ = _models()
        piped = fused_trainer.FusedTrainStep(mc, N_RAYS, perturb=False, use_graph=True, allreduce_chunks=3)
        la, lb, lc = [], [], []
        for _ in range(4):
            plain.step(o, d, tgt); la.append(plain.last_stats()[0])
            synced.step(o, d, tgt); lb.append(synced.last_stats()[0])
            piped.step(o, d, tgt); lc.append(piped.last_stats()[0])     # chunked all-reduce pipelined with Adam
        np.testing.assert_allclose(lb, la, rtol=2e-2)
        np.testing.assert_allclose(lc, la, rtol=2e-2)
        assert lb[-1] < lb[0] and lc[-1] < lc[0]
        assert int(piped.step_count) == 4
    finally:
        dist.destroy_process_group()


def test_fused_step_at_scale_config_shapes():
    """BASELINE.json configs[4] shapes in miniature: 2^22-entry hash table (79 M parameters, 302 MB fp32) and a
    150 k-ray batch (~2.7 M samples): exercises 64-bit indexing, the multi-trip ray scan and the TMA activation stores."""
    from customnerf_b200 import trainer, fused_trainer, synthetic as syn
    model = trainer.build_scene_model(torch.device("cuda"), log2_hashmap_size=22, desired_resolution=2048, seed=1)
    assert model.pos_en.embeddings.shape[0] == 39625280
    n = 150000
    o, d = syn.random_rays(n, seed=5)
    tgt = syn.bear_color(o + d * 1.5)
    fs = fused_trainer.FusedTrainStep(model, n, use_graph=True)
    losses = []
    for _ in range(4):
        fs.step(o.cuda(), d.cuda(), tgt.cuda())
        loss, samples, used = fs.last_stats()
        assert samples == used > 500000 and np.isfinite(loss)
        losses.append(loss)
    assert losses[-1] < losses[0]


def test_pipelined_update_matches_the_serial_step():
    """pipeline_update=True runs step k's Adam on a second stream next to step k + 1's march; after flush() the model
    must be where the serial trainer is after the same number of steps (same arithmetic; only atomics order differs)."""
    from customnerf_b200 import fused_trainer
    ma, mb = _models()
    o, d, tgt = _batch()
    serial = fused_trainer.FusedTrainStep(ma, N_RAYS, perturb=False, use_graph=True)
    piped = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False, use_graph=True, pipeline_update=True)
    ls, lp = [], []
    for _ in range(6):
        serial.step(o, d, tgt); ls.append(serial.last_stats()[0])
        piped.step(o, d, tgt); lp.append(piped.last_stats()[0])
    np.testing.assert_allclose(lp, ls, rtol=2e-2)
    assert int(piped.step_count) == 5 and int(serial.step_count) == 6      # one update still pending
    piped.flush()
    assert int(piped.step_count) == 6
    # parameters after the same six updates: Adam's sign-like first steps amplify atomics-order noise on entries whose
    # gradient is ~0, so compare the bulk of the table, not every entry
    pa, pb = serial.params_flat.float(), piped.params_flat.float()
    close = ((pa - pb).abs() <= 2e-3 + 1e-2 * pa.abs()).float().mean().item()
    assert close > 0.97, close
    piped.step(o, d, tgt)                                                  # primes again after a flush
    assert np.isfinite(piped.last_stats()[0])


@pytest.mark.parametrize("rgb_w", [1.0, 0.25])
def test_mask_loss_term_matches_the_autograd_path(rgb_w):
    """train_rgb / train_conf: loss = rgb_w * MSE(image, target) + w * MSE(render_mask, gt_mask) (nerf/utils_init_nerf.py:224-234) with
    render_mask = sum_i w_i mask_i composited from the 4th field output.  Fused step (mask as a 4th composited channel)
    against autograd over the drop-in ops (run_cuda + _lgie_composites): loss rel 1e-5, gradients 1e-2 of the largest."""
    import torch.nn.functional as F
    from customnerf_b200 import trainer, fused_trainer
    w = 0.5
    opt = trainer.make_opt(train_conf=w)
    dev = torch.device("cuda")
    ma = trainer.build_scene_model(dev, log2_hashmap_size=15, desired_resolution=512, opt=opt, seed=3)
    mb = trainer.build_scene_model(dev, log2_hashmap_size=15, desired_resolution=512, opt=opt, seed=3)
    with torch.no_grad():
        ma.pos_en.embeddings.uniform_(-0.5, 0.5)
        mb.pos_en.embeddings.copy_(ma.pos_en.embeddings)
    o, d, tgt = _batch()
    gt_mask = (torch.rand(N_RAYS, generator=torch.Generator().manual_seed(9)) > 0.5).float().cuda()
    ma.train()
    with torch.autocast("cuda", dtype=torch.float16):
        out = ma.render(o[None], d[None], staged=False, perturb=False, force_all_rays=True, **vars(opt))
        loss_ref = (rgb_w * F.mse_loss(out["image"].reshape(-1, 3), tgt, reduction="sum") / (3.0 * N_RAYS)
                    + w * F.mse_loss(out["render_mask"].reshape(-1), gt_mask, reduction="sum") / N_RAYS)
    (loss_ref * fused_trainer.LOSS_SCALE).backward()
    fs = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False, use_graph=False, mask_weight=w, rgb_weight=rgb_w)
    fs._alloc_samples(fs._round_cap(fs.measure_samples(o, d)))
    fs.set_batch(o, d, tgt, gt_mask)
    fs.forward_backward()
    loss, samples, used = fs.last_stats()
    assert samples == used > 1000
    ref = float(loss_ref.detach())
    assert abs(loss - ref) <= 1e-5 * abs(ref) + 1e-7, (loss, ref)
    np.testing.assert_allclose(fs.render_mask.cpu().numpy(), out["render_mask"].reshape(-1).detach().float().cpu().numpy(), atol=2e-4)
    for name, off, n in fs.layout:
        mod, attr = name.split(".")
        g_ref = getattr(getattr(ma, mod), attr).grad.reshape(-1).float().cpu().numpy()
        g = fs.grads_flat[off:off + n].cpu().numpy()
        scale = np.abs(g_ref).max()
        assert np.abs(g - g_ref).max() <= 1e-2 * scale, (name, np.abs(g - g_ref).max(), scale)
    # the mask head really receives gradient: rows 3 of the colour head's last layer
    gr = fs.grads_flat[fs.layout[3][1]:].cpu().numpy()[64 * 96:].reshape(16, 64)
    assert np.abs(gr[3]).sum() > 0


def test_dynamic_loss_scale_skips_the_step_on_a_non_finite_gradient():
    """GradScaler semantics on the device (the reference trains under torch.cuda.amp.GradScaler, utils_init_nerf.py:100,612-629):
    a non-finite gradient -> the optimiser step is skipped (parameters, moments and the step count untouched, the gradient
    reset), the loss scale is halved; clean steps count towards doubling it.  No host synchronisation is involved: the same
    sequence runs inside the captured graph."""
    from customnerf_b200 import fused_trainer
    _, mb = _models()
    o, d, tgt = _batch()
    for use_graph in (False, True):
        fs = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False, use_graph=use_graph, lr=1e-3)
        fs.scaler[6] = 3                                    # growth interval: 3 clean steps
        for _ in range(2):
            fs.step(o, d, tgt)
        loss, _, _ = fs.last_stats()
        scale, skipped, steps = fs.scaler_state()
        assert (scale, skipped, steps) == (128.0, 0, 2) and np.isfinite(loss)
        p0, m0, v0 = fs.params_flat.clone(), fs.exp_avg.clone(), fs.exp_avg_sq.clone()
        bad = tgt.clone()
        bad[:, 1] = float("inf")                            # a poisoned channel: the loss and every gradient become inf / NaN
        fs.step(o, d, bad)
        fs.last_stats()
        scale, skipped, steps = fs.scaler_state()
        assert (scale, skipped, steps) == (64.0, 1, 2), (scale, skipped, steps)
        assert torch.equal(fs.params_flat, p0) and torch.equal(fs.exp_avg, m0) and torch.equal(fs.exp_avg_sq, v0)
        assert float(fs.grads_flat.abs().max()) == 0.0 and torch.isfinite(fs.params_flat).all()
        for _ in range(3):                                  # three clean steps: the update resumes, then the scale doubles
            fs.step(o, d, tgt)
        loss2, _, _ = fs.last_stats()
        scale, skipped, steps = fs.scaler_state()
        assert (scale, skipped, steps) == (128.0, 1, 5), (scale, skipped, steps)
        assert not torch.equal(fs.params_flat, p0) and torch.isfinite(fs.params_flat).all() and np.isfinite(loss2)
        assert loss2 < loss


def test_dynamic_loss_scale_changes_nothing_while_gradients_are_finite():
    """with finite gradients the scaled step is the constant-scale step: same parameters after 4 steps (the unscale factor
    1 / 128 is the same number either way)"""
    from customnerf_b200 import fused_trainer
    outs = []
    for dyn in (True, False):
        _, mb = _models()
        o, d, tgt = _batch()
        fs = fused_trainer.FusedTrainStep(mb, N_RAYS, perturb=False, use_graph=False, lr=1e-3, dynamic_loss_scale=dyn)
        for _ in range(4):
            fs.step(o, d, tgt)
        fs.last_stats()
        outs.append((fs.params_flat.clone(), int(fs.step_count)))
    assert outs[0][1] == outs[1][1] == 4
    # fp32 atomics in the scatter: order differs run to run, so compare to the run-to-run noise level, not bit for bit
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 2e-3
