"""-m gpu: ray generation on the device (csrc/raygen.cu, nerf/provider_utils.py: get_rays) against the golden vectors of
the reference's get_rays, against the CPU oracle at the bench image size, and inside the fused train step.

Tolerance: directions fp32 abs 1e-6 (unit vectors; the rotation's three products may be summed in another order than
torch's bmm), origins bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import cpu_ops

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_kernel_matches_reference_vectors():
    from customnerf_b200.nerf import provider_utils as pu
    G = np.load(os.path.join(HERE, "golden", "ref_get_rays.npz"))
    for name in ("full", "subset", "offset0"):
        H, W, N = [int(v) for v in G[name + "_HWN"]]
        fx, fy, cx, cy = [float(v) for v in G[name + "_intrinsics"]]
        inds = torch.from_numpy(G[name + "_inds"]).cuda() if N > 0 else None
        o, d = pu.ray_kernel(torch.from_numpy(G[name + "_poses"]).cuda(), fx, fy, cx, cy, H, W, inds, tuple(G[name + "_offset"]))
        assert np.array_equal(o.cpu().numpy(), G[name + "_rays_o"]), name
        err = np.abs(d.cpu().numpy() - G[name + "_rays_d"]).max()
        assert err <= 1e-6, (name, err)


def test_get_rays_api_full_image_and_subsets():
    from customnerf_b200 import synthetic as syn
    from customnerf_b200.nerf import get_rays
    H, W = 105, 142
    pose, intr = syn.camera_pose(H, W, view=3)
    poses = torch.stack([pose, syn.camera_pose(H, W, view=4)[0]]).cuda()
    out = get_rays(poses, intr, H, W)
    o_ref, d_ref = cpu_ops.get_rays(poses.cpu().numpy(), intr, H, W)
    assert set(out) == {"rays_o", "rays_d"} and out["rays_d"].shape == (2, H * W, 3)
    assert np.array_equal(out["rays_o"].cpu().numpy(), o_ref)
    assert np.abs(out["rays_d"].cpu().numpy() - d_ref).max() <= 1e-6
    # the synthetic camera's rays are the same rays (bench.py builds its batches with camera_rays)
    o_syn, d_syn = syn.camera_rays(H, W, view=3)
    assert np.abs(out["rays_d"][0].cpu().numpy() - d_syn.numpy()).max() <= 2e-6
    assert np.abs(out["rays_o"][0].cpu().numpy() - o_syn.numpy()).max() <= 1e-6
    # random pixel subset (:263-267) and error-map importance sampling (:268-282)
    torch.manual_seed(0)
    sub = get_rays(poses, intr, H, W, N=500)
    assert sub["inds"].shape == (2, 500) and sub["rays_d"].shape == (2, 500, 3)
    _, d_sub = cpu_ops.get_rays(poses.cpu().numpy(), intr, H, W, sub["inds"].cpu().numpy())
    assert np.abs(sub["rays_d"].cpu().numpy() - d_sub).max() <= 1e-6
    em = torch.rand(2, 128 * 128)
    imp = get_rays(poses, intr, H, W, N=300, error_map=em)
    assert imp["inds_coarse"].shape == (2, 300) and int(imp["inds"].max()) < H * W and int(imp["inds"].min()) >= 0
    _, d_imp = cpu_ops.get_rays(poses.cpu().numpy(), intr, H, W, imp["inds"].cpu().numpy())
    assert np.abs(imp["rays_d"].cpu().numpy() - d_imp).max() <= 1e-6


def test_bad_arguments():
    from customnerf_b200 import _lib as L
    lib = L.lib()
    p = torch.eye(4, device="cuda")[None].contiguous()
    o = torch.empty(1, 10, 3, device="cuda")
    # all-pixels mode needs N == H * W
    rc = lib.nb200_get_rays(L.ptr(p), L.f32(1), L.f32(1), L.f32(0), L.f32(0), L.u32(4), L.u32(4), L.u32(1), L.u32(10), L.ptr(None),
                            L.f32(.5), L.f32(.5), L.ptr(o), L.ptr(o.clone()), L.stream())
    assert rc == -3
    with pytest.raises(RuntimeError):
        from customnerf_b200.nerf import get_rays
        get_rays(torch.eye(4)[None], (1, 1, 0, 0), 4, 4)            # CPU tensors: no fallback


def test_fused_step_from_pose_equals_step_from_rays():
    """FusedTrainStep.step(pose=...) generates the rays inside the step's graph: same loss sequence and parameters as
    feeding the same camera's rays (the two ray sets agree to 2e-6, so sample sets may differ by a handful of samples)"""
    from customnerf_b200 import trainer, fused_trainer, synthetic as syn
    dev = torch.device("cuda")
    H, W = 48, 64
    pose, intr = syn.camera_pose(H, W, view=1)
    o, d = cpu_ops.get_rays(pose[None].numpy(), intr, H, W)
    o, d = torch.from_numpy(o[0]).to(dev), torch.from_numpy(d[0]).to(dev)
    tgt = syn.bear_color(o.cpu() + d.cpu() * 1.5).to(dev)
    losses = {}
    for mode in ("rays", "pose"):
        model = trainer.build_scene_model(dev, log2_hashmap_size=15, desired_resolution=512, seed=3)
        fs = fused_trainer.FusedTrainStep(model, H * W, perturb=False, raygen=dict(H=H, W=W, intrinsics=intr))
        ls = []
        for it in range(4):
            if mode == "rays":
                fs.step(o, d, tgt)
            else:
                p_h, t_h = fs.pinned_pose_batch()
                p_h.copy_(pose); t_h.copy_(tgt.cpu())
                fs.step(pose=p_h, target=t_h)
            ls.append(fs.last_stats()[0])
        losses[mode] = ls
        if mode == "pose":
            assert np.abs(fs.rays_d.cpu().numpy() - d.cpu().numpy()).max() <= 1e-6
    np.testing.assert_allclose(losses["pose"], losses["rays"], rtol=2e-3)
    assert losses["pose"][-1] < losses["pose"][0]
