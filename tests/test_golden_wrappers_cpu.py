"""oracle/cpu_ops.py -- the restated raymarching wrappers the GPU parity tests use as their checker -- against golden vectors
minted by running the reference's OWN Python wrappers (raymarching/raymarching.py:20-430, unmodified) on the CPU with the
compiled module they bind replaced by shims over the C oracle (tests/golden/make_golden_wrappers.py).  The arithmetic on both
sides is the C oracle's, so everything here is exact: what is pinned is the wrappers' logic -- output sizes and the alignment
quirk, the step counter, the noise draw, flattening of [B, N, 3] inputs, the autograd Function that ignores grad_depth, the
in-place inference ops."""
import os

import numpy as np
import torch

from oracle import cpu_ops
from golden.make_golden_wrappers import scene, field

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_wrappers.npz"))
AABB = np.array([-2, -2, -2, 2, 2, 2], np.float32)


def test_utility_wrappers():
    grid, thr, o, d = scene()
    n0, f0 = cpu_ops.near_far_from_aabb(o.numpy()[None], d.numpy()[None], AABB)                 # default min_near 0.2
    n1, f1 = cpu_ops.near_far_from_aabb(o.numpy(), d.numpy(), AABB, 0.01)
    assert np.array_equal(np.stack([n0, f0]), G["nf_default"]) and np.array_equal(np.stack([n1, f1]), G["nf_001"])
    assert np.array_equal(cpu_ops.sph_from_ray(o.numpy(), d.numpy(), 2.5), G["sph"])
    ind = cpu_ops.morton3D(G["morton_coords"].reshape(-1, 3))
    assert np.array_equal(ind.reshape(G["morton_indices"].shape), G["morton_indices"])
    assert np.array_equal(cpu_ops.morton3D_invert(ind).reshape(G["morton_back"].shape), G["morton_back"])
    assert np.array_equal(G["morton_back"], G["morton_coords"])
    assert abs(float(G["packbits_thresh"]) - thr) < 1e-12
    bits = cpu_ops.packbits(grid.numpy(), thr)
    assert int(np.unpackbits(bits).sum()) == int(G["packbits_sum"]) and np.array_equal(bits[:4096], G["packbits_head"])


def _march(tag, **kw):
    grid, thr, o, d = scene()
    bits = cpu_ops.packbits(grid.numpy(), thr)
    n0, f0 = cpu_ops.near_far_from_aabb(o.numpy(), d.numpy(), AABB)
    counter = np.zeros(2, np.int32)
    out = cpu_ops.march_rays_train(o.numpy(), d.numpy(), 2, bits, 2, 128, n0, f0, counter, **kw)
    return out, counter, (bits, n0, f0, o, d)


def test_march_rays_train_sizes_alignment_counter_and_noise():
    for tag, kw in (("all", dict(mean_count=-1, noises=None, align=128, force_all_rays=True)),
                    ("perturb", dict(mean_count=-1, noises="rand", align=128, force_all_rays=False)),
                    ("budget", dict(mean_count=1000, noises=None, align=128, force_all_rays=False))):
        if kw["noises"] == "rand":
            torch.manual_seed(3)
            kw["noises"] = torch.rand(192).numpy()            # the wrapper's draw (raymarching.py:214-215), same generator state
        (xyzs, dirs, deltas, rays), counter, _ = _march(tag, **kw)
        assert xyzs.shape == G["mt_%s_xyzs" % tag].shape, tag   # "all": count rounded UP past a multiple (m += align - m % align)
        assert np.array_equal(counter, G["mt_%s_counter" % tag]) and np.array_equal(rays, G["mt_%s_rays" % tag]), tag
        for got, key in ((xyzs, "xyzs"), (dirs, "dirs"), (deltas, "deltas")):
            assert np.array_equal(got, G["mt_%s_%s" % (tag, key)]), (tag, key)
    assert G["mt_budget_xyzs"].shape[0] == 1024                 # mean_count 1000 -> 1024 rows, rays past the budget dropped
    assert int(G["mt_all_counter"][0]) % 128 != 0 and G["mt_all_xyzs"].shape[0] % 128 == 0


def test_composite_train_autograd_function_ignores_grad_depth():
    from oracle import torch_ref
    (xyzs, dirs, deltas, rays), _, _ = _march("all", mean_count=-1, noises=None, align=128, force_all_rays=True)
    sig, rgb = field(torch.from_numpy(xyzs), torch.from_numpy(dirs))
    sig, rgb = sig.clone().requires_grad_(), rgb.clone().requires_grad_()
    ws, depth, image = torch_ref.composite_rays_train(sig, rgb, torch.from_numpy(deltas), torch.from_numpy(rays), 1e-4)
    for got, key in ((ws, "ws"), (depth, "depth"), (image, "image")):
        assert np.array_equal(got.detach().numpy(), G["ct_" + key]), key
    loss = (ws * torch.from_numpy(G["ct_gws"])).sum() + (depth * torch.from_numpy(G["ct_gdepth"])).sum() + \
        (image * torch.from_numpy(G["ct_gimg"])).sum()
    loss.backward()
    assert np.array_equal(sig.grad.numpy(), G["ct_grad_sigmas"]) and np.array_equal(rgb.grad.numpy(), G["ct_grad_rgbs"])


def test_inference_rounds_march_rays_composite_rays():
    _, _, (bits, n0, f0, o, d) = _march("all", mean_count=-1, noises=None, align=128, force_all_rays=True)
    N = o.shape[0]
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    alive, rays_t = np.arange(N, dtype=np.int32), n0.copy()
    for rnd in range(3):
        n_alive = alive.shape[0]
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = cpu_ops.march_rays(n_alive, n_step, alive, rays_t, o.numpy(), d.numpy(), 2, bits, 2, 128, n0, f0, 128, None,
                                                0, 1024)
        assert list(xyzs.shape) == list(G["inf%d_shape" % rnd])          # n_alive * n_step rounded up PAST a multiple of 128
        s_, c_ = field(torch.from_numpy(xyzs), torch.from_numpy(dirs))
        cpu_ops.composite_rays(n_alive, n_step, alive, rays_t, s_.numpy(), c_.numpy(), deltas, ws, depth, image, 1e-4)
        assert np.array_equal(alive, G["inf%d_alive" % rnd]) and np.array_equal(rays_t, G["inf%d_rays_t" % rnd])
        for got, key in ((image, "image"), (ws, "ws"), (depth, "depth")):
            assert np.array_equal(got, G["inf%d_%s" % (rnd, key)]), (rnd, key)
        alive = np.ascontiguousarray(alive[alive >= 0])
