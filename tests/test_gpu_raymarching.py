"""-m gpu: CUDA raymarching ops (through the C ABI) vs the CPU oracle.  Integer outputs bit-exact."""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import cpu_ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rm():
    from customnerf_b200 import raymarching
    return raymarching


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_morton_all_cells_bit_exact(rm):
    ar = torch.arange(128, dtype=torch.int32)
    c = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), -1).reshape(-1, 3)
    got = rm.morton3D(c.cuda())
    want = cpu_ops.morton3D(c.numpy())
    assert np.array_equal(got.cpu().numpy(), want)
    back = rm.morton3D_invert(got)
    assert np.array_equal(back.cpu().numpy(), c.numpy())
    assert rm.morton3D(torch.zeros(0, 3, dtype=torch.int32).cuda()).shape == (0,)      # empty input


def test_packbits_bit_exact(rm, scene):
    grid = scene["grid"]
    for thr in (scene["thresh"], 0.0, 10.0, float(grid[0, 12345])):
        got = rm.packbits(_cu(grid), thr)
        assert np.array_equal(got.cpu().numpy(), cpu_ops.packbits(grid, thr))
    # pre-allocated output is filled in place
    out = torch.zeros(2 * 128 ** 3 // 8, dtype=torch.uint8).cuda()
    ret = rm.packbits(_cu(grid), scene["thresh"], out)
    assert ret.data_ptr() == out.data_ptr() and np.array_equal(out.cpu().numpy(), scene["bitfield"])


def test_near_far_bit_exact(rm, scene):
    rng = np.random.RandomState(0)
    o = np.concatenate([scene["rays_o"], rng.uniform(-3, 3, (4096, 3)).astype(np.float32)])
    d = np.concatenate([scene["rays_d"], rng.normal(size=(4096, 3)).astype(np.float32)])
    d[-1] = [0, 0, 1]           # zero components -> inf reciprocals
    d[-2] = [1, 0, 0]
    for mn in (0.2, 0.01):
        n, f = rm.near_far_from_aabb(_cu(o), _cu(d), _cu(scene["aabb"]), mn)
        n0, f0 = cpu_ops.near_far_from_aabb(o, d, scene["aabb"], mn)
        assert np.array_equal(n.cpu().numpy(), n0, equal_nan=True)
        assert np.array_equal(f.cpu().numpy(), f0, equal_nan=True)


def test_sph_from_ray(rm, scene):
    o, d = scene["rays_o"][:1000] * 0.3, scene["rays_d"][:1000]
    got = rm.sph_from_ray(_cu(o), _cu(d), 2.0).cpu().numpy()
    assert_close(got, cpu_ops.sph_from_ray(o, d, 2.0), 1e-4, 1e-5)


@pytest.mark.parametrize("perturb", [False, True])
def test_march_rays_train_bit_exact(rm, scene, perturb):
    s = scene
    N = s["rays_o"].shape[0]
    noises = np.random.RandomState(5).uniform(0, 1, N).astype(np.float32) if perturb else None
    counter = torch.zeros(2, dtype=torch.int32).cuda()
    xyzs, dirs, deltas, rays = rm.march_rays_train(_cu(s["rays_o"]), _cu(s["rays_d"]), 2.0, _cu(s["bitfield"]), 2, 128,
                                                   _cu(s["nears"]), _cu(s["fars"]), counter, -1, perturb, 128, True, 0,
                                                   1024, noises=None if noises is None else _cu(noises))
    c0 = np.zeros(2, np.int32)
    x0, d0, l0, r0 = cpu_ops.march_rays_train(s["rays_o"], s["rays_d"], 2.0, s["bitfield"], 2, 128, s["nears"],
                                              s["fars"], c0, -1, noises, 128, True, 0, 1024)
    assert np.array_equal(counter.cpu().numpy(), c0)
    assert np.array_equal(rays.cpu().numpy(), r0)                     # ids, scan offsets and counts
    assert xyzs.shape == x0.shape and deltas.shape == l0.shape        # m + (128 - m % 128) rule
    assert np.array_equal(xyzs.cpu().numpy(), x0)
    assert np.array_equal(dirs.cpu().numpy(), d0)
    assert np.array_equal(deltas.cpu().numpy(), l0)
    assert c0[0] > 200000


def test_march_rays_train_mean_count_budget_drops_rays(rm, scene):
    s = scene
    sel = slice(6000, 9000)
    counter = torch.zeros(2, dtype=torch.int32).cuda()
    xyzs, dirs, deltas, rays = rm.march_rays_train(_cu(s["rays_o"][sel]), _cu(s["rays_d"][sel]), 2.0,
                                                   _cu(s["bitfield"]), 2, 128, _cu(s["nears"][sel]), _cu(s["fars"][sel]),
                                                   counter, 20000, False, 128, False)
    c0 = np.zeros(2, np.int32)
    x0, d0, l0, r0 = cpu_ops.march_rays_train(s["rays_o"][sel], s["rays_d"][sel], 2.0, s["bitfield"], 2, 128,
                                              s["nears"][sel], s["fars"][sel], c0, 20000, None, 128, False)
    assert xyzs.shape == x0.shape == (20000 + 128 - 20000 % 128, 3)
    assert c0[0] > xyzs.shape[0]                                      # the budget really is exceeded
    assert np.array_equal(rays.cpu().numpy(), r0)
    assert np.array_equal(xyzs.cpu().numpy(), x0) and np.array_equal(deltas.cpu().numpy(), l0)


def test_march_edge_cases(rm, scene):
    s = scene
    # no rays at all
    e = torch.zeros(0, 3).cuda()
    x, d, l, r = rm.march_rays_train(e, e, 2.0, _cu(s["bitfield"]), 2, 128, torch.zeros(0).cuda(), torch.zeros(0).cuda(),
                                     None, -1, False, 128, True)
    assert x.shape == (128, 3) and r.shape == (0, 3)
    # rays that miss the box (near = far = FLT_MAX), a fully occupied grid and max_steps truncation
    o = np.array([[-2.5, 0.0, 0.0], [5, 5, 5]], np.float32)
    dd = np.array([[1.0, 0.0, 0.0], [1, 0, 0]], np.float32)
    aabb = np.array([-2, -2, -2, 2, 2, 2], np.float32)
    n0, f0 = cpu_ops.near_far_from_aabb(o, dd, aabb)
    full = np.full(2 * 128 ** 3 // 8, 255, np.uint8)
    for ms in (1024, 64):
        cnt = torch.zeros(2, dtype=torch.int32).cuda()
        x, d, l, r = rm.march_rays_train(_cu(o), _cu(dd), 2.0, _cu(full), 2, 128, _cu(n0), _cu(f0), cnt, -1, False, 128,
                                         True, 0, ms)
        c0 = np.zeros(2, np.int32)
        x0, d0_, l0, r0 = cpu_ops.march_rays_train(o, dd, 2.0, full, 2, 128, n0, f0, c0, -1, None, 128, True, 0, ms)
        assert np.array_equal(r.cpu().numpy(), r0) and r0[0, 2] == ms and r0[1, 2] == 0
        assert np.array_equal(x.cpu().numpy(), x0) and np.array_equal(l.cpu().numpy(), l0)


def test_march_dt_gamma_cone_stepping(rm, scene):
    s = scene
    sel = slice(7000, 8000)
    cnt = torch.zeros(2, dtype=torch.int32).cuda()
    out = rm.march_rays_train(_cu(s["rays_o"][sel]), _cu(s["rays_d"][sel]), 2.0, _cu(s["bitfield"]), 2, 128,
                              _cu(s["nears"][sel]), _cu(s["fars"][sel]), cnt, -1, False, 128, True, 1.0 / 128, 1024)
    c0 = np.zeros(2, np.int32)
    ref = cpu_ops.march_rays_train(s["rays_o"][sel], s["rays_d"][sel], 2.0, s["bitfield"], 2, 128, s["nears"][sel],
                                   s["fars"][sel], c0, -1, None, 128, True, 1.0 / 128, 1024)
    for a, b in zip(out, ref):
        assert np.array_equal(a.cpu().numpy(), b)


def _random_segments(rng, counts):
    M = int(sum(counts))
    sig = rng.uniform(0, 60, M).astype(np.float32)
    rgb = rng.uniform(0, 1, (M, 3)).astype(np.float32)
    dl = np.stack([np.full(M, 0.0033829, np.float32), rng.uniform(0.003, 0.05, M).astype(np.float32)], -1)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    order = rng.permutation(len(counts))
    rays = np.stack([order, offs[order], np.asarray(counts)[order]], -1).astype(np.int32)
    return sig, rgb, dl, rays


@pytest.mark.parametrize("T_thresh", [1e-4, 0.0])
def test_composite_train_forward_backward(rm, T_thresh):
    rng = np.random.RandomState(0)
    counts = np.concatenate([[0, 1, 31, 32, 33, 64, 65, 300, 1024], rng.randint(0, 150, 4000)])
    sig, rgb, dl, rays = _random_segments(rng, counts)
    sig_t = _cu(sig).requires_grad_()
    rgb_t = _cu(rgb).requires_grad_()
    ws, depth, img = rm.composite_rays_train(sig_t, rgb_t, _cu(dl), _cu(rays), T_thresh)
    ws0, depth0, img0 = cpu_ops.composite_rays_train_forward(sig, rgb, dl, rays, T_thresh)
    # tolerance: rel 1e-4, abs 1e-6 -- the prefix products are a shuffle tree here,
    # a sequential loop in the reference
    assert_close(ws.detach().cpu().numpy(), ws0, 1e-4, 2e-6, "weights_sum")
    assert_close(depth.detach().cpu().numpy(), depth0, 1e-4, 2e-6, "depth")
    assert_close(img.detach().cpu().numpy(), img0, 1e-4, 2e-6, "image")
    g_ws = rng.normal(size=ws0.shape).astype(np.float32)
    g_img = rng.normal(size=img0.shape).astype(np.float32)
    torch.autograd.backward([ws, img], [_cu(g_ws), _cu(g_img)])
    gs0, gc0 = cpu_ops.composite_rays_train_backward(g_ws, g_img, sig, rgb, dl, rays, ws0, img0, T_thresh)
    assert_close(rgb_t.grad.cpu().numpy(), gc0, 1e-4, 2e-6, "grad_rgbs")
    assert_close(sig_t.grad.cpu().numpy(), gs0, 2e-4, 1e-5 * np.abs(gs0).max(), "grad_sigmas")


def test_composite_overflowing_ray_is_zeroed(rm):
    rng = np.random.RandomState(1)
    sig, rgb, dl, rays = _random_segments(rng, [10, 20, 30])
    rays[rays[:, 0] == 2, 1] += 100                                  # segment beyond M -> outputs zero (:521-528)
    ws, depth, img = rm.composite_rays_train(_cu(sig), _cu(rgb), _cu(dl), _cu(rays), 1e-4)
    ws0, depth0, img0 = cpu_ops.composite_rays_train_forward(sig, rgb, dl, rays, 1e-4)
    assert ws0[2] == 0 and ws.cpu().numpy()[2] == 0
    assert_close(img.cpu().numpy(), img0, 1e-4, 2e-6)


def test_inference_march_and_composite_loop(rm, scene):
    """the run_cuda eval loop (renderer.py:651-688) op by op against the oracle, with a synthetic field"""
    s = scene
    sel = slice(7000, 7600)
    o, d, nr, fr = s["rays_o"][sel], s["rays_d"][sel], s["nears"][sel], s["fars"][sel]
    N = o.shape[0]

    def field(x):   # deterministic stand-in for the network
        sig = 30.0 * (1 + np.sin(7 * x.sum(-1)))
        rgb = 0.5 + 0.5 * np.cos(5 * x)
        return sig.astype(np.float32), rgb.astype(np.float32)

    ws_g = torch.zeros(N).cuda(); dp_g = torch.zeros(N).cuda(); im_g = torch.zeros(N, 3).cuda()
    ws_c = np.zeros(N, np.float32); dp_c = np.zeros(N, np.float32); im_c = np.zeros((N, 3), np.float32)
    alive_g = torch.arange(N, dtype=torch.int32).cuda(); t_g = _cu(nr).clone()
    alive_c = np.arange(N, dtype=np.int32); t_c = nr.copy()
    step = 0
    while step < 1024:
        n_alive = alive_c.shape[0]
        assert alive_g.shape[0] == n_alive
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xg, dg, lg = rm.march_rays(n_alive, n_step, alive_g, t_g, _cu(o), _cu(d), 2.0, _cu(s["bitfield"]), 2, 128,
                                   _cu(nr), _cu(fr), 128, False, 0, 1024)
        xc, dc, lc = cpu_ops.march_rays(n_alive, n_step, alive_c, t_c, o, d, 2.0, s["bitfield"], 2, 128, nr, fr, 128)
        assert np.array_equal(xg.cpu().numpy(), xc) and np.array_equal(lg.cpu().numpy(), lc)
        assert np.array_equal(dg.cpu().numpy(), dc)
        sig, rgb = field(xc)
        rm.composite_rays(n_alive, n_step, alive_g, t_g, _cu(sig), _cu(rgb), lg, ws_g, dp_g, im_g, 1e-4)
        cpu_ops.composite_rays(n_alive, n_step, alive_c, t_c, sig, rgb, lc, ws_c, dp_c, im_c, 1e-4)
        assert np.array_equal(alive_g.cpu().numpy(), alive_c)        # -1 marks
        assert_close(t_g.cpu().numpy(), t_c, 1e-6, 0, "rays_t")
        alive_g = alive_g[alive_g >= 0]
        alive_c = np.ascontiguousarray(alive_c[alive_c >= 0])
        step += n_step
    assert_close(ws_g.cpu().numpy(), ws_c, 1e-4, 1e-6, "weights_sum")
    assert_close(im_g.cpu().numpy(), im_c, 1e-4, 1e-6, "image")
    assert_close(dp_g.cpu().numpy(), dp_c, 1e-4, 1e-6, "depth")
    assert ws_c.max() > 0.5
