"""Known-answer and self-consistency tests of the CPU oracle (oracle/).  No GPU.

The reference ships no tests or golden vectors, so the oracle is pinned three ways:
 (1) hand-computed known answers derived from the reference source (this file),
 (2) golden vectors minted from the reference's own CUDA extensions on the GPU box (tests/golden/, checked in
     test_golden_cpu.py),
 (3) live comparison against those extensions (tests/test_gpu_ref_ext.py, -m gpu).
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import cpu_ops, torch_ref


# ------------------------------------------------------------------------------------------ morton / packbits
def test_morton_known_answers():
    c = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [2, 0, 0], [3, 5, 7], [127, 127, 127],
                  [127, 0, 0], [0, 127, 0]], np.int32)
    # x -> bits 0,3,6..., y -> bits 1,4,7..., z -> bits 2,5,8... (raymarching.cu:65-71)
    want = []
    for x, y, z in c:
        m = 0
        for b in range(10):
            m |= ((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2)
        want.append(m)
    got = cpu_ops.morton3D(c)
    assert got.tolist() == want
    assert got[7] == 128 ** 3 - 1
    assert (cpu_ops.morton3D_invert(got) == c).all()


def test_morton_bijection_full_grid():
    ar = np.arange(128, dtype=np.int32)
    c = np.stack(np.meshgrid(ar, ar, ar, indexing="ij"), -1).reshape(-1, 3)
    m = cpu_ops.morton3D(c)
    assert np.array_equal(np.sort(m), np.arange(128 ** 3, dtype=np.int32))
    assert np.array_equal(cpu_ops.morton3D_invert(m), c)


def test_packbits_known_answers():
    g = np.zeros((1, 16), np.float32)
    g[0, [0, 3, 9, 15]] = 2.0
    g[0, 4] = 1.0            # equal to the threshold: strict '>' leaves the bit clear (raymarching.cu:285)
    out = cpu_ops.packbits(g, 1.0)
    assert out.tolist() == [0b00001001, 0b10000010]


# ------------------------------------------------------------------------------------------ near / far
def test_near_far_known_answers():
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    o = np.array([[-3, 0, 0], [0, 0, 0], [-3, 5, 0], [0.5, 0.25, -4]], np.float32)
    d = np.array([[1, 0, 0], [0, 1, 0], [1, 0, 0], [0, 0, 2]], np.float32)
    nears, fars = cpu_ops.near_far_from_aabb(o, d, aabb, min_near=0.2)
    assert nears[0] == 2.0 and fars[0] == 4.0
    assert nears[1] == np.float32(0.2) and fars[1] == 1.0        # origin inside: near clamps to min_near
    assert nears[2] == fars[2] == np.finfo(np.float32).max       # miss -> FLT_MAX for both (:121-124)
    assert nears[3] == 1.5 and fars[3] == 2.5


# ------------------------------------------------------------------------------------------ marching
def _full_bitfield(C=1, H=128, value=0xFF):
    return np.full(C * H ** 3 // 8, value, np.uint8)


def test_march_fully_occupied_single_ray():
    """every step is occupied: count = number of dt_min steps from near to far, deltas all dt_min"""
    o = np.array([[-2.0, 0.01, 0.02]], np.float32)
    d = np.array([[1.0, 0.0, 0.0]], np.float32)
    nears, fars = cpu_ops.near_far_from_aabb(o, d, np.array([-1, -1, -1, 1, 1, 1], np.float32))
    counter = np.zeros(2, np.int32)
    xyzs, dirs, deltas, rays = cpu_ops.march_rays_train(o, d, 1.0, _full_bitfield(), 1, 128, nears, fars, counter,
                                                        force_all_rays=True, max_steps=1024)
    dt = np.float32(2 * np.float32(1.7320508075688772) / np.float32(1024))
    t, n = nears[0], 0
    while t < fars[0] and n < 1024:
        t = np.float32(t + dt)
        n += 1
    assert rays.tolist() == [[0, 0, n]]
    assert counter.tolist() == [n, 1]
    assert xyzs.shape == (n, 3) and (deltas[:, 0] == dt).all()
    assert np.allclose(xyzs[0], [-1.0, 0.01, 0.02]) and (dirs == d[0]).all()
    assert_close(deltas[:, 1], np.full(n, dt), 1e-5, 0, "depth deltas")


def test_march_empty_grid_and_alignment_rule():
    o = np.array([[-2.0, 0.0, 0.0], [0.0, -2.0, 0.3]], np.float32)
    d = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], np.float32)
    nears, fars = cpu_ops.near_far_from_aabb(o, d, np.array([-1, -1, -1, 1, 1, 1], np.float32))
    counter = np.zeros(2, np.int32)
    xyzs, dirs, deltas, rays = cpu_ops.march_rays_train(o, d, 1.0, _full_bitfield(value=0), 1, 128, nears, fars,
                                                        counter, align=128, force_all_rays=True)
    assert counter.tolist() == [0, 2] and rays[:, 2].tolist() == [0, 0]
    # m = 0 -> m += 128 - 0 % 128 = 128 rows of zeros (raymarching.py:226-230, Appendix B12)
    assert xyzs.shape == (128, 3) and not xyzs.any()


def test_march_counts_match_train_and_are_scan_ordered(scene):
    s = scene
    sel = slice(0, 2000)
    counts = cpu_ops.march_rays_count(s["rays_o"][sel], s["rays_d"][sel], 2.0, s["bitfield"], 2, 128, s["nears"][sel],
                                      s["fars"][sel])
    counter = np.array([5, 0], np.int32)        # non-zero point counter: offsets continue from it (:405)
    xyzs, dirs, deltas, rays = cpu_ops.march_rays_train(s["rays_o"][sel], s["rays_d"][sel], 2.0, s["bitfield"], 2, 128,
                                                        s["nears"][sel], s["fars"][sel], counter, align=128,
                                                        force_all_rays=True)
    assert np.array_equal(rays[:, 2], counts)
    assert np.array_equal(rays[:, 0], np.arange(2000))
    assert np.array_equal(rays[:, 1], 5 + np.concatenate([[0], np.cumsum(counts)[:-1]]))
    assert counter.tolist() == [5 + counts.sum(), 2000]
    assert counts.sum() > 1000 and counts.max() <= 1024
    # every emitted point lies in an occupied cell of the level-0/1 grid and inside the box
    assert np.abs(xyzs).max() <= 2.0


def test_march_inference_consumes_train_samples(scene):
    """marching n_step at a time from rays_t reproduces the training samples of the same ray"""
    s = scene
    idx = np.where(cpu_ops.march_rays_count(s["rays_o"], s["rays_d"], 2.0, s["bitfield"], 2, 128, s["nears"],
                                            s["fars"]) > 20)[0][:4]
    o, d, nr, fr = s["rays_o"][idx], s["rays_d"][idx], s["nears"][idx], s["fars"][idx]
    xyzs_t, _, deltas_t, rays = cpu_ops.march_rays_train(o, d, 2.0, s["bitfield"], 2, 128, nr, fr, force_all_rays=True)
    n_alive, n_step = len(idx), 8
    rays_alive = np.arange(n_alive, dtype=np.int32)
    xyzs, dirs, deltas = cpu_ops.march_rays(n_alive, n_step, rays_alive, nr.copy(), o, d, 2.0, s["bitfield"], 2, 128,
                                            nr, fr, align=128)
    assert xyzs.shape[0] == n_alive * n_step + (128 - (n_alive * n_step) % 128)
    for k in range(n_alive):
        off = rays[k, 1]
        assert np.array_equal(xyzs[k * n_step:(k + 1) * n_step], xyzs_t[off:off + n_step])
        assert np.array_equal(deltas[k * n_step:(k + 1) * n_step, 0], deltas_t[off:off + n_step, 0])


# ------------------------------------------------------------------------------------------ compositing
def _ref_composite_np(sig, rgb, dl, T_thresh):
    T, ws, d, t, img = 1.0, 0.0, 0.0, 0.0, np.zeros(3)
    for i in range(len(sig)):
        a = 1 - np.exp(-float(sig[i]) * float(dl[i, 0]))
        w = a * T
        img += w * rgb[i].astype(np.float64)
        t += float(dl[i, 1])
        d += w * t
        ws += w
        T *= 1 - a
        if T < T_thresh:
            break
    return ws, d, img


def test_composite_forward_against_float64_loop():
    rng = np.random.RandomState(0)
    counts = [0, 1, 5, 33, 64, 100, 3]
    M = sum(counts)
    sig = rng.uniform(0, 40, M).astype(np.float32)
    rgb = rng.uniform(0, 1, (M, 3)).astype(np.float32)
    dl = np.stack([np.full(M, 0.0034, np.float32), rng.uniform(0.003, 0.05, M).astype(np.float32)], -1)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    # rays rows deliberately permuted: outputs are scattered by ray id (raymarching.cu:572-576)
    order = np.array([3, 0, 6, 1, 5, 2, 4])
    rays = np.stack([order, offs[order], np.array(counts)[order]], -1).astype(np.int32)
    ws, depth, img = cpu_ops.composite_rays_train_forward(sig, rgb, dl, rays, 1e-4)
    for rid in range(len(counts)):
        o, c = offs[rid], counts[rid]
        w0, d0, i0 = _ref_composite_np(sig[o:o + c], rgb[o:o + c], dl[o:o + c], 1e-4)
        assert_close(ws[rid], w0, 1e-5, 1e-6, "ws")
        assert_close(depth[rid], d0, 1e-5, 1e-6, "depth")
        assert_close(img[rid], i0, 1e-5, 1e-6, "image")


def test_composite_early_out_includes_breaking_sample():
    sig = np.array([1e4, 5.0, 5.0], np.float32)       # first sample already drives T below the threshold
    rgb = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    dl = np.full((3, 2), 0.01, np.float32)
    rays = np.array([[0, 0, 3]], np.int32)
    ws, depth, img = cpu_ops.composite_rays_train_forward(sig, rgb, dl, rays, 1e-4)
    assert_close(img[0], [1, 0, 0], 1e-6, 1e-6)
    gs, gc = cpu_ops.composite_rays_train_backward(np.ones(1, np.float32), np.ones((1, 3), np.float32), sig, rgb, dl,
                                                   rays, ws, img, 1e-4)
    assert gc[0, 0] > 0 and not gc[1:].any() and not gs[1:].any()     # rows after the break stay zero


def test_composite_backward_against_autograd():
    """grad_sigmas / grad_rgbs of the restatement == autograd of the same scan in float64 (T_thresh = 0)"""
    rng = np.random.RandomState(1)
    counts = [7, 40, 1]
    M = sum(counts)
    sig = rng.uniform(0, 30, M)
    rgb = rng.uniform(0, 1, (M, 3))
    dl = np.stack([np.full(M, 0.0034), rng.uniform(0.003, 0.05, M)], -1)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    rays = np.stack([np.arange(3), offs, counts], -1).astype(np.int32)
    g_ws = rng.normal(size=3)
    g_img = rng.normal(size=(3, 3))
    s_t = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    c_t = torch.tensor(rgb, dtype=torch.float64, requires_grad=True)
    loss = 0
    for r in range(3):
        o, c = offs[r], counts[r]
        a = 1 - torch.exp(-s_t[o:o + c] * torch.tensor(dl[o:o + c, 0]))
        T = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64), 1 - a]), 0)[:-1]
        w = a * T
        loss = loss + (w.sum() * g_ws[r]) + ((w[:, None] * c_t[o:o + c]).sum(0) * torch.tensor(g_img[r])).sum()
    loss.backward()
    ws, depth, img = cpu_ops.composite_rays_train_forward(sig, rgb, dl, rays, 0.0)
    gs, gc = cpu_ops.composite_rays_train_backward(g_ws, g_img, sig, rgb, dl, rays, ws, img, 0.0)
    assert_close(gc, c_t.grad.numpy(), 1e-4, 1e-6, "grad_rgbs")
    assert_close(gs, s_t.grad.numpy(), 2e-4, 1e-5 * np.abs(s_t.grad.numpy()).max(), "grad_sigmas")


def test_composite_inference_matches_training_on_one_ray():
    rng = np.random.RandomState(2)
    n_step = 8
    sig = rng.uniform(0, 20, n_step).astype(np.float32)
    rgb = rng.uniform(0, 1, (n_step, 3)).astype(np.float32)
    dl = np.stack([np.full(n_step, 0.0034, np.float32), np.full(n_step, 0.0034, np.float32)], -1)
    rays = np.array([[0, 0, n_step]], np.int32)
    ws_t, depth_t, img_t = cpu_ops.composite_rays_train_forward(sig, rgb, dl, rays, 0.0)
    rays_alive = np.array([0], np.int32)
    rays_t = np.array([0.0], np.float32)
    ws, depth, img = np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros((1, 3), np.float32)
    cpu_ops.composite_rays(1, n_step, rays_alive, rays_t, sig, rgb, dl, ws, depth, img, 0.0)
    assert rays_alive[0] == 0 and abs(rays_t[0] - n_step * 0.0034) < 1e-6
    assert_close(ws, ws_t, 1e-5, 1e-6)
    assert_close(img, img_t, 1e-5, 1e-6)
    assert_close(depth, depth_t, 1e-5, 1e-6)
    # a zero delta terminates the ray (:1042)
    dl2 = dl.copy(); dl2[3:] = 0
    rays_alive = np.array([0], np.int32)
    cpu_ops.composite_rays(1, n_step, rays_alive, np.array([0.0], np.float32), sig, rgb, dl2, np.zeros(1, np.float32),
                           np.zeros(1, np.float32), np.zeros((1, 3), np.float32), 0.0)
    assert rays_alive[0] == -1


# ------------------------------------------------------------------------------------------ grid encoder
def test_grid_offsets_table_values():
    offs, s = cpu_ops.grid_offsets(desired_resolution=2048)
    assert offs[-1] == 6119864 and abs(s - 1.381913) < 1e-6
    assert offs[1] == 4920 and offs[2] - offs[1] == 13824             # 17^3 -> 4920 (x8), 24^3
    offs, s = cpu_ops.grid_offsets(log2_hashmap_size=21, desired_resolution=8192)
    assert offs[-1] == 23967296
    offs, s = cpu_ops.grid_offsets(log2_hashmap_size=22, desired_resolution=2048)
    assert offs[-1] == 39625280


def test_grid_encode_trilinear_known_answer():
    """single dense level, table value = linear function of the vertex => interpolation reproduces it exactly"""
    offs, s = cpu_ops.grid_offsets(num_levels=1, per_level_scale=1, base_resolution=4, log2_hashmap_size=19)
    assert offs.tolist() == [0, 128]                 # (4+1)^3 = 125 -> 128
    emb = np.zeros((128, 2), np.float32)
    for z in range(5):
        for y in range(5):
            for x in range(5):
                emb[x + 5 * y + 25 * z] = [x + 10 * y + 100 * z, 1.0]
    # scale = 2^0 * 4 - 1 = 3; pos = x*3 + 0.5
    x = np.array([[0.25, 0.5, 0.75], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [1.0001, 0.5, 0.5], [-0.1, 0.2, 0.3]], np.float32)
    out, _ = cpu_ops.grid_encode_forward(x, emb, offs, 1, 4)
    pos = x[:3] * 3 + 0.5
    want = pos[:, 0] + 10 * pos[:, 1] + 100 * pos[:, 2]
    assert_close(out[:3, 0], want, 1e-6, 1e-5)
    assert_close(out[:3, 1], np.ones(3), 1e-6, 1e-6)
    assert not out[3:].any()                         # out of [0,1] -> zeros (gridencoder.cu:110-135)


def test_hash_index_known_answer():
    """level with (res+1)^3 > 2^T uses x ^ y*2654435761 ^ z*805459861 mod size (gridencoder.cu:50-84)"""
    offs, s = cpu_ops.grid_offsets(num_levels=1, per_level_scale=1, base_resolution=64, log2_hashmap_size=10)
    assert offs.tolist() == [0, 1024]
    emb = np.arange(2048, dtype=np.float32).reshape(1024, 2)
    x = np.array([[10.0 / 63, 20.0 / 63, 30.0 / 63]], np.float32)      # pos = x*63+0.5 -> cell (10,20,30), frac .5
    out, _ = cpu_ops.grid_encode_forward(x, emb, offs, 1, 64)
    acc = 0.0
    pos = (x[0].astype(np.float64) * 63 + 0.5)
    pg = np.floor(pos).astype(np.int64); fr = pos - pg
    for idx in range(8):
        w, c = 1.0, []
        for dd in range(3):
            b = (idx >> dd) & 1
            w *= fr[dd] if b else 1 - fr[dd]
            c.append(int(pg[dd]) + b)
        h = (c[0] ^ (c[1] * 2654435761 & 0xFFFFFFFF) ^ (c[2] * 805459861 & 0xFFFFFFFF)) % 1024
        acc += w * emb[h, 0]
    assert_close(out[0, 0], acc, 1e-5, 1e-3)


@pytest.mark.parametrize("gridtype,log2T,res", [(0, 19, 2048), (1, 21, 8192), (0, 14, 512)])
def test_grid_encode_c_vs_torch_restatement(gridtype, log2T, res):
    offs, s = cpu_ops.grid_offsets(log2_hashmap_size=log2T, desired_resolution=res)
    rng = np.random.RandomState(0)
    emb = rng.uniform(-1, 1, (offs[-1], 2)).astype(np.float32)
    x = rng.uniform(0, 1, (512, 3)).astype(np.float32)
    x[:4] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1.5, 0.5, 0.5]]
    S = np.float32(np.log2(s))
    sc = np.array([torch_ref.level_scale_f32(l, S, 16) for l in range(16)], np.float32)
    out, _ = cpu_ops.grid_encode_forward(x, emb, offs, s, 16, gridtype=gridtype, scales=sc)
    e = torch.from_numpy(emb).requires_grad_()
    out_t = torch_ref.grid_encode(torch.from_numpy(x), e, offs.tolist(), s, 16, gridtype=gridtype, scales=sc)
    assert_close(out, out_t.detach().numpy(), 1e-5, 1e-6, "forward")
    g = torch.from_numpy(rng.normal(size=out.shape).astype(np.float32))
    out_t.backward(g)
    ge, _ = cpu_ops.grid_encode_backward(g.numpy(), x, emb.shape, offs, s, 16, gridtype=gridtype, scales=sc)
    assert_close(ge, e.grad.numpy(), 1e-4, 1e-5, "backward")


def test_grid_encode_dy_dx_finite_difference():
    offs, s = cpu_ops.grid_offsets(num_levels=4, per_level_scale=1.5, base_resolution=8, log2_hashmap_size=12)
    rng = np.random.RandomState(3)
    emb = rng.uniform(-1, 1, (offs[-1], 2)).astype(np.float32)
    x = rng.uniform(0.1, 0.9, (16, 3)).astype(np.float32)
    out, dy_dx = cpu_ops.grid_encode_forward(x, emb, offs, 1.5, 8, calc_grad_inputs=True)
    g = rng.normal(size=out.shape).astype(np.float32)
    _, gi = cpu_ops.grid_encode_backward(g, x, emb.shape, offs, 1.5, 8, dy_dx=dy_dx)
    eps = 1e-3
    for d in range(3):
        xp, xm = x.copy(), x.copy()
        xp[:, d] += eps; xm[:, d] -= eps
        fp, _ = cpu_ops.grid_encode_forward(xp, emb, offs, 1.5, 8)
        fm, _ = cpu_ops.grid_encode_forward(xm, emb, offs, 1.5, 8)
        fd = ((fp - fm) / (2 * eps) * g).sum(-1)
        # piecewise-trilinear: FD is exact unless a cell boundary lies inside the stencil
        ok = np.abs(fd - gi[:, d]) < 2e-2 * (1 + np.abs(fd))
        assert ok.mean() > 0.8


def test_grid_encode_half_mode_close_to_float():
    offs, s = cpu_ops.grid_offsets(desired_resolution=2048)
    rng = np.random.RandomState(0)
    emb = rng.uniform(-1, 1, (offs[-1], 2)).astype(np.float32)
    x = rng.uniform(0, 1, (256, 3)).astype(np.float32)
    out, _ = cpu_ops.grid_encode_forward(x, emb.astype(np.float16).astype(np.float32), offs, s, 16)
    outh, _ = cpu_ops.grid_encode_forward(x, emb, offs, s, 16, half=True)
    assert_close(outh, out, 4e-3, 2e-3, "fp16 double-rounding stays within a few half ulps")


def test_grad_total_variation_constant_table_is_zero():
    offs, s = cpu_ops.grid_offsets(num_levels=2, per_level_scale=2, base_resolution=4, log2_hashmap_size=12)
    emb = np.ones((offs[-1], 2), np.float32)
    x = np.random.RandomState(0).uniform(0, 1, (64, 3)).astype(np.float32)
    g = cpu_ops.grad_total_variation(x, emb, np.zeros_like(emb), offs, 1e-3, 2, 4)
    assert not g.any()


# ------------------------------------------------------------------------------------------ field / renderer
def test_mlp_param_layout_and_padding_rule():
    assert torch_ref.mlp_layer_shapes(32, 64, 64, 2) == [(64, 32), (64, 64), (64, 64)]
    assert torch_ref.mlp_layer_shapes(64, 1, 64, 1) == [(64, 64), (16, 64)]
    assert torch_ref.mlp_layer_shapes(91, 4, 64, 1) == [(64, 96), (16, 64)]
    p = torch_ref.mlp_init(91, 4)
    assert p.numel() == 64 * 96 + 16 * 64
    x = torch.randn(5, 91)
    W1 = p[:64 * 96].view(64, 96)
    W2 = p[64 * 96:].view(16, 64)
    want = torch.sigmoid(torch.relu(x @ W1[:, :91].t() + W1[:, 91:].sum(1)) @ W2[:4].t())
    got = torch_ref.mlp_forward(x, p, 91, 4, output_activation="Sigmoid")
    assert_close(got.numpy(), want.numpy(), 1e-5, 1e-6)


def test_dense_renderer_runs_and_backpropagates():
    torch.manual_seed(0)
    opt = torch_ref.default_opt(train_conf=0.01, soft_mask=True)
    net = torch_ref.NeRFNetwork(opt, encoder_kwargs=dict(log2_hashmap_size=14, desired_resolution=256, gridtype="hash"))
    from customnerf_b200 import synthetic as syn
    o, d = syn.random_rays(64)
    res = net.render(o[None], d[None], num_steps=16, upsample_steps=16, perturb=True)
    assert res["image"].shape == (1, 64, 3) and res["fg"]["image"].shape == (1, 64, 3)
    assert res["render_mask"].shape == (1, 64, 1)
    (res["image"].mean() + res["render_mask"].mean()).backward()
    assert net.pos_en.embeddings.grad.abs().sum() > 0
    assert net.rgb_network.params.grad.abs().sum() > 0
    # fg + bg sigma partition: with a hard mask the two composites use disjoint densities
    opt.soft_mask = False
    res = net.render(o[None], d[None], num_steps=16, upsample_steps=16, perturb=False)
    assert (res["fg"]["weights_sum"] <= 1 + 1e-5).all()


def test_occupancy_renderer_and_grid_update_cpu(scene):
    torch.manual_seed(0)
    opt = torch_ref.default_opt(cuda_ray=True)
    net = torch_ref.NeRFNetwork(opt, encoder_kwargs=dict(log2_hashmap_size=14, desired_resolution=256, gridtype="hash"))
    net.density_bitfield = torch.from_numpy(scene["bitfield"])
    o = torch.from_numpy(scene["rays_o"][7000:7064])
    d = torch.from_numpy(scene["rays_d"][7000:7064])
    net.train()
    res = net.render(o[None], d[None], perturb=False, force_all_rays=True)
    assert res["image"].shape == (1, 64, 3)
    res["image"].sum().backward()
    assert net.pos_en.embeddings.grad.abs().sum() > 0
    net.eval()
    res_e = net.render(o[None], d[None], perturb=False)
    # train and eval paths composite the same samples (T_thresh early-out aside)
    assert_close(res_e["image"].numpy(), res["image"].detach().numpy(), 1e-3, 2e-4, "train vs eval image")


def test_lgie_oracle_matches_torch_autograd_of_the_same_composition():
    """oracle.cpu_ops.composite_lgie_* (the checker of the gated composite kernels) against torch autograd through a
    plain-torch restatement of the composite (weights = alpha * exclusive cumprod(1 - alpha), no early termination at
    T_thresh = 0) on ragged rays: fp32 rel 1e-4."""
    import torch
    from oracle import cpu_ops
    rng = np.random.RandomState(5)
    counts = np.array([7, 0, 33, 1, 20], np.int32)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays = np.stack([np.arange(5, dtype=np.int32), offs, counts], 1)
    M = int(counts.sum())
    sig = rng.uniform(0, 30, M).astype(np.float32)
    rgb = rng.uniform(0, 1, (M, 3)).astype(np.float32)
    msk = np.clip(rng.normal(0.5, 0.02, M), 0, 1).astype(np.float32)
    dl = np.stack([rng.uniform(0.002, 0.01, M), rng.uniform(0.002, 0.01, M)], 1).astype(np.float32)
    g_ws, g_img, g_m = rng.randn(5).astype(np.float32), rng.randn(5, 3).astype(np.float32), rng.randn(5).astype(np.float32)
    for variant in (0, 1, 2):
        for soft, dbg, dmf in ((True, True, False), (False, False, False), (True, False, True)):
            ws, depth, img, rm = cpu_ops.composite_lgie_forward(variant, sig, rgb, msk, dl, rays, 0.0, soft, 0.5)
            ds, dc, dm = cpu_ops.composite_lgie_backward(variant, g_ws, g_img, g_m, sig, rgb, msk, dl, rays, 0.0, soft, 0.5, dbg, dmf)
            ts, tc, tm = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (sig, rgb, msk))
            e = torch.sigmoid((tm - 0.5) * 100) if soft else (tm > 0.5).double()
            gate = torch.ones_like(tm) if variant == 0 else e if variant == 1 else 1 - e
            s_in, c_in = ts, tc
            if variant == 0 and dbg:
                ep = tm >= 0.5
                s_in = torch.where(ep, ts, ts.detach())
                c_in = torch.where(ep[:, None], tc, tc.detach())
            sv = s_in * gate
            loss = 0.0
            for n, (o, c) in enumerate(zip(offs, counts)):
                if c == 0:
                    continue
                sl = slice(int(o), int(o + c))
                alpha = 1 - torch.exp(-sv[sl] * torch.tensor(dl[sl, 0], dtype=torch.float64))
                T = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64), 1 - alpha[:-1]]), 0)
                w = alpha * T
                wm = w.detach() if dmf else w
                np.testing.assert_allclose(float(w.sum().detach()), ws[n], rtol=1e-4, atol=1e-6)
                np.testing.assert_allclose((w[:, None] * tc[sl]).sum(0).detach().numpy(), img[n], rtol=1e-4, atol=1e-6)
                np.testing.assert_allclose(float((w * tm[sl]).sum().detach()), rm[n], rtol=1e-4, atol=1e-6)
                loss = loss + g_ws[n] * w.sum() + (torch.tensor(g_img[n], dtype=torch.float64) * (w[:, None] * c_in[sl]).sum(0)).sum() \
                    + g_m[n] * (wm * tm[sl]).sum()
            loss.backward()
            for got, want, what in ((ds, ts.grad, "sigma"), (dc, tc.grad, "rgb"), (dm, tm.grad, "mask")):
                want = want.numpy()
                assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), (variant, soft, dbg, dmf, what,
                                                                                           np.abs(got - want).max())
