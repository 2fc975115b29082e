"""-m gpu: live three-way parity at full size -- reference CUDA extensions (oracle/_ref, unmodified sources built by
oracle/build_ref.py) vs this repo's CUDA path vs the CPU oracle.

This is the pin SURVEY.md section 8(c) asks for: per-ray sample counts equal to the reference's kernel on >= 1 M
rays, encode / composite within the stated tolerances, on the sizes BASELINE.json's configs use.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import cpu_ops, ref_ext

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_ext.available() if torch.cuda.is_available() else True,
                                 reason="oracle/_ref not built (needs /root/reference at build time)")]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _ref_march(rm, o, d, bf, nears, fars, noises, bound=2.0, C=2, H=128, max_steps=1024, dt_gamma=0.0):
    N = o.shape[0]
    M = N * max_steps if N <= 20000 else 64           # for the 1M-ray count test only the counts are needed
    xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
    rays = torch.empty(N, 3, dtype=torch.int32, device="cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    rm.march_rays_train(o, d, bf, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter,
                        noises)
    torch.cuda.synchronize()
    return xyzs, dirs, deltas, rays, counter


def test_march_counts_bit_exact_on_one_million_rays(scene):
    from customnerf_b200 import raymarching as mine, synthetic as syn
    rm = ref_ext.raymarching()
    N = 1 << 20
    o, d = syn.random_rays(N, seed=3, device="cuda")
    bf = cu(scene["bitfield"])
    aabb = cu(scene["aabb"])
    nears, fars = mine.near_far_from_aabb(o, d, aabb)
    n_ref = torch.empty(N, device="cuda"); f_ref = torch.empty(N, device="cuda")
    rm.near_far_from_aabb(o, d, aabb, N, 0.2, n_ref, f_ref)
    assert torch.equal(nears, n_ref) and torch.equal(fars, f_ref)                      # bit-exact
    noises = torch.rand(N, device="cuda")
    _, _, _, rays_ref, counter_ref = _ref_march(rm, o, d, bf, nears, fars, noises)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    _, _, _, rays = mine.march_rays_train(o, d, 2.0, bf, 2, 128, nears, fars, counter, -1, True, 128, True, 0, 1024,
                                          noises=noises)
    rr = rays_ref.cpu().numpy()
    rr = rr[np.argsort(rr[:, 0], kind="stable")]
    mm = rays.cpu().numpy()
    assert np.array_equal(mm[:, 0], np.arange(N))
    assert np.array_equal(mm[:, 2], rr[:, 2]), "per-ray counts differ from the reference kernel"
    assert np.array_equal(counter.cpu().numpy(), counter_ref.cpu().numpy())
    assert int(counter[0]) > 5_000_000


@pytest.mark.parametrize("dt_gamma", [0.0, 1.0 / 256])
def test_march_samples_bit_exact_on_scene_image(scene, dt_gamma):
    from customnerf_b200 import raymarching as mine
    rm = ref_ext.raymarching()
    s = scene
    o, d, bf, nr, fr = cu(s["rays_o"]), cu(s["rays_d"]), cu(s["bitfield"]), cu(s["nears"]), cu(s["fars"])
    N = o.shape[0]
    noises = torch.rand(N, device="cuda")
    xr, dr, lr, rays_ref, counter_ref = _ref_march(rm, o, d, bf, nr, fr, noises, dt_gamma=dt_gamma)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xm, dm, lm, rays = mine.march_rays_train(o, d, 2.0, bf, 2, 128, nr, fr, counter, -1, True, 128, True, dt_gamma, 1024,
                                             noises=noises)
    assert torch.equal(counter, counter_ref)
    rr = rays_ref.cpu().numpy()
    rr = rr[np.argsort(rr[:, 0], kind="stable")]
    mm = rays.cpu().numpy()
    assert np.array_equal(mm[:, 2], rr[:, 2])
    # the reference's segments are a permutation of the scan layout: gather them in ray order
    idx = np.concatenate([np.arange(r[1], r[1] + r[2]) for r in rr if r[2] > 0])
    idx = torch.from_numpy(idx).cuda()
    m = int(counter[0])
    assert torch.equal(xm[:m], xr[idx]) and torch.equal(dm[:m], dr[idx]) and torch.equal(lm[:m], lr[idx])
    # and the oracle agrees with both
    c0 = np.zeros(2, np.int32)
    x0, d0, l0, r0 = cpu_ops.march_rays_train(s["rays_o"], s["rays_d"], 2.0, s["bitfield"], 2, 128, s["nears"], s["fars"],
                                              c0, -1, noises.cpu().numpy(), 128, True, dt_gamma, 1024)
    assert np.array_equal(r0, mm) and np.array_equal(x0, xm.cpu().numpy()) and np.array_equal(l0, lm.cpu().numpy())


def test_composite_train_against_reference_kernels(scene):
    from customnerf_b200 import raymarching as mine
    rm = ref_ext.raymarching()
    s = scene
    o, d, bf, nr, fr = cu(s["rays_o"]), cu(s["rays_d"]), cu(s["bitfield"]), cu(s["nears"]), cu(s["fars"])
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = mine.march_rays_train(o, d, 2.0, bf, 2, 128, nr, fr, counter, -1, False, 128, True)
    M, N = xyzs.shape[0], rays.shape[0]
    g = torch.Generator(device="cuda").manual_seed(0)
    sig = (torch.rand(M, device="cuda", generator=g) * 80).requires_grad_()
    rgb = torch.rand(M, 3, device="cuda", generator=g).requires_grad_()
    ws, depth, img = mine.composite_rays_train(sig, rgb, deltas, rays, 1e-4)
    ws_r = torch.empty(N, device="cuda"); dp_r = torch.empty(N, device="cuda"); im_r = torch.empty(N, 3, device="cuda")
    rm.composite_rays_train_forward(sig.detach(), rgb.detach(), deltas, rays, M, N, 1e-4, ws_r, dp_r, im_r)
    assert_close(ws.detach().cpu().numpy(), ws_r.cpu().numpy(), 1e-4, 1e-6, "weights_sum")
    assert_close(depth.detach().cpu().numpy(), dp_r.cpu().numpy(), 1e-4, 1e-6, "depth")
    assert_close(img.detach().cpu().numpy(), im_r.cpu().numpy(), 1e-4, 1e-6, "image")
    g_ws = torch.randn(N, device="cuda", generator=g); g_im = torch.randn(N, 3, device="cuda", generator=g)
    torch.autograd.backward([ws, img], [g_ws, g_im])
    gs_r = torch.zeros(M, device="cuda"); gc_r = torch.zeros(M, 3, device="cuda")
    rm.composite_rays_train_backward(g_ws, g_im, sig.detach(), rgb.detach(), deltas, rays, ws_r, im_r, M, N, 1e-4, gs_r, gc_r)
    assert_close(rgb.grad.cpu().numpy(), gc_r.cpu().numpy(), 1e-4, 1e-6, "grad_rgbs")
    gsr = gs_r.cpu().numpy()
    assert_close(sig.grad.cpu().numpy(), gsr, 2e-4, 1e-5 * np.abs(gsr).max(), "grad_sigmas")


@pytest.mark.parametrize("cfg", [dict(log2_hashmap_size=19, desired_resolution=2048, gridtype="hash"),
                                 dict(log2_hashmap_size=21, desired_resolution=8192, gridtype="tiled")])
@pytest.mark.parametrize("half", [False, True])
def test_grid_encoder_against_reference_kernels(scene, cfg, half):
    """B = 262 144 points sampled along the scene's rays (the real access pattern), fp32 and the AMP fp16 path"""
    from customnerf_b200 import raymarching as mine
    from customnerf_b200.gridencoder import GridEncoder
    ge = ref_ext.gridencoder()
    s = scene
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, _, _, _ = mine.march_rays_train(cu(s["rays_o"]), cu(s["rays_d"]), 2.0, cu(s["bitfield"]), 2, 128,
                                          cu(s["nears"]), cu(s["fars"]), counter, -1, False, 128, True)
    B = 262144
    xb = xyzs[:B].contiguous()
    x01 = ((xb + 2) / (2 * 2)).contiguous()          # the very expression GridEncoder.forward evaluates (grid.py:156)
    torch.manual_seed(0)
    enc = GridEncoder(**cfg).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    dt = torch.float16 if half else torch.float32
    S = float(np.log2(enc.per_level_scale))
    e_ref = enc.embeddings.detach().to(dt)
    out_ref = torch.empty(16, B, 2, dtype=dt, device="cuda")
    ge.grid_encode_forward(x01, e_ref, enc.offsets, out_ref, B, 3, 2, 16, 16, S, 16, None, enc.gridtype_id, False, 0)
    out_ref = out_ref.permute(1, 0, 2).reshape(B, 32)
    with torch.autocast("cuda", dtype=torch.float16, enabled=half):
        out = enc(xb, bound=2)
    if half:
        assert_close(out.float().detach().cpu().numpy(), out_ref.float().cpu().numpy(), 2e-3, 1e-3, "forward fp16")
    else:
        assert_close(out.detach().cpu().numpy(), out_ref.cpu().numpy(), 1e-4, 1e-6, "forward fp32")
    g = torch.randn(B, 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)).to(dt)
    out.backward(g)
    gemb_ref = torch.zeros_like(e_ref)
    ge.grid_encode_backward(g.view(B, 16, 2).permute(1, 0, 2).contiguous(), x01, e_ref, enc.offsets, gemb_ref, B, 3, 2,
                            16, 16, S, 16, None, None, enc.gridtype_id, False, 0)
    mine_g = enc.embeddings.grad.cpu().numpy()
    ref_g = gemb_ref.float().cpu().numpy()
    if half:
        # The reference sums rounded __half2 atomics (gridencoder.cu:324-330), in a run-to-run varying order: the
        # hottest coarse cells receive thousands of adds and their fp16 accumulators lose a few percent (measured:
        # 2-5 % at a handful of level-0 cells).  Against the reference the bound is therefore its own error: every
        # entry within rel 3e-2 (+ abs 2e-2 * max|g|) except at most 8 level-0/1 cells, which must stay within 10 %.
        err = np.abs(mine_g - ref_g)
        bad = err > 2e-2 * np.abs(ref_g).max() + 3e-2 * np.abs(ref_g)
        rows = np.nonzero(bad.any(-1))[0]
        assert len(rows) <= 8 and (rows < int(enc.offsets[2])).all(), (len(rows), rows[:10])
        assert (err <= 0.1 * np.abs(ref_g).max() + 0.1 * np.abs(ref_g)).all()
        # ... while against the fp64-accumulated oracle this implementation (fp32 accumulation) meets the
        # north_star bound (rel 1e-2) with room to spare
        from customnerf_b200.gridencoder import level_scales
        sc = level_scales(16, enc.per_level_scale, 16).cpu().numpy()
        g0, _ = cpu_ops.grid_encode_backward(g.float().cpu().numpy(), x01.cpu().numpy(), tuple(enc.embeddings.shape),
                                             enc.offsets.cpu().numpy(), enc.per_level_scale, 16,
                                             gridtype=enc.gridtype_id, scales=sc)
        assert_close(mine_g, g0, 1e-3, 1e-5 * np.abs(g0).max(), "grad fp16 path vs fp64 oracle")
    else:
        assert_close(mine_g, ref_g, 1e-4, 1e-5 * np.abs(ref_g).max(), "grad fp32")


@pytest.mark.parametrize("cfg", [dict(D=3, C=2, L=16, log2=19, res=2048, gridtype="hash"),
                                 dict(D=3, C=2, L=16, log2=21, res=8192, gridtype="tiled"),
                                 dict(D=2, C=4, L=8, log2=14, res=512, gridtype="hash"),
                                 dict(D=3, C=8, L=6, log2=15, res=256, gridtype="hash", align_corners=True)])
def test_total_variation_gradient_against_the_reference_kernel(cfg):
    """grad_total_variation (gridencoder.cu:505-644, grid.py:171-192) -- this repo's point-major kernel (csrc/grid_generic.cuh:
    k_gen_tv) against the reference's kernel_grad_tv on the same points and table.  Both accumulate with fp32 atomics into the
    same cells: rel 1e-4 with an absolute floor of 1e-6 of the largest entry (the order of the additions differs)."""
    from customnerf_b200.gridencoder import GridEncoder
    ge = ref_ext.gridencoder()
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=cfg["D"], num_levels=cfg["L"], level_dim=cfg["C"], base_resolution=16,
                      log2_hashmap_size=cfg["log2"], desired_resolution=cfg["res"], gridtype=cfg["gridtype"],
                      align_corners=cfg.get("align_corners", False)).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    B = 200000
    x = torch.rand(B, cfg["D"], device="cuda")
    x[:7] = torch.tensor([0.0, 1.0, 0.5, 1e-7, 1 - 1e-7, 0.25, 0.75], device="cuda")[:, None]     # faces and cell boundaries
    x[7] = -0.1                                                                                   # outside: contributes nothing
    enc.embeddings.grad = torch.zeros_like(enc.embeddings)
    enc.grad_total_variation(weight=1e-3, inputs=x * 2 - 1, bound=1)
    got = enc.embeddings.grad.clone()
    want = torch.zeros_like(enc.embeddings)
    xin = ((x * 2 - 1) + 1) / 2                 # the wrapper's own mapping to [0, 1] (grid.py:185), same expression as the product's
    ge.grad_total_variation(xin.contiguous(), enc.embeddings.detach(), want, enc.offsets, 1e-3, B, cfg["D"], cfg["C"], cfg["L"],
                            float(np.log2(enc.per_level_scale)), int(enc.base_resolution), enc.gridtype_id, bool(enc.align_corners))
    torch.cuda.synchronize()
    w = want.cpu().numpy()
    assert np.abs(w).max() > 0
    assert_close(got.cpu().numpy(), w, 1e-4, 1e-6 * np.abs(w).max(), "TV gradient vs kernel_grad_tv")
