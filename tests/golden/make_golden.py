"""Mint golden vectors from the UNMODIFIED reference CUDA extensions (oracle/_ref, built by oracle/build_ref.py).

Run on the GPU box:   python tests/golden/make_golden.py        -> tests/golden/ref_ext_vectors.npz
The reference ships no golden vectors of its own; these are outputs of its real kernels
(gridencoder/src/gridencoder.cu, raymarching/src/raymarching.cu compiled for sm_100a with only the -std flag
patched) on small seeded inputs.  Inputs are stored next to the outputs, so the checks
(tests/test_golden_cpu.py for the oracle, tests/test_gpu_golden.py for the CUDA path) need nothing else.

Canonicalisation: march_rays_train reserves output slots with atomics (raymarching.cu:405-406), so ``rays`` row
order and offsets differ run to run.  The stored form is sorted by ray id with each ray's segment gathered in that
order, i.e. the exclusive-scan layout.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_ext  # noqa: E402
from customnerf_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(HERE, "ref_ext_vectors.npz")


def exact_table(rows, C):
    """values in [-1, 1) that are exact in fp16 and fp32 on every platform (11-bit fractions)"""
    i = np.arange(rows * C, dtype=np.uint64)
    h = (i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    return (((h >> np.uint64(21)).astype(np.float32) / 1024.0) - 1.0).reshape(rows, C)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main():
    assert torch.cuda.is_available()
    assert ref_ext.available(), "oracle/_ref/*.so missing: run python oracle/build_ref.py where /root/reference exists"
    ge, rm = ref_ext.gridencoder(), ref_ext.raymarching()
    rng = np.random.RandomState(20261017)
    G = {}

    # ---------------- morton / packbits / near-far
    coords = rng.randint(0, 128, (64, 3)).astype(np.int32)
    ind = torch.empty(64, dtype=torch.int32, device="cuda")
    rm.morton3D(cu(coords), 64, ind)
    back = torch.empty(64, 3, dtype=torch.int32, device="cuda")
    rm.morton3D_invert(ind, 64, back)
    G["morton_coords"], G["morton_indices"], G["morton_back"] = coords, ind.cpu().numpy(), back.cpu().numpy()

    pg = rng.uniform(0, 2, (2, 512)).astype(np.float32)
    pg[0, :8] = 1.0
    bits = torch.empty(128, dtype=torch.uint8, device="cuda")
    rm.packbits(cu(pg), 128, 1.0, bits)
    G["packbits_grid"], G["packbits_thresh"], G["packbits_out"] = pg, np.float32(1.0), bits.cpu().numpy()

    grid = syn.density_grid(2, 128)
    thr = min(float(grid.mean()), 10.0)
    bf = torch.empty(2 * 128 ** 3 // 8, dtype=torch.uint8, device="cuda")
    rm.packbits(grid.cuda(), bf.numel(), thr, bf)
    G["scene_bitfield"] = bf.cpu().numpy()

    o, d = syn.camera_rays(105, 142)
    o, d = o.numpy(), d.numpy()
    aabb = np.array([-2, -2, -2, 2, 2, 2], np.float32)
    N = o.shape[0]
    nears = torch.empty(N, device="cuda"); fars = torch.empty(N, device="cuda")
    rm.near_far_from_aabb(cu(o), cu(d), cu(aabb), N, 0.2, nears, fars)
    # pick 192 rays that hit the bear + 64 that do not (incl. the image corners)
    sel = np.concatenate([np.arange(6000, 6000 + 142 * 40, 30)[:192], np.arange(0, 64)])
    o_s, d_s = o[sel], d[sel]
    extra_o = rng.uniform(-3, 3, (32, 3)).astype(np.float32)
    extra_d = rng.normal(size=(32, 3)).astype(np.float32)
    nf_o, nf_d = np.concatenate([o_s, extra_o]), np.concatenate([d_s, extra_d])
    n2 = torch.empty(len(nf_o), device="cuda"); f2 = torch.empty(len(nf_o), device="cuda")
    rm.near_far_from_aabb(cu(nf_o), cu(nf_d), cu(aabb), len(nf_o), 0.2, n2, f2)
    G["nf_rays_o"], G["nf_rays_d"], G["nf_aabb"] = nf_o, nf_d, aabb
    G["nf_nears"], G["nf_fars"] = n2.cpu().numpy(), f2.cpu().numpy()

    # ---------------- march_rays_train (perturbed and not)
    Ns = len(sel)
    nr, fr = nears.cpu().numpy()[sel], fars.cpu().numpy()[sel]
    G["march_rays_o"], G["march_rays_d"], G["march_nears"], G["march_fars"] = o_s, d_s, nr, fr
    for tag, noises in (("np", np.zeros(Ns, np.float32)), ("pt", rng.uniform(0, 1, Ns).astype(np.float32))):
        M = Ns * 1024
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda")
        deltas = torch.zeros(M, 2, device="cuda")
        rays = torch.empty(Ns, 3, dtype=torch.int32, device="cuda")
        counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        rm.march_rays_train(cu(o_s), cu(d_s), bf, 2.0, 0.0, 1024, Ns, 2, 128, M, cu(nr), cu(fr), xyzs, dirs, deltas, rays,
                            counter, cu(noises))
        rays = rays.cpu().numpy()
        order = np.argsort(rays[:, 0], kind="stable")
        rays = rays[order]
        xs, ds, ls = xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy()
        segx = np.concatenate([xs[r[1]:r[1] + r[2]] for r in rays])
        segd = np.concatenate([ds[r[1]:r[1] + r[2]] for r in rays])
        segl = np.concatenate([ls[r[1]:r[1] + r[2]] for r in rays])
        G["march_%s_noises" % tag] = noises
        G["march_%s_counts" % tag] = rays[:, 2].copy()
        G["march_%s_counter" % tag] = counter.cpu().numpy()
        G["march_%s_xyzs" % tag], G["march_%s_dirs" % tag], G["march_%s_deltas" % tag] = segx, segd, segl
        print("march", tag, "total samples", int(counter[0]), "hit rays", int((rays[:, 2] > 0).sum()))

    # ---------------- composite_rays_train fwd / bwd on the unperturbed samples
    counts = G["march_np_counts"]
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays_c = np.stack([np.arange(Ns, dtype=np.int32), offs, counts], -1).astype(np.int32)
    M = int(counts.sum())
    x = G["march_np_xyzs"]
    sig = (25.0 * (1 + np.sin(9 * x.sum(-1)))).astype(np.float32)
    rgb = (0.5 + 0.5 * np.cos(4 * x)).astype(np.float32)
    dl = G["march_np_deltas"]
    ws = torch.empty(Ns, device="cuda"); dp = torch.empty(Ns, device="cuda"); im = torch.empty(Ns, 3, device="cuda")
    rm.composite_rays_train_forward(cu(sig), cu(rgb), cu(dl), cu(rays_c), M, Ns, 1e-4, ws, dp, im)
    g_ws = rng.normal(size=Ns).astype(np.float32); g_im = rng.normal(size=(Ns, 3)).astype(np.float32)
    gs = torch.zeros(M, device="cuda"); gc = torch.zeros(M, 3, device="cuda")
    rm.composite_rays_train_backward(cu(g_ws), cu(g_im), cu(sig), cu(rgb), cu(dl), cu(rays_c), ws, im, M, Ns, 1e-4, gs, gc)
    G.update(comp_sigmas=sig, comp_rgbs=rgb, comp_deltas=dl, comp_rays=rays_c, comp_ws=ws.cpu().numpy(),
             comp_depth=dp.cpu().numpy(), comp_image=im.cpu().numpy(), comp_g_ws=g_ws, comp_g_image=g_im,
             comp_grad_sigmas=gs.cpu().numpy(), comp_grad_rgbs=gc.cpu().numpy())

    # ---------------- inference: one march_rays + composite_rays round
    n_alive, n_step = 64, 4
    alive = np.arange(0, 128, 2, dtype=np.int32)
    xyzs = torch.zeros(n_alive * n_step, 3, device="cuda"); dirs = torch.zeros_like(xyzs)
    deltas = torch.zeros(n_alive * n_step, 2, device="cuda")
    rays_t = cu(nr.copy())
    rm.march_rays(n_alive, n_step, cu(alive), rays_t, cu(o_s), cu(d_s), 2.0, 0.0, 1024, 2, 128, bf, cu(nr), cu(fr), xyzs,
                  dirs, deltas, cu(np.zeros(n_alive, np.float32)))
    xi = xyzs.cpu().numpy()
    sig_i = (25.0 * (1 + np.sin(9 * xi.sum(-1)))).astype(np.float32)
    rgb_i = (0.5 + 0.5 * np.cos(4 * xi)).astype(np.float32)
    ws_i = torch.zeros(Ns, device="cuda"); dp_i = torch.zeros(Ns, device="cuda"); im_i = torch.zeros(Ns, 3, device="cuda")
    alive_t = cu(alive.copy())
    rm.composite_rays(n_alive, n_step, 1e-4, alive_t, rays_t, cu(sig_i), cu(rgb_i), deltas, ws_i, dp_i, im_i)
    G.update(inf_alive=alive, inf_n_step=np.int32(n_step), inf_xyzs=xi, inf_dirs=dirs.cpu().numpy(),
             inf_deltas=deltas.cpu().numpy(), inf_sigmas=sig_i, inf_rgbs=rgb_i, inf_alive_out=alive_t.cpu().numpy(),
             inf_rays_t_out=rays_t.cpu().numpy(), inf_ws=ws_i.cpu().numpy(), inf_depth=dp_i.cpu().numpy(),
             inf_image=im_i.cpu().numpy())

    # ---------------- grid encoder: small hash + tiled tables, fp32 and fp16
    from oracle import cpu_ops
    for tag, gridtype, log2T, res in (("hash", 0, 12, 512), ("tiled", 1, 13, 1024)):
        offs_g, pls = cpu_ops.grid_offsets(num_levels=8, log2_hashmap_size=log2T, desired_resolution=res)
        emb = exact_table(int(offs_g[-1]), 2)
        B = 256
        xin = rng.uniform(0, 1, (B, 3)).astype(np.float32)
        xin[:4] = [[0, 0, 0], [1, 1, 1], [1.01, 0.5, 0.5], [0.5, 0.5, 0.5]]
        S = float(np.log2(pls))
        grad = (rng.randint(-8, 9, (B, 16)) / 8.0).astype(np.float32)          # exact in fp16
        for dt_tag, dt in (("f32", torch.float32), ("f16", torch.float16)):
            e = cu(emb).to(dt)
            out = torch.empty(8, B, 2, dtype=dt, device="cuda")
            ge.grid_encode_forward(cu(xin), e, cu(offs_g), out, B, 3, 2, 8, 8, S, 16, None, gridtype, False, 0)
            g_lbc = cu(grad).to(dt).view(B, 8, 2).permute(1, 0, 2).contiguous()
            gemb = torch.zeros_like(e)
            ge.grid_encode_backward(g_lbc, cu(xin), e, cu(offs_g), gemb, B, 3, 2, 8, 8, S, 16, None, None, gridtype,
                                    False, 0)
            G["grid_%s_%s_out" % (tag, dt_tag)] = out.permute(1, 0, 2).reshape(B, 16).float().cpu().numpy()
            G["grid_%s_%s_gemb" % (tag, dt_tag)] = gemb.float().cpu().numpy()
        from customnerf_b200.gridencoder import level_scales
        G["grid_%s_scales" % tag] = level_scales(8, pls, 16).cpu().numpy()     # device exp2f values (see orc_locate)
        G["grid_%s_offsets" % tag], G["grid_%s_pls" % tag] = offs_g, np.float64(pls)
        G["grid_%s_inputs" % tag], G["grid_%s_grad" % tag] = xin, grad
        G["grid_%s_log2T" % tag], G["grid_%s_res" % tag] = np.int32(log2T), np.int32(res)

    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
