"""Checks shared by test_golden_cpu.py (oracle) and test_gpu_golden.py (CUDA path): run an implementation on the
inputs stored in tests/golden/ref_ext_vectors.npz and compare with the reference kernels' stored outputs."""
import os

import numpy as np

from conftest import assert_close

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ext_vectors.npz")


def load():
    if not os.path.isfile(GOLDEN):
        return None
    return dict(np.load(GOLDEN))


def exact_table(rows, C):
    i = np.arange(rows * C, dtype=np.uint64)
    h = (i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    return (((h >> np.uint64(21)).astype(np.float32) / 1024.0) - 1.0).reshape(rows, C)


def check_integer_ops(G, ops):
    assert np.array_equal(ops.morton3D(G["morton_coords"]), G["morton_indices"])
    assert np.array_equal(ops.morton3D_invert(G["morton_indices"]), G["morton_back"])
    assert np.array_equal(ops.packbits(G["packbits_grid"], float(G["packbits_thresh"])), G["packbits_out"])
    n, f = ops.near_far_from_aabb(G["nf_rays_o"], G["nf_rays_d"], G["nf_aabb"], 0.2)
    assert np.array_equal(n, G["nf_nears"]) and np.array_equal(f, G["nf_fars"])


def check_march(G, ops):
    for tag in ("np", "pt"):
        counter = np.zeros(2, np.int32)
        noises = G["march_%s_noises" % tag] if tag == "pt" else None
        xyzs, dirs, deltas, rays = ops.march_rays_train(G["march_rays_o"], G["march_rays_d"], 2.0, G["scene_bitfield"],
                                                        2, 128, G["march_nears"], G["march_fars"], counter, noises)
        counts = G["march_%s_counts" % tag]
        assert np.array_equal(rays[:, 0], np.arange(len(counts)))
        assert np.array_equal(rays[:, 2], counts), "per-ray sample counts must be bit-exact"
        assert np.array_equal(counter, G["march_%s_counter" % tag])
        m = int(counts.sum())
        assert np.array_equal(rays[:, 1], np.concatenate([[0], np.cumsum(counts)[:-1]]))
        assert np.array_equal(xyzs[:m], G["march_%s_xyzs" % tag])
        assert np.array_equal(dirs[:m], G["march_%s_dirs" % tag])
        assert np.array_equal(deltas[:m], G["march_%s_deltas" % tag])
        assert not xyzs[m:].any() and xyzs.shape[0] == m + (128 - m % 128)


def check_composite(G, ops):
    ws, depth, image = ops.composite_forward(G["comp_sigmas"], G["comp_rgbs"], G["comp_deltas"], G["comp_rays"], 1e-4)
    assert_close(ws, G["comp_ws"], 1e-4, 1e-6, "weights_sum")
    assert_close(depth, G["comp_depth"], 1e-4, 1e-6, "depth")
    assert_close(image, G["comp_image"], 1e-4, 1e-6, "image")
    gs, gc = ops.composite_backward(G["comp_g_ws"], G["comp_g_image"], G["comp_sigmas"], G["comp_rgbs"],
                                    G["comp_deltas"], G["comp_rays"], G["comp_ws"], G["comp_image"], 1e-4)
    assert_close(gc, G["comp_grad_rgbs"], 1e-4, 1e-6, "grad_rgbs")
    assert_close(gs, G["comp_grad_sigmas"], 2e-4, 1e-5 * np.abs(G["comp_grad_sigmas"]).max(), "grad_sigmas")


def check_inference(G, ops):
    n_step = int(G["inf_n_step"])
    alive = G["inf_alive"].copy()
    n_alive = len(alive)
    rays_t = G["march_nears"].copy()
    xyzs, dirs, deltas = ops.march_rays(n_alive, n_step, alive, rays_t, G["march_rays_o"], G["march_rays_d"], 2.0,
                                        G["scene_bitfield"], 2, 128, G["march_nears"], G["march_fars"])
    k = n_alive * n_step
    assert np.array_equal(xyzs[:k], G["inf_xyzs"]) and np.array_equal(deltas[:k], G["inf_deltas"])
    assert np.array_equal(dirs[:k], G["inf_dirs"])
    Ns = len(G["march_nears"])
    ws, depth, image = np.zeros(Ns, np.float32), np.zeros(Ns, np.float32), np.zeros((Ns, 3), np.float32)
    alive, rays_t, ws, depth, image = ops.composite_rays(n_alive, n_step, alive, rays_t, G["inf_sigmas"], G["inf_rgbs"],
                                                         G["inf_deltas"], ws, depth, image, 1e-4)
    assert np.array_equal(alive, G["inf_alive_out"])
    assert_close(rays_t, G["inf_rays_t_out"], 1e-6, 0, "rays_t")
    assert_close(ws, G["inf_ws"], 1e-4, 1e-6, "ws")
    assert_close(image, G["inf_image"], 1e-4, 1e-6, "image")
    assert_close(depth, G["inf_depth"], 1e-4, 1e-6, "depth")


def check_grid(G, ops):
    for tag, gridtype in (("hash", 0), ("tiled", 1)):
        offs = G["grid_%s_offsets" % tag]
        emb = exact_table(int(offs[-1]), 2)
        pls = float(G["grid_%s_pls" % tag])
        x, grad, sc = G["grid_%s_inputs" % tag], G["grid_%s_grad" % tag], G["grid_%s_scales" % tag]
        out, gemb = ops.grid(x, emb, offs, pls, gridtype, grad, sc, half=False)
        assert_close(out, G["grid_%s_f32_out" % tag], 1e-4, 1e-6, tag + " forward fp32")
        g0 = G["grid_%s_f32_gemb" % tag]
        assert_close(gemb, g0, 1e-4, 1e-6 * np.abs(g0).max(), tag + " grad fp32")
        out, gemb = ops.grid(x, emb, offs, pls, gridtype, grad, sc, half=True)
        assert_close(out, G["grid_%s_f16_out" % tag], 2e-3, 1e-3, tag + " forward fp16")
        g0 = G["grid_%s_f16_gemb" % tag]
        # the reference accumulates __half2 atomics: rel 1e-2 with an abs floor of 1e-3 * max|g| (north_star)
        assert_close(gemb, g0, 1e-2, 4e-3 * np.abs(g0).max(), tag + " grad fp16")
