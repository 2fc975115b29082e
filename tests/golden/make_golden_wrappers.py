"""Mint golden vectors from the reference's raymarching Python wrappers -- /root/reference/raymarching/raymarching.py:20-430,
the operator surface the renderer calls (allocation and alignment rules, the step counter, noise draws, the autograd
Function of composite_rays_train that ignores grad_depth, shape handling) -- imported UNMODIFIED and run on the CPU:

  * the compiled ``_raymarching`` module they bind (bindings.cpp:5-20, CUDA only) is replaced by shims that hand the same
    tensors to the C oracle (oracle/nerf_oracle.c), one shim per pybind entry, same argument order;
  * ``Tensor.cuda()`` is the identity while they run (every wrapper moves its inputs to the GPU first).

So the arithmetic is the oracle's (itself pinned to the reference's CUDA kernels by ref_ext_vectors.npz) and what these vectors
pin is the wrappers' own logic; tests/test_golden_wrappers_cpu.py checks oracle/cpu_ops.py (the restated wrappers the GPU
parity tests use as their checker) against them.

Run in the build container:  python tests/golden/make_golden_wrappers.py  ->  tests/golden/ref_wrappers.npz
"""
import ctypes as C
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import cpu_ops  # noqa: E402
from customnerf_b200 import synthetic as syn  # noqa: E402

u32, f32 = C.c_uint32, C.c_float


def P(t):
    if t is None:
        return None
    assert t.is_contiguous() and t.device.type == "cpu"
    return C.c_void_p(t.data_ptr())


def backend_shims():
    L = cpu_ops.lib()
    m = types.ModuleType("_raymarching")
    m.near_far_from_aabb = lambda o, d, aabb, N, min_near, nears, fars: L.orc_near_far_from_aabb(P(o), P(d), P(aabb), u32(N), f32(min_near), P(nears), P(fars))
    m.sph_from_ray = lambda o, d, radius, N, coords: L.orc_sph_from_ray(P(o), P(d), f32(radius), u32(N), P(coords))
    m.morton3D = lambda coords, N, indices: L.orc_morton3D(P(coords.contiguous()), u32(N), P(indices))
    m.morton3D_invert = lambda indices, N, coords: L.orc_morton3D_invert(P(indices.contiguous()), u32(N), P(coords))
    m.packbits = lambda grid, N, thresh, bitfield: L.orc_packbits(P(grid), u32(N), f32(np.float32(thresh)), P(bitfield))
    m.march_rays_train = lambda o, d, grid, bound, dt_gamma, max_steps, N, Cc, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises: \
        L.orc_march_rays_train(P(o), P(d), P(grid), f32(bound), f32(dt_gamma), u32(max_steps), u32(N), u32(Cc), u32(H), u32(M), P(nears),
                               P(fars), P(xyzs), P(dirs), P(deltas), P(rays), P(counter), P(noises))
    m.composite_rays_train_forward = lambda s, c, dl, rays, M, N, T, ws, depth, image: \
        L.orc_composite_rays_train_forward(P(s), P(c), P(dl), P(rays), u32(M), u32(N), f32(T), P(ws), P(depth), P(image))
    m.composite_rays_train_backward = lambda gws, gimg, s, c, dl, rays, ws, image, M, N, T, gs, gc: \
        L.orc_composite_rays_train_backward(P(gws), P(gimg), P(s), P(c), P(dl), P(rays), P(ws), P(image), u32(M), u32(N), f32(T), P(gs), P(gc))
    m.march_rays = lambda n_alive, n_step, alive, rt, o, d, bound, dt_gamma, max_steps, Cc, H, grid, near, far, xyzs, dirs, deltas, noises: \
        L.orc_march_rays(u32(n_alive), u32(n_step), P(alive), P(rt), P(o), P(d), f32(bound), f32(dt_gamma), u32(max_steps), u32(Cc), u32(H),
                         P(grid), P(near), P(far), P(xyzs), P(dirs), P(deltas), P(noises))
    m.composite_rays = lambda n_alive, n_step, T, alive, rt, s, c, dl, ws, depth, image: \
        L.orc_composite_rays(u32(n_alive), u32(n_step), f32(T), P(alive), P(rt), P(s), P(c), P(dl), P(ws), P(depth), P(image))
    return m


def load_reference():
    sys.modules["_raymarching"] = backend_shims()
    pkg = types.ModuleType("raymarching")
    pkg.__path__ = ["/root/reference/raymarching"]
    sys.modules["raymarching"] = pkg
    return importlib.import_module("raymarching.raymarching")


def scene():
    grid = syn.density_grid(2, 128)
    o, d = syn.camera_rays(12, 16)
    return grid, min(float(grid.mean()), 10.0), o.contiguous(), d.contiguous()


def field(x, d):
    return syn.bear_density(x), (syn.bear_color(x) * (0.5 + 0.5 * d[:, :1].abs())).contiguous()


def main():
    rm = load_reference()
    keep = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    G = {}
    try:
        grid, thr, o, d = scene()
        aabb = torch.tensor([-2, -2, -2, 2, 2, 2], dtype=torch.float32)
        # near / far: default min_near and an explicit one; [B, N, 3] inputs flatten
        n0, f0 = rm.near_far_from_aabb(o[None], d[None], aabb)
        n1, f1 = rm.near_far_from_aabb(o, d, aabb, 0.01)
        G["nf_default"], G["nf_001"] = torch.stack([n0, f0]).numpy(), torch.stack([n1, f1]).numpy()
        G["sph"] = rm.sph_from_ray(o, d, 2.5).numpy()
        coords = torch.randint(0, 128, (36, 3), generator=torch.Generator().manual_seed(1))   # [N, 3]: the wrapper takes N = shape[0]
        ind = rm.morton3D(coords)
        G["morton_coords"], G["morton_indices"], G["morton_back"] = coords.numpy(), ind.numpy(), rm.morton3D_invert(ind).numpy()
        bits = rm.packbits(grid, thr)
        G["packbits_thresh"], G["packbits_sum"] = np.float32(thr), np.int64(np.unpackbits(bits.numpy()).sum())
        G["packbits_head"] = bits.numpy()[:4096].copy()
        # march_rays_train: (a) everything, aligned to 128; (b) perturbed (the wrapper draws torch.rand(N)); (c) a mean_count budget
        for tag, kw, seed in (("all", dict(mean_count=-1, perturb=False, align=128, force_all_rays=True), None),
                              ("perturb", dict(mean_count=-1, perturb=True, align=128, force_all_rays=False), 3),
                              ("budget", dict(mean_count=1000, perturb=False, align=128, force_all_rays=False), None)):
            counter = torch.zeros(2, dtype=torch.int32)
            if seed is not None:
                torch.manual_seed(seed)
            xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, 2, bits, 2, 128, n0, f0, counter, kw["mean_count"], kw["perturb"],
                                                           kw["align"], kw["force_all_rays"], 0, 1024)
            G["mt_%s_xyzs" % tag], G["mt_%s_dirs" % tag], G["mt_%s_deltas" % tag] = xyzs.numpy(), dirs.numpy(), deltas.numpy()
            G["mt_%s_rays" % tag], G["mt_%s_counter" % tag] = rays.numpy(), counter.numpy().copy()
            if tag == "all":
                keep_march = (xyzs, dirs, deltas, rays)
        # composite_rays_train through autograd (grad_depth is ignored by the reference, raymarching.py:274-289)
        xyzs, dirs, deltas, rays = keep_march
        sig, rgb = field(xyzs, dirs)
        sig, rgb = sig.clone().requires_grad_(), rgb.clone().requires_grad_()
        ws, depth, image = rm.composite_rays_train(sig, rgb, deltas, rays, 1e-4)
        g = torch.Generator().manual_seed(2)
        gws, gdepth, gimg = torch.randn(ws.shape, generator=g), torch.randn(depth.shape, generator=g), torch.randn(image.shape, generator=g)
        ((ws * gws).sum() + (depth * gdepth).sum() + (image * gimg).sum()).backward()
        for k, v in (("ws", ws), ("depth", depth), ("image", image), ("gws", gws), ("gdepth", gdepth), ("gimg", gimg), ("grad_sigmas", sig.grad),
                     ("grad_rgbs", rgb.grad)):
            G["ct_" + k] = v.detach().numpy()
        # the inference loop of run_cuda over march_rays / composite_rays (three rounds)
        N = o.shape[0]
        ws, depth, image = torch.zeros(N), torch.zeros(N), torch.zeros(N, 3)
        alive, rays_t = torch.arange(N, dtype=torch.int32), n0.clone()
        for rnd in range(3):
            n_alive = alive.shape[0]
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = rm.march_rays(n_alive, n_step, alive, rays_t, o, d, 2, bits, 2, 128, n0, f0, 128, False, 0, 1024)
            s_, c_ = field(xyzs, dirs)
            rm.composite_rays(n_alive, n_step, alive, rays_t, s_, c_, deltas, ws, depth, image, 1e-4)
            G["inf%d_shape" % rnd] = np.array(xyzs.shape, np.int64)
            G["inf%d_alive" % rnd], G["inf%d_rays_t" % rnd] = alive.numpy().copy(), rays_t.numpy().copy()
            G["inf%d_image" % rnd], G["inf%d_ws" % rnd], G["inf%d_depth" % rnd] = image.numpy().copy(), ws.numpy().copy(), depth.numpy().copy()
            alive = alive[alive >= 0]
    finally:
        torch.Tensor.cuda = keep
    np.savez_compressed(os.path.join(HERE, "ref_wrappers.npz"), **G)
    print("wrote ref_wrappers.npz:", {k: v.shape for k, v in G.items() if k.startswith("mt_") and k.endswith(("xyzs", "counter"))})


if __name__ == "__main__":
    main()
