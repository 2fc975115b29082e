#!/usr/bin/env python
"""k_field_backward / k_field_forward time against the number of rows: slope = steady-state cost per 128-row tile, intercept =
fixed cost per launch (prologue, weight-gradient flush + slab reduce, tail).  Random inputs, weights of a fresh network."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from customnerf_b200 import trainer, _lib as L  # noqa: E402

dev = torch.device("cuda")
lib = L.lib()
model = trainer.build_scene_model(dev, log2_hashmap_size=15, desired_resolution=512)
nb = int(lib.nb200_field_weight_image_bytes())
w_fwd, w_bwd = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
L.check(lib.nb200_field_pack_weights(L.ptr(model.network.params.detach()), L.ptr(model.density_network.params.detach()),
                                     L.ptr(model.rgb_network.params.detach()), L.ptr(w_fwd), L.ptr(w_bwd), L.stream()), "pack")
lib.nb200_field_wgrad_scratch_bytes.restype = C.c_uint32
wg = torch.empty(int(lib.nb200_field_wgrad_scratch_bytes()) // 4, dtype=torch.float32, device=dev)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = [torch.zeros(n, device=dev) for n in (model.network.params.numel(), model.density_network.params.numel(), model.rgb_network.params.numel())]


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


for M in [int(v) for v in os.environ.get("PROBE_M", "18944,37888,75776,151552,269696,539392,1078784,2697216").split(",")]:
    x_en = (torch.randn(M, 32, device=dev) * 0.3).half()
    xyz, dirs = torch.rand(M, 3, device=dev) * 2 - 1, torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
    sigma, sarg = torch.empty(M, device=dev), torch.empty(M, device=dev)
    rgba, act = torch.empty(M, 4, dtype=torch.half, device=dev), torch.empty(5, M, 64, dtype=torch.half, device=dev)
    d_sigma, d_rgba, d_x_en = torch.randn(M, device=dev) * 1e-3, torch.randn(M, 4, device=dev) * 1e-3, torch.empty(M, 32, dtype=torch.half, device=dev)

    def fwd():
        L.check(lib.nb200_field_forward(L.ptr(x_en), L.ptr(xyz), L.ptr(dirs), L.ptr(w_fwd), L.ptr(sigma), L.ptr(sarg), L.ptr(rgba),
                                        L.ptr(act), L.u32(M), L.ptr(None), L.stream()), "fwd")

    def bwd():
        L.check(lib.nb200_field_backward(L.ptr(d_sigma), L.ptr(d_rgba), L.ptr(sarg), L.ptr(rgba), L.ptr(x_en), L.ptr(dirs), L.ptr(act),
                                         L.ptr(w_bwd), L.ptr(d_x_en), L.ptr(g[0]), L.ptr(g[1]), L.ptr(g[2]), L.u32(M), L.ptr(None),
                                         L.ptr(wg if os.environ.get("PROBE_ATOMICS", "0") != "1" else None), L.ptr(None), L.stream()), "bwd")
    tf, tb = timeit(fwd), timeit(bwd)
    tiles = (M + 127) // 128
    print("M %8d  tiles %6d (%.1f per SM)  forward %8.1f us  backward %8.1f us   per tile and SM: fwd %.2f  bwd %.2f us"
          % (M, tiles, tiles / 148.0, tf, tb, tf / (tiles / 148.0), tb / (tiles / 148.0)))
    del x_en, xyz, dirs, sigma, sarg, rgba, act, d_sigma, d_rgba, d_x_en
