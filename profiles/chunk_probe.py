#!/usr/bin/env python
"""Do the field kernels (tcgen05, smem-heavy, latency-bound) and the grid-encoder kernels (L2-bound, no smem) overlap when
launched on two streams?  Diagnostic for a chunk-pipelined step: field^T(chunk c+1) beside encode^T(chunk c), encode(chunk
c+1) beside field(chunk c).  The data dependencies are ignored here (both kernels run on the buffers of a finished step)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from customnerf_b200 import trainer, fused_trainer, synthetic as syn, _lib as L  # noqa: E402

dev = torch.device("cuda")
model = trainer.build_scene_model(dev)
o, d = syn.camera_rays(105, 142)
fs = fused_trainer.FusedTrainStep(model, o.shape[0], use_graph=False)
for _ in range(3):
    fs.step(o.to(dev), d.to(dev), syn.bear_color(o + d * 1.5).to(dev))
torch.cuda.synchronize()
side = torch.cuda.Stream()
lib, p = fs.lib, fs.plan
V = C.c_void_p
chk = fused_trainer._check


def fieldb(st):
    chk(lib.nb200_field_backward(V(p.d_sigma), V(p.d_rgba), V(p.sigma_arg), V(p.rgba), V(p.x_en), V(p.dirs), V(p.act), V(p.w_bwd),
                                 V(p.d_x_en), V(p.g_trunk), V(p.g_density), V(p.g_rgb), C.c_uint32(p.M_cap), V(p.m_eff),
                                 V(p.wg_scratch), None, st), "fieldb")


def encb(st):
    chk(lib.nb200_fs_encode_backward_levels(V(p.d_x_en), V(p.xyzs), C.c_float(p.bound), V(p.offsets), V(p.g_table), C.c_uint32(p.M_cap),
                                            C.c_uint32(p.L), C.c_float(p.S), C.c_uint32(p.base_res), C.c_uint32(p.gridtype), 0, 0,
                                            V(p.m_eff), C.c_uint32(0), C.c_uint32(p.L), st), "encb")


def encf(st):
    chk(lib.nb200_fs_encode_forward(V(p.xyzs), C.c_float(p.bound), V(p.table), V(p.offsets), V(p.x_en), C.c_uint32(p.M_cap),
                                    C.c_uint32(p.L), C.c_float(p.S), C.c_uint32(p.base_res), C.c_uint32(p.gridtype), 0, 0, V(p.m_eff), st), "encf")


def fieldf(st):
    chk(lib.nb200_field_forward(V(p.x_en), V(p.xyzs), V(p.dirs), V(p.w_fwd), V(p.sigma), V(p.sigma_arg), V(p.rgba), V(p.act),
                                C.c_uint32(p.M_cap), V(p.m_eff), st), "fieldf")


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def pair(first, second):
    def run():
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        first(L.stream())
        with torch.cuda.stream(side):
            second(L.stream())
        main.wait_stream(side)
    return run


def serial(x, y):
    return lambda: (x(L.stream()), y(L.stream()))


print("rows capacity %d" % p.M_cap)
for name, x, y in (("field^T | encode^T", fieldb, encb), ("encode | field", encf, fieldf)):
    tx, ty = timeit(lambda: x(L.stream())), timeit(lambda: y(L.stream()))
    print("%-20s alone %.1f + %.1f us, serial %.1f us, two streams (first issued first) %.1f us, (second first) %.1f us"
          % (name, tx, ty, timeit(serial(x, y)), timeit(pair(x, y)), timeit(pair(y, x))))
