#!/usr/bin/env python
"""Do the optimiser sweep and the ray march overlap when launched on two streams?  (diagnostic for pipeline_update)"""
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
fs.step(o.to(dev), d.to(dev), syn.bear_color(o + d * 1.5).to(dev))
torch.cuda.synchronize()
side = torch.cuda.Stream(priority=0)
hi = torch.cuda.Stream(priority=-1)
lib = fs.lib


def march(st):
    fused_trainer._check(lib.nb200_train_phase(C.byref(fs.plan), C.c_int(1), st), "march")


def update(st):
    fused_trainer._check(lib.nb200_train_update(C.byref(fs.plan), st), "update")


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def both():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        update(L.stream())
    march(L.stream())
    main.wait_stream(side)


def both_march_first():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    march(L.stream())
    with torch.cuda.stream(side):
        update(L.stream())
    main.wait_stream(side)


def both_prio():
    main = torch.cuda.current_stream()
    side.wait_stream(main); hi.wait_stream(main)
    with torch.cuda.stream(side):
        update(L.stream())
    with torch.cuda.stream(hi):
        march(L.stream())
    main.wait_stream(side); main.wait_stream(hi)


def both_update_prio():
    main = torch.cuda.current_stream()
    hi.wait_stream(main)
    with torch.cuda.stream(hi):
        update(L.stream())
    march(L.stream())
    main.wait_stream(hi)


def adam_only(st):
    p = fs.plan
    fused_trainer._check(lib.nb200_fused_adam(C.c_void_p(p.params_flat), C.c_void_p(p.grads_flat), C.c_void_p(p.exp_avg),
                                              C.c_void_p(p.exp_avg_sq), C.c_uint64(p.n_params), C.c_uint64(p.n_table_params),
                                              C.c_void_p(p.hyper), C.c_int(1), st), "adam")


def both_adam_only_prio():
    """only the sweep kernel (no one-thread hyper kernel in front of it) on the high-priority stream, issued first"""
    main = torch.cuda.current_stream()
    hi.wait_stream(main)
    with torch.cuda.stream(hi):
        adam_only(L.stream())
    march(L.stream())
    main.wait_stream(hi)


print("NB200_ADAM_CTAS_PER_SM =", os.environ.get("NB200_ADAM_CTAS_PER_SM"))
print("march alone   %.1f us" % timeit(lambda: march(L.stream())))
print("update alone  %.1f us" % timeit(lambda: update(L.stream())))
print("serial        %.1f us" % timeit(lambda: (update(L.stream()), march(L.stream()))))
print("two streams (update issued first) %.1f us" % timeit(both))
print("two streams (march issued first)  %.1f us" % timeit(both_march_first))
print("two streams, march on a high-priority stream %.1f us" % timeit(both_prio))
print("two streams, update on a high-priority stream, issued first %.1f us" % timeit(both_update_prio))
print("adam alone    %.1f us" % timeit(lambda: adam_only(L.stream())))
print("two streams, sweep kernel only on a high-priority stream, issued first %.1f us" % timeit(both_adam_only_prio))
