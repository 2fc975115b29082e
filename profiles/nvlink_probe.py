#!/usr/bin/env python
"""NVLink rates between the GPUs of this box as kernels see them (torchrun, one process per GPU): SM-issued pull (loads from
the next rank's memory) and push (stores into it), one rank alone and all ranks at once, next to a copy-engine
cudaMemcpy of the same bytes.  MEASURED_PEAKS.json has no NVLink figure; this is the roofline the peer-memory update kernel
is held against."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from customnerf_b200 import parallel, _lib as L  # noqa: E402

rank, local_rank, world = parallel.init_from_env()
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
lib = L.lib()
NB = 256 << 20
peer = parallel.PeerMemory(NB // 8, dev)          # [params | grads] = 2 x NB / 8 floats = NB bytes, used as one raw buffer
local = torch.zeros(NB // 4, dtype=torch.float32, device=dev)
nxt = (rank + 1) % world
remote_ptr = peer.bases[nxt]
remote = torch.as_tensor(parallel._DevArray(remote_ptr, NB // 4, "<f4", peer), device=dev)
mine = peer.bases[rank]


def timed(fn, active, reps=10):
    """all ranks enter together; only ranks in `active` do work; returns max over active ranks of us per rep"""
    torch.cuda.synchronize(); dist.barrier()
    us = 0.0
    if rank in active:
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / reps * 1e3
    t = torch.tensor([us], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sm_copy(dst, src, ctas):
    return lambda: L.check(lib.nb200_stream_copy_probe(C.c_void_p(dst), C.c_void_p(src), C.c_uint64(NB), C.c_uint32(ctas), L.stream()), "copy")


out = {"world": world, "bytes": NB}
everyone, one = set(range(world)), {0}
for name, ctas in (("148x2", 296), ("148x8", 1184)):
    for who, active in (("one_rank", one), ("all_ranks", everyone)):
        out["pull_%s_%s_gbs" % (name, who)] = round(NB / timed(sm_copy(local.data_ptr(), remote_ptr, ctas), active) / 1e3, 1)
        out["push_%s_%s_gbs" % (name, who)] = round(NB / timed(sm_copy(remote_ptr, local.data_ptr(), ctas), active) / 1e3, 1)
out["local_copy_gbs"] = round(NB / timed(sm_copy(local.data_ptr(), mine, 1184), everyone) / 1e3, 1)
for who, active in (("one_rank", one), ("all_ranks", everyone)):
    out["copy_engine_pull_%s_gbs" % who] = round(NB / timed(lambda: local.copy_(remote, non_blocking=True), active) / 1e3, 1)
    out["copy_engine_push_%s_gbs" % who] = round(NB / timed(lambda: remote.copy_(local, non_blocking=True), active) / 1e3, 1)


# pull and push at the same time on two streams (what the update kernel does: gradient loads + parameter stores)
s2 = torch.cuda.Stream()
half = NB // 2


def both():
    s2.wait_stream(torch.cuda.current_stream())
    L.check(lib.nb200_stream_copy_probe(C.c_void_p(local.data_ptr()), C.c_void_p(remote_ptr), C.c_uint64(half), C.c_uint32(296), L.stream()), "pull")
    with torch.cuda.stream(s2):
        L.check(lib.nb200_stream_copy_probe(C.c_void_p(remote_ptr + half), C.c_void_p(local.data_ptr() + half), C.c_uint64(half), C.c_uint32(296), L.stream()), "push")
    torch.cuda.current_stream().wait_stream(s2)


us = timed(both, everyone)
out["pull_plus_push_all_ranks_gbs_per_direction"] = round(NB / us / 1e3, 1)     # each link direction carries half + half = NB
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
