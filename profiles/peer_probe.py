#!/usr/bin/env python
"""The peer-memory update kernel alone (torchrun, one process per GPU): microseconds per launch and NVLink GB/s each way
for the bench's parameter count.  Knobs: NB200_PEER_UNROLL, NB200_PEER_CTAS_PER_SM."""
import ctypes as C
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
n = int(os.environ.get("PROBE_N", 12262256))
mc = os.environ.get("PROBE_NVLS", "0") == "1"
peer = parallel.PeerMemory(n, dev, multicast=mc)
m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
hyper = torch.zeros(16, device=dev)
sched = torch.tensor([5e-3, 5e-4, 0.9, 0.99, 1e-15, 1.0 / 128.0, 1.0, 0.0], device=dev)
step = torch.zeros(1, dtype=torch.int32, device=dev)
L.check(lib.nb200_adam_hyper(L.ptr(step), L.ptr(sched), L.ptr(hyper), L.stream()), "hyper")
plan = peer.plan(n - 22528, m, v, hyper)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(reps, do_flush):
    tot = 0.0
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.check(lib.nb200_peer_reduce_adam_bcast(C.byref(plan), L.stream()), "peer")
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


run(5, False)
dist.barrier()
res = []
for fl in (False, True):
    us = run(30, fl)
    t = torch.tensor([us], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res.append(float(t[0]))
    dist.barrier()
if rank == 0:
    wire = n * 4 * (world - 1) / world
    print("world %d nvls %d unroll %s ctas/sm %s grid %d: %.1f us back-to-back (%.0f GB/s each way), %.1f us after an L2 flush; status %d"
          % (world, int(mc), os.environ.get("NB200_PEER_UNROLL", "default"), os.environ.get("NB200_PEER_CTAS_PER_SM", "default"),
             peer.grid, res[0], wire / res[0] / 1e3, res[1], int(peer.status[0])), flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
