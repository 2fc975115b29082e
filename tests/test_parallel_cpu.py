"""Ray-sharded data parallelism on CPU: gloo backend, world size 2 (the N > 1 host logic of customnerf_b200/parallel.py
and the loss normalisation the sharded train step relies on).  The GPU path differs only in the backend (NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from customnerf_b200 import parallel
    r, lr, w = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and dist.is_initialized()

    # a toy "field": table + mlp parameters, per-ray loss = (w . x_ray - y_ray)^2 averaged over ALL rays
    g = torch.Generator().manual_seed(0)
    n_rays = 1001
    X, Y = torch.randn(n_rays, 6, generator=g), torch.randn(n_rays, generator=g)
    table = torch.nn.Parameter(torch.randn(4, generator=g))
    mlp = torch.nn.Parameter(torch.randn(2, generator=g))

    def local_loss(idx):
        pred = X[idx] @ torch.cat([table, mlp])
        return ((pred - Y[idx]) ** 2).sum() / n_rays           # local sum / N_total, as TrainStep / nb200_mse_loss_grad

    for mode in ("interleaved", "contiguous"):
        idx = parallel.shard_rays(n_rays, rank, world, mode)
        counts = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([idx.numel()]))
        assert sum(int(c) for c in counts) == n_rays            # the shards partition the rays
        table.grad = mlp.grad = None
        local_loss(idx).backward()
        sync = parallel.FlatGradSync([table, mlp])
        sync()
        assert table.grad.data_ptr() == sync.flat.data_ptr()    # gradients alias the flat all-reduce buffer
        got = sync.flat.clone()
        # single-process gradient on the whole batch
        t2, m2 = table.detach().clone().requires_grad_(), mlp.detach().clone().requires_grad_()
        full = ((X @ torch.cat([t2, m2]) - Y) ** 2).mean()
        full.backward()
        want = torch.cat([t2.grad, m2.grad])
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    # a parameter without gradient contributes zeros, not stale data
    table.grad, mlp.grad = None, torch.ones(2) * (rank + 1)
    sync = parallel.FlatGradSync([table, mlp])
    sync()
    np.testing.assert_allclose(sync.flat.numpy(), [0, 0, 0, 0, 3, 3])
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_ray_sharded_gradients_match_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_shard_rays_partitions():
    from customnerf_b200 import parallel
    for n, w in ((14910, 8), (7, 2), (5, 8)):
        for mode in ("interleaved", "contiguous"):
            parts = [parallel.shard_rays(n, r, w, mode) for r in range(w)]
            allidx = torch.cat(parts).sort().values
            assert torch.equal(allidx, torch.arange(n)), (n, w, mode)


def test_peer_slices_partition_and_match_the_library():
    """parallel.peer_slice == nb200_peer_slice (host arithmetic of csrc/peer_update.cu); slices tile [0, n)"""
    import ctypes as C
    from customnerf_b200 import parallel, _lib
    lib = _lib.lib()
    lib.nb200_peer_grid.restype = C.c_uint32
    for n in (12262256, 4, 8, 1000036, 79250560 + 22528):
        for w in range(1, 9):
            end = 0
            for r in range(w):
                lo, hi = C.c_uint64(), C.c_uint64()
                lib.nb200_peer_slice(C.c_uint64(n), C.c_uint32(w), C.c_uint32(r), C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == parallel.peer_slice(n, w, r)
                assert lo.value == end and hi.value >= lo.value and lo.value % 4 == 0
                end = hi.value
            assert end == n
            g = lib.nb200_peer_grid(C.c_uint64(n), C.c_uint32(w), C.c_uint32(148))
            assert 1 <= g <= 148 * 8


def _peer_schedule_worker(rank, world, port, out_dir):
    """the schedule of k_peer_reduce_adam_bcast restated with gloo: owner sums slice r of every rank's gradient in rank
    order, Adam on the slice, new parameters to every replica == all-reduce + Adam on every rank"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from customnerf_b200 import parallel
    parallel.init_from_env(backend="gloo")
    n = 1036
    g0 = torch.Generator().manual_seed(1)
    p_repl = torch.randn(n, generator=g0)
    p_ref = p_repl.clone().requires_grad_()
    opt = torch.optim.Adam([p_ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    lo, hi = parallel.peer_slice(n, world, rank)
    m, v = torch.zeros(n), torch.zeros(n)
    gr = torch.Generator().manual_seed(10 + rank)
    for t in range(1, 4):
        grad = torch.randn(n, generator=gr)
        # reference: all-reduce + Adam everywhere
        total = grad.clone()
        dist.all_reduce(total)
        p_ref.grad = total
        opt.step()
        # peer schedule: every rank exposes its gradient; the owner reduces its slice in rank order
        everyone = [torch.zeros(n) for _ in range(world)]
        dist.all_gather(everyone, grad)
        gs = sum(e[lo:hi] for e in everyone[1:]) + everyone[0][lo:hi] if world > 1 else everyone[0][lo:hi]
        m[lo:hi] = m[lo:hi] + 0.1 * (gs - m[lo:hi])
        v[lo:hi] = 0.99 * v[lo:hi] + 0.01 * gs * gs
        new = p_repl[lo:hi] - (1e-2 / (1 - 0.9 ** t)) * m[lo:hi] / (v[lo:hi].sqrt() / (1 - 0.99 ** t) ** 0.5 + 1e-15)
        # broadcast: slices may be ragged, so ship (lo, hi, values)
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, new))
        for a, b, vals in parts:
            p_repl[a:b] = vals
        np.testing.assert_allclose(p_repl.numpy(), p_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert float(m[:lo].abs().sum()) == 0.0 and float(m[hi:].abs().sum()) == 0.0     # moments only where owned
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_peer_update_schedule_equals_allreduce_plus_adam(tmp_path):
    port = _free_port()
    mp.spawn(_peer_schedule_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def _gather_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from customnerf_b200 import parallel
    parallel.init_from_env(backend="gloo")
    for n in (1036, 8, 4):                      # ragged last slice; fewer float4 groups than ranks
        full = torch.arange(n, dtype=torch.float32) + 1
        lo, hi = parallel.peer_slice(n, world, rank)
        vec = torch.zeros(n)
        vec[lo:hi] = full[lo:hi]                 # every rank holds only what it owns
        parallel.gather_owned_slices(vec, world, rank)
        assert torch.equal(vec, full), (n, rank)
    # the split update (FusedTrainStep.update_ranges): rank r owns the r-th 1/world of EVERY range; views are gathered in place
    n, cut = 2072, 1040
    full = torch.arange(n, dtype=torch.float32) + 1
    vec = torch.zeros(n)
    for a, b in ((cut, n), (0, cut)):
        lo, hi = parallel.peer_slice(b - a, world, rank)
        vec[a + lo:a + hi] = full[a + lo:a + hi]
    for a, b in ((cut, n), (0, cut)):
        parallel.gather_owned_slices(vec[a:b], world, rank)
    assert torch.equal(vec, full), rank
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_gather_owned_slices_rebuilds_the_full_vector(tmp_path):
    port = _free_port()
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_peer_plan_is_filled_from_the_mapped_bases():
    """PeerMemory.plan() on a stand-in object (no GPU, no IPC): pointers of every rank = base, base + 4n, base + 8n"""
    import ctypes as C
    from customnerf_b200 import parallel, _lib
    pm = object.__new__(parallel.PeerMemory)
    pm.C, pm.lib = C, _lib.lib()
    pm.lib.nb200_peer_plan_bytes.restype = C.c_uint32
    pm.world, pm.rank, pm.grid, pm.n = 4, 2, 296, 1000
    pm.sig_off, pm.mc_base = 2 * pm.n * 4, 0
    pm.bases = [0x10000000 * (r + 1) for r in range(4)]
    pm.epoch, pm.status = torch.zeros(pm.grid + 1, dtype=torch.int32), torch.zeros(1, dtype=torch.int32)
    m, v, h = torch.zeros(pm.n), torch.zeros(pm.n), torch.zeros(16)
    p = pm.plan(600, m, v, h)
    assert (p.world, p.rank, p.grid, p.n, p.split) == (4, 2, 296, 1000, 600)
    for r in range(4):
        assert (p.params[r], p.grads[r], p.signals[r]) == (pm.bases[r], pm.bases[r] + 4000, pm.bases[r] + 8000)
    assert p.params[4] is None and p.mc_params is None and p.mc_grads is None
    assert (p.exp_avg, p.exp_avg_sq, p.hyper, p.epoch, p.status) == (m.data_ptr(), v.data_ptr(), h.data_ptr(),
                                                                      pm.epoch.data_ptr(), pm.status.data_ptr())
    pm.mc_base = 0x7000000000
    q = pm.plan(600, m, v, h, status=torch.zeros(1, dtype=torch.int32))
    assert (q.mc_params, q.mc_grads) == (pm.mc_base, pm.mc_base + 4000) and q.status != p.status
    # launch shape of the narrow update and the loss-scaler words of every rank
    pm.threads, pm.unroll, pm.scaler_off = 512, 2, 9000
    z = pm.plan(600, m, v, h, use_scaler=True)
    assert (z.threads, z.unroll) == (512, 2)
    assert [z.scalers[r] for r in range(4)] == [b + 9000 for b in pm.bases] and z.scalers[4] is None
    assert p.scalers[0] is None and (p.threads, p.unroll) == (0, 0)


def test_stats_block_reports_kernel_and_peer_status():
    """FusedTrainStep._parse_stats: words 4 (peer update timed out) and 7 (a field kernel's mbarrier wait timed out) raise"""
    import pytest
    from customnerf_b200 import fused_trainer
    fs = object.__new__(fused_trainer.FusedTrainStep)
    s = torch.zeros(8, dtype=torch.int32)
    s[0], s[2] = 1234, 1200
    s[3:4] = torch.tensor([0.25]).view(torch.int32)
    assert fs._parse_stats(s) == (0.25, 1234, 1200)
    for word, what in ((4, "peer-memory update"), (7, "field kernel")):
        t = s.clone()
        t[word] = 2
        with pytest.raises(RuntimeError, match=what):
            fs._parse_stats(t)
