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
