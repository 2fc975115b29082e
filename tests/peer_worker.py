"""Worker of tests/test_gpu_peer_update.py (one process per GPU, launched by torch.distributed.run): the one-kernel
NVLink peer-memory update (csrc/peer_update.cu) against NCCL all-reduce + nb200_fused_adam, then a few sharded
FusedTrainStep steps in both modes.  Prints PEER_OK on rank 0 when everything matched."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from customnerf_b200 import parallel, _lib as L, trainer, fused_trainer, synthetic as syn  # noqa: E402


def adam_case(rank, world, dev, multicast=False):
    lib = L.lib()
    n, split = 1_000_036, 600_004            # not a multiple of world * 4 * 256: ragged last slice
    peer = parallel.PeerMemory(n, dev, multicast=multicast)
    g = torch.Generator(device=dev).manual_seed(7)             # same seed on every rank: identical initial parameters
    p0 = torch.randn(n, device=dev, generator=g)
    peer.params.copy_(p0)
    m_p, v_p = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    p_n, m_n, v_n = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    g_n = torch.zeros(n, device=dev)
    hyper = torch.zeros(16, device=dev)
    sched = torch.tensor([5e-3, 5e-4, 0.9, 0.99, 1e-15, 1.0 / 128.0, 0.1, 100.0], device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    plan = peer.plan(split, m_p, v_p, hyper)
    lo, hi = peer.owned()
    gr = torch.Generator(device=dev).manual_seed(100 + rank)   # a different gradient on every rank
    for it in range(5):
        grad = torch.randn(n, device=dev, generator=gr) * 128.0
        peer.grads.copy_(grad)
        g_n.copy_(grad)
        torch.cuda.synchronize()
        dist.barrier()
        L.check(lib.nb200_adam_hyper(L.ptr(step), L.ptr(sched), L.ptr(hyper), L.stream()), "adam_hyper")
        L.check(lib.nb200_peer_reduce_adam_bcast(C.byref(plan), L.stream()), "peer_reduce_adam_bcast")
        dist.all_reduce(g_n)
        L.check(lib.nb200_fused_adam(L.ptr(p_n), L.ptr(g_n), L.ptr(m_n), L.ptr(v_n), C.c_uint64(n), C.c_uint64(split),
                                     L.ptr(hyper), C.c_int(1), L.stream()), "fused_adam")
        torch.cuda.synchronize()
        dist.barrier()
        L.check(lib.nb200_peer_rank_barrier(C.byref(plan), L.stream()), "peer_rank_barrier")
        torch.cuda.synchronize()
        assert int(peer.status[0]) == 0, "barrier timed out"
        assert float(peer.grads.abs().max()) == 0.0, "the local gradient was not reset"
        if world == 2 and not multicast:          # a + b is commutative: the two-rank sum equals NCCL's bit for bit
            assert torch.equal(peer.params, p_n), (it, float((peer.params - p_n).abs().max()))
            assert torch.equal(m_p[lo:hi], m_n[lo:hi]) and torch.equal(v_p[lo:hi], v_n[lo:hi])
        else:                   # W > 2: summation order differs from NCCL's tree/ring -- fp32 rel 1e-5 on the update
            np.testing.assert_allclose(peer.params.cpu().numpy(), p_n.cpu().numpy(), rtol=0, atol=1e-6)
        # the moments outside the owned slice are never touched
        assert float(m_p[:lo].abs().max() if lo else 0.0) == 0.0 and float(m_p[hi:].abs().max() if hi < n else 0.0) == 0.0
        # every replica holds the same parameters, bit for bit
        mine = peer.params.clone()
        ref = mine.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(mine, ref), "replicas diverged"
    peer_params_final = peer.params.clone()
    del plan
    return peer, peer_params_final


def train_case(rank, world, dev):
    """sharded FusedTrainStep: peer update (also in its split form: two scatter launches, two update parts) vs NCCL all-reduce,
    same rays, no perturbation"""
    losses = {}
    finals = {}
    keep = []
    for mode in ("nccl", "peer", "peer_split"):
        model = trainer.build_scene_model(dev, log2_hashmap_size=15, desired_resolution=512, seed=3)
        with torch.no_grad():
            g = torch.Generator(device=dev).manual_seed(11)
            model.pos_en.embeddings.copy_(torch.rand(model.pos_en.embeddings.shape, device=dev, generator=g) - 0.5)
        o, d = syn.camera_rays(105, 142)
        idx = parallel.shard_rays(4096, rank, world) + 5000
        o, d = o[idx].contiguous().to(dev), d[idx].contiguous().to(dev)
        tgt = syn.bear_color(o.cpu() + d.cpu() * 1.5).to(dev)
        peer = parallel.PeerMemory(fused_trainer.flat_parameter_count(model), dev) if mode != "nccl" else None
        sync = (lambda flat: dist.all_reduce(flat)) if mode == "nccl" else None
        fs = fused_trainer.FusedTrainStep(model, o.shape[0], world_size=world, grad_sync=sync, peer=peer, perturb=False,
                                          use_graph=True, pipeline_update=(mode != "nccl"),
                                          split_level=10 if mode == "peer_split" else 0)
        assert bool(fs.split_level) == (mode == "peer_split")
        ls = []
        for it in range(6):
            fs.step(o, d, tgt)
            ls.append(fs.last_stats()[0])
        fs.flush()
        torch.cuda.synchronize()
        dist.barrier()
        losses[mode] = ls
        finals[mode] = fs.params_flat.clone()
        keep.append((fs, peer))
    a = np.array(losses["nccl"])
    assert a[-1] < a[0], "training did not reduce the loss"
    for mode in ("peer", "peer_split"):
        # the gradients come from fp32 atomics whose order differs run to run: losses agree to rel 1e-4 over 6 steps
        np.testing.assert_allclose(np.array(losses[mode]), a, rtol=1e-4, err_msg=mode)
        # Adam moves every touched parameter by ~lr whatever |g| is, so an entry whose gradient is rounding noise may step the
        # other way: bound the mean drift, not the maximum
        diff = float((finals["nccl"] - finals[mode]).abs().mean())
        assert diff < 1e-5, (mode, diff)
    return keep


def edit_case(rank, world, dev):
    """sharded FusedEditStep (LGIE editing step): peer update vs NCCL all-reduce, same rays, no perturbation"""
    from customnerf_b200 import fused_edit
    losses, keep = {}, []
    for mode in ("nccl", "peer"):
        model = trainer.build_scene_model(dev, log2_hashmap_size=15, desired_resolution=512, seed=3,
                                          opt=trainer.make_opt(train_conf=0.01, soft_mask=True, detach_bg=True))
        with torch.no_grad():
            g = torch.Generator(device=dev).manual_seed(11)
            model.pos_en.embeddings.copy_(torch.rand(model.pos_en.embeddings.shape, device=dev, generator=g) - 0.5)
            model.rgb_network.params[-16 * 64:].view(16, 64)[3] *= 12.0        # a mask head that splits fg / bg
        o, d = syn.camera_rays(105, 142)
        idx = parallel.shard_rays(4096, rank, world) + 5000
        o, d = o[idx].contiguous().to(dev), d[idx].contiguous().to(dev)
        tgt = syn.bear_color(o.cpu() + d.cpu() * 1.5).to(dev)
        inv = 1.0 / 4096

        def loss_fn(out):
            return inv * (((out["image"].reshape(-1, 3) - tgt) ** 2).sum() + ((out["fg"]["image"].reshape(-1, 3) - 0.5 * tgt) ** 2).sum() +
                          ((out["bg"]["image"].reshape(-1, 3) - 0.5 * tgt) ** 2).sum() + ((out["render_mask"].reshape(-1) - 0.5) ** 2).sum())
        peer = parallel.PeerMemory(fused_trainer.flat_parameter_count(model), dev) if mode == "peer" else None
        sync = (lambda flat: dist.all_reduce(flat)) if mode == "nccl" else None
        fs = fused_edit.FusedEditStep(model, o.shape[0], loss_fn, world_size=world, grad_sync=sync, peer=peer, perturb=False)
        ls = []
        for it in range(6):
            fs.step(o, d, tgt)
            ls.append(fs.last_stats()[0])
        torch.cuda.synchronize()
        dist.barrier()
        losses[mode] = ls
        keep.append((fs, peer))
    a, b = np.array(losses["nccl"]), np.array(losses["peer"])
    np.testing.assert_allclose(b, a, rtol=1e-3)
    assert a[-1] < a[0], "the editing step did not reduce its loss"
    return keep




def scaler_case(rank, world, dev):
    """GradScaler semantics across ranks: a non-finite gradient on ONE rank skips the optimiser step on EVERY rank (the found-inf
    bits travel through peer memory), the loss scale halves everywhere, the replicas stay bit-identical, the step count
    stays; sample-buffer growth is a joint decision (every rank sees the same maximum sample count)"""
    model = trainer.build_scene_model(dev, log2_hashmap_size=15, desired_resolution=512, seed=3)
    o, d = syn.camera_rays(105, 142)
    idx = parallel.shard_rays(4096, rank, world) + 5000
    o, d = o[idx].contiguous().to(dev), d[idx].contiguous().to(dev)
    tgt = syn.bear_color(o.cpu() + d.cpu() * 1.5).to(dev)
    peer = parallel.PeerMemory(fused_trainer.flat_parameter_count(model), dev)
    fs = fused_trainer.FusedTrainStep(model, o.shape[0], world_size=world, peer=peer, perturb=False, use_graph=True)
    for _ in range(2):
        fs.step(o, d, tgt)
    fs.last_stats()
    torch.cuda.synchronize(); dist.barrier()
    assert fs.scaler_state() == (128.0, 0, 2), fs.scaler_state()
    caps = [None] * world
    dist.all_gather_object(caps, fs.m_cap)
    assert len(set(caps)) == 1, "sample-buffer capacity differs across ranks: %r" % (caps,)
    peak = [None] * world
    dist.all_gather_object(peak, int(fs.stats_host[5]))
    assert len(set(peak)) == 1 and peak[0] > 0, "the published maximum sample count differs across ranks: %r" % (peak,)
    p0, m0 = fs.params_flat.clone(), fs.exp_avg.clone()
    bad = tgt.clone()
    if rank == world - 1:
        bad[:, 1] = float("inf")                    # only the last rank's batch is poisoned
    fs.step(o, d, bad)
    fs.last_stats()
    torch.cuda.synchronize(); dist.barrier()
    assert fs.scaler_state() == (64.0, 1, 2), (rank, fs.scaler_state())
    assert torch.equal(fs.params_flat, p0) and torch.equal(fs.exp_avg, m0), "rank %d took the step" % rank
    assert float(fs.grads_flat.abs().max()) == 0.0
    fs.step(o, d, tgt)
    fs.last_stats()
    torch.cuda.synchronize(); dist.barrier()
    assert fs.scaler_state() == (64.0, 1, 3) and not torch.equal(fs.params_flat, p0)
    mine = fs.params_flat.clone()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(mine, ref) and torch.isfinite(mine).all(), "replicas diverged after the skipped step"
    return fs, peer


def main():
    rank, local_rank, world = parallel.init_from_env()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    keep = [adam_case(rank, world, dev)]
    try:
        keep.append(adam_case(rank, world, dev, multicast=True))     # NVSwitch multicast form, where the fabric has it
        nvls = "ok"
    except RuntimeError as e:
        if "multicast" not in str(e):
            raise
        nvls = "unavailable"
    keep.append(train_case(rank, world, dev))
    keep.append(edit_case(rank, world, dev))
    keep.append(scaler_case(rank, world, dev))
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("PEER_OK world=%d nvls=%s" % (world, nvls), flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
