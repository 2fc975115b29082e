"""Ray-sharded data parallelism: one process per GPU, a full replica of the hash table /
MLPs / occupancy grid on every rank, rays split across ranks, and ONE all-reduce(sum) per step over a single flat
buffer holding every gradient (hash table + the three MLPs).  Nothing else is exchanged.

The reference's own DDP wrap is dead code (nerf/utils_init_nerf.py:76-78; init_process_group is never called).
"""
import ctypes as _C
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """(rank, local_rank, world_size); initialises torch.distributed when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend)
    return rank, local_rank, world


def shard_rays(n_rays, rank, world_size, mode="interleaved"):
    """Indices of the rays owned by ``rank``.

    interleaved (default): ray i -> rank i % world.  Neighbouring pixels have similar sample counts, so a strided
    split balances the per-rank sample totals far better than contiguous image rows (SURVEY.md 8(e) caveat)."""
    if mode == "interleaved":
        return torch.arange(min(rank, n_rays), n_rays, world_size)
    per = (n_rays + world_size - 1) // world_size
    return torch.arange(min(n_rays, rank * per), min(n_rays, (rank + 1) * per))


class FlatGradSync:
    """All-reduce every gradient as one flat fp32 buffer (one collective per step)."""

    def __init__(self, params, group=None):
        self.params = list(params)
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def __call__(self, params=None):
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        for p, v in zip(self.params, self.views):
            p.grad = v            # gradients now alias the flat buffer: later steps accumulate straight into it


# ------------------------------------------------------------------------------------------------ NVLink peer-memory update
def peer_slice(n, world_size, rank):
    """elements [lo, hi) of a flat n-element vector that ``rank`` owns in the peer-memory update (the arithmetic of
    nb200_peer_slice / k_peer_reduce_adam_bcast: float4 groups, ceil(n4 / world) per rank)"""
    n4 = n // 4
    per = (n4 + world_size - 1) // world_size
    lo = min(rank * per, n4)
    return lo * 4, min(lo + per, n4) * 4


def gather_owned_slices(vec, world_size, rank, group=None):
    """In place: every rank's copy of the flat vector ``vec`` receives slice r from rank r (the slices of peer_slice).
    The peer-memory update keeps Adam's moments only where they are owned; a checkpoint wants the whole vectors."""
    if world_size == 1:
        return vec
    n = vec.numel()
    per = max(peer_slice(n, world_size, r)[1] - peer_slice(n, world_size, r)[0] for r in range(world_size))
    lo, hi = peer_slice(n, world_size, rank)
    mine = torch.zeros(per, dtype=vec.dtype, device=vec.device)
    mine[:hi - lo].copy_(vec[lo:hi])
    parts = [torch.empty_like(mine) for _ in range(world_size)]
    dist.all_gather(parts, mine, group=group)
    for r, part in enumerate(parts):
        a, b = peer_slice(n, world_size, r)
        vec[a:b].copy_(part[:b - a])
    return vec


class _DevArray:
    """a raw device allocation seen through __cuda_array_interface__ (torch.as_tensor wraps it without a copy)"""

    def __init__(self, ptr, numel, typestr, owner):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner         # keeps the allocation alive as long as a tensor made from this object lives


class PeerPlan(_C.Structure):
    """mirror of nb200_peer_plan (include/nerf_b200.h)"""
    _fields_ = [("world", _C.c_uint32), ("rank", _C.c_uint32), ("grid", _C.c_uint32), ("threads", _C.c_uint32),
                ("n", _C.c_uint64), ("split", _C.c_uint64), ("params", _C.c_void_p * 8), ("grads", _C.c_void_p * 8),
                ("signals", _C.c_void_p * 8), ("exp_avg", _C.c_void_p), ("exp_avg_sq", _C.c_void_p), ("hyper", _C.c_void_p),
                ("epoch", _C.c_void_p), ("status", _C.c_void_p), ("mc_params", _C.c_void_p), ("mc_grads", _C.c_void_p),
                ("unroll", _C.c_uint32), ("pad", _C.c_uint32), ("scalers", _C.c_void_p * 8)]


class PeerMemory:
    """The peer-visible buffers of the one-kernel multi-GPU update (csrc/peer_update.cu): on every rank ONE cudaMalloc
    allocation [params n | grads n | flag words], its cudaIpc handle exchanged through ``torch.distributed`` (plumbing
    only: the step itself contains no collective call) and the peers' allocations mapped into this process.

    ``params`` / ``grads``: fp32 [n] tensors over this rank's allocation -- FusedTrainStep places the model's parameters
    and the flat gradient there.  ``plan(...)`` fills the nb200_peer_plan the kernel takes.

    ``multicast=True``: the allocation comes from torch's symmetric-memory allocator instead (cuMem + an NVSwitch
    multicast object bound over every rank's copy; again plumbing only) and the kernel reduces / broadcasts through the
    multicast mapping (multimem.ld_reduce / multimem.st: the sum happens inside the switch).  Raises when the fabric has
    no multicast support."""

    def __init__(self, n_params, device, group=None, multicast=False, shape=(64, 512, 2)):
        """shape = (CTAs, threads per CTA, float4 groups in flight per thread and rank) of the update kernel, or None for the
        wide default (2 CTAs of 256 threads per SM).  The narrow default keeps the kernel on <= 64 SMs with deep loads: on its
        own it is ~8 % slower than the wide grid (126 vs 117 us at 2 GPUs), but the next step's ray march then runs BESIDE it
        (a wide grid and the march's single resident wave exclude each other): 0.483 vs 0.534 ms/step at 2 GPUs
        (profiles/r02l_peer_narrow_2gpu.txt)."""
        import ctypes as C
        from . import _lib as L
        self.lib, self.C, self.L = L.lib(), C, L
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise RuntimeError("PeerMemory: at most 8 ranks (one NVSwitch domain), got %d" % self.world)
        if n_params % 4:
            raise RuntimeError("PeerMemory: parameter count %d is not a multiple of 4" % n_params)
        self.n = int(n_params)
        self.device = torch.device(device)
        lib = self.lib
        lib.nb200_peer_grid.restype = C.c_uint32
        lib.nb200_peer_signal_bytes.restype = C.c_uint64
        lib.nb200_peer_handle_bytes.restype = C.c_uint32
        lib.nb200_peer_plan_bytes.restype = C.c_uint32
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.grid = int(lib.nb200_peer_grid(C.c_uint64(self.n), C.c_uint32(self.world), C.c_uint32(sms)))
        self.threads, self.unroll = 0, 0
        if shape is not None and self.world > 1 and not os.environ.get("NB200_PEER_GRID"):
            self.grid, self.threads, self.unroll = min(self.grid, int(shape[0])), int(shape[1]), int(shape[2])
        self.sig_off = 2 * self.n * 4
        # [params | grads | flag words | loss-scaler words (8 x 4 B, 64-byte slot)]: the peers read each other's found-inf flags
        self.scaler_off = self.sig_off + (int(lib.nb200_peer_signal_bytes(C.c_uint32(self.grid))) + 63) // 64 * 64
        self.bytes = self.scaler_off + 64
        self.base = C.c_void_p()
        self.imported = []
        self.mc_base = 0
        if multicast and self.world > 1:
            self._init_symmetric(group)
            return
        with torch.cuda.device(self.device):
            L.check(lib.nb200_peer_alloc(C.byref(self.base), C.c_uint64(self.bytes)), "peer_alloc")
            handle = (C.c_char * int(lib.nb200_peer_handle_bytes()))()
            L.check(lib.nb200_peer_export(self.base, handle), "peer_export")
            bases = [None] * self.world
            if self.world > 1:
                # every rank must have the same grid (same SM count) -- the flag slots are per CTA
                mine = (bytes(handle.raw), self.grid, self.n)
                got = [None] * self.world
                dist.all_gather_object(got, mine, group=group)
                for r, (h, g, n) in enumerate(got):
                    if (g, n) != (self.grid, self.n):
                        raise RuntimeError("PeerMemory: rank %d disagrees on the plan (grid %d / n %d vs %d / %d)"
                                           % (r, g, n, self.grid, self.n))
                    if r == self.rank:
                        continue
                    p = C.c_void_p()
                    buf = C.create_string_buffer(h, len(h))
                    L.check(lib.nb200_peer_import(buf, C.byref(p)), "peer_import (cudaIpcOpenMemHandle)")
                    self.imported.append(p)
                    bases[r] = p.value
            bases[self.rank] = self.base.value
        self.bases = bases
        self.params = torch.as_tensor(_DevArray(self.base.value, self.n, "<f4", self), device=self.device)
        self.grads = torch.as_tensor(_DevArray(self.base.value + 4 * self.n, self.n, "<f4", self), device=self.device)
        self.scaler = torch.as_tensor(_DevArray(self.base.value + self.scaler_off, 8, "<i4", self), device=self.device)
        self.epoch = torch.zeros(self.grid + 1, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)               # nobody signals into a buffer that is not mapped and zeroed yet

    def _init_symmetric(self, group):
        import torch.distributed._symmetric_memory as symm
        with torch.cuda.device(self.device):
            self._symm_buf = symm.empty(self.bytes // 4, dtype=torch.float32, device=self.device)
            self._symm_buf.zero_()
            torch.cuda.synchronize(self.device)
            self._symm = symm.rendezvous(self._symm_buf, group if group is not None else dist.group.WORLD)
            mc = int(self._symm.multicast_ptr or 0)
            if not mc:
                raise RuntimeError("PeerMemory(multicast=True): this fabric / driver exposes no NVSwitch multicast mapping")
            self.mc_base = mc
            self.bases = [int(p) for p in self._symm.buffer_ptrs]
            assert self.bases[self.rank] == self._symm_buf.data_ptr()
            self.params, self.grads = self._symm_buf[:self.n], self._symm_buf[self.n:2 * self.n]
            self.scaler = self._symm_buf[self.scaler_off // 4:self.scaler_off // 4 + 8].view(torch.int32)
            self.epoch = torch.zeros(self.grid + 1, dtype=torch.int32, device=self.device)
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
            got = [None] * self.world
            dist.all_gather_object(got, (self.grid, self.n), group=group)
            if any(g != (self.grid, self.n) for g in got):
                raise RuntimeError("PeerMemory: the ranks disagree on the plan: %r" % (got,))
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)

    def plan(self, n_table_params, exp_avg, exp_avg_sq, hyper, status=None, use_scaler=False):
        """use_scaler: the update is skipped on every rank when any rank's found-inf flag (``self.scaler``: the device-side
        dynamic loss scaler, csrc/adam.cuh) is up"""
        C = self.C
        p = PeerPlan()
        p.world, p.rank, p.grid, p.n, p.split = self.world, self.rank, self.grid, self.n, int(n_table_params)
        p.threads, p.unroll = getattr(self, "threads", 0), getattr(self, "unroll", 0)
        for r, b in enumerate(self.bases):
            p.params[r], p.grads[r], p.signals[r] = b, b + 4 * self.n, b + self.sig_off
        p.exp_avg, p.exp_avg_sq, p.hyper = exp_avg.data_ptr(), exp_avg_sq.data_ptr(), hyper.data_ptr()
        p.epoch = self.epoch.data_ptr()
        p.status = (status if status is not None else self.status).data_ptr()
        if self.mc_base:
            p.mc_params, p.mc_grads = self.mc_base, self.mc_base + 4 * self.n
        if use_scaler:
            for r, b in enumerate(self.bases):
                p.scalers[r] = b + self.scaler_off
        assert C.sizeof(p) == int(self.lib.nb200_peer_plan_bytes()), "nb200_peer_plan layout mismatch"
        return p

    def owned(self):
        return peer_slice(self.n, self.world, self.rank)

    def close(self):
        """unmap the peers' allocations and free this rank's (no tensor over ``params`` / ``grads`` may be used after)"""
        with torch.cuda.device(self.device):
            for p in self.imported:
                self.lib.nb200_peer_release(p)
            self.imported = []
            if self.base:
                self.lib.nb200_peer_free(self.base)
                self.base = self.C.c_void_p()
