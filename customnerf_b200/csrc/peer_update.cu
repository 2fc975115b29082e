// peer_update.cu -- the optimiser update of the ray-sharded multi-GPU step as ONE kernel over NVLink peer memory.
//
// Path: every rank holds a replica of [hash table | MLPs] and, after its backward pass, a
// gradient of its own rays; the reference's (dead) DDP wrap would all-reduce that gradient (nerf/utils_init_nerf.py:
// 76-78) and then run torch.optim.Adam on every rank (main.py:182).  The NCCL form of that is all-reduce(49 MB) followed
// by a full Adam sweep per rank.  Here the two are one kernel, k_peer_reduce_adam_bcast:
//
//   * the flat parameter vector is cut into `world` slices; rank r OWNS slice r;
//   * reduce:    the owner reads slice r of every rank's gradient straight out of the peers' HBM (NVLink loads from
//                cudaIpc-mapped buffers) and sums them in rank order 0..W-1 (a fixed order: deterministic);
//   * Adam:      the owner updates slice r (the moments exist only where they are owned: 1/W of the Adam traffic);
//   * broadcast: the owner stores the new parameters into every rank's replica (NVLink stores);
//   * the gradient is reset locally once the peers have read it.
//
// Per rank and step that moves (W-1)/W of the vector in and the same out over NVLink -- the volume of a ring all-reduce
// -- but in two hops instead of 2(W-1), with the optimiser's arithmetic riding under the transfers, no staging copies,
// and the Adam sweep cut to 1/W.  Inbound (gradient reads) and outbound (parameter stores) share no link direction.
//
// Synchronisation is per CTA, the way flag barriers are done over peer memory: CTA b of every rank handles the same
// relative offsets of its own slice, so it only ever has to agree with CTA b of the other ranks.  Each CTA runs a start
// barrier (every peer's backward has finished: its update kernel is running) and an end barrier (every peer has read my
// gradient and delivered my parameters); a barrier is one st.release.sys per peer + one ld.acquire.sys spin per peer on
// monotonically increasing epoch words.  CTAs are dispatched in index order, so the lowest unfinished index on any rank
// is always resident and the barriers cannot dead-lock; a spin that lasts longer than kSpinTimeoutNs (a peer died)
// raises a status word instead of hanging the GPU.
#include "adam.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr unsigned long long kSpinTimeoutNs = 10000000000ull;    // 10 s

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// signal words of one rank: [phase 3][CTA grid][writer NB200_PEER_MAX]; phases 0 / 1 are the update kernel's start / end
// barriers, phase 2 (CTA 0 only) belongs to the standalone rank barrier
__device__ __forceinline__ uint32_t *signal_slot(uint32_t *base, uint32_t grid, uint32_t phase, uint32_t cta, uint32_t writer) {
    return base + ((size_t)(phase * grid + cta) * NB200_PEER_MAX + writer);
}

// Release / acquire across the CTA: bar.sync orders every thread's earlier peer stores before the signalling threads'
// st.release.sys (causality order is transitive through the barrier), so ONE system-scope release per peer publishes the
// whole CTA's stores -- a membar.sys in every thread costs tens of microseconds per barrier and buys nothing.
template <int W>
__device__ __forceinline__ void peer_barrier(const nb200_peer_plan &pl, uint32_t phase, uint32_t epoch) {
    __syncthreads();
    if (threadIdx.x < W) {
        const uint32_t peer = threadIdx.x;
        st_release_sys(signal_slot(pl.signals[peer], pl.grid, phase, blockIdx.x, pl.rank), epoch);
        const uint32_t *mine = signal_slot(pl.signals[pl.rank], pl.grid, phase, blockIdx.x, peer);
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (globaltimer_ns() - t0 > kSpinTimeoutNs) { atomicOr(pl.status, 1u + phase); break; }
        }
    }
    __syncthreads();
}

constexpr int kPeerThreads = 256;

// float4 groups of a slice that one CTA handles: a contiguous chunk, a multiple of the CTA width
__host__ __device__ __forceinline__ uint64_t peer_chunk(uint64_t per, uint32_t grid) {
    const uint64_t c = (per + grid - 1) / grid;
    return (c + kPeerThreads - 1) / kPeerThreads * kPeerThreads;
}

// NVSwitch multicast (NVLS) forms: one load returns the sum over every rank's copy, reduced inside the switch; one store
// lands in every rank's copy.  Per rank and step the wire then carries n (1 + 1/W) floats each way instead of 2 n (W-1)/W.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4 *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4 *mc, const float4 &v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// W ranks, U float4 groups per thread and trip (U * W gradient loads in flight per thread before the first add);
// MC: reduce and broadcast through the multicast mappings (plan.mc_grads / mc_params) instead of W-1 peer loads / stores
template <int W, int U, bool MC>
__global__ void __launch_bounds__(512, 1)
k_peer_reduce_adam_bcast(const nb200_peer_plan pl) {
    const uint32_t epoch = pl.epoch[blockIdx.x] + 1u;
    const uint64_t n4 = pl.n / 4, split4 = pl.split / 4;
    const uint64_t per = (n4 + W - 1) / W;                               // float4 groups per slice
    const uint64_t chunk = peer_chunk(per, gridDim.x);
    const uint64_t rel0 = (uint64_t)blockIdx.x * chunk;                  // this CTA's chunk, relative to a slice start
    const uint64_t lo = (uint64_t)pl.rank * per < n4 ? (uint64_t)pl.rank * per : n4;
    const uint64_t hi = lo + per < n4 ? lo + per : n4;
    const uint64_t c_lo = lo + rel0 < hi ? lo + rel0 : hi, c_hi = c_lo + chunk < hi ? c_lo + chunk : hi;
    const AdamHyper *hy = (const AdamHyper *)pl.hyper;
    const AdamConst h0(hy[0]), h1(hy[1]);

    peer_barrier<W>(pl, 0, epoch);                                       // every rank's gradient is complete

    // dynamic loss scaling (adam.cuh): a non-finite gradient on ANY rank skips the step on EVERY rank -- each CTA ORs the
    // ranks' found-inf flags of this iteration (final since the start barrier; all ranks run the same iteration)
    bool skip = false;
    if (pl.scalers[pl.rank]) {
        const uint32_t slot = kScalerFlag0 + (pl.scalers[pl.rank][kScalerIter] & 1u);
#pragma unroll
        for (int q = 0; q < W; q++) skip |= (*(volatile const uint32_t *)(pl.scalers[q] + slot) & kScalerInfBit) != 0u;
    }

    float4 *const p_own = (float4 *)pl.params[pl.rank];
    float4 *const m_own = (float4 *)pl.exp_avg, *const v_own = (float4 *)pl.exp_avg_sq;
    for (uint64_t base = c_lo + threadIdx.x; base < c_hi && !skip; base += (uint64_t)blockDim.x * U) {
        float4 gq[U][MC ? 1 : W], p[U], m[U], v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = base + (uint64_t)u * blockDim.x;
            if (i < c_hi) {
                if (MC) {
                    gq[u][0] = multimem_ld_reduce_add((const float4 *)pl.mc_grads + i);           // summed in the switch
                } else {
#pragma unroll
                    for (int q = 0; q < W; q++) gq[u][q] = __ldcg((const float4 *)pl.grads[q] + i);   // W-1 of W over NVLink
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = base + (uint64_t)u * blockDim.x;
            if (i < c_hi) { p[u] = p_own[i]; m[u] = __ldcs(m_own + i); v[u] = __ldcs(v_own + i); }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = base + (uint64_t)u * blockDim.x;
            if (i < c_hi) {
                float4 g = gq[u][0];
#pragma unroll
                for (int q = 1; q < (MC ? 1 : W); q++) { g.x += gq[u][q].x; g.y += gq[u][q].y; g.z += gq[u][q].z; g.w += gq[u][q].w; }
                const AdamConst &h = i < split4 ? h0 : h1;
                adam1(p[u].x, g.x, m[u].x, v[u].x, h); adam1(p[u].y, g.y, m[u].y, v[u].y, h);
                adam1(p[u].z, g.z, m[u].z, v[u].z, h); adam1(p[u].w, g.w, m[u].w, v[u].w, h);
                if (MC) {
                    multimem_st((float4 *)pl.mc_params + i, p[u]);                                // every replica, one store
                } else {
#pragma unroll
                    for (int q = 0; q < W; q++) ((float4 *)pl.params[q])[i] = p[u];               // every replica
                }
                __stcs(m_own + i, m[u]); __stcs(v_own + i, v[u]);
            }
        }
    }

    peer_barrier<W>(pl, 1, epoch);      // CTA b of every peer has read my gradient and delivered its parameters to me

    // reset the local gradient: exactly the groups CTA b of some rank has read (the same relative chunk of every slice)
    float4 *const g_own = (float4 *)pl.grads[pl.rank];
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int s = 0; s < W; s++) {
        const uint64_t slo = (uint64_t)s * per < n4 ? (uint64_t)s * per : n4;
        const uint64_t shi = slo + per < n4 ? slo + per : n4;
        const uint64_t z_lo = slo + rel0 < shi ? slo + rel0 : shi, z_hi = z_lo + chunk < shi ? z_lo + chunk : shi;
        for (uint64_t i = z_lo + threadIdx.x; i < z_hi; i += blockDim.x) g_own[i] = zero;
    }
    if (threadIdx.x == 0) pl.epoch[blockIdx.x] = epoch;
}

// Standalone rank barrier: one CTA; returns when every rank's stream has reached its own call.  (bench.py aligns the ranks
// with it after the L2 flush that sits between timed steps, so that one rank's flush is not billed to a peer's step.)
__global__ void __launch_bounds__(32)
k_peer_rank_barrier(const nb200_peer_plan pl) {
    const uint32_t epoch = pl.epoch[pl.grid] + 1u;
    if (threadIdx.x < pl.world) {
        const uint32_t peer = threadIdx.x;
        st_release_sys(signal_slot(pl.signals[peer], pl.grid, 2, 0, pl.rank), epoch);
        const uint32_t *mine = signal_slot(pl.signals[pl.rank], pl.grid, 2, 0, peer);
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (globaltimer_ns() - t0 > kSpinTimeoutNs) { atomicOr(pl.status, 4u); break; }
        }
    }
    __syncwarp();
    if (threadIdx.x == 0) pl.epoch[pl.grid] = epoch;
}

template <int W>
int launch_peer(const nb200_peer_plan *pl, cudaStream_t st) {
    static int unroll_env = -1;         // float4 groups per thread and trip (tuning knob; default: W * U = 8 loads in flight)
    if (unroll_env < 0) { const char *e = getenv("NB200_PEER_UNROLL"); unroll_env = e ? atoi(e) : 0; }
    static int threads_env = -1;        // threads per CTA (tuning knob): a thin CTA leaves the SM's registers to a co-running kernel
    if (threads_env < 0) { const char *e = getenv("NB200_PEER_THREADS"); threads_env = e ? atoi(e) : 0; }
    const int want_t = threads_env > 0 ? threads_env : (int)pl->threads;
    const int T = (want_t == 64 || want_t == 128 || want_t == 512) ? want_t : kPeerThreads;
    const bool mc = pl->mc_grads && pl->mc_params;
    const int want_u = unroll_env > 0 ? unroll_env : (int)pl->unroll;
    const int U = want_u > 0 ? want_u : (mc || W <= 2 ? 4 : W <= 4 ? 2 : 1);
    if (mc) {
        if (U >= 4) k_peer_reduce_adam_bcast<W, 4, true><<<pl->grid, T, 0, st>>>(*pl);
        else if (U >= 2) k_peer_reduce_adam_bcast<W, 2, true><<<pl->grid, T, 0, st>>>(*pl);
        else k_peer_reduce_adam_bcast<W, 1, true><<<pl->grid, T, 0, st>>>(*pl);
    } else {
        bool done = false;
        if constexpr (W <= 4) {             // W x 4 float4 gradient registers only fit for few ranks
            if (U >= 4) { k_peer_reduce_adam_bcast<W, 4, false><<<pl->grid, T, 0, st>>>(*pl); done = true; }
        }
        if (!done) {
            if (U >= 2) k_peer_reduce_adam_bcast<W, 2, false><<<pl->grid, T, 0, st>>>(*pl);
            else k_peer_reduce_adam_bcast<W, 1, false><<<pl->grid, T, 0, st>>>(*pl);
        }
    }
    NB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" {

uint32_t nb200_peer_plan_bytes(void) { return (uint32_t)sizeof(nb200_peer_plan); }
uint32_t nb200_peer_handle_bytes(void) { return (uint32_t)sizeof(cudaIpcMemHandle_t); }
uint64_t nb200_peer_signal_bytes(uint32_t grid) { return (uint64_t)3 * grid * NB200_PEER_MAX * sizeof(uint32_t); }

// CTAs of the update kernel for n parameters on `world` ranks with `sms` SMs each: host arithmetic only, so that every
// rank derives the same grid (the signal slots are per CTA)
uint32_t nb200_peer_grid(uint64_t n, uint32_t world, uint32_t sms) {
    if (world == 0) return 0;
    static int per_sm_env = -1;         // CTAs per SM (tuning knob); every CTA must be resident: the barriers are per CTA
    if (per_sm_env < 0) { const char *e = getenv("NB200_PEER_CTAS_PER_SM"); per_sm_env = e ? atoi(e) : 0; }
    static int grid_env = -1;           // absolute CTA count (tuning knob): a NARROW grid leaves most SMs to a co-running kernel
    if (grid_env < 0) { const char *e = getenv("NB200_PEER_GRID"); grid_env = e ? atoi(e) : 0; }
    const uint64_t per = (n / 4 + world - 1) / world;
    const uint64_t want = (per + kPeerThreads - 1) / kPeerThreads;
    const uint64_t cap = grid_env > 0 ? (uint64_t)grid_env : (uint64_t)sms * (per_sm_env > 0 ? per_sm_env : 2);
    const uint64_t g = want < cap ? want : cap;
    return (uint32_t)(g ? g : 1);
}

// elements [lo, hi) of the flat vector that `rank` owns (host arithmetic; the kernel uses the same formula)
void nb200_peer_slice(uint64_t n, uint32_t world, uint32_t rank, uint64_t *lo, uint64_t *hi) {
    const uint64_t n4 = n / 4, per = world ? (n4 + world - 1) / world : 0;
    const uint64_t a = (uint64_t)rank * per < n4 ? (uint64_t)rank * per : n4;
    const uint64_t b = a + per < n4 ? a + per : n4;
    if (lo) *lo = a * 4;
    if (hi) *hi = b * 4;
}

// Peer-visible device memory.  cudaIpc handles exist only for whole cudaMalloc allocations, so this is the one place
// where the library allocates: the caller frees with nb200_peer_free.
int nb200_peer_alloc(void **ptr, uint64_t bytes) {
    if (!ptr || bytes == 0) return NB200_E_BAD_ARG;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    return (int)e;
}
int nb200_peer_free(void *ptr) { return ptr ? (int)cudaFree(ptr) : 0; }
int nb200_peer_export(void *ptr, void *handle) {
    if (!ptr || !handle) return NB200_E_BAD_ARG;
    return (int)cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle, ptr);
}
int nb200_peer_import(const void *handle, void **ptr) {
    if (!ptr || !handle) return NB200_E_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
int nb200_peer_release(void *ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : 0; }

int nb200_peer_rank_barrier(const nb200_peer_plan *pl, void *stream) {
    if (!pl || pl->world == 0 || pl->world > NB200_PEER_MAX || pl->rank >= pl->world || pl->grid == 0 || !pl->epoch || !pl->status)
        return NB200_E_BAD_ARG;
    for (uint32_t q = 0; q < pl->world; q++) if (!pl->signals[q]) return NB200_E_BAD_ARG;
    k_peer_rank_barrier<<<1, 32, 0, nb_stream(stream)>>>(*pl);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_peer_reduce_adam_bcast(const nb200_peer_plan *pl, void *stream) {
    if (!pl || pl->world == 0 || pl->world > NB200_PEER_MAX || pl->rank >= pl->world || pl->grid == 0) return NB200_E_BAD_ARG;
    if ((pl->n & 3u) || (pl->split & 3u) || pl->split > pl->n) return NB200_E_BAD_ARG;
    if (!pl->exp_avg || !pl->exp_avg_sq || !pl->hyper || !pl->epoch || !pl->status) return NB200_E_BAD_ARG;
    uintptr_t al = reinterpret_cast<uintptr_t>(pl->exp_avg) | reinterpret_cast<uintptr_t>(pl->exp_avg_sq);
    for (uint32_t q = 0; q < pl->world; q++) {
        if (!pl->params[q] || !pl->grads[q] || !pl->signals[q]) return NB200_E_BAD_ARG;
        al |= reinterpret_cast<uintptr_t>(pl->params[q]) | reinterpret_cast<uintptr_t>(pl->grads[q]);
    }
    if ((pl->mc_grads == nullptr) != (pl->mc_params == nullptr)) return NB200_E_BAD_ARG;
    for (uint32_t q = 0; q < pl->world; q++)
        if ((pl->scalers[q] == nullptr) != (pl->scalers[pl->rank] == nullptr)) return NB200_E_BAD_ARG;
    al |= reinterpret_cast<uintptr_t>(pl->mc_grads) | reinterpret_cast<uintptr_t>(pl->mc_params);
    if (al & 15u) return NB200_E_BAD_ARG;
    if (pl->n == 0) return 0;
    cudaStream_t st = nb_stream(stream);
    switch (pl->world) {
        case 1: return launch_peer<1>(pl, st);
        case 2: return launch_peer<2>(pl, st);
        case 3: return launch_peer<3>(pl, st);
        case 4: return launch_peer<4>(pl, st);
        case 5: return launch_peer<5>(pl, st);
        case 6: return launch_peer<6>(pl, st);
        case 7: return launch_peer<7>(pl, st);
        default: return launch_peer<8>(pl, st);
    }
}

}  // extern "C"
