// optim.cu -- the table-wide passes of a train step, fused.
//
// The reference trains with torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15) under a GradScaler (main.py:182,
// nerf/utils_init_nerf.py:100,612-629): per step that is a zero-fill of every gradient (grid.py:83), an unscale
// pass, an inf check with a D2H sync, and the Adam pass -- four sweeps over the 12.2 M-entry hash table
// (~0.6 GB of traffic at BASELINE.json configs[1]).  Here it is ONE sweep: read p, g, m, v; write p, m, v and
// g = 0 (32 B per parameter), with the loss-scale reciprocal folded into the gradient load.
//
// Arithmetic follows torch's fused Adam (aten/src/ATen/native/cuda/fused_adam_utils.cuh, non-amsgrad, no weight
// decay): m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).
#include "adam.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

// n4 = number of float4 groups; elements [0, split) use group 0, [split, n) group 1.  split % 4 == 0 is required
// when vectorised (the launcher checks).
template <int U>
__global__ void __launch_bounds__(512)
k_fused_adam_deep(float4 *__restrict__ param, float4 *__restrict__ grad, float4 *__restrict__ exp_avg,
                  float4 *__restrict__ exp_avg_sq, uint64_t n4, uint64_t split4, const AdamHyper *__restrict__ hyper,
                  int zero_grad) {
    // the same sweep on a NARROW grid (few SMs, wide CTAs, U float4 groups of every vector in flight per thread): enough
    // requests in flight to pull the HBM stream from a fraction of the SMs, so that a latency-bound kernel of the next step
    // (the ray march) can run beside it on the rest -- the pipelined update of FusedTrainStep
    const AdamConst h0(hyper[0]), h1(hyper[1]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (hyper[0].skip != 0.0f) {
        if (zero_grad)
            for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += stride * U) {
        float4 p[U], g[U], m[U], v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = base + (uint64_t)u * stride;
            if (i < n4) { p[u] = param[i]; g[u] = grad[i]; m[u] = __ldcs(exp_avg + i); v[u] = __ldcs(exp_avg_sq + i); }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = base + (uint64_t)u * stride;
            if (i < n4) {
                const AdamConst &h = i < split4 ? h0 : h1;
                adam1(p[u].x, g[u].x, m[u].x, v[u].x, h); adam1(p[u].y, g[u].y, m[u].y, v[u].y, h);
                adam1(p[u].z, g[u].z, m[u].z, v[u].z, h); adam1(p[u].w, g[u].w, m[u].w, v[u].w, h);
                param[i] = p[u]; __stcs(exp_avg + i, m[u]); __stcs(exp_avg_sq + i, v[u]);
                if (zero_grad) grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_fused_adam(float4 *__restrict__ param, float4 *__restrict__ grad, float4 *__restrict__ exp_avg,
             float4 *__restrict__ exp_avg_sq, uint64_t n4, uint64_t split4, const AdamHyper *__restrict__ hyper,
             int zero_grad) {
    const AdamConst h0(hyper[0]), h1(hyper[1]);
    if (hyper[0].skip != 0.0f) {            // a non-finite gradient this step: no update, only the gradient reset
        if (zero_grad)
            for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x)
                grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const AdamConst &h = i < split4 ? h0 : h1;
        // the moments are touched once per step: streaming loads / stores keep them from displacing the table and its
        // gradient (both L2-resident between the encode kernels and this sweep)
        float4 p = param[i], g = grad[i], m = __ldcs(exp_avg + i), v = __ldcs(exp_avg_sq + i);
        adam1(p.x, g.x, m.x, v.x, h); adam1(p.y, g.y, m.y, v.y, h);
        adam1(p.z, g.z, m.z, v.z, h); adam1(p.w, g.w, m.w, v.w, h);
        param[i] = p; __stcs(exp_avg + i, m); __stcs(exp_avg_sq + i, v);
        if (zero_grad) grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(256)
k_fused_adam_tail(float *__restrict__ param, float *__restrict__ grad, float *__restrict__ exp_avg,
                  float *__restrict__ exp_avg_sq, uint64_t begin, uint64_t n, uint64_t split,
                  const AdamHyper *__restrict__ hyper, int zero_grad) {
    const uint64_t i = begin + threadIdx.x;
    if (i >= n) return;
    const AdamConst h(hyper[i < split ? 0 : 1]);
    if (hyper[0].skip != 0.0f) { if (zero_grad) grad[i] = 0.0f; return; }
    float p = param[i], g = grad[i], m = exp_avg[i], v = exp_avg_sq[i];
    adam1(p, g, m, v, h);
    param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
    if (zero_grad) grad[i] = 0.0f;
}

// Per-step hyper-parameters computed on the device so that a captured step never reads host memory that the host may
// already have advanced: t = ++(*step); lr_g = lr0_g * decay_base^min((t-1)/decay_iters, 1)  (the LambdaLR of
// main.py:189 evaluated before this optimiser step); bias corrections in double precision as torch computes them.
// sched = {lr0 group 0, lr0 group 1, beta1, beta2, eps, grad_scale, decay_base, decay_iters}
__global__ void k_adam_hyper(int32_t *__restrict__ step, const float *__restrict__ sched, AdamHyper *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int32_t t = *step + 1;
    *step = t;
    const double b1 = sched[2], b2 = sched[3];
    double decay = 1.0;
    if (sched[7] > 0.0f) decay = pow((double)sched[6], fmin((double)(t - 1) / (double)sched[7], 1.0));
    const float bc1 = (float)(1.0 - pow(b1, (double)t)), bc2s = (float)sqrt(1.0 - pow(b2, (double)t));
    for (int g = 0; g < 2; g++) {
        AdamHyper h;
        h.lr = (float)((double)sched[g] * decay);
        h.beta1 = sched[2]; h.beta2 = sched[3]; h.eps = sched[4];
        h.bc1 = bc1; h.bc2_sqrt = bc2s; h.grad_scale = sched[5]; h.skip = 0.0f;
        out[g] = h;
    }
}

// The same for a step under the device-side loss scaler (adam.cuh): t = *step + 1 WITHOUT committing it (the step count only
// advances when the update is taken: nb200_scaler_commit), grad_scale = 1 / scaler.scale, skip = this rank's found-inf flag
// (local_skip; the peer-memory update ORs every rank's flag itself and ignores it).
__global__ void k_adam_hyper_scaled(const int32_t *__restrict__ step, const float *__restrict__ sched, AdamHyper *__restrict__ out,
                                    uint32_t *__restrict__ scaler, int local_skip, const int32_t *__restrict__ samples) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (samples) atomicOr(scaler_flag(scaler), (uint32_t)max(*samples, 0) & ~kScalerInfBit);   // this step's sample count, for the peers
    const int32_t t = *step + 1;
    const double b1 = sched[2], b2 = sched[3];
    // LambdaLR (main.py:189) is stepped after EVERY train step, skipped or not (utils_init_nerf.py:626-629): its epoch is the
    // scaler's iteration count, not the optimiser's step count
    const double epoch = (double)scaler[kScalerIter];
    double decay = 1.0;
    if (sched[7] > 0.0f) decay = pow((double)sched[6], fmin(epoch / (double)sched[7], 1.0));
    const float bc1 = (float)(1.0 - pow(b1, (double)t)), bc2s = (float)sqrt(1.0 - pow(b2, (double)t));
    const float scale = __uint_as_float(scaler[kScalerScale]);
    const float skip = (local_skip && (*scaler_flag(scaler) & kScalerInfBit)) ? 1.0f : 0.0f;
    for (int g = 0; g < 2; g++) {
        AdamHyper h;
        h.lr = (float)((double)sched[g] * decay);
        h.beta1 = sched[2]; h.beta2 = sched[3]; h.eps = sched[4];
        h.bc1 = bc1; h.bc2_sqrt = bc2s; h.grad_scale = 1.0f / scale; h.skip = skip;
        out[g] = h;
    }
}

__global__ void k_scaler_commit(ScalerCommit c) {
    if (threadIdx.x == 0 && blockIdx.x == 0) scaler_commit(c);
}

// loss = sum (image - target)^2 * inv_n ; g_image = 2 (image - target) * inv_n * loss_scale     (one thread per ray)
__global__ void __launch_bounds__(256)
k_mse_loss_grad(const float *__restrict__ image, const float *__restrict__ target, uint32_t N, float inv_n,
                float loss_scale, float *__restrict__ loss, float *__restrict__ g_image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    if (n < N) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float d = image[n * 3 + c] - target[n * 3 + c];
            acc += d * d;
            g_image[n * 3 + c] = 2.0f * d * inv_n * loss_scale;
        }
    }
    acc = nb_warp_sum(acc);
    __shared__ float part[8];
    if (nb_lane() == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = part[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
        if (threadIdx.x == 0) atomicAdd(loss, v * inv_n);
    }
}

}  // namespace

extern "C" {

int nb200_fused_adam(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, uint64_t split,
                     const float *hyper, int zero_grad, void *stream) {
    return nb200_fused_adam_cfg(param, grad, exp_avg, exp_avg_sq, n, split, hyper, zero_grad, 0, 0, 0, stream);
}

int nb200_fused_adam_cfg(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, uint64_t split,
                         const float *hyper, int zero_grad, uint32_t cfg_grid, uint32_t cfg_threads, uint32_t cfg_unroll,
                         void *stream) {
    if (n == 0) return 0;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper || split > n) return NB200_E_BAD_ARG;
    const uintptr_t al = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq);
    if ((al & 15u) || (split & 3u)) return NB200_E_BAD_ARG;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t n4 = n / 4;
    cudaStream_t st = nb_stream(stream);
    if (n4) {
        static int per_sm = -1;         // tuning knob: CTAs per SM (a thin sweep leaves room for a co-running kernel)
        if (per_sm < 0) { const char *e = getenv("NB200_ADAM_CTAS_PER_SM"); per_sm = e ? atoi(e) : 0; if (per_sm <= 0 || per_sm > 8) per_sm = 8; }
        static int env_grid = -1, env_thr = 512, env_u = 2;     // tuning knobs: override the caller's shape
        if (env_grid < 0) {
            const char *e = getenv("NB200_ADAM_GRID"); env_grid = e ? atoi(e) : 0;
            e = getenv("NB200_ADAM_THREADS"); if (e) env_thr = atoi(e);
            e = getenv("NB200_ADAM_UNROLL"); if (e) env_u = atoi(e);
        }
        int deep_grid = env_grid > 0 ? env_grid : (int)cfg_grid;
        int deep_thr = env_grid > 0 ? env_thr : (int)cfg_threads, deep_u = env_grid > 0 ? env_u : (int)cfg_unroll;
        if (deep_thr != 256 && deep_thr != 512) deep_thr = 512;
        const uint32_t want = nb_div_up(n4, 256);
        const uint32_t grid = want < (uint32_t)sms * per_sm ? want : (uint32_t)sms * per_sm;
        if (deep_grid > 0 && n4 > (uint64_t)deep_grid * deep_thr * 4) {
#define NB_ADAM_DEEP(UU) k_fused_adam_deep<UU><<<deep_grid, deep_thr, 0, st>>>((float4 *)param, (float4 *)grad, (float4 *)exp_avg, \
                                                                              (float4 *)exp_avg_sq, n4, split / 4, (const AdamHyper *)hyper, zero_grad)
            if (deep_u >= 4) NB_ADAM_DEEP(4); else if (deep_u >= 2) NB_ADAM_DEEP(2); else NB_ADAM_DEEP(1);
#undef NB_ADAM_DEEP
        } else
        k_fused_adam<<<grid, 256, 0, st>>>((float4 *)param, (float4 *)grad, (float4 *)exp_avg, (float4 *)exp_avg_sq, n4,
                                           split / 4, (const AdamHyper *)hyper, zero_grad);
        NB_LAUNCH_CHECK();
    }
    if (n4 * 4 < n) {
        k_fused_adam_tail<<<1, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n4 * 4, n, split, (const AdamHyper *)hyper,
                                             zero_grad);
        NB_LAUNCH_CHECK();
    }
    return 0;
}

int nb200_adam_hyper(int32_t *step, const float *sched, float *hyper, void *stream) {
    if (!step || !sched || !hyper) return NB200_E_BAD_ARG;
    k_adam_hyper<<<1, 32, 0, nb_stream(stream)>>>(step, sched, (AdamHyper *)hyper);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_adam_hyper_scaled(const int32_t *step, const float *sched, float *hyper, uint32_t *scaler, int local_skip,
                            const int32_t *samples, void *stream) {
    if (!step || !sched || !hyper || !scaler) return NB200_E_BAD_ARG;
    k_adam_hyper_scaled<<<1, 32, 0, nb_stream(stream)>>>(step, sched, (AdamHyper *)hyper, scaler, local_skip, samples);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_scaler_commit(int32_t *step, uint32_t *scaler, uint32_t *const *peer_scalers, uint32_t world, int32_t *max_samples,
                        void *stream) {
    if (!step || !scaler || world > NB200_PEER_MAX) return NB200_E_BAD_ARG;
    ScalerCommit c;
    memset(&c, 0, sizeof(c));
    c.step = step; c.scaler = scaler; c.max_samples = max_samples; c.world = peer_scalers ? world : 0;
    for (uint32_t q = 0; q < c.world; q++) c.peers[q] = peer_scalers[q];
    k_scaler_commit<<<1, 32, 0, nb_stream(stream)>>>(c);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_mse_loss_grad(const float *image, const float *target, uint32_t N, float inv_n, float loss_scale, float *loss,
                        float *g_image, void *stream) {
    if (N == 0) return 0;
    if (!image || !target || !loss || !g_image) return NB200_E_BAD_ARG;
    k_mse_loss_grad<<<nb_div_up(N, 256), 256, 0, nb_stream(stream)>>>(image, target, N, inv_n, loss_scale, loss, g_image);
    NB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// ---- L2 bandwidth probes (measurement aid: MEASURED_PEAKS.json has no L2 figure) -------------------
namespace {
// every thread streams float4s of a small (L2-resident) buffer `reps` times
__global__ void __launch_bounds__(256)
k_l2_stream(const float4 *__restrict__ buf, uint64_t n4, uint32_t reps, float *__restrict__ sink) {
    float acc = 0.0f;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t r = 0; r < reps; r++)
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 v = __ldcg(buf + i);           // .cg: L2 only, the L1 never serves these
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123456.789f) *sink = acc;                // keep the loads alive
}
// every thread gathers 8-byte words at hashed positions (one 32-byte sector per request, like a hashed grid level)
__global__ void __launch_bounds__(256)
k_l2_gather(const float2 *__restrict__ buf, uint32_t mask, uint32_t per_thread, float *__restrict__ sink) {
    float acc = 0.0f;
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
#pragma unroll 8
    for (uint32_t j = 0; j < per_thread; j++) {
        h = h * 1664525u + 1013904223u;
        const float2 v = __ldcg(buf + ((h >> 7) & mask));
        acc += v.x + v.y;
    }
    if (acc == 123456.789f) *sink = acc;
}
// plain float4 copy, 4 groups in flight per thread: dst / src may be peer mappings (NVLink pull or push)
__global__ void __launch_bounds__(256)
k_stream_copy(float4 *__restrict__ dst, const float4 *__restrict__ src, uint64_t n4) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 a = __ldcg(src + i), b = __ldcg(src + i + stride), c = __ldcg(src + i + 2 * stride), d = __ldcg(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n4; i += stride) dst[i] = __ldcg(src + i);
}
}  // namespace

extern "C" {
// SM-issued copy of `bytes` (16-byte aligned) on `ctas` CTAs of 256 threads; with one side a peer mapping this measures
// what NVLink gives loads (pull) or stores (push) issued by SMs -- the roofline of the peer-memory update kernel
int nb200_stream_copy_probe(void *dst, const void *src, uint64_t bytes, uint32_t ctas, void *stream) {
    if (!dst || !src || (bytes & 15u) || ctas == 0) return NB200_E_BAD_ARG;
    k_stream_copy<<<ctas, 256, 0, nb_stream(stream)>>>((float4 *)dst, (const float4 *)src, bytes / 16);
    NB_LAUNCH_CHECK();
    return 0;
}
// streams `bytes` (16-byte aligned, should fit the L2) `reps` times; returns after enqueueing.  bytes_moved = bytes * reps
int nb200_l2_stream_probe(const void *buf, uint64_t bytes, uint32_t reps, float *sink, void *stream) {
    if (!buf || !sink || (bytes & 15u)) return NB200_E_BAD_ARG;
    k_l2_stream<<<148 * 8, 256, 0, nb_stream(stream)>>>((const float4 *)buf, bytes / 16, reps, sink);
    NB_LAUNCH_CHECK();
    return 0;
}
// n_threads x per_thread random 8-byte gathers over the first `words` (power of two) float2 words of buf
int nb200_l2_gather_probe(const void *buf, uint32_t words, uint32_t n_threads, uint32_t per_thread, float *sink, void *stream) {
    if (!buf || !sink || (words & (words - 1)) || n_threads == 0) return NB200_E_BAD_ARG;
    k_l2_gather<<<nb_div_up(n_threads, 256), 256, 0, nb_stream(stream)>>>((const float2 *)buf, words - 1, per_thread, sink);
    NB_LAUNCH_CHECK();
    return 0;
}
}  // extern "C"
