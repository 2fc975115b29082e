// adam.cuh -- the Adam arithmetic shared by the local sweep (optim.cu) and the peer-memory update (peer_update.cu).
// Follows torch's fused Adam (aten/src/ATen/native/cuda/fused_adam_utils.cuh, non-amsgrad, no weight decay):
// m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).
#pragma once
#include "common.cuh"

struct AdamHyper {          // 8 floats per parameter group, written on the device before every step (k_adam_hyper*)
    float lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale;
    float skip;             // != 0: a gradient of this step was not finite -- leave p, m, v alone (GradScaler.step semantics)
};

// Dynamic loss scaling on the device -- torch.cuda.amp.GradScaler as the reference trains with it
// (nerf/utils_init_nerf.py:100,612-629): scaler state, 8 x 32 bit:
//   [0] scale f32        [1] growth tracker i32        [2] iteration u32 (every update, taken or skipped)
//   [3] skipped steps u32        [4], [5] status words of even / odd iterations: bit 31 = found-inf (set by the backward
//   kernels), bits 0..30 = the step's sample count (so that ray-sharded ranks agree on buffer growth)
//   [6] growth interval i32 (2000)        [7] unused
// The backward kernels of iteration `it` raise bit 31 of word [4 + (it & 1)]; the update skips the optimiser step when any
// rank's bit is up and nb200_scaler_commit halves the scale (else counts towards doubling it) and clears the OTHER word for
// iteration it + 1 -- two alternating words so that a fast rank never clears one a slow peer has yet to read.  (bit 31 so
// that an all-reduce(MAX) of the words -- the NCCL form of the exchange -- preserves it.)
constexpr uint32_t kScalerScale = 0, kScalerTracker = 1, kScalerIter = 2, kScalerSkipped = 3, kScalerFlag0 = 4, kScalerInterval = 6;
constexpr uint32_t kScalerInfBit = 0x80000000u;
__device__ __forceinline__ uint32_t *scaler_flag(uint32_t *scaler) { return scaler + kScalerFlag0 + (scaler[kScalerIter] & 1u); }
__device__ __forceinline__ void scaler_raise(uint32_t *scaler) { atomicOr(scaler_flag(scaler), kScalerInfBit); }

// per-group constants hoisted out of the sweep (two of the three divisions of the update are per-step constants)
struct AdamConst {
    float grad_scale, w1, beta2, w2, inv_bc2_sqrt, eps, step_size;
    __device__ __forceinline__ explicit AdamConst(const AdamHyper &h)
        : grad_scale(h.grad_scale), w1(1.0f - h.beta1), beta2(h.beta2), w2(1.0f - h.beta2),
          inv_bc2_sqrt(1.0f / h.bc2_sqrt), eps(h.eps), step_size(h.lr / h.bc1) {}
};

// Every contraction is spelled out (explicit FMAs / separately rounded products), so that the local sweep and the
// peer-memory kernel -- two different loop bodies around this function -- round identically: bit-equal results.
__device__ __forceinline__ void adam1(float &p, float &g, float &m, float &v, const AdamConst &h) {
    const float gr = __fmul_rn(g, h.grad_scale);
    m = __fmaf_rn(h.w1, __fsub_rn(gr, m), m);           // torch lerp(m, g, w) for w < 0.5
    v = __fmaf_rn(h.beta2, v, __fmul_rn(__fmul_rn(h.w2, gr), gr));
    const float denom = __fmaf_rn(sqrtf(v), h.inv_bc2_sqrt, h.eps);
    p = __fmaf_rn(-h.step_size, __fdividef(m, denom), p);   // 2-ulp division: the update is <= lr, its error ~1e-10
}

// GradScaler.update() (torch/amp/grad_scaler.py: backoff 0.5, growth 2.0 every `interval` clean steps) + the optimiser's own
// step count; ONE thread.  peers[q] = the scaler words of rank q (world > 1, peer-memory update): the decision is the OR of
// every rank's found-inf bit, *max_samples the maximum of their sample counts (the same numbers on every rank).
struct ScalerCommit {
    int32_t *step; uint32_t *scaler; int32_t *max_samples; uint32_t world, pad;
    uint32_t *peers[NB200_PEER_MAX];
};
__device__ __forceinline__ void scaler_commit(const ScalerCommit &c) {
    uint32_t *scaler = c.scaler;
    const uint32_t it = scaler[kScalerIter], slot = kScalerFlag0 + (it & 1u);
    uint32_t word = scaler[slot], found = word & kScalerInfBit, smax = word & ~kScalerInfBit;
    for (uint32_t q = 0; q < c.world; q++)
        if (c.peers[q]) {
            word = *(volatile uint32_t *)(c.peers[q] + slot);
            found |= word & kScalerInfBit;
            smax = max(smax, word & ~kScalerInfBit);
        }
    if (c.max_samples) *c.max_samples = (int32_t)smax;      // the same number on every rank: buffer growth is a joint decision
    float scale = __uint_as_float(scaler[kScalerScale]);
    int32_t tracker = (int32_t)scaler[kScalerTracker];
    if (found) {
        scale *= 0.5f; tracker = 0; scaler[kScalerSkipped] += 1u;
    } else {
        *c.step += 1;
        const int32_t interval = (int32_t)scaler[kScalerInterval];
        if (++tracker >= interval && interval > 0) { scale *= 2.0f; tracker = 0; }
    }
    scaler[kScalerScale] = __float_as_uint(scale);
    scaler[kScalerTracker] = (uint32_t)tracker;
    scaler[kScalerFlag0 + ((it + 1u) & 1u)] = 0u;           // the word of iteration it + 1 (nobody reads it any more)
    scaler[kScalerIter] = it + 1u;
}
