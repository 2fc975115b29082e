// adam.cuh -- the Adam arithmetic shared by the local sweep (optim.cu) and the peer-memory update (peer_update.cu).
// Follows torch's fused Adam (aten/src/ATen/native/cuda/fused_adam_utils.cuh, non-amsgrad, no weight decay):
// m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).
#pragma once
#include "common.cuh"

struct AdamHyper {          // 8 floats per parameter group, written by the host before every step
    float lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale, pad;
};

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
