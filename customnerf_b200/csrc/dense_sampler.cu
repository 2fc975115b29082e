// dense_sampler.cu -- the sampler of the dense (non-cuda_ray) renderer as two kernels.
//
// Replaces the torch-op sequence of NeRFRenderer.run before the field network is called (nerf/renderer.py:297-367) and
// sample_pdf (:21-55): stratified coarse samples -> [density query] -> coarse weights -> inverse-CDF importance samples ->
// merge.  The reference builds the merged sample set with a cumprod, a cumsum, a searchsorted, four gathers, a cat, a
// sort and another gather, and evaluates the density of the new points a second time only to throw it away (:353).
// Here one CTA owns a ray: the prefix products / sums run sequentially in fp32 (the order torch's CPU kernels use), the
// 2 x 64 depths are sorted by a bitonic network in shared memory, and the output is written in the layout of the
// occupancy path -- xyzs / dirs / deltas [N * T, ...] and rays = (n, n * T, T) -- so that everything after the sampler
// (fused field kernels, composite kernels with their LGIE variants, loss, backward) is the code of that path:
//   deltas[:, 0] = z[k+1] - z[k] (last: sample_dist)      what weights_sum_i composites with (:420-424)
//   deltas[:, 1] = ori_z[k] - ori_z[k-1]                  so that the composite kernel's running t is ori_z = clamp((z - near) /
//                                                         (far - near), 0, 1), the dense path's depth coordinate (:431-432)
// and composite_rays_train with T_thresh = 0 evaluates the dense formula (no early termination).
#include "common.cuh"

namespace {

constexpr uint32_t kMaxDense = 256;     // coarse + importance samples per ray

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// z_c[n, i] = near + (far - near) * lin[i] (+ (noise - 0.5) * sample_dist); xyz = clamp(o + d * z) -- every product / sum
// rounded on its own, as the chain of torch kernels rounds them (:306-317)
__global__ void __launch_bounds__(256)
k_dense_coarse(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ nears,
               const float *__restrict__ fars, const float *__restrict__ aabb, const float *__restrict__ lin,
               const float *__restrict__ noise, uint32_t N, uint32_t S, float *__restrict__ z_c, float *__restrict__ xyzs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * S) return;
    const uint32_t n = t / S, i = t - n * S;
    const float near = nears[n], far = fars[n], span = __fsub_rn(far, near);
    float z = __fadd_rn(near, __fmul_rn(span, lin[i]));
    // sample_dist = (fars - nears) / num_steps: a tensor divided by a Python scalar is, in torch's CUDA kernel, a multiplication
    // by the fp32 reciprocal (exact for the power-of-two sample counts the reference uses)
    if (noise) z = __fadd_rn(z, __fmul_rn(__fsub_rn(noise[t], 0.5f), __fmul_rn(span, __fdiv_rn(1.0f, (float)S))));
    z_c[t] = z;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float p = __fadd_rn(rays_o[n * 3 + a], __fmul_rn(rays_d[n * 3 + a], z));
        xyzs[(size_t)t * 3 + a] = fminf(fmaxf(p, aabb[a]), aabb[3 + a]);
    }
}

// one CTA per ray
__global__ void __launch_bounds__(256)
k_dense_importance(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ nears,
                   const float *__restrict__ fars, const float *__restrict__ aabb, const float *__restrict__ z_c,
                   const float *__restrict__ sigma_c, const float *__restrict__ u_in, uint32_t u_per_ray, uint32_t N, uint32_t S,
                   uint32_t S_up, float *__restrict__ z_all, float *__restrict__ xyzs, float *__restrict__ dirs,
                   float *__restrict__ deltas, int32_t *__restrict__ rays) {
    __shared__ float z[kMaxDense];          // coarse depths, later all depths (sorted)
    __shared__ float w[kMaxDense];          // coarse weights
    __shared__ float cdf[kMaxDense];
    __shared__ float mid[kMaxDense];
    __shared__ float oz[kMaxDense];
    const uint32_t n = blockIdx.x, tid = threadIdx.x, T = S + S_up;
    const float near = nears[n], far = fars[n], span = __fsub_rn(far, near);
    const float sample_dist = __fmul_rn(span, __fdiv_rn(1.0f, (float)S));      // as torch's CUDA kernel evaluates tensor / scalar
    if (tid < S) z[tid] = z_c[(size_t)n * S + tid];
    __syncthreads();
    // alphas of the coarse samples (:330-335)
    float alpha = 0.0f, delta = 0.0f;
    if (tid < S) {
        delta = tid + 1 < S ? __fsub_rn(z[tid + 1], z[tid]) : sample_dist;
        alpha = __fsub_rn(1.0f, expf(-__fmul_rn(delta, sigma_c[(size_t)n * S + tid])));
        w[tid] = alpha;
        if (tid + 1 < S) mid[tid] = __fadd_rn(z[tid], __fmul_rn(0.5f, delta));     // z_vals_mid (:338)
    }
    __syncthreads();
    if (tid == 0 && S_up > 0) {
        // weights = alphas * cumprod([1, 1 - alphas + 1e-15])[:-1], sequential like torch.cumprod on the CPU (:334-336)
        float Tr = 1.0f;
        for (uint32_t i = 0; i < S; i++) {
            const float a = w[i];
            w[i] = __fmul_rn(a, Tr);
            Tr = __fmul_rn(Tr, __fadd_rn(__fsub_rn(1.0f, a), 1e-15f));
        }
        // sample_pdf(z_mid, weights[1:-1]) (:21-31): pdf = (w + 1e-5) / sum, cdf = [0, cumsum(pdf)]  (S - 1 entries)
        float sum = 0.0f;
        for (uint32_t i = 1; i + 1 < S; i++) sum = __fadd_rn(sum, __fadd_rn(w[i], 1e-5f));
        float c = 0.0f;
        cdf[0] = 0.0f;
        for (uint32_t i = 1; i + 1 < S; i++) {
            c = __fadd_rn(c, __fdiv_rn(__fadd_rn(w[i], 1e-5f), sum));
            cdf[i] = c;
        }
    }
    __syncthreads();
    // inverse CDF (:41-55): thread j draws importance sample j
    if (tid < S_up) {
        const uint32_t nc = S - 1;                                  // entries of cdf / z_mid
        const float u = u_in[(size_t)(u_per_ray ? n : 0) * S_up + tid];
        uint32_t lo = 0, hi = nc;                                   // searchsorted(cdf, u, right=True): first index with cdf > u
        while (lo < hi) {
            const uint32_t m = (lo + hi) >> 1;
            if (cdf[m] <= u) lo = m + 1; else hi = m;
        }
        const uint32_t below = lo > 0 ? lo - 1 : 0, above = lo < nc - 1 ? lo : nc - 1;
        float denom = __fsub_rn(cdf[above], cdf[below]);
        if (denom < 1e-5f) denom = 1.0f;
        const float tt = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
        oz[tid] = __fadd_rn(mid[below], __fmul_rn(tt, __fsub_rn(mid[above], mid[below])));
    }
    __syncthreads();
    // all depths, sorted (:360-361): bitonic network over the next power of two, padded with +inf
    uint32_t P = 1;
    while (P < T) P <<= 1;
    for (uint32_t i = tid; i < P; i += blockDim.x) {
        if (i >= S) z[i] = i < T ? oz[i - S] : __int_as_float(0x7f800000);
    }
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < P; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const float a = z[i], b = z[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { z[i] = b; z[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // ori_z = clamp((z - near) / (far - near), 0, 1) (:431-432)
    for (uint32_t i = tid; i < T; i += blockDim.x) oz[i] = clampf(__fdiv_rn(__fsub_rn(z[i], near), span), 0.0f, 1.0f);
    __syncthreads();
    const float o0 = rays_o[n * 3], o1 = rays_o[n * 3 + 1], o2 = rays_o[n * 3 + 2];
    const float d0 = rays_d[n * 3], d1 = rays_d[n * 3 + 1], d2 = rays_d[n * 3 + 2];
    for (uint32_t i = tid; i < T; i += blockDim.x) {
        const size_t s = (size_t)n * T + i;
        const float zi = z[i];
        z_all[s] = zi;
        xyzs[s * 3] = fminf(fmaxf(__fadd_rn(o0, __fmul_rn(d0, zi)), aabb[0]), aabb[3]);
        xyzs[s * 3 + 1] = fminf(fmaxf(__fadd_rn(o1, __fmul_rn(d1, zi)), aabb[1]), aabb[4]);
        xyzs[s * 3 + 2] = fminf(fmaxf(__fadd_rn(o2, __fmul_rn(d2, zi)), aabb[2]), aabb[5]);
        dirs[s * 3] = d0; dirs[s * 3 + 1] = d1; dirs[s * 3 + 2] = d2;
        deltas[s * 2] = i + 1 < T ? __fsub_rn(z[i + 1], zi) : sample_dist;
        deltas[s * 2 + 1] = i > 0 ? __fsub_rn(oz[i], oz[i - 1]) : oz[0];
    }
    if (tid == 0) { rays[n * 3] = (int32_t)n; rays[n * 3 + 1] = (int32_t)(n * T); rays[n * 3 + 2] = (int32_t)T; }
}

}  // namespace

extern "C" {

int nb200_dense_coarse(const float *rays_o, const float *rays_d, const float *nears, const float *fars, const float *aabb,
                       const float *lin, const float *noise, uint32_t N, uint32_t S, float *z_c, float *xyzs, void *stream) {
    if (N == 0 || S == 0) return 0;
    if (!rays_o || !rays_d || !nears || !fars || !aabb || !lin || !z_c || !xyzs || (uint64_t)N * S > 0x7fffffffull) return NB200_E_BAD_ARG;
    k_dense_coarse<<<nb_div_up((uint64_t)N * S, 256), 256, 0, nb_stream(stream)>>>(rays_o, rays_d, nears, fars, aabb, lin, noise, N, S,
                                                                                 z_c, xyzs);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_dense_importance(const float *rays_o, const float *rays_d, const float *nears, const float *fars, const float *aabb,
                           const float *z_c, const float *sigma_c, const float *u, int u_per_ray, uint32_t N, uint32_t S,
                           uint32_t S_up, float *z_all, float *xyzs, float *dirs, float *deltas, int32_t *rays, void *stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !nears || !fars || !aabb || !z_c || !z_all || !xyzs || !dirs || !deltas || !rays) return NB200_E_BAD_ARG;
    if (S < 3 || S + S_up > kMaxDense || (S_up > 0 && (!sigma_c || !u)) || (uint64_t)N * (S + S_up) > 0x7fffffffull) return NB200_E_BAD_ARG;
    const uint32_t T = S + S_up;
    k_dense_importance<<<N, T > 128 ? 256 : 128, 0, nb_stream(stream)>>>(rays_o, rays_d, nears, fars, aabb, z_c, sigma_c, u,
                                                                       u_per_ray ? 1u : 0u, N, S, S_up, z_all, xyzs, dirs, deltas, rays);
    NB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
