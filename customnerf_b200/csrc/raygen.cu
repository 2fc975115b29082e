// raygen.cu -- camera rays on the device: the arithmetic of get_rays
// (nerf/provider_utils.py:238-302) for the nerfstudio pin-hole model the reference's loader uses (provider.py:344-470).
//
// The reference builds rays with ~15 torch ops per batch (meshgrid, gather, stack, normalise, bmm); a train step then
// receives [N,3] origins and directions.  Here one kernel turns a 4x4 camera-to-world pose into the step's rays, so what
// crosses PCIe per step is the pose (64 B) and the target pixels instead of 24 B of ray per pixel, and the kernel is the
// first compute node of the step's CUDA graph.
//
//   pixel p -> (i, j) = (p % W + off_x, p / W + off_y)                     provider_utils.py:258-260
//   dir = (xs, ys, 1) / sqrt(max(xs^2 + ys^2 + 1, 1e-20)),  xs = (i - cx) / fx,  ys = (j - cy) / fy     :289-293, :125-126
//   rays_d = R dir  (directions @ poses[:, :3, :3]^T),  rays_o = poses[:, :3, 3]                          :294-297
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_get_rays(const float *__restrict__ poses, float fx, float fy, float cx, float cy, uint32_t W, uint32_t HW, uint32_t N,
           const int64_t *__restrict__ inds, float off_x, float off_y, float *__restrict__ rays_o,
           float *__restrict__ rays_d) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (n >= N) return;
    const float *P = poses + (size_t)b * 16;
    int64_t p = inds ? inds[(size_t)b * N + n] : (int64_t)n;
    if (p < 0) p = 0;
    if (p >= (int64_t)HW) p = HW - 1;
    const float i = (float)((uint32_t)p % W) + off_x, j = (float)((uint32_t)p / W) + off_y;
    const float xs = __fdiv_rn(__fsub_rn(i, cx), fx), ys = __fdiv_rn(__fsub_rn(j, cy), fy);
    // sum(x * x, -1): torch reduces the three products in order ((xs^2 + ys^2) + 1), products rounded separately
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(xs, xs), __fmul_rn(ys, ys)), 1.0f);
    const float nrm = __fsqrt_rn(fmaxf(s, 1e-20f));
    const float dx = __fdiv_rn(xs, nrm), dy = __fdiv_rn(ys, nrm), dz = __fdiv_rn(1.0f, nrm);
    const size_t o = ((size_t)b * N + n) * 3;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        rays_d[o + k] = fmaf(dz, P[k * 4 + 2], fmaf(dy, P[k * 4 + 1], dx * P[k * 4 + 0]));
        rays_o[o + k] = P[k * 4 + 3];
    }
}

}  // namespace

extern "C" int nb200_get_rays(const float *poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, uint32_t B,
                              uint32_t N, const int64_t *inds, float off_x, float off_y, float *rays_o, float *rays_d,
                              void *stream) {
    if (B == 0 || N == 0) return 0;
    if (!poses || !rays_o || !rays_d || H == 0 || W == 0 || fx == 0.0f || fy == 0.0f) return NB200_E_BAD_ARG;
    if (!inds && N != H * W) return NB200_E_BAD_ARG;
    const dim3 grid(nb_div_up(N, 256), B);
    k_get_rays<<<grid, 256, 0, nb_stream(stream)>>>(poses, fx, fy, cx, cy, W, H * W, N, inds, off_x, off_y, rays_o, rays_d);
    NB_LAUNCH_CHECK();
    return 0;
}
