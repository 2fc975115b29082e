// common.cuh -- shared device helpers for libnerf_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nerf_b200.h"

#define NB_LAUNCH_CHECK()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

static inline uint32_t nb_div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
static inline cudaStream_t nb_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- scalar <-> storage type conversion -----------------------------------------------------
template <typename T> __device__ __forceinline__ float nb_to_float(T v);
template <> __device__ __forceinline__ float nb_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float nb_to_float<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T nb_from_float(float v);
template <> __device__ __forceinline__ float nb_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half nb_from_float<__half>(float v) { return __float2half_rn(v); }

// ---- warp helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nb_lane() { return threadIdx.x & 31u; }

__device__ __forceinline__ int nb_warp_incl_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)nb_lane() >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ float nb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- morton (raymarching.cu:56-81) ----------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t nb_expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t nb_morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return nb_expand_bits(x) | (nb_expand_bits(y) << 1) | (nb_expand_bits(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t nb_morton3D_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}
