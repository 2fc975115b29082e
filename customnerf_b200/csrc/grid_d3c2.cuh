// grid_d3c2.cuh -- the D = 3, C = 2 multiresolution grid gather shared by the standalone encoder kernels (gridencoder.cu)
// and the fused encode + field-network kernels (field_fused.cu): per-level constants, row addressing and the 8-corner
// interpolation of one level.  Index arithmetic is uint32 with wrap-around exactly as gridencoder/src/gridencoder.cu:50-84 of
// the reference; both users go through ge_level_gather(), so the features they produce are bit-identical.
#pragma once
#include "common.cuh"

namespace {

__device__ __forceinline__ float ge_level_scale(uint32_t level, float S, uint32_t H) {
    return exp2f(level * S) * H - 1.0f;      // gridencoder.cu:138 (same expression => same FFMA contraction)
}
__device__ __forceinline__ float ge_smoothstep(float v) { return v * v * (3.0f - 2.0f * v); }
__device__ __forceinline__ float ge_smoothstep_d(float v) { return 6 * v * (1.0f - v); }

constexpr uint32_t kMaxFastLevels = 32;
constexpr uint32_t P1 = 2654435761u, P2 = 805459861u;

// optional input transform and device-side row count of the fused train step (nb200_fs_*): x01 = (x + add) * mul is
// what GridEncoder.forward's `(inputs + bound) / (2 * bound)` evaluates to in torch (a tensor / python-scalar division
// is a multiplication by the fp32 reciprocal), fused here so the normalised copy of the sample positions never exists.
struct InXform {
    float add, mul;                 // mul == 0: identity (inputs are already in [0, 1])
    const int32_t *count_dev;       // when non-null only rows < min(B, *count_dev) are processed
    uint32_t level_begin = 0;       // scatter only: levels [level_begin, max_level) (the ray-sharded step scatters the table in two
                                    // launches so that the update of the first part runs beside the second launch)
    __device__ __forceinline__ float operator()(float x) const { return mul != 0.0f ? __fmul_rn(__fadd_rn(x, add), mul) : x; }
};

struct LevelInfo {
    uint32_t offset;     // first row of the level
    uint32_t size;       // rows in the level (hashmap_size)
    uint32_t m1, m2;     // dense strides of y and z (0 when the reference's stride loop has stopped)
    uint32_t mask;       // size-1 when size is a power of two, else 0
    uint32_t use_hash;
    float scale;
    uint32_t pad;
};

__device__ __forceinline__ void ge_fill_level_info(LevelInfo *info, const int32_t *__restrict__ offsets, uint32_t nlev,
                                                   float S, uint32_t H, uint32_t gridtype, bool align_corners) {
    for (uint32_t l = threadIdx.x; l < nlev; l += blockDim.x) {
        LevelInfo li;
        li.offset = (uint32_t)offsets[l];
        li.size = (uint32_t)(offsets[l + 1] - offsets[l]);
        li.scale = ge_level_scale(l, S, H);
        const uint32_t resolution = (uint32_t)ceilf(li.scale) + 1;
        const uint32_t r1 = align_corners ? resolution : resolution + 1;
        // replay of the stride loop of get_grid_index (gridencoder.cu:71-75) for D = 3
        uint32_t stride = 1;
        stride *= r1;                                   // d = 0 always executes (1 <= size)
        li.m1 = 0; li.m2 = 0;
        if (stride <= li.size) {
            li.m1 = stride; stride *= r1;
            if (stride <= li.size) { li.m2 = stride; stride *= r1; }
        }
        li.use_hash = (gridtype == 0 && stride > li.size) ? 1u : 0u;
        li.mask = ((li.size & (li.size - 1)) == 0) ? li.size - 1 : 0u;
        li.pad = 0;
        info[l] = li;
    }
}

__device__ __forceinline__ uint32_t ge_row_d3(const LevelInfo &li, uint32_t x, uint32_t y, uint32_t z) {
    uint32_t raw = li.use_hash ? (x ^ (y * P1) ^ (z * P2)) : (x + y * li.m1 + z * li.m2);
    if (li.mask) return raw & li.mask;
    return raw < li.size ? raw : raw % li.size;
}

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<__half> { using type = __half2; };
__device__ __forceinline__ float2 ge_ld2(const float *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }
__device__ __forceinline__ float2 ge_ld2(const __half *p) { return __half22float2(__ldg(reinterpret_cast<const __half2 *>(p))); }

// fp32 master table read as if it had been cast to fp16 first (the autocast path of grid.py:45-46 without the copy)
__device__ __forceinline__ float2 ge_ld2_round_half(const float *p) {
    return __half22float2(__float22half2_rn(__ldg(reinterpret_cast<const float2 *>(p))));
}


// the 8 corners of level `li` around x (already in [0, 1]^3): r0, r1 = the two interpolated features (fp32 accumulation in
// corner order 0..7, weights multiplied x, y, z -- gridencoder.cu:166-187).  kRoundHalf: fp32 table entries are rounded
// to fp16 as they are loaded (the autocast path of grid.py:45-46 without the table copy).
template <typename TE, bool kRoundHalf>
__device__ __forceinline__ void ge_level_gather(const LevelInfo &li, const TE *__restrict__ grid, float x0, float x1, float x2,
                                                float half_off, uint32_t interp, float &r0, float &r1) {
    float p0 = x0 * li.scale + half_off, p1 = x1 * li.scale + half_off, p2 = x2 * li.scale + half_off;
    const uint32_t g0 = (uint32_t)floorf(p0), g1 = (uint32_t)floorf(p1), g2 = (uint32_t)floorf(p2);
    p0 -= (float)g0; p1 -= (float)g1; p2 -= (float)g2;
    if (interp == 1) { p0 = ge_smoothstep(p0); p1 = ge_smoothstep(p1); p2 = ge_smoothstep(p2); }
    const TE *lg = grid + (size_t)li.offset * 2;
    float2 v[8];
#pragma unroll
    for (uint32_t idx = 0; idx < 8; idx++) {
        const uint32_t row = ge_row_d3(li, g0 + (idx & 1u), g1 + ((idx >> 1) & 1u), g2 + ((idx >> 2) & 1u));
        if constexpr (kRoundHalf) v[idx] = ge_ld2_round_half(lg + (size_t)row * 2);
        else v[idx] = ge_ld2(lg + (size_t)row * 2);
    }
    r0 = 0.0f; r1 = 0.0f;
#pragma unroll
    for (uint32_t idx = 0; idx < 8; idx++) {
        float w = 1;
        w *= (idx & 1u) ? p0 : 1 - p0;
        w *= (idx & 2u) ? p1 : 1 - p1;
        w *= (idx & 4u) ? p2 : 1 - p2;
        r0 += w * v[idx].x;
        r1 += w * v[idx].y;
    }
}

// ge_row_d3 with the (never taken on the reference's table layouts: dense levels index below their size, capped levels have
// power-of-two sizes) modulo kept out of line: 32 inlined copies of the division sequence per 4-level group cost ~2400
// instructions of I-cache footprint in ge_gather4
__device__ __noinline__ uint32_t ge_mod_slow(uint32_t raw, uint32_t size) { return raw % size; }
__device__ __forceinline__ uint32_t ge_row_d3_compact(const LevelInfo &li, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t raw = li.use_hash ? (x ^ (y * P1) ^ (z * P2)) : (x + y * li.m1 + z * li.m2);
    if (li.mask) return raw & li.mask;
    if (raw >= li.size) return ge_mod_slow(raw, li.size);
    return raw;
}

// Four consecutive levels at once for the latency-bound callers (field_fused.cu: few gather warps per SM): all 32 row
// indices first, then the 32 loads back to back (32 independent requests in flight per thread), then the interpolation --
// per level the very expression sequence of ge_level_gather, so the results are bit-identical to it.
template <typename TE, bool kRoundHalf>
__device__ __forceinline__ void ge_gather4(const LevelInfo *__restrict__ info4, const TE *__restrict__ grid, float x0, float x1,
                                           float x2, float half_off, uint32_t interp, float (&res)[8]) {
    float p[4][3];
    const TE *src[4][8];
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
        const LevelInfo li = info4[j];
        float p0 = x0 * li.scale + half_off, p1 = x1 * li.scale + half_off, p2 = x2 * li.scale + half_off;
        const uint32_t g0 = (uint32_t)floorf(p0), g1 = (uint32_t)floorf(p1), g2 = (uint32_t)floorf(p2);
        p0 -= (float)g0; p1 -= (float)g1; p2 -= (float)g2;
        if (interp == 1) { p0 = ge_smoothstep(p0); p1 = ge_smoothstep(p1); p2 = ge_smoothstep(p2); }
        p[j][0] = p0; p[j][1] = p1; p[j][2] = p2;
        const TE *lg = grid + (size_t)li.offset * 2;
#pragma unroll
        for (uint32_t idx = 0; idx < 8; idx++)
            src[j][idx] = lg + (size_t)ge_row_d3_compact(li, g0 + (idx & 1u), g1 + ((idx >> 1) & 1u), g2 + ((idx >> 2) & 1u)) * 2;
    }
    float2 v[4][8];
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
#pragma unroll
        for (uint32_t idx = 0; idx < 8; idx++) {
            if constexpr (kRoundHalf) v[j][idx] = ge_ld2_round_half(src[j][idx]);
            else v[j][idx] = ge_ld2(src[j][idx]);
        }
    }
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
        const float p0 = p[j][0], p1 = p[j][1], p2 = p[j][2];
        float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
        for (uint32_t idx = 0; idx < 8; idx++) {
            float w = 1;
            w *= (idx & 1u) ? p0 : 1 - p0;
            w *= (idx & 2u) ? p1 : 1 - p1;
            w *= (idx & 4u) ? p2 : 1 - p2;
            r0 += w * v[j][idx].x;
            r1 += w * v[j][idx].y;
        }
        res[2 * j] = r0; res[2 * j + 1] = r1;
    }
}

}  // namespace
