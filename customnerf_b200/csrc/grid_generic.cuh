// grid_generic.cuh -- the multiresolution grid encoding for every shape the product's fast path does not cover
// (D in 2..5, C in {1, 2, 4, 8}, either output layout, optional input gradients) and the total-variation gradient.
//
// Same design as the D = 3, C = 2 fast path (grid_d3c2.cuh), generalised: ONE THREAD PER POINT that walks the levels (the
// reference launches a thread per (point, level): gridencoder/src/gridencoder.cu:87-244, 247-339, 505-609), per-level
// constants -- offset, size, the strides its index loop would produce, scale -- computed once per CTA into shared memory,
// every corner of a cell gathered ONCE: the interpolated features and all D input derivatives come out of the same 2^D
// loads (the reference gathers 2^D corners for the features and 2 x 2^(D-1) more per derivative axis, :166-241),
// fp32 accumulation whatever the storage type.
//
// Index arithmetic is uint32 with wrap-around as gridencoder.cu:50-84 defines it: the stride of axis d is the product of
// (resolution [+ 1]) over the axes before it while that product does not exceed the level's size -- later axes drop out
// (stride 0); a hashed level (gridtype 0 and the product has outgrown the size) XORs coordinate x prime instead; the row is
// the index modulo the level size.
#pragma once
#include "common.cuh"

namespace {

constexpr uint32_t kGenMaxLevels = 32;

template <uint32_t D> struct GenLevel {
    uint32_t offset, size, mask, hashed, resolution;
    uint32_t stride[D];
    float scale;
};

__device__ __forceinline__ float gen_level_scale(uint32_t level, float S, uint32_t H) { return exp2f(level * S) * H - 1.0f; }

template <uint32_t D>
__device__ __forceinline__ void gen_fill_levels(GenLevel<D> *lv, const int32_t *__restrict__ offsets, uint32_t nlev, float S,
                                                uint32_t H, uint32_t gridtype, bool align_corners) {
    for (uint32_t l = threadIdx.x; l < nlev; l += blockDim.x) {
        GenLevel<D> g;
        g.offset = (uint32_t)offsets[l];
        g.size = (uint32_t)(offsets[l + 1] - offsets[l]);
        g.scale = gen_level_scale(l, S, H);
        g.resolution = (uint32_t)ceilf(g.scale) + 1;
        const uint32_t side = align_corners ? g.resolution : g.resolution + 1;
        uint32_t prod = 1;
        bool open = true;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            open = open && prod <= g.size;
            g.stride[d] = open ? prod : 0u;
            if (open) prod *= side;
        }
        g.hashed = (gridtype == 0 && prod > g.size) ? 1u : 0u;
        g.mask = ((g.size & (g.size - 1)) == 0) ? g.size - 1 : 0u;
        lv[l] = g;
    }
}

template <uint32_t D>
__device__ __forceinline__ uint32_t gen_row(const GenLevel<D> &g, const uint32_t (&cell)[D]) {
    constexpr uint32_t prime[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t raw = 0;
    if (g.hashed) {
#pragma unroll
        for (uint32_t d = 0; d < D; d++) raw ^= cell[d] * prime[d];
    } else {
#pragma unroll
        for (uint32_t d = 0; d < D; d++) raw += cell[d] * g.stride[d];
    }
    if (g.mask) return raw & g.mask;
    return raw < g.size ? raw : raw % g.size;
}

// element (b, level, channel) of a [B, L*C] (BLC) or [L, B, C] (LBC) tensor
__device__ __forceinline__ size_t gen_at(int layout, uint32_t b, uint32_t l, uint32_t B, uint32_t L, uint32_t C) {
    return layout == NB200_LAYOUT_LBC ? ((size_t)l * B + b) * C : ((size_t)b * L + l) * C;
}

// ---- forward (+ optional d out / d in), one thread per point
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(128)
k_gen_encode(const float *__restrict__ inputs, const T *__restrict__ table, const int32_t *__restrict__ offsets,
             T *__restrict__ outputs, T *__restrict__ dy_dx, uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H,
             uint32_t gridtype, bool align_corners, uint32_t interp, int layout) {
    __shared__ GenLevel<D> lv[kGenMaxLevels];
    gen_fill_levels<D>(lv, offsets, max_level, S, H, gridtype, align_corners);
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float x[D];
    bool outside = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = inputs[(size_t)b * D + d];
        outside |= (x[d] < 0.0f) | (x[d] > 1.0f);               // NaN passes, as in the reference (:110)
    }
    const float shift = align_corners ? 0.0f : 0.5f;
    for (uint32_t l = 0; l < max_level; l++) {
        const GenLevel<D> g = lv[l];
        float val[C], der[D][C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) {
            val[c] = 0.0f;
#pragma unroll
            for (uint32_t d = 0; d < D; d++) der[d][c] = 0.0f;
        }
        if (!outside) {
            uint32_t base[D];
            float f[D], slope[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                const float p = x[d] * g.scale + shift;
                const float fl = floorf(p);
                base[d] = (uint32_t)fl;
                f[d] = p - (float)base[d];
                slope[d] = 1.0f;
                if (interp == 1) { slope[d] = 6.0f * f[d] * (1.0f - f[d]); f[d] = f[d] * f[d] * (3.0f - 2.0f * f[d]); }
            }
            const T *rows = table + (size_t)g.offset * C;
#pragma unroll
            for (uint32_t corner = 0; corner < (1u << D); corner++) {
                uint32_t cell[D];
                float wd[D], w = 1.0f;
#pragma unroll
                for (uint32_t d = 0; d < D; d++) {
                    const bool hi = (corner >> d) & 1u;
                    cell[d] = base[d] + (hi ? 1u : 0u);
                    wd[d] = hi ? f[d] : 1.0f - f[d];
                    w *= wd[d];
                }
                const T *e = rows + (size_t)gen_row<D>(g, cell) * C;
                float v[C];
#pragma unroll
                for (uint32_t c = 0; c < C; c++) v[c] = nb_to_float<T>(e[c]);
#pragma unroll
                for (uint32_t c = 0; c < C; c++) val[c] += w * v[c];
                if (dy_dx) {
                    // d/dx_a of prod_d w_d = (+-1) * prod_{d != a} w_d: the same corner values serve every axis
#pragma unroll
                    for (uint32_t a = 0; a < D; a++) {
                        float wa = ((corner >> a) & 1u) ? 1.0f : -1.0f;
#pragma unroll
                        for (uint32_t d = 0; d < D; d++) if (d != a) wa *= wd[d];
#pragma unroll
                        for (uint32_t c = 0; c < C; c++) der[a][c] += wa * v[c];
                    }
                }
            }
            if (dy_dx) {
#pragma unroll
                for (uint32_t a = 0; a < D; a++)
#pragma unroll
                    for (uint32_t c = 0; c < C; c++) der[a][c] *= g.scale * slope[a];
            }
        }
        T *out = outputs + gen_at(layout, b, l, B, L, C);
#pragma unroll
        for (uint32_t c = 0; c < C; c++) out[c] = nb_from_float<T>(val[c]);
        if (dy_dx) {
            T *dd = dy_dx + ((size_t)b * L + l) * D * C;            // [B, L, D, C]
#pragma unroll
            for (uint32_t a = 0; a < D; a++)
#pragma unroll
                for (uint32_t c = 0; c < C; c++) dd[a * C + c] = nb_from_float<T>(der[a][c]);
        }
    }
}

// ---- backward: scatter of the feature gradients into the fp32 table gradient, one thread per point
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(128)
k_gen_scatter(const T *__restrict__ grad, const float *__restrict__ inputs, const int32_t *__restrict__ offsets,
              float *__restrict__ grad_table, uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H, uint32_t gridtype,
              bool align_corners, uint32_t interp, int layout) {
    __shared__ GenLevel<D> lv[kGenMaxLevels];
    gen_fill_levels<D>(lv, offsets, max_level, S, H, gridtype, align_corners);
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = inputs[(size_t)b * D + d];
        if (x[d] < 0.0f || x[d] > 1.0f) return;                     // points outside the unit cube contribute nothing (:283-287)
    }
    const float shift = align_corners ? 0.0f : 0.5f;
    for (uint32_t l = 0; l < max_level; l++) {
        const GenLevel<D> g = lv[l];
        const T *gin = grad + gen_at(layout, b, l, B, L, C);
        float gv[C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) gv[c] = nb_to_float<T>(gin[c]);
        uint32_t base[D];
        float f[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            const float p = x[d] * g.scale + shift;
            base[d] = (uint32_t)floorf(p);
            f[d] = p - (float)base[d];
            if (interp == 1) f[d] = f[d] * f[d] * (3.0f - 2.0f * f[d]);
        }
        float *rows = grad_table + (size_t)g.offset * C;
#pragma unroll
        for (uint32_t corner = 0; corner < (1u << D); corner++) {
            uint32_t cell[D];
            float w = 1.0f;
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                const bool hi = (corner >> d) & 1u;
                cell[d] = base[d] + (hi ? 1u : 0u);
                w *= hi ? f[d] : 1.0f - f[d];
            }
            float *e = rows + (size_t)gen_row<D>(g, cell) * C;
            if constexpr (C % 2 == 0) {
#pragma unroll
                for (uint32_t c = 0; c < C; c += 2) atomicAdd(reinterpret_cast<float2 *>(e + c), make_float2(w * gv[c], w * gv[c + 1]));
            } else {
#pragma unroll
                for (uint32_t c = 0; c < C; c++) atomicAdd(e + c, w * gv[c]);
            }
        }
    }
}

// ---- backward with respect to the inputs: grad_inputs[b, a] = sum_{l, c} grad[b, l, c] * dy_dx[b, l, a, c], one thread per point
template <typename T>
__global__ void __launch_bounds__(128)
k_gen_input_grad(const T *__restrict__ grad, const T *__restrict__ dy_dx, T *__restrict__ grad_inputs, uint32_t B, uint32_t D,
                 uint32_t C, uint32_t L, int layout) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float acc[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (uint32_t l = 0; l < L; l++) {
        const T *gin = grad + gen_at(layout, b, l, B, L, C);
        const T *dd = dy_dx + ((size_t)b * L + l) * D * C;
        for (uint32_t c = 0; c < C; c++) {
            const float gv = nb_to_float<T>(gin[c]);
            for (uint32_t a = 0; a < D; a++) acc[a] += gv * nb_to_float<T>(dd[a * C + c]);
        }
    }
    for (uint32_t a = 0; a < D; a++) grad_inputs[(size_t)b * D + a] = nb_from_float<T>(acc[a]);
}

// ---- total-variation gradient (GridEncoder.grad_total_variation, grid.py:171-192): for the cell every point falls in, at
// every level: grad[cell, c] += weight / (2 D) * sum_n (v - v_n) / sqrt(sum_n (v - v_n)^2 + 1e-9) over the cell's 2 D axis
// neighbours n that exist (coordinate + 1 while the coordinate is below the level's resolution, coordinate - 1 while it is
// above 0).  One thread per point; the 2 D + 1 rows are gathered once and reused for all channels.
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(128)
k_gen_tv(const float *__restrict__ inputs, const float *__restrict__ table, float *__restrict__ grad, const int32_t *__restrict__ offsets,
         float weight, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align_corners) {
    __shared__ GenLevel<D> lv[kGenMaxLevels];
    gen_fill_levels<D>(lv, offsets, L, S, H, gridtype, align_corners);
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = inputs[(size_t)b * D + d];
        if (x[d] < 0.0f || x[d] > 1.0f) return;
    }
    const float shift = align_corners ? 0.0f : 0.5f, k = weight / (2 * D);
    for (uint32_t l = 0; l < L; l++) {
        const GenLevel<D> g = lv[l];
        uint32_t cell[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) cell[d] = (uint32_t)floorf(x[d] * g.scale + shift);
        const float *rows = table + (size_t)g.offset * C;
        const size_t self = (size_t)gen_row<D>(g, cell) * C;
        float v[C], lin[C], sq[C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) { v[c] = rows[self + c]; lin[c] = 0.0f; sq[c] = 0.0f; }
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            const uint32_t at = cell[d];
#pragma unroll
            for (int side = 0; side < 2; side++) {
                if (side == 0 ? !(at < g.resolution) : !(at > 0)) continue;
                cell[d] = side == 0 ? at + 1 : at - 1;
                const float *nb = rows + (size_t)gen_row<D>(g, cell) * C;
#pragma unroll
                for (uint32_t c = 0; c < C; c++) { const float dv = v[c] - nb[c]; lin[c] += dv; sq[c] += dv * dv; }
            }
            cell[d] = at;
        }
        float *out = grad + (size_t)g.offset * C + self;
#pragma unroll
        for (uint32_t c = 0; c < C; c++) atomicAdd(out + c, k * lin[c] * rsqrtf(sq[c] + 1e-9f));
    }
}

}  // namespace
