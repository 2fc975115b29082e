// gridencoder.cu -- multiresolution hash / tiled grid encoding for sm_100a.
//
// Replaces gridencoder/src/gridencoder.cu of the reference (per-entry file:line map in include/nerf_b200.h).
//
// Two families of kernels, both ONE THREAD PER POINT walking the levels (the reference launches a thread per (point, level)):
//   * generic  <T, D, C> (grid_generic.cuh): any D in 2..5, C in {1,2,4,8}, both output layouts, input gradients (every
//     corner gathered once, features and all D derivatives from the same loads), total-variation gradient.
//   * d3c2 fast path (D=3, C=2, layout [B, L*C], L <= 32 -- the only shape nerf/network_grid.py:95 and
//     nerf/encoding.py:55-58 ever build):
//       - a warp holds 32 consecutive samples of a ray: at coarse levels their 8 corners fall in the same
//         cells and the gathers coalesce into a few sectors,
//       - each thread owns one [L*C] output row and writes it with 16/32-byte vector stores straight in the
//         MLP-ready layout (no [L,B,C] -> [B,L*C] permute copy, grid.py:63),
//       - per-level constants (offset, size, strides, scale) are computed once per CTA into shared memory,
//       - backward: adjacent lanes that share a base cell are summed with a segmented shuffle reduction and
//         only the run head issues the (float2) atomic -- the warp-aggregated scatter that replaces the
//         reference's per-element __half2 atomics (gridencoder.cu:324-337).  Accumulation is fp32.
//
// Index arithmetic is uint32 with wrap-around exactly as gridencoder.cu:50-84; scale is computed on the
// device with the reference's expression exp2f(level * S) * H - 1.0f so fine-level positions are identical.
#include "common.cuh"
#include "grid_d3c2.cuh"
#include "grid_generic.cuh"
#include "wgrad_reduce.cuh"
#include <stdlib.h>

namespace {



// one thread per point; outputs [B, L*2].  TE = table storage type, T = value / output type.
template <typename TE, typename T>
__global__ void __launch_bounds__(256)
k_grid_fwd_d3c2(const float *__restrict__ inputs, const TE *__restrict__ grid, const int32_t *__restrict__ offsets,
                T *__restrict__ outputs, uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H,
                uint32_t gridtype, bool align_corners, uint32_t interp, InXform xf) {
    __shared__ LevelInfo info[kMaxFastLevels];
    if (xf.count_dev) B = min(B, (uint32_t)max(*xf.count_dev, 0));
    if (blockIdx.x * blockDim.x >= B) return;
    ge_fill_level_info(info, offsets, max_level, S, H, gridtype, align_corners);
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float x0 = xf(inputs[(size_t)b * 3]), x1 = xf(inputs[(size_t)b * 3 + 1]), x2 = xf(inputs[(size_t)b * 3 + 2]);
    const bool oob = (x0 < 0 || x0 > 1) || (x1 < 0 || x1 > 1) || (x2 < 0 || x2 > 1);
    T *out = outputs + (size_t)b * L * 2;
    const float half_off = align_corners ? 0.0f : 0.5f;

    for (uint32_t l0 = 0; l0 < max_level; l0 += 4) {
        float res[8];
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
            const uint32_t l = l0 + j;
            float r0 = 0.0f, r1 = 0.0f;
            if (l < max_level && !oob)
                ge_level_gather<TE, (sizeof(TE) == 4 && sizeof(T) == 2)>(info[l], grid, x0, x1, x2, half_off, interp, r0, r1);
            res[j * 2] = r0; res[j * 2 + 1] = r1;
        }
        // rows are (L*2) elements; L*2*sizeof(T) is a multiple of 16 bytes only when L % 4 == 0 (f16) / L % 2 == 0 (f32)
        const uint32_t nvalid = min(4u, L - l0);   // levels >= max_level but < L are the caller's to zero (grid.py:52)
        if (l0 + 4 <= max_level && (L % 4 == 0)) {
            if constexpr (sizeof(T) == 2) {
                __half2 h0 = __floats2half2_rn(res[0], res[1]), h1 = __floats2half2_rn(res[2], res[3]);
                __half2 h2 = __floats2half2_rn(res[4], res[5]), h3 = __floats2half2_rn(res[6], res[7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
                pk.z = *reinterpret_cast<uint32_t *>(&h2); pk.w = *reinterpret_cast<uint32_t *>(&h3);
                *reinterpret_cast<uint4 *>(out + l0 * 2) = pk;
            } else {
                *reinterpret_cast<float4 *>(out + l0 * 2) = make_float4(res[0], res[1], res[2], res[3]);
                *reinterpret_cast<float4 *>(out + l0 * 2 + 4) = make_float4(res[4], res[5], res[6], res[7]);
            }
        } else {
            for (uint32_t j = 0; j < nvalid && l0 + j < max_level; j++) {
                out[(l0 + j) * 2] = nb_from_float<T>(res[j * 2]);
                out[(l0 + j) * 2 + 1] = nb_from_float<T>(res[j * 2 + 1]);
            }
        }
    }
}

// one thread per point; grad [B, L*2]; fp32 float2 atomics; optional warp aggregation
template <typename T, bool kAgg>
__global__ void __launch_bounds__(256)
k_grid_bwd_d3c2(const T *__restrict__ grad, const float *__restrict__ inputs, const int32_t *__restrict__ offsets,
                float *__restrict__ grad_grid, uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H,
                uint32_t gridtype, bool align_corners, uint32_t interp, InXform xf, uint32_t agg_max_heads,
                NbWgradRed red = NbWgradRed{}) {
    __shared__ LevelInfo info[kMaxFastLevels];
    // fused train step: the LAST red.blocks blocks of this launch are the slab reduction of the field backward's weight
    // gradients (wgrad_reduce.cuh) -- independent of the scatter; dispatched last, they run in the scatter's draining tail
    if (blockIdx.x >= gridDim.x - red.blocks) {
        __shared__ float part[4][64];
        wgrad_reduce_block(red, blockIdx.x - (gridDim.x - red.blocks), part);
        return;
    }
    const uint32_t bid = blockIdx.x;
    if (xf.count_dev) B = min(B, (uint32_t)max(*xf.count_dev, 0));
    if (bid * blockDim.x >= B) return;
    ge_fill_level_info(info, offsets, max_level, S, H, gridtype, align_corners);
    __syncthreads();
    const uint32_t b = bid * blockDim.x + threadIdx.x;
    const uint32_t lane = nb_lane();
    const bool inb = b < B;            // keep whole warps alive for the shuffles
    float x0 = -1.0f, x1 = -1.0f, x2 = -1.0f;
    if (inb) { x0 = xf(inputs[(size_t)b * 3]); x1 = xf(inputs[(size_t)b * 3 + 1]); x2 = xf(inputs[(size_t)b * 3 + 2]); }
    const bool oob = (x0 < 0 || x0 > 1) || (x1 < 0 || x1 > 1) || (x2 < 0 || x2 > 1);   // !inb => oob
    const T *g = grad + (size_t)(inb ? b : 0) * L * 2;
    const float half_off = align_corners ? 0.0f : 0.5f;

    for (uint32_t l = xf.level_begin; l < max_level; l++) {
        const LevelInfo li = info[l];
        float g0v = 0.0f, g1v = 0.0f;
        if (!oob) { g0v = nb_to_float<T>(g[l * 2]); g1v = nb_to_float<T>(g[l * 2 + 1]); }
        float p0 = x0 * li.scale + half_off, p1 = x1 * li.scale + half_off, p2 = x2 * li.scale + half_off;
        const uint32_t c0 = (uint32_t)floorf(p0), c1 = (uint32_t)floorf(p1), c2 = (uint32_t)floorf(p2);
        p0 -= (float)c0; p1 -= (float)c1; p2 -= (float)c2;
        if (interp == 1) { p0 = ge_smoothstep(p0); p1 = ge_smoothstep(p1); p2 = ge_smoothstep(p2); }
        float2 v[8];
#pragma unroll
        for (uint32_t idx = 0; idx < 8; idx++) {
            float w = 1;
            w *= (idx & 1u) ? p0 : 1 - p0;
            w *= (idx & 2u) ? p1 : 1 - p1;
            w *= (idx & 4u) ? p2 : 1 - p2;
            v[idx] = make_float2(w * g0v, w * g1v);
        }
        bool emit = !oob;
        if (kAgg) {
            // runs of adjacent lanes (consecutive samples of a ray) that share the base cell share all 8 corners
            const uint32_t q0 = __shfl_up_sync(0xffffffffu, c0, 1), q1 = __shfl_up_sync(0xffffffffu, c1, 1),
                           q2 = __shfl_up_sync(0xffffffffu, c2, 1);
            const bool poob = __shfl_up_sync(0xffffffffu, (int)oob, 1) != 0;
            const bool head = (lane == 0) || oob || poob || (q0 != c0) || (q1 != c1) || (q2 != c2);
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            if ((uint32_t)__popc(heads) <= agg_max_heads) {     // warp-uniform: enough atomics disappear to pay for the shuffles
                const uint32_t nh = (lane == 31) ? 0u : (heads >> (lane + 1));
                const uint32_t run_last = nh ? lane + (uint32_t)__ffs(nh) - 1u : 31u;
                // log-step segmented reduction, only as many rounds as the longest run of this warp needs
                const uint32_t max_run = __reduce_max_sync(0xffffffffu, head ? run_last - lane + 1u : 0u);
                for (uint32_t o = 1; o < max_run; o <<= 1) {
                    const bool take = (lane + o) <= run_last;
#pragma unroll
                    for (uint32_t idx = 0; idx < 8; idx++) {
                        const float ax = __shfl_down_sync(0xffffffffu, v[idx].x, o);
                        const float ay = __shfl_down_sync(0xffffffffu, v[idx].y, o);
                        if (take) { v[idx].x += ax; v[idx].y += ay; }
                    }
                }
                emit = head && !oob;
            }
        }
        if (emit) {
            // The two corners of a cell that differ in x only sit in rows r and r ^ 1 for half of all cells: on a hashed
            // level x enters the hash with prime 1, so for even x the neighbour's row is the same hash with bit 0 flipped; on
            // a dense level the rows are r and r + 1.  Such a pair is one aligned 16-byte word of the gradient table: ONE
            // RED.ADD.F32x4 instead of two F32x2 -- a quarter fewer L2 reductions per sample (measured: 128 -> 107 us; the same
            // trick on the forward gathers, a 16-byte load plus a predicated 8-byte one, was slower: 52 -> 58 us)
            float *lg = grad_grid + (size_t)li.offset * 2;
#pragma unroll
            for (uint32_t pr = 0; pr < 4; pr++) {
                const uint32_t ra = ge_row_d3(li, c0, c1 + (pr & 1u), c2 + (pr >> 1));
                const uint32_t rb = ge_row_d3(li, c0 + 1u, c1 + (pr & 1u), c2 + (pr >> 1));
                const float2 va = v[2 * pr], vb = v[2 * pr + 1];
                if ((ra ^ rb) == 1u) {
                    const float4 q = (ra & 1u) ? make_float4(vb.x, vb.y, va.x, va.y) : make_float4(va.x, va.y, vb.x, vb.y);
                    atomicAdd(reinterpret_cast<float4 *>(lg + (size_t)(ra & ~1u) * 2), q);
                } else {
                    atomicAdd(reinterpret_cast<float2 *>(lg + (size_t)ra * 2), va);
                    atomicAdd(reinterpret_cast<float2 *>(lg + (size_t)rb * 2), vb);
                }
            }
        }
    }
}

__global__ void k_level_scales(float *scales, uint32_t L, float S, uint32_t H) {
    const uint32_t l = threadIdx.x + blockIdx.x * blockDim.x;
    if (l < L) scales[l] = ge_level_scale(l, S, H);
}

__global__ void k_cast_f32_f16(const float *__restrict__ src, __half *__restrict__ dst, uint64_t n) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + i));
        __half2 a = __floats2half2_rn(v.x, v.y), b2 = __floats2half2_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&a); pk.y = *reinterpret_cast<uint32_t *>(&b2);
        *reinterpret_cast<uint2 *>(dst + i) = pk;
    } else {
        for (uint64_t j = i; j < n; j++) dst[j] = __float2half_rn(src[j]);
    }
}

// warp-aggregate a level only when at most this many of the 32 lanes start a new cell run (tunable: NB200_GE_AGG_MAXHEADS)
static uint32_t ge_agg_max_heads() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("NB200_GE_AGG_MAXHEADS"); v = e ? atoi(e) : 32; if (v < 0 || v > 32) v = 32; }
    return (uint32_t)v;
}

// ---- dispatch helpers ------------------------------------------------------------------------------
template <typename T, uint32_t D>
int launch_fwd_c(const float *inputs, const T *emb, const int32_t *offsets, T *out, uint32_t B, uint32_t C, uint32_t L,
                 uint32_t max_level, float S, uint32_t H, T *dy_dx, uint32_t gridtype, bool ac, uint32_t interp,
                 int layout, cudaStream_t st) {
    if (max_level > kGenMaxLevels) return NB200_E_BAD_ARG;
    const uint32_t grid = nb_div_up(B, 128);
#define NB_GEN_FWD(CC) k_gen_encode<T, D, CC><<<grid, 128, 0, st>>>(inputs, emb, offsets, out, dy_dx, B, L, max_level, S, H, gridtype, ac, interp, layout)
    switch (C) {
        case 1: NB_GEN_FWD(1); break;
        case 2: NB_GEN_FWD(2); break;
        case 4: NB_GEN_FWD(4); break;
        case 8: NB_GEN_FWD(8); break;
        default: return NB200_E_BAD_DIM;
    }
#undef NB_GEN_FWD
    return 0;
}

template <typename T>
int launch_fwd(const float *inputs, const T *emb, const int32_t *offsets, T *out, uint32_t B, uint32_t D, uint32_t C,
               uint32_t L, uint32_t max_level, float S, uint32_t H, T *dy_dx, uint32_t gridtype, bool ac,
               uint32_t interp, int layout, cudaStream_t st) {
    if (D == 3 && C == 2 && layout == NB200_LAYOUT_BLC && !dy_dx && L <= kMaxFastLevels) {
        k_grid_fwd_d3c2<T, T><<<nb_div_up(B, 256), 256, 0, st>>>(inputs, emb, offsets, out, B, L, max_level, S, H, gridtype, ac, interp, InXform{0.0f, 0.0f, nullptr});
        return 0;
    }
    switch (D) {
        case 2: return launch_fwd_c<T, 2>(inputs, emb, offsets, out, B, C, L, max_level, S, H, dy_dx, gridtype, ac, interp, layout, st);
        case 3: return launch_fwd_c<T, 3>(inputs, emb, offsets, out, B, C, L, max_level, S, H, dy_dx, gridtype, ac, interp, layout, st);
        case 4: return launch_fwd_c<T, 4>(inputs, emb, offsets, out, B, C, L, max_level, S, H, dy_dx, gridtype, ac, interp, layout, st);
        case 5: return launch_fwd_c<T, 5>(inputs, emb, offsets, out, B, C, L, max_level, S, H, dy_dx, gridtype, ac, interp, layout, st);
        default: return NB200_E_BAD_DIM;
    }
}

template <typename T, uint32_t D>
int launch_bwd_c(const T *grad, const float *inputs, const int32_t *offsets, float *gg, uint32_t B, uint32_t C,
                 uint32_t L, uint32_t max_level, float S, uint32_t H, uint32_t gridtype, bool ac, uint32_t interp,
                 int layout, cudaStream_t st) {
    if (max_level > kGenMaxLevels) return NB200_E_BAD_ARG;
    const uint32_t grid = nb_div_up(B, 128);
#define NB_GEN_BWD(CC) k_gen_scatter<T, D, CC><<<grid, 128, 0, st>>>(grad, inputs, offsets, gg, B, L, max_level, S, H, gridtype, ac, interp, layout)
    switch (C) {
        case 1: NB_GEN_BWD(1); break;
        case 2: NB_GEN_BWD(2); break;
        case 4: NB_GEN_BWD(4); break;
        case 8: NB_GEN_BWD(8); break;
        default: return NB200_E_BAD_DIM;
    }
#undef NB_GEN_BWD
    return 0;
}

template <typename T>
int launch_bwd(const T *grad, const float *inputs, const int32_t *offsets, float *gg, uint32_t B, uint32_t D, uint32_t C,
               uint32_t L, uint32_t max_level, float S, uint32_t H, const T *dy_dx, T *grad_inputs, uint32_t gridtype,
               bool ac, uint32_t interp, int layout, int agg, cudaStream_t st) {
    int rc = 0;
    if (D == 3 && C == 2 && layout == NB200_LAYOUT_BLC && L <= kMaxFastLevels) {
        const uint32_t nblk = nb_div_up(B, 256);
        if (agg) k_grid_bwd_d3c2<T, true><<<nblk, 256, 0, st>>>(grad, inputs, offsets, gg, B, L, max_level, S, H, gridtype, ac, interp, InXform{0.0f, 0.0f, nullptr}, ge_agg_max_heads());
        else k_grid_bwd_d3c2<T, false><<<nblk, 256, 0, st>>>(grad, inputs, offsets, gg, B, L, max_level, S, H, gridtype, ac, interp, InXform{0.0f, 0.0f, nullptr}, 0);
    } else {
        switch (D) {
            case 2: rc = launch_bwd_c<T, 2>(grad, inputs, offsets, gg, B, C, L, max_level, S, H, gridtype, ac, interp, layout, st); break;
            case 3: rc = launch_bwd_c<T, 3>(grad, inputs, offsets, gg, B, C, L, max_level, S, H, gridtype, ac, interp, layout, st); break;
            case 4: rc = launch_bwd_c<T, 4>(grad, inputs, offsets, gg, B, C, L, max_level, S, H, gridtype, ac, interp, layout, st); break;
            case 5: rc = launch_bwd_c<T, 5>(grad, inputs, offsets, gg, B, C, L, max_level, S, H, gridtype, ac, interp, layout, st); break;
            default: return NB200_E_BAD_DIM;
        }
    }
    if (rc) return rc;
    if (dy_dx && grad_inputs)
        k_gen_input_grad<T><<<nb_div_up(B, 128), 128, 0, st>>>(grad, dy_dx, grad_inputs, B, D, C, L, layout);
    return 0;
}

template <uint32_t D>
int launch_tv_c(const float *inputs, const float *emb, float *grad, const int32_t *offsets, float weight, uint32_t B,
                uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    if (L > kGenMaxLevels) return NB200_E_BAD_ARG;
    const uint32_t grid = nb_div_up(B, 128);
#define NB_GEN_TV(CC) k_gen_tv<D, CC><<<grid, 128, 0, st>>>(inputs, emb, grad, offsets, weight, B, L, S, H, gridtype, ac)
    switch (C) {
        case 1: NB_GEN_TV(1); break;
        case 2: NB_GEN_TV(2); break;
        case 4: NB_GEN_TV(4); break;
        case 8: NB_GEN_TV(8); break;
        default: return NB200_E_BAD_DIM;
    }
#undef NB_GEN_TV
    return 0;
}

}  // namespace

extern "C" {

int nb200_grid_encode_forward(const float *inputs, const void *embeddings, const int32_t *offsets, void *outputs,
                              uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                              void *dy_dx, uint32_t gridtype, int align_corners, uint32_t interp,
                              int emb_dtype, int layout, void *stream) {
    if (B == 0 || max_level == 0) return 0;
    if (!inputs || !embeddings || !offsets || !outputs || max_level > L) return NB200_E_BAD_ARG;
    if (layout != NB200_LAYOUT_LBC && layout != NB200_LAYOUT_BLC) return NB200_E_BAD_ARG;
    int rc;
    if (emb_dtype == NB200_F32)
        rc = launch_fwd<float>(inputs, (const float *)embeddings, offsets, (float *)outputs, B, D, C, L, max_level, S, H,
                               (float *)dy_dx, gridtype, align_corners != 0, interp, layout, nb_stream(stream));
    else if (emb_dtype == NB200_F16)
        rc = launch_fwd<__half>(inputs, (const __half *)embeddings, offsets, (__half *)outputs, B, D, C, L, max_level, S,
                                H, (__half *)dy_dx, gridtype, align_corners != 0, interp, layout, nb_stream(stream));
    else if (emb_dtype == NB200_F32_AS_F16) {
        // fast path only: fp32 table, values rounded to fp16 on load, fp16 outputs
        if (!(D == 3 && C == 2 && layout == NB200_LAYOUT_BLC && !dy_dx && L <= kMaxFastLevels)) return NB200_E_BAD_DTYPE;
        k_grid_fwd_d3c2<float, __half><<<nb_div_up(B, 256), 256, 0, nb_stream(stream)>>>(
            inputs, (const float *)embeddings, offsets, (__half *)outputs, B, L, max_level, S, H, gridtype,
            align_corners != 0, interp, InXform{0.0f, 0.0f, nullptr});
        rc = 0;
    } else
        return NB200_E_BAD_DTYPE;
    if (rc) return rc;
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_grid_encode_backward(const void *grad, const float *inputs, const int32_t *offsets, float *grad_embeddings,
                               uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                               const void *dy_dx, void *grad_inputs, uint32_t gridtype, int align_corners,
                               uint32_t interp, int grad_dtype, int layout, int agg, void *stream) {
    if (B == 0 || max_level == 0) return 0;
    if (!grad || !inputs || !offsets || !grad_embeddings || max_level > L) return NB200_E_BAD_ARG;
    if (layout != NB200_LAYOUT_LBC && layout != NB200_LAYOUT_BLC) return NB200_E_BAD_ARG;
    int rc;
    if (grad_dtype == NB200_F32)
        rc = launch_bwd<float>((const float *)grad, inputs, offsets, grad_embeddings, B, D, C, L, max_level, S, H,
                               (const float *)dy_dx, (float *)grad_inputs, gridtype, align_corners != 0, interp, layout,
                               agg, nb_stream(stream));
    else if (grad_dtype == NB200_F16)
        rc = launch_bwd<__half>((const __half *)grad, inputs, offsets, grad_embeddings, B, D, C, L, max_level, S, H,
                                (const __half *)dy_dx, (__half *)grad_inputs, gridtype, align_corners != 0, interp,
                                layout, agg, nb_stream(stream));
    else
        return NB200_E_BAD_DTYPE;
    if (rc) return rc;
    NB_LAUNCH_CHECK();
    return 0;
}

// fused train step (D = 3, C = 2): raw sample positions in [-bound, bound], fp32 master table rounded to fp16 per load,
// fp16 [M, L*2] features / feature gradients, rows limited by a device-side count
int nb200_fs_encode_forward(const float *xyz, float bound, const float *table, const int32_t *offsets, void *x_en,
                            uint32_t M_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                            uint32_t interp, const int32_t *count_dev, void *stream) {
    if (M_cap == 0 || L == 0) return 0;
    if (!xyz || !table || !offsets || !x_en || L > kMaxFastLevels || !(bound > 0.0f)) return NB200_E_BAD_ARG;
    const InXform xf{bound, 1.0f / (2.0f * bound), count_dev};
    k_grid_fwd_d3c2<float, __half><<<nb_div_up(M_cap, 256), 256, 0, nb_stream(stream)>>>(
        xyz, table, offsets, (__half *)x_en, M_cap, L, L, S, H, gridtype, align_corners != 0, interp, xf);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_fs_encode_backward(const void *d_x_en, const float *xyz, float bound, const int32_t *offsets, float *grad_table,
                             uint32_t M_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                             uint32_t interp, const int32_t *count_dev, void *stream) {
    return nb200_fs_encode_backward_levels(d_x_en, xyz, bound, offsets, grad_table, M_cap, L, S, H, gridtype, align_corners, interp,
                                           count_dev, 0, L, stream);
}

// the same scatter restricted to levels [level_begin, level_end)
int nb200_fs_encode_backward_levels(const void *d_x_en, const float *xyz, float bound, const int32_t *offsets, float *grad_table,
                                    uint32_t M_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                    uint32_t interp, const int32_t *count_dev, uint32_t level_begin, uint32_t level_end,
                                    void *stream) {
    return nb_fs_encode_backward_red(d_x_en, xyz, bound, offsets, grad_table, M_cap, L, S, H, gridtype, align_corners, interp,
                                     count_dev, level_begin, level_end, nullptr, stream);
}

int nb200_grad_total_variation(const float *inputs, const float *embeddings, float *grad, const int32_t *offsets,
                               float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                               uint32_t gridtype, int align_corners, void *stream) {
    if (B == 0 || L == 0) return 0;
    if (!inputs || !embeddings || !grad || !offsets) return NB200_E_BAD_ARG;
    int rc;
    cudaStream_t st = nb_stream(stream);
    switch (D) {
        case 2: rc = launch_tv_c<2>(inputs, embeddings, grad, offsets, weight, B, C, L, S, H, gridtype, align_corners != 0, st); break;
        case 3: rc = launch_tv_c<3>(inputs, embeddings, grad, offsets, weight, B, C, L, S, H, gridtype, align_corners != 0, st); break;
        case 4: rc = launch_tv_c<4>(inputs, embeddings, grad, offsets, weight, B, C, L, S, H, gridtype, align_corners != 0, st); break;
        case 5: rc = launch_tv_c<5>(inputs, embeddings, grad, offsets, weight, B, C, L, S, H, gridtype, align_corners != 0, st); break;
        default: return NB200_E_BAD_DIM;
    }
    if (rc) return rc;
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_grid_level_scales(float *scales, uint32_t L, float S, uint32_t H, void *stream) {
    if (L == 0) return 0;
    k_level_scales<<<nb_div_up(L, 64), 64, 0, nb_stream(stream)>>>(scales, L, S, H);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_cast_f32_to_f16(const float *src, void *dst, uint64_t n, void *stream) {
    if (n == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) || (reinterpret_cast<uintptr_t>(dst) & 7u)) return NB200_E_BAD_ARG;
    k_cast_f32_f16<<<nb_div_up((n + 3) / 4, 256), 256, 0, nb_stream(stream)>>>(src, (__half *)dst, n);
    NB_LAUNCH_CHECK();
    return 0;
}

const char *nb200_error_string(int code) {
    if (code == 0) return "ok";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    switch (code) {
        case NB200_E_BAD_DIM: return "GridEncoding: C must be 1, 2, 4, or 8 and D must be 2, 3, 4 or 5.";
        case NB200_E_BAD_DTYPE: return "unsupported dtype (f32 and f16 are built)";
        case NB200_E_BAD_ARG: return "bad argument (null pointer, misaligned buffer or inconsistent sizes)";
        case NB200_E_SCRATCH: return "scratch buffer too small";
        default: return "unknown nb200 error";
    }
}

int nb200_version(void) { return 1; }

}  // extern "C"

int nb_fs_encode_backward_red(const void *d_x_en, const float *xyz, float bound, const int32_t *offsets, float *grad_table,
                              uint32_t M_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                              uint32_t interp, const int32_t *count_dev, uint32_t level_begin, uint32_t level_end,
                              const NbWgradRed *red, void *stream) {
    const NbWgradRed r = red ? *red : NbWgradRed{};
    if ((M_cap == 0 || L == 0 || level_begin >= level_end) && r.blocks == 0) return 0;
    if (!d_x_en || !xyz || !offsets || !grad_table || L > kMaxFastLevels || !(bound > 0.0f) || level_end > L) return NB200_E_BAD_ARG;
    const InXform xf{bound, 1.0f / (2.0f * bound), count_dev, level_begin};
    k_grid_bwd_d3c2<__half, true><<<nb_div_up(M_cap, 256) + r.blocks, 256, 0, nb_stream(stream)>>>(
        (const __half *)d_x_en, xyz, offsets, grad_table, M_cap, L, level_end, S, H, gridtype, align_corners != 0, interp, xf,
        ge_agg_max_heads(), r);
    NB_LAUNCH_CHECK();
    return 0;
}
