// field_common.cuh -- pieces shared by the field-network kernels (field_mlp.cu: MLP-only forward / backward; field_fused.cu:
// grid gather + MLP in one kernel): the packed weight images, the operand-tile helpers, the epilogues and the view-direction
// embedding.  See field_mlp.cu for the layer mapping and the numerics contract.
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include "wgrad_reduce.cuh"

// Kernel status word (include/nerf_b200.h: nb200_set_kernel_status_word): a device word of the CURRENT device into which the
// field kernels OR a bit when one of their bounded mbarrier waits times out (a descriptor / protocol mistake must neither
// hang the GPU nor pass silently).  Without a registered word the forward kernels poison sigma[0] and the backward kernel
// d_x_en[0] with NaN instead.  Defined in field_mlp.cu.
uint32_t *nb_kernel_status_word();

namespace {

// ---- packed weight images (bytes) ---------------------------------------------------------------------
// forward: B operand rows = output neuron n, cols = input k
constexpr uint32_t F_W1V = 0;        // 64 rows: cols 0..31 trunk layer 0 (W1), cols 32..63 colour layer 0 view part
constexpr uint32_t F_W2 = 8192;      // trunk layer 1   [64 x 64]
constexpr uint32_t F_W3 = 16384;     // trunk layer 2   [64 x 64]
constexpr uint32_t F_WD1 = 24576;    // density layer 0 [64 x 64]
constexpr uint32_t F_WR1F = 32768;   // colour layer 0, fea part [64 x 64]
constexpr uint32_t F_WD2 = 40960;    // density layer 1 [16 x 64]
constexpr uint32_t F_WR2 = 43008;    // colour layer 1  [16 x 64]
constexpr uint32_t F_BYTES = 45056;
// backward (dgrad): B operand rows = input k, cols = output neuron n (transposes)
constexpr uint32_t B_W1T = 0;        // [32 x 64]
constexpr uint32_t B_W2T = 4096;     // [64 x 64]
constexpr uint32_t B_W3T = 12288;
constexpr uint32_t B_WD1T = 20480;
constexpr uint32_t B_WR1FT = 28672;
constexpr uint32_t B_W16T = 36864;   // 64 rows: cols 0..15 = Wr2^T, cols 16..31 = Wd2^T
constexpr uint32_t B_BYTES = 45056;

// (the flat tcnn-layout parameter offsets T_* / D_* / R_* and the slab layout live in wgrad_reduce.cuh)

__device__ __forceinline__ void put(uint8_t *img, uint32_t row, uint32_t col, float v) {
    *reinterpret_cast<__half *>(img + umma::sw128_offset(row, col >> 3) + (col & 7u) * 2) = __float2half_rn(v);
}


// ---- shared-memory map of the forward kernel ------------------------------------------------------------
constexpr uint32_t S_W = 0;                       // weights, F_BYTES
constexpr uint32_t S_XV = F_BYTES;                // cols 0..31 x_en, cols 32..63 [view_en(27) | ones(5)]
constexpr uint32_t S_H0 = S_XV + 16384;
constexpr uint32_t S_H1 = S_H0 + 16384;
constexpr uint32_t S_FEA = S_H1 + 16384;
constexpr uint32_t S_FWD_BYTES = S_FEA + 16384;   // 110592
constexpr uint32_t kTmemCols = 64;


__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ---- asynchronous global -> shared copies (LDGSTS): the activation tiles of the NEXT tile stream in behind the MMAs
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src, bool valid) {
    const uint32_t sz = valid ? 16u : 0u;                   // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// K-major A/B descriptor of K-step ks (16 halves) inside a 64-wide swizzled tile
__device__ __forceinline__ uint64_t kdesc(uint32_t tile_addr, uint32_t ks) {
    return umma::make_desc(tile_addr + ks * 32, 16, 1024, umma::kLayoutSW128);
}

constexpr uint32_t kStatusFieldFwdTimeout = 1u, kStatusFieldBwdTimeout = 2u, kStatusFieldFusedTimeout = 4u;

// accumulator row (64 fp32) -> optional ReLU -> fp16 -> row of a swizzled tile (+ optional global copy)
template <bool kRelu>
__device__ __noinline__ void epilogue_row64(uint32_t tmem_row_addr, uint8_t *tile, uint32_t row, __half *gdst) {
    uint32_t a[32], b[32];
    umma::tmem_ld32(tmem_row_addr, a);
    umma::tmem_ld32(tmem_row_addr + 32, b);
    const uint32_t tile_s = umma::smem_u32(tile);
    umma::tmem_ld_wait();
#pragma unroll
    for (uint32_t c = 0; c < 8; c++) {
        uint32_t *src = (c < 4) ? (a + c * 8) : (b + (c - 4) * 8);
        uint32_t q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            q[j] = pack_h2(__uint_as_float(src[2 * j]), __uint_as_float(src[2 * j + 1]));
            if (kRelu) {        // ReLU on the packed pair (rounding is monotone and max(., 0) drops NaN: same result as before rounding)
                const __half2 r2 = __hmax2(*reinterpret_cast<const __half2 *>(&q[j]), __float2half2_rn(0.0f));
                q[j] = *reinterpret_cast<const uint32_t *>(&r2);
            }
        }
        const uint4 pk = make_uint4(q[0], q[1], q[2], q[3]);
        // shared-space store on the 32-bit shared address (a generic ST.E pays an address-space check per access)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tile_s + umma::sw128_offset(row, c)), "r"(pk.x), "r"(pk.y),
                     "r"(pk.z), "r"(pk.w) : "memory");
        if (gdst) *reinterpret_cast<uint4 *>(gdst + c * 8) = pk;
    }
}

// direction encoding get_embedder(4) of one row (nerf/base.py:42-77): e = [d, sin(2^k d), cos(2^k d)] (27 lanes, input
// first, then per frequency sin then cos) followed by five 1.0 lanes (the padded inputs 91..95 of the colour head).
// sin / cos of 2^k d: one sincosf per component (the accurate libdevice routine, ~1 ulp; |d| <= 1 for unit directions, so
// its fast path), then the double-angle identities.  Three doublings multiply the base error by 8: measured max deviation
// from sin/cos(2^k d) 1e-6 (the fast intrinsic __sincosf gave 4.9e-6 and was replaced); asserted <= 2e-6 through
// nb200_freq_embed in tests/test_gpu_field_mlp.py -- two orders below the fp16 rounding of the operand.  One sincosf body
// instead of twelve in the instruction stream.
__device__ __forceinline__ void view_embed32(float d0, float d1, float d2, float (&e)[32]) {
    e[0] = d0; e[1] = d1; e[2] = d2;
    float s0, c0, s1, c1, s2, c2;
    sincosf(d0, &s0, &c0); sincosf(d1, &s1, &c1); sincosf(d2, &s2, &c2);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        e[3 + 6 * k] = s0; e[4 + 6 * k] = s1; e[5 + 6 * k] = s2;
        e[6 + 6 * k] = c0; e[7 + 6 * k] = c1; e[8 + 6 * k] = c2;
        const float t0 = 2.0f * s0 * c0, t1 = 2.0f * s1 * c1, t2 = 2.0f * s2 * c2;
        c0 = 1.0f - 2.0f * s0 * s0; c1 = 1.0f - 2.0f * s1 * s1; c2 = 1.0f - 2.0f * s2 * s2;
        s0 = t0; s1 = t1; s2 = t2;
    }
#pragma unroll
    for (int j = 27; j < 32; j++) e[j] = 1.0f;
}

// ... written as chunks 4..7 of row `row` of the XV tile
__device__ __noinline__ void write_view_chunks(uint8_t *xv, uint32_t row, float d0, float d1, float d2, bool valid) {
    float e[32];
    view_embed32(d0, d1, d2, e);
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        uint4 pk = make_uint4(pack_h2(e[c * 8], e[c * 8 + 1]), pack_h2(e[c * 8 + 2], e[c * 8 + 3]),
                              pack_h2(e[c * 8 + 4], e[c * 8 + 5]), pack_h2(e[c * 8 + 6], e[c * 8 + 7]));
        if (!valid) pk = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(xv + umma::sw128_offset(row, 4 + c)) = pk;
    }
}


// cuTensorMapEncodeTiled through the runtime's driver entry point table (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

}  // namespace
