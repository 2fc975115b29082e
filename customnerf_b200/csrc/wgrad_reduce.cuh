// wgrad_reduce.cuh -- the per-CTA weight-gradient slabs of the field backward kernel (field_mlp.cu) and their reduction into
// the flat tcnn-layout gradients.  Shared by field_mlp.cu (the slab flush, the stand-alone reduce kernel) and gridencoder.cu
// (the fused train step runs the reduction as extra blocks of the table-scatter launch: it depends on nothing but the slabs
// and only the optimiser depends on it, so its launch gap + run time -- 10 us at configs[1] -- leave the critical path).
#pragma once
#include "common.cuh"
#include "adam.cuh"

// what a reduction needs (plain struct at global scope: it crosses translation units as a kernel / function argument)
struct NbWgradRed {
    const float *slabs;           // [grid][kWgradFloats]
    const int32_t *count_dev;     // rows the backward really processed (device), or null
    float *g_trunk, *g_density, *g_rgb;
    uint32_t *scaler;             // loss-scaler words or null
    uint32_t grid, M, blocks;     // CTAs of the backward launch, its row capacity, reduce blocks (0: disabled)
};
// nb200_field_backward with the reduction optional (field_mlp.cu); reduce = false leaves the slabs in wg_scratch
extern "C" int nb_field_backward_launch(const float *d_sigma, const float *d_rgba, const float *sigma_arg, const void *rgba,
                                        const void *x_en, const float *dirs, const void *act, const void *bwd_img, void *d_x_en,
                                        float *g_trunk, float *g_density, float *g_rgb, uint32_t M, const int32_t *count_dev,
                                        float *wg_scratch, uint32_t *scaler, bool reduce, void *stream);
// nb200_fs_encode_backward_levels whose launch also carries the blocks of a slab reduction (gridencoder.cu); red may be null
int nb_fs_encode_backward_red(const void *d_x_en, const float *xyz, float bound, const int32_t *offsets, float *grad_table,
                              uint32_t M_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                              uint32_t interp, const int32_t *count_dev, uint32_t level_begin, uint32_t level_end,
                              const NbWgradRed *red, void *stream);
uint32_t nb_wgrad_reduce_blocks();                  // blocks of 256 threads a reduction takes
uint32_t nb_field_backward_grid(uint32_t M);        // CTAs nb200_field_backward launches for M rows on the current device

namespace {

// flat tcnn-layout parameter offsets (elements)
constexpr uint32_t T_W1 = 0, T_W2 = 64 * 32, T_W3 = 64 * 32 + 64 * 64;      // trunk:   [64x32][64x64][64x64]
constexpr uint32_t D_W1 = 0, D_W2 = 64 * 64;                                 // density: [64x64][16x64]
constexpr uint32_t R_W1 = 0, R_W2 = 64 * 96;                                 // colour:  [64x96][16x64]
constexpr uint32_t kTrunkFloats = 64 * 32 + 2 * 64 * 64, kDensityFloats = 64 * 64 + 16 * 64, kRgbFloats = 64 * 96 + 16 * 64;
constexpr uint32_t kWgradFloats = kTrunkFloats + kDensityFloats + kRgbFloats;   // 22528 floats per slab
// per-CTA weight-gradient slab: one column-major block per accumulator (floats)
constexpr uint32_t kSlabW1 = 0, kSlabW2 = kSlabW1 + 32 * 64, kSlabW3 = kSlabW2 + 64 * 64, kSlabPair = kSlabW3 + 64 * 64,
                   kSlabR1V = kSlabPair + 64 * 128, kSlabR2 = kSlabR1V + 32 * 64, kSlabD2 = kSlabR2 + 64 * 16;
static_assert(kSlabD2 + 64 * 16 == kWgradFloats, "slab blocks must tile kWgradFloats");

// One block (256 threads = 64 slab elements x 4 slab groups; a warp reads 128 contiguous bytes of every slab it visits) of the
// reduction slabs -> += flat gradients.  Padded output rows of the two heads (tcnn pads 1 -> 16 and 4 -> 16 outputs) are
// dropped.  `part`: 4 x 64 floats of shared memory.
__device__ __forceinline__ void wgrad_reduce_block(const NbWgradRed &r, uint32_t block, float (*part)[64]) {
    const uint32_t Mrows = r.count_dev ? min(r.M, (uint32_t)max(*r.count_dev, 0)) : r.M;
    const uint32_t nslab = min(r.grid, (Mrows + 127) / 128);
    const uint32_t pl = threadIdx.x & 63u, cg = threadIdx.x >> 6, i = block * 64 + pl;
    float acc = 0.0f;
    if (i < kWgradFloats) {
#pragma unroll 8
        for (uint32_t c = cg; c < nslab; c += 4) acc += __ldg(r.slabs + (size_t)c * kWgradFloats + i);
    }
    part[cg][pl] = acc;
    __syncthreads();
    if (cg != 0 || i >= kWgradFloats || nslab == 0) return;
    const float sum = (part[0][pl] + part[1][pl]) + (part[2][pl] + part[3][pl]);
    // slab element -> parameter: block, column c, row n (output neuron)
    float *dst = nullptr;
    if (i < kSlabW2) { const uint32_t c = i / 64, n = i % 64; dst = r.g_trunk + T_W1 + n * 32 + c; }
    else if (i < kSlabW3) { const uint32_t e = i - kSlabW2, c = e / 64, n = e % 64; dst = r.g_trunk + T_W2 + n * 64 + c; }
    else if (i < kSlabPair) { const uint32_t e = i - kSlabW3, c = e / 64, n = e % 64; dst = r.g_trunk + T_W3 + n * 64 + c; }
    else if (i < kSlabR1V) {        // rows 0..63 = colour layer 0 (fea columns 27..90), rows 64..127 = density layer 0
        const uint32_t e = i - kSlabPair, c = e / 128, n = e % 128;
        dst = n < 64 ? r.g_rgb + R_W1 + n * 96 + 27 + c : r.g_density + D_W1 + (n - 64) * 64 + c;
    } else if (i < kSlabR2) {       // colour layer 0, view columns: internal column c -> input lane c (c < 27) or 91 + (c - 27)
        const uint32_t e = i - kSlabR1V, c = e / 64, n = e % 64;
        dst = r.g_rgb + R_W1 + n * 96 + (c < 27 ? c : 91 + (c - 27));
    } else if (i < kSlabD2) {       // colour head: rows 0..3 real
        const uint32_t e = i - kSlabR2, c = e / 16, n = e % 16;
        if (n < 4) dst = r.g_rgb + R_W2 + n * 64 + c;
    } else {                        // density head: row 0 real
        const uint32_t e = i - kSlabD2, c = e / 16, n = e % 16;
        if (n == 0) dst = r.g_density + D_W2 + c;
    }
    // every gradient tile of the step (head gradients, dHR .. dH1) is an operand of some weight-gradient GEMM: an fp16
    // overflow anywhere in the backward chain shows up here as inf / NaN (inf x 0 included) -- GradScaler's found_inf.
    // Only entries that map to a parameter are looked at.
    if (!dst) return;
    if (r.scaler && !isfinite(sum)) scaler_raise(r.scaler);
    *dst += sum;
}

}  // namespace
