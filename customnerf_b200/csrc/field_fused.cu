// field_fused.cu -- grid encoding + field network in ONE kernel: the hash-grid features never round-trip HBM.
//
// BASELINE.json north_star, kernel 3: "a fused 64-wide density+color MLP on tcgen05 tensor cores with weights in shared
// memory ... fused with the encode output so features never round-trip HBM".  Replaces the pair
//   GridEncoder.forward      gridencoder/grid.py:27-69,151-168 + gridencoder/src/gridencoder.cu:87-244  (writes [M,32])
//   NeRFNetwork.forward      nerf/network_grid.py:159-177 (three tcnn launches re-reading it)
// and, in density-only form, NeRFNetwork.density (:179-193) as the occupancy update (renderer.py:1658-1715) and the coarse
// pass of the dense renderer (renderer.py:310-320) call it.
//
// One CTA = 256 threads, two per SM.  Warp-specialised:
//   warps 4..7  PRODUCERS  thread = sample: 16 levels x 8 corner gathers of the fp32 master table (rounded to fp16 on load,
//               ge_level_gather -- the very function of the standalone encoder, so the features are bit-identical),
//               packed to fp16 and written straight into the 128B-swizzled A-operand tile (XV) together with the view-
//               direction embedding; the gathers of tile t+1 run while the consumers are busy with tile t.
//   warps 0..3  CONSUMERS  thread = row = TMEM lane: the layer chain of field_mlp.cu (one thread issues tcgen05.mma, all
//               128 run the tcgen05.ld epilogues that write the next layer's operand tile).
// Hand-over through two mbarriers: `full` (128 producer arrivals: XV + POS of the tile are written and fenced to the async
// proxy) and `empty` (128 consumer arrivals).  XV is read by the MMAs of the FIRST stage only: the view-direction part
// of colour layer 0 is issued together with trunk layer 0 into a second TMEM accumulator (columns 64..127) and the
// fea x Wr1f part is accumulated onto it five stages later -- so XV is released after one stage and a single buffer
// suffices (2 CTAs x 112 KB of shared memory per SM).
//
// Variants: kColor = false -> trunk + density head only (5 layers); kSave -> training forward (x_en and sigma_arg written
// once for the backward pass, the five activation planes leave through TMA stores as in field_mlp.cu); kSrc = SRC_OCC ->
// the sample positions are the jittered cell centres of the occupancy grid in Morton order and sigma lands in
// tmp_grid[cascade][morton].
#include "common.cuh"
#include "umma.cuh"
#include "field_common.cuh"
#include "grid_d3c2.cuh"
#include <string.h>

namespace {

constexpr uint32_t kFusedThreads = 256;
constexpr uint32_t kFusedLevels = 16;
constexpr uint32_t SF_POS = S_FWD_BYTES;            // [128] float4: sample position + output row (bits), after the FEA tile
constexpr uint32_t SF_BYTES = SF_POS + 128 * 16;
constexpr uint32_t kNoRow = 0xffffffffu;
enum { SRC_XYZ = 0, SRC_OCC = 1 };

struct FusedArgs {
    const float *xyz;         // SRC_XYZ: [M,3] positions in [-bound, bound];  SRC_OCC: [G^3,3] cell centres in [-1,1], x-major
    const float *dirs;        // [M,3] (kColor)
    const float *noise;       // SRC_OCC: [cascade, G^3, 3] uniform [0,1), x-major cell order, or null (no jitter)
    const float *table;       // fp32 master table [rows, 2]
    const int32_t *offsets;   // [17]
    const uint8_t *wimg;      // forward weight image (nb200_field_pack_weights)
    float *sigma;             // [M]
    float *sigma_arg;         // [M]   (kSave)
    __half *rgba;             // [M,4] (kColor)
    __half *x_en;             // [M,32] (kSave)
    const int32_t *count_dev;
    float S, bound;
    uint32_t H, gridtype, align_corners, interp;
    uint32_t M, G, cascade, pad;
    uint32_t *status;         // kernel status word or null (field_common.cuh)
};

__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

// accumulator row (64 fp32 columns starting at taddr) -> optional ReLU -> fp16 -> row of a swizzled tile; two passes of 32
// columns (this kernel runs under a 128-register cap: 256 threads x 2 CTAs per SM)
template <bool kRelu>
__device__ __forceinline__ void epilogue_row64_2pass(uint32_t taddr, uint8_t *tile, uint32_t row) {
#pragma unroll
    for (uint32_t h = 0; h < 2; h++) {
        uint32_t a[32];
        umma::tmem_ld32(taddr + 32 * h, a);
        umma::tmem_ld_wait();
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
            uint32_t q[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                q[j] = pack_h2(__uint_as_float(a[c * 8 + 2 * j]), __uint_as_float(a[c * 8 + 2 * j + 1]));
                if (kRelu) {
                    const __half2 r2 = __hmax2(*reinterpret_cast<const __half2 *>(&q[j]), __float2half2_rn(0.0f));
                    q[j] = *reinterpret_cast<const uint32_t *>(&r2);
                }
            }
            *reinterpret_cast<uint4 *>(tile + umma::sw128_offset(row, h * 4 + c)) = make_uint4(q[0], q[1], q[2], q[3]);
        }
    }
}

template <bool kColor, bool kSave, int kSrc>
__global__ void __launch_bounds__(kFusedThreads, 2)
k_field_fused(const FusedArgs p, const __grid_constant__ CUtensorMap act_map) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_mma, bar_full, bar_empty;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t fail_s;
    __shared__ LevelInfo info[kFusedLevels];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t Mrows = p.count_dev ? min(p.M, (uint32_t)max(*p.count_dev, 0)) : p.M;
    const uint32_t ntiles = (Mrows + 127) / 128;
    if (blockIdx.x >= ntiles) return;
    constexpr uint32_t kCols = kColor ? 128u : 64u;

    // one-time: weights -> smem, level constants, TMEM, barriers
    for (uint32_t i = tid; i < F_BYTES / 16; i += kFusedThreads)
        cp_async16(umma::smem_u32(smem + S_W) + i * 16, p.wimg + (size_t)i * 16, true);
    cp_async_commit();
    ge_fill_level_info(info, p.offsets, kFusedLevels, p.S, p.H, p.gridtype, p.align_corners != 0);
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, kCols);
    if (tid == 0) {
        umma::mbar_init(&bar_mma, 1);
        umma::mbar_init(&bar_full, 128);
        umma::mbar_init(&bar_empty, 128);
        umma::mbar_fence_init();
        fail_s = 0;
    }
    cp_async_wait<0>();
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    float4 *pos_s = reinterpret_cast<float4 *>(smem + SF_POS);

    // =================================================================================================== producers
    if (warp >= 4) {
        const uint32_t ptid = tid - 128;
        uint32_t ephase = 1;                                   // a fresh mbarrier passes a wait on parity 1
        const InXform xf{p.bound, 1.0f / (2.0f * p.bound), nullptr};
        const float half_off = p.align_corners ? 0.0f : 0.5f;
        const uint32_t G3 = p.G * p.G * p.G;
        uint8_t *xv = smem + S_XV;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const uint32_t g = tile * 128 + ptid;
            const bool valid = g < Mrows;
            float px = 0.0f, py = 0.0f, pz = 0.0f, d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
            if (valid) {
                if (kSrc == SRC_XYZ) {
                    px = __ldg(p.xyz + (size_t)g * 3); py = __ldg(p.xyz + (size_t)g * 3 + 1); pz = __ldg(p.xyz + (size_t)g * 3 + 2);
                } else {
                    // update_extra_state (renderer.py:1680-1690): cell (x, y, z) <- Morton index, position =
                    // centre * (bound_c - hgs) + (2 u - 1) * hgs with hgs = bound_c / G, bound_c = min(2^cas, bound);
                    // every product / sum rounded separately, as the chain of torch kernels rounds them
                    const uint32_t cas = g / G3, m = g - cas * G3;
                    const uint32_t cx = nb_morton3D_invert(m), cy = nb_morton3D_invert(m >> 1), cz = nb_morton3D_invert(m >> 2);
                    const size_t n = ((size_t)cx * p.G + cy) * p.G + cz;
                    const float bc = fminf((float)(1u << cas), p.bound), hgs = bc / (float)p.G, span = bc - hgs;
                    px = __fmul_rn(__ldg(p.xyz + n * 3), span); py = __fmul_rn(__ldg(p.xyz + n * 3 + 1), span);
                    pz = __fmul_rn(__ldg(p.xyz + n * 3 + 2), span);
                    if (p.noise) {
                        const float *u = p.noise + ((size_t)cas * G3 + n) * 3;
                        px = __fadd_rn(px, __fmul_rn(__fsub_rn(__fmul_rn(__ldg(u), 2.0f), 1.0f), hgs));
                        py = __fadd_rn(py, __fmul_rn(__fsub_rn(__fmul_rn(__ldg(u + 1), 2.0f), 1.0f), hgs));
                        pz = __fadd_rn(pz, __fmul_rn(__fsub_rn(__fmul_rn(__ldg(u + 2), 2.0f), 1.0f), hgs));
                    }
                }
                if (kColor) {
                    d0 = __ldg(p.dirs + (size_t)g * 3); d1 = __ldg(p.dirs + (size_t)g * 3 + 1); d2 = __ldg(p.dirs + (size_t)g * 3 + 2);
                }
            }
            // ---- the 16 x 8 corner gathers of this sample, four levels (32 loads) in flight at a time
            uint4 xe[4];
            {
                float x0 = xf(px), x1 = xf(py), x2 = xf(pz);
                const bool live = valid && !((x0 < 0 || x0 > 1) || (x1 < 0 || x1 > 1) || (x2 < 0 || x2 > 1));
                if (!live) { x0 = x1 = x2 = 0.5f; }          // branch-free gather: dead rows read a valid cell and are zeroed below
#pragma unroll
                for (uint32_t l0 = 0; l0 < kFusedLevels; l0 += 4) {
                    float res[8];
                    ge_gather4<float, true>(info + l0, p.table, x0, x1, x2, half_off, p.interp, res);
                    const uint4 pk = make_uint4(pack_h2(res[0], res[1]), pack_h2(res[2], res[3]), pack_h2(res[4], res[5]),
                                                pack_h2(res[6], res[7]));
                    xe[l0 >> 2] = live ? pk : make_uint4(0, 0, 0, 0);
                }
            }
            // ---- hand the row over once the consumers have released the tile
            if (!umma::mbar_wait(&bar_empty, ephase)) fail_s = 1;
            ephase ^= 1u;
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(xv + umma::sw128_offset(ptid, c)) = xe[c];
            if (kColor) write_view_chunks(xv, ptid, d0, d1, d2, valid);
            pos_s[ptid] = make_float4(px, py, pz, __uint_as_float(valid ? g : kNoRow));
            if (kSave && valid) {
                uint4 *dst = reinterpret_cast<uint4 *>(p.x_en + (size_t)g * 32);
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) dst[c] = xe[c];
            }
            umma::fence_proxy_async();
            mbar_arrive(&bar_full);
        }
        return;
    }

    // =================================================================================================== consumers
    const uint32_t tmem = tmem_base_s;
    const uint32_t trow = tmem + ((warp * 32u) << 16);
    const uint32_t sW = umma::smem_u32(smem + S_W), sXV = umma::smem_u32(smem + S_XV), sH0 = umma::smem_u32(smem + S_H0),
                   sH1 = umma::smem_u32(smem + S_H1), sFEA = umma::smem_u32(smem + S_FEA);
    constexpr uint32_t ID64 = umma::make_idesc_f16(128, 64, 0, 0), ID16 = umma::make_idesc_f16(128, 16, 0, 0);
    uint32_t phase = 0, fphase = 0;

    auto mma_run = [&](uint32_t dcol, uint32_t a, uint32_t ak, uint32_t b, uint32_t bk, uint32_t nk, uint32_t idesc, bool first) {
        for (uint32_t k = 0; k < nk; k++)
            umma::mma_f16_ss(tmem + dcol, kdesc(a, ak + k), kdesc(b, bk + k), idesc, !(first && k == 0));
    };
    auto sync_mma = [&]() {
        if (tid == 0) umma::commit(&bar_mma);
        if (!umma::mbar_wait(&bar_mma, phase)) fail_s = 1;
        phase ^= 1;
        umma::fence_after_sync();
    };
    auto publish = [&]() {
        umma::fence_proxy_async();
        umma::fence_before_sync();
        bar_consumers();
        umma::fence_after_sync();
    };
    auto store_tile = [&](uint32_t tile_smem, uint32_t pl, uint32_t row0) {
        if (kSave && tid == 0) {
            umma::tma_store_3d(&act_map, tile_smem, 0, (int32_t)row0, (int32_t)pl);
            umma::tma_store_commit();
        }
    };
    auto drain1 = [&]() { if (kSave && tid == 0) umma::tma_store_wait_read<1>(); };
    auto drain0 = [&]() { if (kSave && tid == 0) umma::tma_store_wait_read<0>(); };

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t row0 = tile * 128;
        // ---- stage 0: trunk layer 0 (h1 = relu(x_en W1^T), K = 32) and, into the second accumulator, the view part of
        //      colour layer 0 ([view | 1] Wr1v^T, K = 32): the only readers of the XV tile
        if (!umma::mbar_wait(&bar_full, fphase)) fail_s = 1;
        fphase ^= 1u;
        umma::fence_after_sync();
        const float4 pos = pos_s[tid];
        const uint32_t orow = __float_as_uint(pos.w);
        const bool valid = orow != kNoRow;
        if (tid == 0) {
            mma_run(0, sXV, 0, sW + F_W1V, 0, 2, ID64, true);
            if (kColor) mma_run(64, sXV, 2, sW + F_W1V, 2, 2, ID64, true);
        }
        sync_mma();
        mbar_arrive(&bar_empty);                            // XV and POS may be refilled for the next tile
        epilogue_row64_2pass<true>(trow, smem + S_H0, tid);
        drain0();                                           // (the previous tile's hr store has left H1)
        publish();
        // ---- trunk layer 1: h2 = relu(h1 W2^T)
        if (tid == 0) mma_run(0, sH0, 0, sW + F_W2, 0, 4, ID64, true);
        store_tile(sH0, 0, row0);
        sync_mma();
        epilogue_row64_2pass<true>(trow, smem + S_H1, tid);
        drain1();
        publish();
        // ---- trunk layer 2: fea = h2 W3^T (no activation)
        if (tid == 0) mma_run(0, sH1, 0, sW + F_W3, 0, 4, ID64, true);
        store_tile(sH1, 1, row0);
        sync_mma();
        epilogue_row64_2pass<false>(trow, smem + S_FEA, tid);
        drain1();                                           // the h1 store has left H0 before hd overwrites it
        publish();
        // ---- density layer 0: hd = relu(fea Wd1^T)
        if (tid == 0) mma_run(0, sFEA, 0, sW + F_WD1, 0, 4, ID64, true);
        store_tile(sFEA, 2, row0);
        sync_mma();
        epilogue_row64_2pass<true>(trow, smem + S_H0, tid);
        drain1();                                           // the h2 store has left H1 before hr overwrites it
        publish();
        // ---- density layer 1: raw = hd Wd2^T (N = 16, lane 0 is the output); sigma = exp(raw + 5 exp(-|x|^2 / 0.08))
        if (tid == 0) mma_run(0, sH0, 0, sW + F_WD2, 0, 4, ID16, true);
        store_tile(sH0, 3, row0);
        sync_mma();
        {
            uint32_t r[16];
            umma::tmem_ld16(trow, r);
            umma::tmem_ld_wait();
            if (valid) {
                const float gauss = 5.0f * expf(-(pos.x * pos.x + pos.y * pos.y + pos.z * pos.z) / (2 * 0.2f * 0.2f));
                const float arg = __uint_as_float(r[0]) + gauss;
                p.sigma[orow] = expf(arg);
                if (kSave) p.sigma_arg[orow] = arg;
            }
        }
        drain1();
        umma::fence_before_sync();
        bar_consumers();
        umma::fence_after_sync();
        if (kColor) {
            // ---- colour layer 0: hr = relu(fea Wr1f^T + [view|1] Wr1v^T): the fea part accumulates onto stage 0's view part
            if (tid == 0) mma_run(64, sFEA, 0, sW + F_WR1F, 0, 4, ID64, false);
            sync_mma();
            epilogue_row64_2pass<true>(trow + 64, smem + S_H1, tid);
            drain1();
            publish();
            // ---- colour layer 1: rgba = sigmoid(hr Wr2^T) (N = 16, lanes 0..3)
            if (tid == 0) mma_run(0, sH1, 0, sW + F_WR2, 0, 4, ID16, true);
            store_tile(sH1, 4, row0);
            sync_mma();
            {
                uint32_t r[16];
                umma::tmem_ld16(trow, r);
                umma::tmem_ld_wait();
                if (valid) {
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) v[j] = 1.0f / (1.0f + expf(-__uint_as_float(r[j])));
                    *reinterpret_cast<uint2 *>(p.rgba + (size_t)orow * 4) = make_uint2(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]));
                }
            }
            drain1();                       // the hd store has left H0 before the next tile's h1 overwrites it
            umma::fence_before_sync();
            bar_consumers();
            umma::fence_after_sync();
        }
    }
    if (kSave && tid == 0) umma::tma_store_wait<0>();
    umma::fence_before_sync();
    bar_consumers();
    if (warp == 0) umma::tmem_dealloc(tmem, kCols);
    if (tid == 0 && fail_s) {           // make a barrier time-out visible: status word, else NaN
        if (p.status) atomicOr(p.status, kStatusFieldFusedTimeout);
        else p.sigma[0] = __int_as_float(0x7fc00000);
    }
}

// the saved activations as a rank-3 tensor [5 planes][M rows][64 halves], boxes of 128 rows, 128B swizzle
int make_act_map_fused(CUtensorMap *map, void *act, uint32_t M) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return NB200_E_BAD_ARG;
    const cuuint64_t dims[3] = {64, M, 5};
    const cuuint64_t strides[2] = {128, (cuuint64_t)M * 128};
    const cuuint32_t box[3] = {64, 128, 1}, estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, act, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : NB200_E_BAD_ARG;
}

template <bool kColor, bool kSave, int kSrc>
int launch_fused(const FusedArgs &a, const CUtensorMap &map, cudaStream_t st) {
    const int smem = (int)SF_BYTES + 1024;
    // per device: the attribute belongs to the function on the CURRENT device (one flag per device ordinal)
    static bool configured[64] = {};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_field_fused<kColor, kSave, kSrc>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t ntiles = (a.M + 127) / 128;
    const uint32_t grid = ntiles < (uint32_t)(2 * sms) ? ntiles : (uint32_t)(2 * sms);
    k_field_fused<kColor, kSave, kSrc><<<grid, kFusedThreads, smem, st>>>(a, map);
    NB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" {

int nb200_field_fused_forward(const float *xyz, const float *dirs, float bound, const float *table, const int32_t *offsets,
                              uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp,
                              const void *fwd_img, float *sigma, float *sigma_arg, void *rgba, void *x_en, void *act,
                              uint32_t M, const int32_t *count_dev, void *stream) {
    if (M == 0) return 0;
    if (!xyz || !table || !offsets || !fwd_img || !sigma || L != kFusedLevels || !(bound > 0.0f)) return NB200_E_BAD_ARG;
    const bool color = rgba != nullptr, save = act != nullptr;
    if (color && !dirs) return NB200_E_BAD_ARG;
    if (save && (!color || !sigma_arg || !x_en || (reinterpret_cast<uintptr_t>(act) & 15u) ||
                 (reinterpret_cast<uintptr_t>(x_en) & 15u))) return NB200_E_BAD_ARG;
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.xyz = xyz; a.dirs = dirs; a.table = table; a.offsets = offsets; a.wimg = (const uint8_t *)fwd_img;
    a.sigma = sigma; a.sigma_arg = sigma_arg; a.rgba = (__half *)rgba; a.x_en = (__half *)x_en; a.count_dev = count_dev;
    a.S = S; a.bound = bound; a.H = H; a.gridtype = gridtype; a.align_corners = align_corners != 0; a.interp = interp; a.M = M;
    a.status = nb_kernel_status_word();
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (save) {
        const int rc = make_act_map_fused(&map, act, M);
        if (rc) return rc;
        return launch_fused<true, true, SRC_XYZ>(a, map, nb_stream(stream));
    }
    if (color) return launch_fused<true, false, SRC_XYZ>(a, map, nb_stream(stream));
    return launch_fused<false, false, SRC_XYZ>(a, map, nb_stream(stream));
}

int nb200_occ_density(const float *cell_xyz, const float *noise, uint32_t G, uint32_t cascade, float bound, const float *table,
                      const int32_t *offsets, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                      uint32_t interp, const void *fwd_img, float *tmp_grid, void *stream) {
    if (!cell_xyz || !table || !offsets || !fwd_img || !tmp_grid || L != kFusedLevels || !(bound > 0.0f) || G == 0 || G > 1024 ||
        cascade == 0 || cascade > 8 || (uint64_t)cascade * G * G * G > 0x7fffffffull)
        return NB200_E_BAD_ARG;
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.xyz = cell_xyz; a.noise = noise; a.table = table; a.offsets = offsets; a.wimg = (const uint8_t *)fwd_img;
    a.sigma = tmp_grid; a.S = S; a.bound = bound; a.H = H; a.gridtype = gridtype; a.align_corners = align_corners != 0;
    a.interp = interp; a.M = cascade * G * G * G; a.G = G; a.cascade = cascade;
    a.status = nb_kernel_status_word();
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    return launch_fused<false, false, SRC_OCC>(a, map, nb_stream(stream));
}

}  // extern "C"
