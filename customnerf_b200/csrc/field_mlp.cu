// field_mlp.cu -- the fused density + colour field network on the sm_100a tensor cores (tcgen05 / TMEM).
//
// Replaces the three tiny-cuda-nn FullyFusedMLP launches of nerf/network_grid.py:159-177 (trunk 32-64-64-64, density
// head 64-64-1, colour head 91(96)-64-4) plus the frequency embedder (nerf/base.py:42-77), the gaussian density bias
// (network_grid.py:150-156), trunc_exp (provider_utils.py:16-29) and the concat -- one kernel, no activation ever
// round-trips HBM between layers.
//
// Mapping: one CTA = 128 threads works on tiles of 128 points (UMMA M = 128).  Every operand tile in shared memory is
// "128 rows x 128 bytes, 128B-swizzled" (umma.cuh): activations are K-major A operands, weights K-major B operands
// resident for the CTA's lifetime (44 KB).  Each layer is: one elected thread issues K/16 tcgen05.mma into the TMEM
// accumulator and commits to an mbarrier; all 128 threads (thread == row == TMEM lane) read their accumulator row with
// tcgen05.ld, apply the activation, round to fp16 and store the row as the next layer's A tile.
//
// Numerics contract (tiny-cuda-nn is un-vendored: parity unpinned): fp16 operands, fp32 accumulation, hidden
// activations rounded to fp16, heads evaluated in fp32.  Bias-free; input lanes 91..95 of the colour head are fed 1.0.
//
// K order of the colour head inside the kernel: [fea(64) | view_en(27) | ones(5)] -- a column permutation of the
// reference's cat([view_en, fea]) that the weight packer applies to the flat tcnn-layout parameter vector.
#include "common.cuh"
#include "umma.cuh"
#include "field_common.cuh"
#include "adam.cuh"
#include <string.h>

namespace {

constexpr uint32_t kPackBlocks = 16;
// every byte of both images that a descriptor can reach is written here (no separate zero fill)
__global__ void k_pack_field_weights(const float *__restrict__ trunk, const float *__restrict__ density,
                                      const float *__restrict__ rgb, uint8_t *__restrict__ fwd, uint8_t *__restrict__ bwd,
                                      ScalerCommit sc) {
    if (blockIdx.x == kPackBlocks) {        // the extra block: GradScaler.update() + the optimiser's step count (adam.cuh)
        if (threadIdx.x == 0) scaler_commit(sc);
        return;
    }
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = kPackBlocks * blockDim.x;
    for (uint32_t i = tid; i < 64 * 64; i += nth) {
        const uint32_t n = i >> 6, k = i & 63;
        if (k < 32) {
            const float w1 = trunk[T_W1 + n * 32 + k];
            put(fwd + F_W1V, n, k, w1);
            put(bwd + B_W1T, k, n, w1);
            // colour layer 0, view part: internal col 32+j <- reference input lane j (j < 27) or 91 + (j - 27) (ones)
            const uint32_t src = (k < 27) ? k : 91 + (k - 27);
            put(fwd + F_W1V, n, 32 + k, rgb[R_W1 + n * 96 + src]);
        }
        const float w2 = trunk[T_W2 + n * 64 + k], w3 = trunk[T_W3 + n * 64 + k];
        const float wd1 = density[D_W1 + n * 64 + k], wr1f = rgb[R_W1 + n * 96 + 27 + k];
        put(fwd + F_W2, n, k, w2);     put(bwd + B_W2T, k, n, w2);
        put(fwd + F_W3, n, k, w3);     put(bwd + B_W3T, k, n, w3);
        put(fwd + F_WD1, n, k, wd1);   put(bwd + B_WD1T, k, n, wd1);
        put(fwd + F_WR1F, n, k, wr1f); put(bwd + B_WR1FT, k, n, wr1f);
        if (n < 16) {
            const float wd2 = density[D_W2 + n * 64 + k], wr2 = rgb[R_W2 + n * 64 + k];
            put(fwd + F_WD2, n, k, wd2);
            put(fwd + F_WR2, n, k, wr2);
            put(bwd + B_W16T, k, n, wr2);
            put(bwd + B_W16T, k, 16 + n, wd2);
        } else if (n >= 32) {
            put(bwd + B_W16T, k, n, 0.0f);          // columns 32..63 of the head tile: never read, kept finite
        }
    }
}

struct FieldFwdArgs {
    const __half *x_en;     // [M, 32]
    const float *xyz;       // [M, 3]  (gaussian density bias)
    const float *dirs;      // [M, 3]
    const uint8_t *wimg;    // forward weight image
    float *sigma;           // [M]
    float *sigma_arg;       // [M] or null: raw + gaussian (the argument of trunc_exp), saved for backward
    __half *rgba;           // [M, 4]
    __half *act;            // [5, M, 64] or null: h1, h2, fea, hd, hr saved for backward
    uint32_t M;             // rows allocated (the stride of `act` planes)
    const int32_t *count_dev;   // when non-null only rows < min(M, *count_dev) are evaluated
    uint32_t *status;           // kernel status word or null (field_common.cuh)
};

// the embedding as the field kernels compute it, fp32 [M, 27] (get_embedder(4) as a device op; also the test hook that
// pins the double-angle evaluation above)
__global__ void __launch_bounds__(256)
k_freq_embed(const float *__restrict__ dirs, float *__restrict__ out, uint32_t M) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= M) return;
    float e[32];
    view_embed32(dirs[(size_t)g * 3], dirs[(size_t)g * 3 + 1], dirs[(size_t)g * 3 + 2], e);
#pragma unroll
    for (int j = 0; j < 27; j++) out[(size_t)g * 27 + j] = e[j];
}

// kColor = false: trunk + density head only (sigma; no view encoding, no colour layers, no saves): the occupancy update's
// chunked density query and any density() call that already holds the features
template <bool kColor>
__global__ void __launch_bounds__(128, 2)
k_field_forward(const FieldFwdArgs p, const __grid_constant__ CUtensorMap act_map) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t fail_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t Mrows = p.count_dev ? min(p.M, (uint32_t)max(*p.count_dev, 0)) : p.M;
    const uint32_t ntiles = (Mrows + 127) / 128;
    if (blockIdx.x >= ntiles) return;

    // this thread's input row of the NEXT tile, prefetched into registers while the current tile is in flight; the first
    // tile's row is requested before anything else so that its latency runs under the one-time setup below
    uint4 nx0, nx1, nx2, nx3;
    float nd0, nd1, nd2, npx, npy, npz;
    auto prefetch_inputs = [&](uint32_t tile) {
        const uint32_t g = tile * 128 + tid;
        nx0 = nx1 = nx2 = nx3 = make_uint4(0, 0, 0, 0);
        nd0 = nd1 = nd2 = npx = npy = npz = 0.0f;
        if (tile < ntiles && g < Mrows) {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.x_en + (size_t)g * 32);
            nx0 = __ldg(src); nx1 = __ldg(src + 1); nx2 = __ldg(src + 2); nx3 = __ldg(src + 3);
            if (kColor) {
                nd0 = __ldg(p.dirs + (size_t)g * 3); nd1 = __ldg(p.dirs + (size_t)g * 3 + 1); nd2 = __ldg(p.dirs + (size_t)g * 3 + 2);
            }
            npx = __ldg(p.xyz + (size_t)g * 3); npy = __ldg(p.xyz + (size_t)g * 3 + 1); npz = __ldg(p.xyz + (size_t)g * 3 + 2);
        }
    };
    prefetch_inputs(blockIdx.x);

    // one-time: weights -> smem (asynchronously, behind the first tile's input loads), TMEM, barrier
    for (uint32_t i = tid; i < F_BYTES / 16; i += 128)
        cp_async16(umma::smem_u32(smem + S_W) + i * 16, p.wimg + (size_t)i * 16, true);
    cp_async_commit();
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemCols);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_fence_init();
        fail_s = 0;
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t trow = tmem + ((warp * 32u) << 16);      // this warp's TMEM lane quarter
    const uint32_t sW = umma::smem_u32(smem + S_W), sXV = umma::smem_u32(smem + S_XV), sH0 = umma::smem_u32(smem + S_H0),
                   sH1 = umma::smem_u32(smem + S_H1), sFEA = umma::smem_u32(smem + S_FEA);
    constexpr uint32_t ID64 = umma::make_idesc_f16(128, 64, 0, 0), ID16 = umma::make_idesc_f16(128, 16, 0, 0);
    uint32_t phase = 0;

    // issue `nk` K-steps of A(tile a, first K-step ak) x B(tile b, first K-step bk)
    auto mma_run = [&](uint32_t a, uint32_t ak, uint32_t b, uint32_t bk, uint32_t nk, uint32_t idesc, bool first) {
        for (uint32_t k = 0; k < nk; k++)
            umma::mma_f16_ss(tmem, kdesc(a, ak + k), kdesc(b, bk + k), idesc, !(first && k == 0));
    };
    // completion of everything issued so far -> all threads; returns after the accumulator is readable
    auto sync_mma = [&]() {
        if (tid == 0) umma::commit(&bar);
        if (!umma::mbar_wait(&bar, phase)) fail_s = 1;
        phase ^= 1;
        umma::fence_after_sync();
    };
    // publish the smem tile just written by the epilogue to the async proxy and order the TMEM loads before the next MMA
    auto publish = [&]() {
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    };

    // saved activations leave through the TMA unit: after a tile has been published, thread 0 hands it to a bulk tensor
    // store (plane `pl` of the [5][M][64] activation tensor, rows row0..row0+127, clipped at M by the tensor map) and the
    // copy proceeds behind the next layers.  Before a barrier that lets a later epilogue overwrite a tile, thread 0 makes
    // sure the stores that still read it have finished reading (at most `kPending` younger stores may be in flight).
    const bool save = p.act != nullptr;
    auto store_tile = [&](uint32_t tile_smem, uint32_t pl, uint32_t row0) {
        if (save && tid == 0) {
            umma::tma_store_3d(&act_map, tile_smem, 0, (int32_t)row0, (int32_t)pl);
            umma::tma_store_commit();
        }
    };
    auto drain1 = [&]() { if (save && tid == 0) umma::tma_store_wait_read<1>(); };
    auto drain0 = [&]() { if (save && tid == 0) umma::tma_store_wait_read<0>(); };

    auto write_xv = [&](bool valid) {
        uint8_t *xv = smem + S_XV;
        *reinterpret_cast<uint4 *>(xv + umma::sw128_offset(tid, 0)) = nx0;
        *reinterpret_cast<uint4 *>(xv + umma::sw128_offset(tid, 1)) = nx1;
        *reinterpret_cast<uint4 *>(xv + umma::sw128_offset(tid, 2)) = nx2;
        *reinterpret_cast<uint4 *>(xv + umma::sw128_offset(tid, 3)) = nx3;
        if (kColor) write_view_chunks(xv, tid, nd0, nd1, nd2, true);
        (void)valid;
    };

    write_xv(true);
    float px = npx, py = npy, pz = npz;                     // this tile's sample position (gaussian density bias)
    cp_async_wait<0>();                                     // the weight image
    publish();

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t g = tile * 128 + tid, row0 = tile * 128;
        const bool valid = g < Mrows;

        // ---- trunk layer 0: h1 = relu(x_en W1^T), K = 32
        if (tid == 0) mma_run(sXV, 0, sW + F_W1V, 0, 2, ID64, true);
        prefetch_inputs(tile + gridDim.x);                  // loads complete behind the next layers
        sync_mma();
        epilogue_row64<true>(trow, smem + S_H0, tid, nullptr);
        drain0();                                           // (the previous tile's hr store has left H1)
        publish();
        // ---- trunk layer 1: h2 = relu(h1 W2^T)
        if (tid == 0) mma_run(sH0, 0, sW + F_W2, 0, 4, ID64, true);
        store_tile(sH0, 0, row0);
        sync_mma();
        epilogue_row64<true>(trow, smem + S_H1, tid, nullptr);
        drain1();
        publish();
        // ---- trunk layer 2: fea = h2 W3^T (no activation)
        if (tid == 0) mma_run(sH1, 0, sW + F_W3, 0, 4, ID64, true);
        store_tile(sH1, 1, row0);
        sync_mma();
        epilogue_row64<false>(trow, smem + S_FEA, tid, nullptr);
        drain1();                                           // the h1 store has left H0 before hd overwrites it
        publish();
        // ---- density layer 0: hd = relu(fea Wd1^T)
        if (tid == 0) mma_run(sFEA, 0, sW + F_WD1, 0, 4, ID64, true);
        store_tile(sFEA, 2, row0);
        sync_mma();
        epilogue_row64<true>(trow, smem + S_H0, tid, nullptr);
        drain1();                                           // the h2 store has left H1 before hr overwrites it
        publish();
        // ---- density layer 1: raw = hd Wd2^T (N = 16, lane 0 is the output); sigma = exp(raw + 5 exp(-|x|^2 / 0.08))
        if (tid == 0) mma_run(sH0, 0, sW + F_WD2, 0, 4, ID16, true);
        store_tile(sH0, 3, row0);
        float cx = px, cy = py, cz = pz;
        if (!kColor) {                      // density only: this is the tile's last MMA -- stage the next tile's inputs under it
            write_xv(true);
            px = npx; py = npy; pz = npz;
        }
        sync_mma();
        {
            uint32_t r[16];
            umma::tmem_ld16(trow, r);
            umma::tmem_ld_wait();
            if (valid) {
                const float gauss = 5.0f * expf(-(cx * cx + cy * cy + cz * cz) / (2 * 0.2f * 0.2f));
                const float arg = __uint_as_float(r[0]) + gauss;
                p.sigma[g] = expf(arg);
                if (p.sigma_arg) p.sigma_arg[g] = arg;
            }
        }
        drain1();
        if (!kColor) {
            umma::fence_proxy_async();      // the XV rows just written feed the next tile's first MMA
            umma::fence_before_sync();
            __syncthreads();
            umma::fence_after_sync();
            continue;
        }
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        // ---- colour layer 0: hr = relu(fea Wr1f^T + [view|1] Wr1v^T), K = 64 + 32
        if (tid == 0) {
            mma_run(sFEA, 0, sW + F_WR1F, 0, 4, ID64, true);
            mma_run(sXV, 2, sW + F_W1V, 2, 2, ID64, false);
        }
        sync_mma();
        epilogue_row64<true>(trow, smem + S_H1, tid, nullptr);
        drain1();
        publish();
        // ---- colour layer 1: rgba = sigmoid(hr Wr2^T) (N = 16, lanes 0..3)
        if (tid == 0) mma_run(sH1, 0, sW + F_WR2, 0, 4, ID16, true);
        store_tile(sH1, 4, row0);
        // the XV tile has been free since colour layer 0 completed: stage the next tile's inputs (and its direction
        // encoding) while the last MMA runs; published by the fence + barrier that end the tile
        write_xv(true);
        px = npx; py = npy; pz = npz;
        sync_mma();
        {
            uint32_t r[16];
            umma::tmem_ld16(trow, r);
            umma::tmem_ld_wait();
            if (valid) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = 1.0f / (1.0f + expf(-__uint_as_float(r[j])));
                *reinterpret_cast<uint2 *>(p.rgba + (size_t)g * 4) = make_uint2(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]));
            }
        }
        drain1();                           // the hd store has left H0 before the next tile's h1 overwrites it
        umma::fence_proxy_async();          // the XV rows written after colour layer 0 feed the next tile's first MMA
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    }
    if (save && tid == 0) umma::tma_store_wait<0>();

    if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
    if (tid == 0 && fail_s) {           // make a barrier time-out visible: status word, else NaN
        if (p.status) atomicOr(p.status, kStatusFieldFwdTimeout);
        else p.sigma[0] = __int_as_float(0x7fc00000);
    }
}


// =====================================================================================================
// backward: dgrad chain + weight gradients, one kernel
// =====================================================================================================
// shared-memory map (bytes).  The four gradient tiles are consecutive so that an MN-major M = 128 A operand whose
// first 64-row block is one tile finds its second block (LBO = 16384) in the next tile.
constexpr uint32_t SB_W = 0;                        // transposed weights, B_BYTES
constexpr uint32_t SB_T16 = B_BYTES;                // cols 0..15 dOr (colour head pre-sigmoid grads), 16..31 dOd
constexpr uint32_t SB_GA = SB_T16 + 16384;          // dHR, later dH2
constexpr uint32_t SB_GB = SB_GA + 16384;           // dHD, later dH1
constexpr uint32_t SB_GC = SB_GB + 16384;           // dFEA
constexpr uint32_t SB_XV = SB_GC + 16384;           // x_en | view
constexpr uint32_t SB_H1 = SB_XV + 16384;
constexpr uint32_t SB_H2 = SB_H1 + 16384;
constexpr uint32_t SB_FEA = SB_H2 + 16384;
constexpr uint32_t SB_HD = SB_FEA + 16384;
constexpr uint32_t SB_HR = SB_HD + 16384;
constexpr uint32_t SB_XV2 = SB_HR + 16384;          // second x_en | view buffer (tiles alternate)
constexpr uint32_t S_BWD_BYTES = SB_XV2 + 16384;    // 45056 + 11 * 16384 = 225280
// TMEM columns
constexpr uint32_t C_DG = 0, C_W1 = 64, C_W2 = 96, C_W3 = 160, C_PAIR = 224, C_R1V = 288, C_D2 = 320, C_R2 = 384, C_DG2 = 448;
constexpr uint32_t kTmemColsBwd = 512;
// (slab layout: wgrad_reduce.cuh)

struct FieldBwdArgs {
    const float *d_sigma;     // [M]
    const float *d_rgba;      // [M, 4]
    const float *sigma_arg;   // [M]
    const __half *rgba;       // [M, 4]
    const __half *x_en;       // [M, 32]
    const float *dirs;        // [M, 3]
    const __half *act;        // [5, M, 64]
    const uint8_t *wimg;      // backward weight image
    __half *d_x_en;           // [M, 32]
    float *g_trunk, *g_density, *g_rgb;   // flat fp32 parameter gradients, ACCUMULATED INTO
    float *slabs;                         // [gridDim.x][kWgradFloats] per-CTA partial sums, or null (atomics)
    uint32_t M;               // rows allocated (the stride of `act` planes)
    const int32_t *count_dev; // when non-null only rows < min(M, *count_dev) are processed
    uint32_t *scaler;         // loss-scaler words (adam.cuh) or null: a feature gradient that leaves the fp16 range raises found-inf
    uint32_t *status;         // kernel status word or null (field_common.cuh)
    uint32_t dbg;             // timing experiments only (NB200_FIELDB_DBG): 1 = skip the weight-gradient MMAs, 2 = skip their flush, 4 = skip the slab reduce launch
};

// MN-major descriptor of K-step ks (16 rows = 2 swizzle atoms) starting at column `col0` (multiple of 8 halves)
__device__ __forceinline__ uint64_t mndesc(uint32_t tile_addr, uint32_t ks, uint32_t col0) {
    return umma::make_desc(tile_addr + ks * 2048 + col0 * 2, 16384, 1024, umma::kLayoutSW128);
}

// The backward CTA has 8 warps: warp w owns the TMEM lane quarter 32 (w % 4) (a hardware restriction of tcgen05.ld) and
// the column half (w / 4) of every 64-wide accumulator, so each epilogue is split over twice the warps of the forward
// kernel and twice as many warps are in flight to hide the TMEM / shared-memory / barrier latencies between the MMAs.
constexpr uint32_t kBwdThreads = 256;

// shared-space 16-byte accesses on 32-bit shared addresses (the generic LD.E / ST.E forms cost an address-space check and a
// descriptor per access on the epilogue's critical path)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// relu'(x) masks of one row and column half as 0xffff / 0 half lanes: loaded and converted BEFORE the wait for the dgrad
// MMAs (the activation tile is in shared memory long before the accumulator is ready), so that nothing but the TMEM load
// sits between the wake-up and the gradient tile's stores
struct ReluMask { uint32_t w[16]; };
__device__ __forceinline__ void load_relu_mask(ReluMask &m, uint32_t act_tile, uint32_t row, uint32_t half) {
    const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        const uint4 v = lds128(act_tile + umma::sw128_offset(row, half * 4 + c));
        const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) m.w[c * 4 + j] = __hgt2_mask(*reinterpret_cast<const __half2 *>(&vw[j]), zero);
    }
}

// dgrad epilogue of one row and one column half: DG (32 fp32) * relu'(activation) -> fp16 -> gradient tile.
// (inlined at its six call sites: the masks stay in registers)
template <bool kMask>
__device__ __forceinline__ void bwd_epilogue_half(uint32_t taddr, const ReluMask &m, uint32_t dst_tile, uint32_t row, uint32_t half) {
    uint32_t a[32];
    umma::tmem_ld32(taddr, a);
    umma::tmem_ld_wait();
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            pk[j] = pack_h2(__uint_as_float(a[c * 8 + 2 * j]), __uint_as_float(a[c * 8 + 2 * j + 1]));
            if (kMask) pk[j] &= m.w[c * 4 + j];
        }
        sts128(dst_tile + umma::sw128_offset(row, half * 4 + c), make_uint4(pk[0], pk[1], pk[2], pk[3]));
    }
}

// 16 accumulator columns of one weight-gradient row -> this CTA's private partial-sum slab (plain 16-byte stores; a
// second small kernel adds the slabs into the flat parameter gradient), or, without a slab, fp32 atomics straight on the
// gradient (every CTA then hits the same 22 k addresses at the same time: ~25 % of the kernel in the r01e profile)
__device__ __noinline__ void bwd_flush16(uint32_t taddr, float *dst, bool on, bool slab) {
    uint32_t r[16];
    umma::tmem_ld16(taddr, r);
    umma::tmem_ld_wait();
    if (on) {
        if (slab) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<uint4 *>(dst + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) atomicAdd(dst + j, __uint_as_float(r[j]));
        }
    }
}

// rows [row0, row0 + 128) of a row-major [M, 64] half plane -> swizzled tile; a warp's copy covers 4 complete lines
__device__ __forceinline__ void load_tile_async(uint32_t tile_smem, const __half *plane, uint32_t row0, uint32_t Mrows,
                                                uint32_t tid) {
#pragma unroll
    for (uint32_t i = 0; i < 1024 / kBwdThreads; i++) {
        const uint32_t q = i * kBwdThreads + tid, row = q >> 3, ch = q & 7u, g = row0 + row;
        const bool ok = g < Mrows;
        cp_async16(tile_smem + umma::sw128_offset(row, ch), plane + (ok ? (size_t)g * 64 + ch * 8 : 0), ok);
    }
}
// rows of the [M, 32] half encoding -> chunks 0..3 of the XV tile
__device__ __forceinline__ void load_xen_async(uint32_t tile_smem, const __half *x_en, uint32_t row0, uint32_t Mrows,
                                               uint32_t tid) {
#pragma unroll
    for (uint32_t i = 0; i < 512 / kBwdThreads; i++) {
        const uint32_t q = i * kBwdThreads + tid, row = q >> 2, ch = q & 3u, g = row0 + row;
        const bool ok = g < Mrows;
        cp_async16(tile_smem + umma::sw128_offset(row, ch), x_en + (ok ? (size_t)g * 32 + ch * 8 : 0), ok);
    }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
k_field_backward(const FieldBwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t half = warp >> 2, row = (warp & 3u) * 32u + (tid & 31u);     // column half and tile row of this thread
    const uint32_t Mrows = p.count_dev ? min(p.M, (uint32_t)max(*p.count_dev, 0)) : p.M;
    const uint32_t ntiles = (Mrows + 127) / 128;
    if (blockIdx.x >= ntiles) return;

    // the weight image streams in asynchronously with the first tile's activations (part of cp.async group G1 below)
    for (uint32_t i = tid; i < B_BYTES / 16; i += kBwdThreads)
        cp_async16(umma::smem_u32(smem + SB_W) + i * 16, p.wimg + (size_t)i * 16, true);
    // the T16 tile is only ever written in its first 4 chunks per row: clear the rest once
    for (uint32_t i = tid; i < 16384 / 16; i += kBwdThreads) reinterpret_cast<uint4 *>(smem + SB_T16)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemColsBwd);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_fence_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t trow = tmem + (((warp & 3u) * 32u) << 16);
    const uint32_t base = umma::smem_u32(smem);
    const uint32_t sW = base + SB_W, sT16 = base + SB_T16, sGA = base + SB_GA, sGB = base + SB_GB, sGC = base + SB_GC,
                   sXV = base + SB_XV, sH1 = base + SB_H1, sH2 = base + SB_H2, sFEA = base + SB_FEA, sHD = base + SB_HD,
                   sHR = base + SB_HR;
    constexpr uint32_t ID64 = umma::make_idesc_f16(128, 64, 0, 0), ID32 = umma::make_idesc_f16(128, 32, 0, 0);
    constexpr uint32_t WG64 = umma::make_idesc_f16(128, 64, 1, 1), WG32 = umma::make_idesc_f16(128, 32, 1, 1);
    uint32_t phase = 0;
    bool first_tile = true;
    const size_t act_stride = (size_t)p.M * 64;

    auto dgrad = [&](uint32_t a, uint32_t ak, uint32_t b, uint32_t bk, uint32_t nk, uint32_t idesc, bool first) {
        for (uint32_t k = 0; k < nk; k++)
            umma::mma_f16_ss(tmem + C_DG, kdesc(a, ak + k), kdesc(b, bk + k), idesc, !(first && k == 0));
    };
    auto dgrad_to = [&](uint32_t dcol, uint32_t a, uint32_t ak, uint32_t b, uint32_t bk, uint32_t nk, uint32_t idesc) {
        for (uint32_t k = 0; k < nk; k++)
            umma::mma_f16_ss(tmem + dcol, kdesc(a, ak + k), kdesc(b, bk + k), idesc, k != 0);
    };
    // D[cols] (+)= A_tile^T (M = 128: this tile and the next) x B_tile[:, col0 : col0 + N], contraction over the 128 points
    auto wgrad = [&](uint32_t dcol, uint32_t a, uint32_t b, uint32_t bcol0, uint32_t idesc) {
        if (p.dbg & 1u) return;
        for (uint32_t k = 0; k < 8; k++)
            umma::mma_f16_ss(tmem + dcol, mndesc(a, k, 0), mndesc(b, k, bcol0), idesc, !(first_tile && k == 0));
    };
    bool timed_out = false;
    auto wait_mma = [&]() {                 // completion of everything committed by the issuing thread so far
        timed_out |= !umma::mbar_wait(&bar, phase);
        phase ^= 1;
        umma::fence_after_sync();
    };
    auto publish = [&]() {
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    };
    ReluMask mk, mk2;                      // relu' masks of the coming epilogue(s), loaded while the MMAs run

    // ---- per-row inputs prefetched one tile ahead into registers: the warps of column half 0 build the head
    //      gradients of their row (T16), the warps of half 1 its direction encoding (XV chunks 4..7)
    uint2 n_rgba = make_uint2(0, 0); float4 n_drgba = make_float4(0, 0, 0, 0);
    float n_dsig = 0, n_sarg = 0, n_d0 = 0, n_d1 = 0, n_d2 = 0;
    auto prefetch_row = [&](uint32_t tile) {
        const uint32_t g = tile * 128 + row;
        const bool ok = tile < ntiles && g < Mrows;
        if (half == 0) {
            n_rgba = make_uint2(0, 0); n_drgba = make_float4(0, 0, 0, 0); n_dsig = 0; n_sarg = 0;
            if (ok) {
                n_rgba = __ldg(reinterpret_cast<const uint2 *>(p.rgba + (size_t)g * 4));
                n_drgba = __ldg(reinterpret_cast<const float4 *>(p.d_rgba + (size_t)g * 4));
                n_dsig = __ldg(p.d_sigma + g); n_sarg = __ldg(p.sigma_arg + g);
            }
        } else {
            n_d0 = n_d1 = n_d2 = 0;
            if (ok) {
                n_d0 = __ldg(p.dirs + (size_t)g * 3); n_d1 = __ldg(p.dirs + (size_t)g * 3 + 1); n_d2 = __ldg(p.dirs + (size_t)g * 3 + 2);
            }
        }
    };
    // head gradients of a row: colour = g * y (1 - y) (sigmoid), density = g * exp(clamp(arg, -15, 15)) (trunc_exp)
    auto write_t16 = [&]() {
        if (half != 0) return;
        const float2 y01 = __half22float2(*reinterpret_cast<const __half2 *>(&n_rgba.x));
        const float2 y23 = __half22float2(*reinterpret_cast<const __half2 *>(&n_rgba.y));
        const float go0 = n_drgba.x * y01.x * (1 - y01.x), go1 = n_drgba.y * y01.y * (1 - y01.y);
        const float go2 = n_drgba.z * y23.x * (1 - y23.x), go3 = n_drgba.w * y23.y * (1 - y23.y);
        const float gd = n_dsig * expf(fminf(fmaxf(n_sarg, -15.0f), 15.0f));
        uint8_t *t16 = smem + SB_T16;
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(t16 + umma::sw128_offset(row, 0)) = make_uint4(pack_h2(go0, go1), pack_h2(go2, go3), 0, 0);
        *reinterpret_cast<uint4 *>(t16 + umma::sw128_offset(row, 1)) = z;
        *reinterpret_cast<uint4 *>(t16 + umma::sw128_offset(row, 2)) = make_uint4(pack_h2(gd, 0.0f), 0, 0, 0);
        *reinterpret_cast<uint4 *>(t16 + umma::sw128_offset(row, 3)) = z;
    };

    // Order of one stage: the issuing thread launches the stage's dgrad MMAs, commits, then launches the stage's weight-
    // gradient MMAs WITHOUT a commit.  The epilogue warps wake on the dgrad commit and work while the tensor pipe runs
    // the weight gradients; tcgen05 MMAs complete in issue order, so the NEXT stage's commit also covers them -- which
    // is when their operands (a gradient tile, an activation tile) may be overwritten / refilled.
    //
    // cp.async groups per tile, in commit order: G1 = {HR, HD, x_en of the next XV buffer}, G2 = {FEA}, G3 = {H2}, G4 = {H1}
    {
        const uint32_t row0 = blockIdx.x * 128;
        load_tile_async(sHR, p.act + 4 * act_stride, row0, Mrows, tid);
        load_tile_async(sHD, p.act + 3 * act_stride, row0, Mrows, tid);
        load_xen_async(sXV, p.x_en, row0, Mrows, tid);
        cp_async_commit();
        load_tile_async(sFEA, p.act + 2 * act_stride, row0, Mrows, tid);
        cp_async_commit();
        load_tile_async(sH2, p.act + act_stride, row0, Mrows, tid);
        cp_async_commit();
        load_tile_async(sH1, p.act, row0, Mrows, tid);
        cp_async_commit();
        prefetch_row(blockIdx.x);
        if (half == 1) write_view_chunks(smem + SB_XV, row, n_d0, n_d1, n_d2, row0 + row < Mrows);
        write_t16();                       // later tiles: written one tile ahead, under the stage-4 MMAs
        prefetch_row(blockIdx.x + gridDim.x);
    }
    const bool issuer_warp = warp == 0;
    uint32_t xv_sel = 0;                   // which XV buffer the current tile uses
    bool bad_dx = false;                   // a feature gradient of this thread's rows left the fp16 range (GradScaler's found_inf)

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t g = tile * 128 + row;
        const bool valid = g < Mrows;
        const uint32_t next = tile + gridDim.x;
        const bool has_next = next < ntiles;
        const uint32_t nrow0 = next * 128;
        const uint32_t sXVc = xv_sel ? base + SB_XV2 : sXV, sXVn = xv_sel ? sXV : base + SB_XV2;
        uint8_t *xv_next = smem + (xv_sel ? SB_XV : SB_XV2);

        cp_async_wait<3>();                // G1 landed (this thread's part); the barrier below covers the other threads
        publish();                         // (also: every warp has read the previous tile's dX accumulator; T16 was written
                                           //  and published during the previous tile, or by the prologue)

        // ---- stages 1 + 2: dHR = (dOr Wr2) * [hr > 0] and dHD = (dOd Wd2) * [hd > 0] are independent (both read only the
        //      head gradients): one commit, one epilogue pass over two accumulators;  wgrad of the two heads (A = T16^T)
        if (issuer_warp && umma::elect_one()) {
            dgrad_to(C_DG, sT16, 0, sW + B_W16T, 0, 1, ID64);
            dgrad_to(C_DG2, sT16, 1, sW + B_W16T, 1, 1, ID64);
            umma::commit(&bar);
            wgrad(C_R2, sT16, sHR, 0, WG64);
            wgrad(C_D2, sT16, sHD, 0, WG64);
        }
        load_relu_mask(mk, sHR, row, half);
        load_relu_mask(mk2, sHD, row, half);
        wait_mma();
        bwd_epilogue_half<true>(trow + C_DG + 32 * half, mk, sGA, row, half);
        bwd_epilogue_half<true>(trow + C_DG2 + 32 * half, mk2, sGB, row, half);
        cp_async_wait<2>();                // G2: FEA (and G1's x_en) for the weight gradients of stage 3
        publish();
        // ---- stage 3: dFEA = dHR Wr1f + dHD Wd1;   wgrad Wr1f | Wd1 (A = [dHR | dHD]^T, B = fea), Wr1v (A = dHR^T, B = view)
        if (issuer_warp && umma::elect_one()) {
            dgrad(sGA, 0, sW + B_WR1FT, 0, 4, ID64, true);
            dgrad(sGB, 0, sW + B_WD1T, 0, 4, ID64, false);
            umma::commit(&bar);
            wgrad(C_PAIR, sGA, sFEA, 0, WG64);
            wgrad(C_R1V, sGA, sXVc, 32, WG32);
        }
        // while the MMAs run: the direction encoding of the NEXT tile (its XV buffer was last read by the previous tile's
        // weight gradients, which the stage-1 commit of this tile has covered)
        if (half == 1) write_view_chunks(xv_next, row, n_d0, n_d1, n_d2, has_next && nrow0 + row < Mrows);
        wait_mma();                        // ... which also means the head weight gradients of stage 2 are done:
        if (has_next) {                    // HR / HD and the other XV buffer are free, refill them for the next tile
            load_tile_async(sHR, p.act + 4 * act_stride, nrow0, Mrows, tid);
            load_tile_async(sHD, p.act + 3 * act_stride, nrow0, Mrows, tid);
            load_xen_async(sXVn, p.x_en, nrow0, Mrows, tid);
        }
        cp_async_commit();                 // G1'
        bwd_epilogue_half<false>(trow + C_DG + 32 * half, mk, sGC, row, half);
        cp_async_wait<2>();                // G3: H2
        publish();
        // ---- stage 4: dH2 = (dFEA W3) * [h2 > 0];   wgrad W3 (A = dFEA^T, B = h2)
        if (issuer_warp && umma::elect_one()) {
            dgrad(sGC, 0, sW + B_W3T, 0, 4, ID64, true);
            umma::commit(&bar);
            wgrad(C_W3, sGC, sH2, 0, WG64);
        }
        load_relu_mask(mk, sH2, row, half);
        // while the MMAs run: the head gradients of the NEXT tile (T16's readers -- stages 1 + 2 and their weight gradients --
        // were covered by the stage-3 commit), then the prefetch of the per-row inputs of the tile after it
        write_t16();
        prefetch_row(next + gridDim.x);
        wait_mma();                        // stage-3 weight gradients done: FEA is free
        if (has_next) load_tile_async(sFEA, p.act + 2 * act_stride, nrow0, Mrows, tid);
        cp_async_commit();                 // G2'
        bwd_epilogue_half<true>(trow + C_DG + 32 * half, mk, sGA, row, half);
        cp_async_wait<2>();                // G4: H1
        publish();
        // ---- stage 5: dH1 = (dH2 W2) * [h1 > 0];   wgrad W2 (A = dH2^T, B = h1)
        if (issuer_warp && umma::elect_one()) {
            dgrad(sGA, 0, sW + B_W2T, 0, 4, ID64, true);
            umma::commit(&bar);
            wgrad(C_W2, sGA, sH1, 0, WG64);
        }
        load_relu_mask(mk, sH1, row, half);
        wait_mma();                        // stage-4 weight gradients done: H2 is free
        if (has_next) load_tile_async(sH2, p.act + act_stride, nrow0, Mrows, tid);
        cp_async_commit();                 // G3'
        bwd_epilogue_half<true>(trow + C_DG + 32 * half, mk, sGB, row, half);
        publish();
        // ---- stage 6: dX = dH1 W1 (N = 32) -> global;   wgrad W1 (A = dH1^T, B = x_en)
        if (issuer_warp && umma::elect_one()) {
            dgrad(sGB, 0, sW + B_W1T, 0, 4, ID32, true);
            umma::commit(&bar);
            wgrad(C_W1, sGB, sXVc, 0, WG32);
        }
        wait_mma();                        // stage-5 weight gradients done: H1 is free
        if (has_next) load_tile_async(sH1, p.act, nrow0, Mrows, tid);
        cp_async_commit();                 // G4'
        {
            uint32_t a[16];
            umma::tmem_ld16(trow + C_DG + 16 * half, a);
            umma::tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; j++) bad_dx |= !(fabsf(__uint_as_float(a[j])) <= 65504.0f);   // inf / NaN / fp16 overflow
                uint4 *dst = reinterpret_cast<uint4 *>(p.d_x_en + (size_t)g * 32) + half * 2;
#pragma unroll
                for (uint32_t c = 0; c < 2; c++)
                    dst[c] = make_uint4(pack_h2(__uint_as_float(a[c * 8]), __uint_as_float(a[c * 8 + 1])),
                                        pack_h2(__uint_as_float(a[c * 8 + 2]), __uint_as_float(a[c * 8 + 3])),
                                        pack_h2(__uint_as_float(a[c * 8 + 4]), __uint_as_float(a[c * 8 + 5])),
                                        pack_h2(__uint_as_float(a[c * 8 + 6]), __uint_as_float(a[c * 8 + 7])));
            }
        }
        // (the W1 weight gradient may still be running: nothing it reads -- GB, this tile's XV buffer -- is written
        //  before the next tile's stage-1 commit has completed; the publish() that heads the next tile is the barrier
        //  between this tile's last accumulator read and the next tile's first MMA)
        first_tile = false;
        xv_sel ^= 1u;
    }
    if (p.scaler && __any_sync(0xffffffffu, bad_dx) && (tid & 31u) == 0) scaler_raise(p.scaler);
    cp_async_wait<0>();
    // drain the last weight-gradient MMAs before reading their accumulators
    if (issuer_warp && umma::elect_one()) umma::commit(&bar);
    wait_mma();

    // ---- flush the weight-gradient accumulators: TMEM lane == output neuron n (row of dW); the two column halves of
    //      every accumulator go to the two warps that own the lane quarter
    if (!first_tile && !(p.dbg & 2u) && p.slabs) {
        // Slab layout (kSlab*): one block per accumulator, COLUMN-major within the block -- element (row n, column c) at
        // off + c * rows + n -- so that a warp's store of one accumulator column covers 32 consecutive floats (TMEM lane ==
        // row == thread: the natural row-major slab made every store instruction touch 32 half-filled sectors, and the
        // misaligned / remapped columns of the colour layer went out as scalar stores: 9.4 us of the kernel at configs[1]).
        // The mapping to the tcnn parameter layout happens once, in k_field_wgrad_reduce.
        float *sl = p.slabs + (size_t)blockIdx.x * kWgradFloats;
        const uint32_t wq = warp & 3u;              // lane quarter: rows 32 wq .. 32 wq + 31
        auto flush32 = [&](uint32_t col, uint32_t off, uint32_t rows, uint32_t row_base) {      // 64-column accumulator
            if (wq * 32u >= row_base + rows || wq * 32u + 32u <= row_base) return;              // (warp-uniform)
            uint32_t r[32];
            umma::tmem_ld32(trow + col + 32 * half, r);
            umma::tmem_ld_wait();
            if (row >= row_base && row < row_base + rows) {
                float *dst = sl + off + (size_t)(32 * half) * rows + (row - row_base);
#pragma unroll
                for (int j = 0; j < 32; j++) dst[(size_t)j * rows] = __uint_as_float(r[j]);
            }
        };
        auto flush16 = [&](uint32_t col, uint32_t off, uint32_t rows) {                         // 32-column accumulator
            if (wq * 32u >= rows) return;
            uint32_t r[16];
            umma::tmem_ld16(trow + col + 16 * half, r);
            umma::tmem_ld_wait();
            float *dst = sl + off + (size_t)(16 * half) * rows + row;
#pragma unroll
            for (int j = 0; j < 16; j++) dst[(size_t)j * rows] = __uint_as_float(r[j]);
        };
        flush16(C_W1, kSlabW1, 64);
        flush32(C_W2, kSlabW2, 64, 0);
        flush32(C_W3, kSlabW3, 64, 0);
        flush32(C_PAIR, kSlabPair, 128, 0);
        flush16(C_R1V, kSlabR1V, 64);
        flush32(C_R2, kSlabR2, 16, 0);              // A = T16: rows 0..15 = colour outputs (4 real)
        flush32(C_D2, kSlabD2, 16, 16);             //          rows 16..31 = density outputs (1 real)
    } else if (!first_tile && !(p.dbg & 2u)) {
        const uint32_t n = row;
        const bool slab = false;
        float *gt = p.g_trunk, *gd = p.g_density, *gr = p.g_rgb;
        auto flush = [&](uint32_t col, uint32_t ncols, float *dst, bool on) {
            const uint32_t hc = ncols / 2;
            for (uint32_t c0 = half * hc; c0 < (half + 1) * hc; c0 += 16) bwd_flush16(trow + col + c0, dst + c0, on, slab);
        };
        // trunk
        flush(C_W1, 32, gt + T_W1 + n * 32, n < 64);
        flush(C_W2, 64, gt + T_W2 + n * 64, n < 64);
        flush(C_W3, 64, gt + T_W3 + n * 64, n < 64);
        // pair: rows 0..63 = colour layer 0 (fea columns 27..90), rows 64..127 = density layer 0.  Columns 27..90 of a
        // 96-float row are not 16-byte aligned: scalar stores / atomics
        {
            float *dst = n < 64 ? gr + R_W1 + n * 96 + 27 : gd + D_W1 + (n - 64) * 64;
            for (uint32_t c0 = half * 32; c0 < (half + 1) * 32; c0 += 16) {
                if (n >= 64) { bwd_flush16(trow + C_PAIR + c0, dst + c0, true, slab); continue; }
                uint32_t r[16];
                umma::tmem_ld16(trow + C_PAIR + c0, r);
                umma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (slab) dst[c0 + j] = __uint_as_float(r[j]);
                    else atomicAdd(dst + c0 + j, __uint_as_float(r[j]));
                }
            }
        }
        // colour layer 0, view columns: internal col j -> lane j (j < 27) or 91 + (j - 27)
        {
            const uint32_t c0 = half * 16;
            uint32_t r[16];
            umma::tmem_ld16(trow + C_R1V + c0, r);
            umma::tmem_ld_wait();
            if (n < 64) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t col = c0 + j, src = (col < 27) ? col : 91 + (col - 27);
                    if (slab) gr[R_W1 + n * 96 + src] = __uint_as_float(r[j]);
                    else atomicAdd(gr + R_W1 + n * 96 + src, __uint_as_float(r[j]));
                }
            }
        }
        // heads: A = T16, rows 0..15 = colour outputs (4 real), rows 16..31 = density outputs (1 real)
        flush(C_R2, 64, gr + R_W2 + n * 64, n < 4);
        flush(C_D2, 64, gd + D_W2, n == 16);
    }
    if (timed_out && (tid & 31u) == 0) {    // a bounded wait gave up: status word, else poison the feature gradient
        if (p.status) atomicOr(p.status, kStatusFieldBwdTimeout);
        else if (tid == 0) p.d_x_en[0] = __ushort_as_half((unsigned short)0x7e00);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, kTmemColsBwd);
}

// slabs -> += flat gradients as a kernel of its own (nb200_field_backward; the fused train step runs the same blocks inside
// its table-scatter launch instead: wgrad_reduce.cuh)
__global__ void __launch_bounds__(256)
k_field_wgrad_reduce(const NbWgradRed r) {
    __shared__ float part[4][64];
    wgrad_reduce_block(r, blockIdx.x, part);
}

constexpr int kMaxDevices = 64;
uint32_t *g_status_word[kMaxDevices] = {};      // per device ordinal; registered by nb200_set_kernel_status_word

template <bool kColor>
int launch_field_forward(const FieldFwdArgs &a, const CUtensorMap &act_map, uint32_t M, cudaStream_t st) {
    // per device: the attribute belongs to the function on the CURRENT device (one flag per device ordinal)
    static bool configured[kMaxDevices] = {};
    const int smem = (int)S_FWD_BYTES + 1024;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_field_forward<kColor>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < kMaxDevices) configured[dev] = true;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t ntiles = (M + 127) / 128;
    const uint32_t grid = ntiles < (uint32_t)(2 * sms) ? ntiles : (uint32_t)(2 * sms);
    k_field_forward<kColor><<<grid, 128, smem, st>>>(a, act_map);
    NB_LAUNCH_CHECK();
    return 0;
}
}  // namespace

uint32_t nb_wgrad_reduce_blocks() { return nb_div_up(kWgradFloats, 64); }
uint32_t nb_field_backward_grid(uint32_t M) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t ntiles = (M + 127) / 128;
    return ntiles < (uint32_t)sms ? ntiles : (uint32_t)sms;
}

uint32_t *nb_kernel_status_word() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    return g_status_word[dev];
}

extern "C" {

int nb200_set_kernel_status_word(uint32_t *word) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= kMaxDevices) return NB200_E_BAD_ARG;
    g_status_word[dev] = word;
    return 0;
}

int nb200_release_kernel_status_word(uint32_t *word) {     // unregisters `word` if it is the registered one
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < kMaxDevices && g_status_word[dev] == word) g_status_word[dev] = nullptr;
    return 0;
}

uint32_t nb200_field_wgrad_scratch_bytes(void) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (uint32_t)sms * kWgradFloats * (uint32_t)sizeof(float);
}

uint32_t nb200_field_weight_image_bytes(void) { return F_BYTES; }

int nb200_freq_embed(const float *dirs, float *out, uint32_t M, void *stream) {
    if (M == 0) return 0;
    if (!dirs || !out) return NB200_E_BAD_ARG;
    k_freq_embed<<<nb_div_up(M, 256), 256, 0, nb_stream(stream)>>>(dirs, out, M);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_field_pack_weights(const float *trunk, const float *density, const float *rgb, void *fwd_img, void *bwd_img,
                             void *stream) {
    if (!trunk || !density || !rgb || !fwd_img || !bwd_img) return NB200_E_BAD_ARG;
    cudaStream_t st = nb_stream(stream);
    ScalerCommit none;
    memset(&none, 0, sizeof(none));
    k_pack_field_weights<<<kPackBlocks, 256, 0, st>>>(trunk, density, rgb, (uint8_t *)fwd_img, (uint8_t *)bwd_img, none);
    NB_LAUNCH_CHECK();
    return 0;
}

// the same launch with one extra block that runs nb200_scaler_commit's work (saves a launch per train step)
int nb200_field_pack_weights_commit(const float *trunk, const float *density, const float *rgb, void *fwd_img, void *bwd_img,
                                    int32_t *step, uint32_t *scaler, uint32_t *const *peer_scalers, uint32_t world,
                                    int32_t *max_samples, void *stream) {
    if (!trunk || !density || !rgb || !fwd_img || !bwd_img || !step || !scaler || world > NB200_PEER_MAX) return NB200_E_BAD_ARG;
    ScalerCommit sc;
    memset(&sc, 0, sizeof(sc));
    sc.step = step; sc.scaler = scaler; sc.max_samples = max_samples; sc.world = peer_scalers ? world : 0;
    for (uint32_t q = 0; q < sc.world; q++) sc.peers[q] = peer_scalers[q];
    k_pack_field_weights<<<kPackBlocks + 1, 256, 0, nb_stream(stream)>>>(trunk, density, rgb, (uint8_t *)fwd_img, (uint8_t *)bwd_img, sc);
    NB_LAUNCH_CHECK();
    return 0;
}

// the saved activations as a rank-3 tensor [5 planes][M rows][64 halves], boxes of 128 rows, 128B swizzle
static int make_act_map(CUtensorMap *map, void *act, uint32_t M) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return NB200_E_BAD_ARG;
    const cuuint64_t dims[3] = {64, M, 5};
    const cuuint64_t strides[2] = {128, (cuuint64_t)M * 128};
    const cuuint32_t box[3] = {64, 128, 1}, estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, act, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : NB200_E_BAD_ARG;
}

int nb200_field_forward(const void *x_en, const float *xyz, const float *dirs, const void *fwd_img, float *sigma,
                        float *sigma_arg, void *rgba, void *act, uint32_t M, const int32_t *count_dev, void *stream) {
    if (M == 0) return 0;
    if (!x_en || !xyz || !fwd_img || !sigma) return NB200_E_BAD_ARG;
    const bool color = rgba != nullptr;         // rgba == NULL: density only (dirs unused, nothing saved)
    if (color ? !dirs : act != nullptr) return NB200_E_BAD_ARG;
    FieldFwdArgs a;
    a.x_en = (const __half *)x_en; a.xyz = xyz; a.dirs = dirs; a.wimg = (const uint8_t *)fwd_img;
    a.sigma = sigma; a.sigma_arg = sigma_arg; a.rgba = (__half *)rgba; a.act = (__half *)act; a.M = M;
    a.count_dev = count_dev;
    a.status = nb_kernel_status_word();
    CUtensorMap act_map;
    memset(&act_map, 0, sizeof(act_map));
    if (act) {
        if (reinterpret_cast<uintptr_t>(act) & 15u) return NB200_E_BAD_ARG;
        const int rc = make_act_map(&act_map, act, M);
        if (rc) return rc;
    }
    return color ? launch_field_forward<true>(a, act_map, M, nb_stream(stream))
                 : launch_field_forward<false>(a, act_map, M, nb_stream(stream));
}

// reduce = false: the slabs stay in wg_scratch (required) for a reduction the caller runs itself (wgrad_reduce.cuh)
int nb_field_backward_launch(const float *d_sigma, const float *d_rgba, const float *sigma_arg, const void *rgba,
                             const void *x_en, const float *dirs, const void *act, const void *bwd_img, void *d_x_en,
                             float *g_trunk, float *g_density, float *g_rgb, uint32_t M, const int32_t *count_dev,
                             float *wg_scratch, uint32_t *scaler, bool reduce, void *stream) {
    if (M == 0) return 0;
    if ((scaler || !reduce) && !wg_scratch) return NB200_E_BAD_ARG;      // the finite check lives in the slab reduction
    if (!d_sigma || !d_rgba || !sigma_arg || !rgba || !x_en || !dirs || !act || !bwd_img || !d_x_en || !g_trunk ||
        !g_density || !g_rgb)
        return NB200_E_BAD_ARG;
    static bool configured[kMaxDevices] = {};
    const int smem = (int)S_BWD_BYTES + 1024;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_field_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < kMaxDevices) configured[dev] = true;
    }
    FieldBwdArgs a;
    a.d_sigma = d_sigma; a.d_rgba = d_rgba; a.sigma_arg = sigma_arg; a.rgba = (const __half *)rgba;
    a.x_en = (const __half *)x_en; a.dirs = dirs; a.act = (const __half *)act; a.wimg = (const uint8_t *)bwd_img;
    a.d_x_en = (__half *)d_x_en; a.g_trunk = g_trunk; a.g_density = g_density; a.g_rgb = g_rgb; a.M = M;
    a.count_dev = count_dev;
    a.scaler = scaler;
    a.status = nb_kernel_status_word();
    { static int dbg = -1; if (dbg < 0) { const char *e = getenv("NB200_FIELDB_DBG"); dbg = e ? atoi(e) : 0; } a.dbg = (uint32_t)dbg; }
    a.slabs = wg_scratch;
    const uint32_t grid = nb_field_backward_grid(M);
    if (wg_scratch && (reinterpret_cast<uintptr_t>(wg_scratch) & 15u)) return NB200_E_BAD_ARG;
    k_field_backward<<<grid, kBwdThreads, smem, nb_stream(stream)>>>(a);
    NB_LAUNCH_CHECK();
    if (wg_scratch && reduce && !(a.dbg & 4u)) {
        const NbWgradRed r{wg_scratch, count_dev, g_trunk, g_density, g_rgb, scaler, grid, M, nb_wgrad_reduce_blocks()};
        k_field_wgrad_reduce<<<r.blocks, 256, 0, nb_stream(stream)>>>(r);
        NB_LAUNCH_CHECK();
    }
    return 0;
}

int nb200_field_backward(const float *d_sigma, const float *d_rgba, const float *sigma_arg, const void *rgba,
                         const void *x_en, const float *dirs, const void *act, const void *bwd_img, void *d_x_en,
                         float *g_trunk, float *g_density, float *g_rgb, uint32_t M, const int32_t *count_dev,
                         float *wg_scratch, uint32_t *scaler, void *stream) {
    return nb_field_backward_launch(d_sigma, d_rgba, sigma_arg, rgba, x_en, dirs, act, bwd_img, d_x_en, g_trunk, g_density, g_rgb,
                                    M, count_dev, wg_scratch, scaler, true, stream);
}

}  // extern "C"
