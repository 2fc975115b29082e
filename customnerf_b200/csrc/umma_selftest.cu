// umma_selftest.cu -- one-CTA tcgen05 GEMM used by the tests to pin the operand-descriptor conventions of umma.cuh
// (K-major and MN-major 128B-swizzled tiles) against a plain matmul before the fused MLP kernels rely on them.
#include "common.cuh"
#include "umma.cuh"

namespace {

// D[128, N] = A[128, K] * B[N, K]^T, fp16 operands, fp32 accumulate.
//   a_mn == 0: A given as [128][K] row-major (K-major operand)     a_mn == 1: A given as [K][128] (MN-major operand)
//   b_mn == 0: B given as [N][K]   row-major (K-major operand)     b_mn == 1: B given as [K][N]   (MN-major operand)
__global__ void __launch_bounds__(128, 1)
k_umma_selftest(const __half *__restrict__ A, const __half *__restrict__ B, float *__restrict__ D, uint32_t N, uint32_t K,
                uint32_t a_mn, uint32_t b_mn, uint32_t *__restrict__ status) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem;                  // up to 128 x 128 halves = 32 KB
    uint8_t *sB = smem + 32768;          // up to 128 x 128 halves = 32 KB
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t M = 128;

    for (uint32_t i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // ---- place A
    if (!a_mn) {
        for (uint32_t i = tid; i < M * (K / 8); i += 128) {           // one 16-byte chunk (8 halves) per iteration
            const uint32_t m = i / (K / 8), c = i % (K / 8);
            const uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)m * K + c * 8);
            *reinterpret_cast<uint4 *>(sA + (c / 8) * (M * 128) + umma::sw128_offset(m, c % 8)) = v;
        }
    } else {
        for (uint32_t i = tid; i < K * (M / 8); i += 128) {
            const uint32_t k = i / (M / 8), c = i % (M / 8);          // c: chunk along M
            const uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)k * M + c * 8);
            *reinterpret_cast<uint4 *>(sA + (c / 8) * (K * 128) + umma::sw128_offset(k, c % 8)) = v;
        }
    }
    // ---- place B
    if (!b_mn) {
        for (uint32_t i = tid; i < N * (K / 8); i += 128) {
            const uint32_t n = i / (K / 8), c = i % (K / 8);
            const uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)n * K + c * 8);
            *reinterpret_cast<uint4 *>(sB + (c / 8) * (N * 128) + umma::sw128_offset(n, c % 8)) = v;
        }
    } else {
        for (uint32_t i = tid; i < K * (N / 8); i += 128) {
            const uint32_t k = i / (N / 8), c = i % (N / 8);
            const uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)k * N + c * 8);
            *reinterpret_cast<uint4 *>(sB + (c / 8) * (K * 128) + umma::sw128_offset(k, c % 8)) = v;
        }
    }
    umma::fence_proxy_async();

    if (warp == 0) umma::tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_fence_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_f16(M, N, a_mn, b_mn);
        const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
        for (uint32_t ks = 0; ks < K / 16; ks++) {
            uint64_t da, db;
            if (!a_mn) da = umma::make_desc(a0 + (ks / 4) * (M * 128) + (ks % 4) * 32, 16, 1024, umma::kLayoutSW128);
            else       da = umma::make_desc(a0 + ks * 2048, K * 128, 1024, umma::kLayoutSW128);
            if (!b_mn) db = umma::make_desc(b0 + (ks / 4) * (N * 128) + (ks % 4) * 32, 16, 1024, umma::kLayoutSW128);
            else       db = umma::make_desc(b0 + ks * 2048, K * 128, 1024, umma::kLayoutSW128);
            umma::mma_f16_ss(tmem_base, da, db, idesc, ks > 0);
        }
        umma::commit(&bar);
    }
    const bool ok = umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    if (!ok) {
        if (tid == 0) *status = 1;
    } else {
        for (uint32_t c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            umma::tmem_ld16(tmem_base + ((warp * 32u) << 16) + c0, r);
            umma::tmem_ld_wait();
            const uint32_t row = warp * 32 + lane;
#pragma unroll
            for (int j = 0; j < 16; j++) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 128);
}

}  // namespace

extern "C" int nb200_umma_selftest(const void *A, const void *B, float *D, uint32_t N, uint32_t K, uint32_t a_mn,
                                   uint32_t b_mn, uint32_t *status, void *stream) {
    if (N % 16 || N < 16 || N > 128 || K % 16 || K < 16 || K > 128) return NB200_E_BAD_ARG;
    const int smem = 65536 + 1024;
    cudaError_t e = cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    k_umma_selftest<<<1, 128, smem, nb_stream(stream)>>>((const __half *)A, (const __half *)B, D, N, K, a_mn, b_mn, status);
    NB_LAUNCH_CHECK();
    return 0;
}
