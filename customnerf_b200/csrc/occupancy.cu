// occupancy.cu -- the occupancy-grid update after the density query, on the device and without a host round trip.
//
// Replaces the tail of NeRFRenderer.update_extra_state (nerf/renderer.py:1700-1715): the boolean-mask EMA-max update
// (nonzero + gather + scatter), torch.mean(...).item(), the Python min() for the threshold, packbits with a host float, and
// step_counter[...].sum().item() -- ~10 torch kernels and two blocking reads -- by three launches:
//   k_occ_ema       grid = max(grid * decay, tmp) where grid >= 0, per-block partial sums of the valid entries (double)
//   k_occ_finalize  mean density (fixed summation order: deterministic, so ray-sharded replicas keep bit-identical bit fields),
//                   threshold = min(mean, density_thresh), mean_count = int(sum(step_counter[:k, 0]) / k)
//   k_packbits_dev  raymarching.cu:279-288 with the threshold read from device memory
// The density query itself is nb200_occ_density (field_fused.cu: cell -> jittered position -> grid gather -> trunk + density head).
#include "common.cuh"

namespace {

constexpr uint32_t kOccBlocks = 592;      // 4 per SM; the partial-sum buffer has this many entries

// rows [row0, row0 + n) of the density query in tmp_grid order (cascade-major, Morton index within a cascade) -> positions.
// update_extra_state (renderer.py:1680-1690): cell (x, y, z) <- Morton index, position = centre * (bound_c - hgs) +
// (2 u - 1) * hgs with hgs = bound_c / G, bound_c = min(2^cas, bound); every product / sum rounded separately, as the chain
// of torch kernels rounds them (the same lines as the SRC_OCC producer of field_fused.cu: the two paths are bit-identical)
__global__ void __launch_bounds__(256)
k_occ_positions(const float *__restrict__ cell_xyz, const float *__restrict__ noise, uint32_t G, float bound, uint32_t row0,
                uint32_t n, float *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = row0 + i, G3 = G * G * G;
    const uint32_t cas = g / G3, m = g - cas * G3;
    const uint32_t cx = nb_morton3D_invert(m), cy = nb_morton3D_invert(m >> 1), cz = nb_morton3D_invert(m >> 2);
    const size_t c = ((size_t)cx * G + cy) * G + cz;
    const float bc = fminf((float)(1u << cas), bound), hgs = bc / (float)G, span = bc - hgs;
    float px = __fmul_rn(__ldg(cell_xyz + c * 3), span), py = __fmul_rn(__ldg(cell_xyz + c * 3 + 1), span),
          pz = __fmul_rn(__ldg(cell_xyz + c * 3 + 2), span);
    if (noise) {
        const float *u = noise + ((size_t)cas * G3 + c) * 3;
        px = __fadd_rn(px, __fmul_rn(__fsub_rn(__fmul_rn(__ldg(u), 2.0f), 1.0f), hgs));
        py = __fadd_rn(py, __fmul_rn(__fsub_rn(__fmul_rn(__ldg(u + 1), 2.0f), 1.0f), hgs));
        pz = __fadd_rn(pz, __fmul_rn(__fsub_rn(__fmul_rn(__ldg(u + 2), 2.0f), 1.0f), hgs));
    }
    out[(size_t)i * 3] = px; out[(size_t)i * 3 + 1] = py; out[(size_t)i * 3 + 2] = pz;
}

// state (8 x 4 bytes): [0] mean density f32, [1] threshold f32, [2] mean_count i32 (unchanged when total_step == 0),
// [3] number of valid cells u32
__global__ void __launch_bounds__(256)
k_occ_ema(float *__restrict__ grid, const float *__restrict__ tmp, uint32_t n, float decay, double *__restrict__ part_sum,
          uint32_t *__restrict__ part_cnt) {
    double acc = 0.0;
    uint32_t cnt = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        if (g >= 0.0f) {                                    // renderer.py:1701-1702 (NaN fails the test, as there)
            if (tmp) { g = fmaxf(__fmul_rn(g, decay), tmp[i]); grid[i] = g; }
            acc += (double)g;
            cnt++;
        }
    }
    __shared__ double s_sum[256];
    __shared__ uint32_t s_cnt[256];
    s_sum[threadIdx.x] = acc; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (uint32_t o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part_sum[blockIdx.x] = s_sum[0]; part_cnt[blockIdx.x] = s_cnt[0]; }
}

__global__ void __launch_bounds__(256)
k_occ_finalize(const double *__restrict__ part_sum, const uint32_t *__restrict__ part_cnt, uint32_t nparts, float density_thresh,
               const int32_t *__restrict__ step_counter, uint32_t total_step, float *__restrict__ state) {
    __shared__ double s_sum[256];
    __shared__ uint32_t s_cnt[256];
    double acc = 0.0;
    uint32_t cnt = 0;
    for (uint32_t i = threadIdx.x; i < nparts; i += 256) { acc += part_sum[i]; cnt += part_cnt[i]; }
    s_sum[threadIdx.x] = acc; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (uint32_t o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const float mean = (float)(s_sum[0] / (double)s_cnt[0]);      // 0 / 0 = NaN, as torch.mean of an empty selection
    state[0] = mean;
    state[1] = (density_thresh < mean) ? density_thresh : mean;   // Python min(mean, thresh), NaN behaviour included
    reinterpret_cast<uint32_t *>(state)[3] = s_cnt[0];
    if (total_step > 0 && step_counter) {
        long long s = 0;
        for (uint32_t k = 0; k < total_step; k++) s += step_counter[2 * k];
        reinterpret_cast<int32_t *>(state)[2] = (int32_t)((double)s / (double)total_step);      // int(sum / total_step)
    }
}

__global__ void k_packbits_dev(const float *__restrict__ grid, uint32_t N, const float *__restrict__ thresh_dev,
                               uint8_t *__restrict__ bitfield) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float thresh = __ldg(thresh_dev);
    const float4 a = __ldg(reinterpret_cast<const float4 *>(grid) + (size_t)n * 2);
    const float4 b = __ldg(reinterpret_cast<const float4 *>(grid) + (size_t)n * 2 + 1);
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;   bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;   bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;  bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;  bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

}  // namespace

extern "C" {

// The density query as encoder + density-only field launches over chunks of the grid (the standalone encoder runs at full
// occupancy with the whole L1; the one-kernel form nb200_occ_density has 8 gather warps per SM): per chunk positions ->
// nb200_fs_encode_forward -> nb200_field_forward(rgba = NULL) writing sigma straight into tmp_grid (rows are in tmp_grid
// order).  chunk_rows samples of scratch: xyz_buf f32 [chunk_rows,3], x_en_buf f16 [chunk_rows,32] -- sized to stay in L2.
int nb200_occ_density_chunked(const float *cell_xyz, const float *noise, uint32_t G, uint32_t cascade, float bound,
                              const float *table, const int32_t *offsets, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                              int align_corners, uint32_t interp, const void *fwd_img, float *tmp_grid, float *xyz_buf,
                              void *x_en_buf, uint32_t chunk_rows, void *stream) {
    if (!cell_xyz || !table || !offsets || !fwd_img || !tmp_grid || !xyz_buf || !x_en_buf || !(bound > 0.0f) || G == 0 ||
        G > 1024 || cascade == 0 || cascade > 8 || (uint64_t)cascade * G * G * G > 0x7fffffffull || chunk_rows == 0)
        return NB200_E_BAD_ARG;
    const uint32_t total = cascade * G * G * G;
    for (uint32_t row0 = 0; row0 < total; row0 += chunk_rows) {
        const uint32_t n = total - row0 < chunk_rows ? total - row0 : chunk_rows;
        k_occ_positions<<<nb_div_up(n, 256), 256, 0, nb_stream(stream)>>>(cell_xyz, noise, G, bound, row0, n, xyz_buf);
        NB_LAUNCH_CHECK();
        int rc;
        if ((rc = nb200_fs_encode_forward(xyz_buf, bound, table, offsets, x_en_buf, n, L, S, H, gridtype, align_corners, interp,
                                          nullptr, stream))) return rc;
        if ((rc = nb200_field_forward(x_en_buf, xyz_buf, nullptr, fwd_img, tmp_grid + row0, nullptr, nullptr, nullptr, n,
                                      nullptr, stream))) return rc;
    }
    return 0;
}

uint32_t nb200_occ_scratch_bytes(void) { return kOccBlocks * (uint32_t)(sizeof(double) + sizeof(uint32_t)); }

int nb200_occ_finalize(float *density_grid, const float *tmp_grid, uint32_t n_cells, float decay, float density_thresh,
                       const int32_t *step_counter, uint32_t total_step, uint8_t *bitfield, float *state, void *scratch,
                       void *stream) {
    if (!density_grid || !bitfield || !state || !scratch || n_cells == 0 || (n_cells & 7u) || total_step > 16) return NB200_E_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(density_grid) & 15u) || (reinterpret_cast<uintptr_t>(scratch) & 7u)) return NB200_E_BAD_ARG;
    cudaStream_t st = nb_stream(stream);
    double *part_sum = (double *)scratch;
    uint32_t *part_cnt = (uint32_t *)(part_sum + kOccBlocks);
    const uint32_t want = nb_div_up(n_cells, 256), blocks = want < kOccBlocks ? want : kOccBlocks;
    k_occ_ema<<<blocks, 256, 0, st>>>(density_grid, tmp_grid, n_cells, decay, part_sum, part_cnt);
    NB_LAUNCH_CHECK();
    k_occ_finalize<<<1, 256, 0, st>>>(part_sum, part_cnt, blocks, density_thresh, step_counter, total_step, state);
    NB_LAUNCH_CHECK();
    k_packbits_dev<<<nb_div_up(n_cells / 8, 256), 256, 0, st>>>(density_grid, n_cells / 8, state + 1, bitfield);
    NB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
