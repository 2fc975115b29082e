/*
 * nerf_b200.h -- C ABI of libnerf_b200.so: the B200 (sm_100a) replacement for CustomNeRF's two native
 * extensions (gridencoder, raymarching) plus the field-network MLP the reference takes from tiny-cuda-nn.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer on the current device unless stated otherwise;
 *   - the caller owns all memory; nothing here allocates (scratch is passed in) or retains pointers;
 *   - sizes are uint32_t, scalars float, exactly as the reference's pybind entry points pass them;
 *   - the last argument is the cudaStream_t to launch on (as void*; NULL = legacy default stream) --
 *     the reference always launches on the legacy default stream;
 *   - the return value is 0 on success, otherwise a cudaError_t (> 0) or one of NB200_E_* (< 0);
 *     nb200_error_string() turns either into text.  The reference returns void and surfaces faults
 *     asynchronously; its TORCH_CHECKs become NB200_E_* codes that the Python layer re-raises as
 *     RuntimeError with the reference's wording.
 *
 * Each entry point names the reference function it replaces as file:line under /root/reference.
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dtype tags for void* tensors */
#define NB200_F32 0
#define NB200_F16 1
#define NB200_F32_AS_F16 2   /* grid_encode_forward fast path only: fp32 table, every value rounded to fp16 as it is
                                loaded, fp16 outputs -- autocast semantics of grid.py:45-46 without the table copy */

/* error codes (negative; positive values are cudaError_t) */
#define NB200_E_BAD_DIM   (-1)   /* "GridEncoding: C must be 1, 2, 4, or 8." / D must be 2..5 (gridencoder.cu:380,397) */
#define NB200_E_BAD_DTYPE (-2)   /* unsupported dtype tag (the reference's f64 instantiations are not built) */
#define NB200_E_BAD_ARG   (-3)   /* null pointer / inconsistent sizes */
#define NB200_E_SCRATCH   (-4)   /* scratch buffer too small */

const char *nb200_error_string(int code);
int nb200_version(void);

/* ============================================================================================
 * gridencoder  (reference: gridencoder/src/gridencoder.h:12-15, bindings.cpp:5-7)
 * ========================================================================================== */

/* Output layout of grid_encode_forward / input layout of grid_encode_backward's grad. */
#define NB200_LAYOUT_LBC 0   /* [L, B, C]: the reference's native layout (gridencoder.cu:384-389)            */
#define NB200_LAYOUT_BLC 1   /* [B, L*C]: what grid.py:63 / :81 produce with an extra permute copy; written
                                directly here so the copy disappears                                        */

/* Replaces grid_encode_forward (gridencoder.cu:447-470 -> kernel_grid :87-244).
 *   inputs      f32 [B, D] in [0,1]            embeddings  emb_dtype [rows, C]
 *   offsets     i32 [L+1]                      outputs     emb_dtype, layout per `layout`
 *   dy_dx       emb_dtype [B, L*D*C] or NULL   S = log2(per_level_scale), H = base resolution
 *   gridtype 0 = hash, 1 = tiled; interp 0 = linear, 1 = smoothstep.
 * Levels >= max_level are not written (caller zero-fills, grid.py:52).
 * fp16 mode accumulates in fp32 and rounds once (the reference rounds twice per corner). */
int nb200_grid_encode_forward(const float *inputs, const void *embeddings, const int32_t *offsets, void *outputs,
                              uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                              void *dy_dx, uint32_t gridtype, int align_corners, uint32_t interp,
                              int emb_dtype, int layout, void *stream);

/* Replaces grid_encode_backward (gridencoder.cu:472-502 -> kernel_grid_backward :247-339,
 * kernel_input_backward :342-368).
 *   grad             grad_dtype, layout per `layout`          grad_embeddings  f32 [rows, C], ACCUMULATED INTO
 *   dy_dx            grad_dtype [B, L*D*C] or NULL            grad_inputs      grad_dtype [B, D] or NULL
 * grad_embeddings is always fp32 (the reference scatters __half2 atomics under AMP); the caller zero-fills it
 * (grid.py:83).  agg: 0 = one atomic per corner, 1 = warp-aggregated scatter (adjacent samples of a ray that share
 * a cell are summed in registers and issue one atomic). */
int nb200_grid_encode_backward(const void *grad, const float *inputs, const int32_t *offsets, float *grad_embeddings,
                               uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                               const void *dy_dx, void *grad_inputs, uint32_t gridtype, int align_corners,
                               uint32_t interp, int grad_dtype, int layout, int agg, void *stream);

/* Replaces grad_total_variation (gridencoder.cu:638-644 -> kernel_grad_tv :505-609).  f32 only
 * (grid.py:171 forces autocast off).  inputs f32 [B, D] in [0,1]; adds into grad f32 [rows, C]. */
int nb200_grad_total_variation(const float *inputs, const float *embeddings, float *grad, const int32_t *offsets,
                               float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                               uint32_t gridtype, int align_corners, void *stream);

/* scales[l] = exp2f(l * S) * H - 1.0f evaluated on the device (gridencoder.cu:138); test aid: lets the CPU
 * oracle use the device's exp2f values.  scales f32 [L]. */
int nb200_grid_level_scales(float *scales, uint32_t L, float S, uint32_t H, void *stream);

/* fp32 -> fp16 copy of the table (the persistent half "shadow" that replaces grid.py:45-46's per-call cast). */
int nb200_cast_f32_to_f16(const float *src, void *dst, uint64_t n, void *stream);

/* ============================================================================================
 * raymarching  (reference: raymarching/src/raymarching.h:7-22, bindings.cpp:5-20)
 * ========================================================================================== */

/* Replaces near_far_from_aabb (raymarching.cu:148-156 -> kernel :91-145). */
int nb200_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N, float min_near,
                             float *nears, float *fars, void *stream);
/* Replaces sph_from_ray (raymarching.cu:201-209 -> kernel :162-199). */
int nb200_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords, void *stream);
/* Replaces morton3D / morton3D_invert (raymarching.cu:229-232, :257-260). */
int nb200_morton3D(const int32_t *coords, uint32_t N, int32_t *indices, void *stream);
int nb200_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords, void *stream);
/* Replaces packbits (raymarching.cu:292-300 -> kernel :267-289).  N = number of output bytes. */
int nb200_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream);

/* march_rays_train (raymarching.cu:482-490 -> kernel :311-480), split in two so the caller can size the outputs
 * exactly instead of allocating N*max_steps rows (raymarching.py:196-209):
 *
 *   nb200_march_rays_train_count : march every ray once, counts -> rays[n] = (n, offset, count) with offsets the
 *       exclusive scan of the counts in ray-id order STARTING AT counter[0]'s entry value (one valid member of
 *       the reference's atomicAdd-ordered output set); then counter[0] += sum, counter[1] += N.
 *       scratch: i32, at least nb200_march_scratch_ints(N) entries.
 *       The count pass also records the parameter t of every sample (up to a per-ray cap) in `scratch`.
 *   nb200_march_rays_train_write : write xyzs/dirs/deltas for rays whose segment fits in M (raymarching.cu:415-416)
 *       by expanding the records of the count pass (`scratch` = the same buffer, same rays; one warp per ray,
 *       coalesced stores), or by marching again when scratch is NULL / a ray has more samples than records.
 *       Rows not covered by a segment are left untouched (caller zero-fills, raymarching.py:206-208).
 *   nb200_march_rays_train       : both, back to back == the reference entry point. */
uint32_t nb200_march_scratch_ints(uint32_t N);
int nb200_march_rays_train_count(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                 const float *nears, const float *fars, const float *noises,
                                 int32_t *rays, int32_t *counter, int32_t *scratch, void *stream);
int nb200_march_rays_train_write(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                 const float *nears, const float *fars, const float *noises, const int32_t *rays,
                                 float *xyzs, float *dirs, float *deltas, const int32_t *scratch, void *stream);
int nb200_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                           uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                           const float *fars, float *xyzs, float *dirs, float *deltas, int32_t *rays, int32_t *counter,
                           const float *noises, int32_t *scratch, void *stream);

/* Replaces composite_rays_train_forward (+_sdf, byte-identical math) (raymarching.cu:660-679 -> kernel :500-577). */
int nb200_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                                       uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *depth,
                                       float *image, void *stream);
/* Replaces composite_rays_train_backward (+_sdf) (raymarching.cu:860-878 -> kernel :691-772).  Every row of a ray's
 * segment is written (zeros after the early-out), rows outside any segment are left untouched. */
int nb200_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image, const float *sigmas,
                                        const float *rgbs, const float *deltas, const int32_t *rays,
                                        const float *weights_sum, const float *image, uint32_t M, uint32_t N,
                                        float T_thresh, float *grad_sigmas, float *grad_rgbs, void *stream);

/* Replaces march_rays (raymarching.cu:992-999 -> kernel :884-989).  Outputs must be zero-filled by the caller. */
int nb200_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                     const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                     uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                     float *xyzs, float *dirs, float *deltas, const float *noises, void *stream);
/* Replaces composite_rays (raymarching.cu:1092-1098 -> kernel :1002-1089).  In-place on rays_alive / rays_t /
 * weights_sum / depth / image. */
int nb200_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t,
                         const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum, float *depth,
                         float *image, void *stream);

/* Device-driven inference rounds: the eval loop of nerf/renderer.py:651-688 without its per-round host read-back.
 * state (device i32[8]) = {n_alive, n_step, step, n_alive * n_step, survivor count of the running compaction, ...}.  One round:
 *   nb200_infer_plan          n_step = max(min(N / n_alive, 8), 1) (:676); n_alive = 0 once step >= max_steps (:667)
 *   nb200_march_rays_dev      nb200_march_rays for the state's n_alive / n_step (launched for N rays; unused sample
 *                             slots are zero-filled by the kernel; noises only applied while step == 0)
 *   nb200_fs_encode_forward + nb200_field_forward with count_dev = &state[3]
 *   nb200_composite_rays_dev  nb200_composite_rays reading the field kernel's rgba f16 [M,4] rows
 *   nb200_compact_alive       rays_alive = rays_alive[rays_alive >= 0] (:685; survivors in arbitrary order, which only
 *                             decides their buffer slots in the next round); n_alive, step += n_step
 * Every kernel is bounded by the state, so a round can be captured in a CUDA graph and replayed. */
int nb200_infer_plan(int32_t *state, uint32_t N, uint32_t max_steps, void *stream);
int nb200_march_rays_dev(const int32_t *state, uint32_t N, const int32_t *rays_alive, const float *rays_t,
                         const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                         uint32_t C, uint32_t H, const uint8_t *grid, const float *fars, float *xyzs, float *dirs,
                         float *deltas, const float *noises, void *stream);
/* nb200_march_rays_dev served from the sample records of ONE whole-ray traversal per frame: call
 * nb200_march_rays_train_count(rays_o, rays_d, grid, ..., nears, fars, noises, rays, counter, scratch) once (the
 * speculative-segment march; rays i32 [N,3], scratch of nb200_march_scratch_ints(N) ints), zero consumed i32 [N], then this
 * entry every round.  Same samples bit for bit: a ray whose rays_t has left the recorded chain (compositing re-accumulates
 * it from the deltas) or whose records were truncated walks the grid as nb200_march_rays_dev does; state[5] counts them. */
int nb200_march_rays_rec(int32_t *state, uint32_t N, const int32_t *rays_alive, const float *rays_t, const float *rays_o,
                         const float *rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                         const uint8_t *grid, const float *fars, float *xyzs, float *dirs, float *deltas, const float *noises,
                         const int32_t *rays, const int32_t *scratch, int32_t *consumed, void *stream);
int nb200_composite_rays_dev(const int32_t *state, uint32_t N, float T_thresh, int32_t *rays_alive, float *rays_t,
                             const float *sigmas, const void *rgba, const float *deltas, float *weights_sum, float *depth,
                             float *image, void *stream);
int nb200_compact_alive(int32_t *state, uint32_t N, int32_t *rays_alive, int32_t *tmp, void *stream);

/* ============================================================================================
 * field network: replaces the three tinycudann.Network (FullyFusedMLP) modules of nerf/network_grid.py:98-139 and the
 * torch glue of NeRFNetwork.forward / density (:159-193): trunk 32-64-64-64 -> density head 64-64-1 -> trunc_exp with
 * the gaussian bias; [view_en(27) | fea(64)] -> colour head 96-64-(3+1) sigmoid.  One fused tcgen05 kernel per
 * direction.  Parameters are the three flat fp32 tcnn-layout vectors (row-major [out_padded, in_padded] per layer).
 * ========================================================================================== */
uint32_t nb200_field_weight_image_bytes(void);
/* fp32 flat params -> fp16 pre-swizzled operand images (forward and transposed/backward), each
 * nb200_field_weight_image_bytes() long and 16-byte aligned.  Run once per optimiser step. */
int nb200_field_pack_weights(const float *trunk, const float *density, const float *rgb, void *fwd_img, void *bwd_img,
                             void *stream);
/* the same launch with one extra block doing nb200_scaler_commit's work (see there): one launch less per train step */
int nb200_field_pack_weights_commit(const float *trunk, const float *density, const float *rgb, void *fwd_img, void *bwd_img,
                                    int32_t *step, uint32_t *scaler, uint32_t *const *peer_scalers, uint32_t world,
                                    int32_t *max_samples, void *stream);
/* x_en f16 [M,32] (grid encoding), xyz f32 [M,3], dirs f32 [M,3] -> sigma f32 [M], rgba f16 [M,4].
 * Optional (training): sigma_arg f32 [M] (argument of trunc_exp) and act f16 [5,M,64] (h1, h2, fea, hd, hr).
 * count_dev (device i32, may be NULL): when given only rows < min(M, *count_dev) are evaluated. */
int nb200_field_forward(const void *x_en, const float *xyz, const float *dirs, const void *fwd_img, float *sigma,
                        float *sigma_arg, void *rgba, void *act, uint32_t M, const int32_t *count_dev, void *stream);
/* rgba == NULL (then act must be NULL, dirs is unused): trunk + density head only -- NeRFNetwork.density (network_grid.py:179-193) */

/* Backward of nb200_field_forward.  d_sigma f32 [M], d_rgba f32 [M,4] are the upstream gradients; sigma_arg, rgba, act
 * are what the forward saved.  Writes d_x_en f16 [M,32] (gradient of the grid encoding, input of
 * nb200_grid_encode_backward) and ACCUMULATES the three flat fp32 parameter gradients (tcnn layout). */
int nb200_field_backward(const float *d_sigma, const float *d_rgba, const float *sigma_arg, const void *rgba,
                         const void *x_en, const float *dirs, const void *act, const void *bwd_img, void *d_x_en,
                         float *g_trunk, float *g_density, float *g_rgb, uint32_t M, const int32_t *count_dev,
                         float *wg_scratch, uint32_t *scaler, void *stream);
/* wg_scratch: nb200_field_wgrad_scratch_bytes() bytes (16-byte aligned) of per-CTA partial weight-gradient sums that a
 * second small kernel adds into g_* (deterministic, no contended atomics); NULL = fp32 atomics straight on g_*.
 * scaler (may be NULL; needs wg_scratch): loss-scaler words whose found-inf bit is raised when a weight gradient is not finite
 * or a feature gradient (d_x_en) leaves the fp16 range. */
uint32_t nb200_field_wgrad_scratch_bytes(void);
/* Kernel status word of the CURRENT device (no reference counterpart: tcnn's kernels cannot time out).  The field kernels
 * wait for their tensor-core work on mbarriers with a bounded spin; when a wait gives up they OR NB200_STATUS_* into
 * `word` (a device uint32 the caller owns and zeroes; NULL unregisters).  Without a registered word the forward kernels
 * write NaN into sigma[0] and the backward kernel into d_x_en[0] instead, so that a time-out can never pass as a result. */
#define NB200_STATUS_FIELD_FWD_TIMEOUT   1u
#define NB200_STATUS_FIELD_BWD_TIMEOUT   2u
#define NB200_STATUS_FIELD_FUSED_TIMEOUT 4u
int nb200_set_kernel_status_word(uint32_t *word);
int nb200_release_kernel_status_word(uint32_t *word);   /* unregisters `word` if it is the registered one (owner going away) */
/* Grid encoding + field network in ONE kernel (csrc/field_fused.cu): the [M,32] hash-grid features are gathered by producer
 * warps straight into the tensor-core operand tile and never round-trip HBM.  Replaces GridEncoder.forward
 * (gridencoder/grid.py:151-168, gridencoder.cu:87-244) followed by NeRFNetwork.forward / .density (nerf/network_grid.py:159-193).
 * xyz f32 [M,3] in [-bound, bound] (normalised to [0,1] inside, grid.py:156); table = fp32 master embeddings, entries
 * rounded to fp16 on load (the autocast path of grid.py:45-46); L must be 16, D = 3, C = 2.
 *   rgba == NULL (dirs may be NULL): density only (trunk + density head)     -> sigma f32 [M]
 *   rgba != NULL, act == NULL:       inference forward                       -> sigma, rgba f16 [M,4]
 *   act != NULL:                     training forward: also writes x_en f16 [M,32], sigma_arg f32 [M], act f16 [5,M,64]
 *                                    (what nb200_field_backward + nb200_fs_encode_backward read)
 * count_dev as nb200_field_forward. */
int nb200_field_fused_forward(const float *xyz, const float *dirs, float bound, const float *table, const int32_t *offsets,
                              uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp,
                              const void *fwd_img, float *sigma, float *sigma_arg, void *rgba, void *x_en, void *act,
                              uint32_t M, const int32_t *count_dev, void *stream);

/* ============================================================================================
 * occupancy-grid update: replaces NeRFRenderer.update_extra_state (nerf/renderer.py:1658-1715).
 * nb200_occ_density: the density query of every cell of every cascade in one launch (same fused kernel, positions
 * generated in place): cell_xyz f32 [G^3,3] = cell centres in [-1,1] in the reference's x-major order (:1678), noise f32
 * [cascade,G^3,3] = the torch.rand_like draws of :1690 (NULL: no jitter); point of cell c in cascade k:
 * cell_xyz[c] * (b_k - h_k) + (2 noise - 1) * h_k with b_k = min(2^k, bound), h_k = b_k / G, products and sums rounded
 * one by one as the reference's chain of torch ops rounds them.  tmp_grid f32 [cascade, G^3] indexed by Morton code (:1696).
 * nb200_occ_finalize: density_grid = max(density_grid * decay, tmp_grid) where density_grid >= 0 (:1700-1702; tmp_grid NULL:
 * skip), mean of the valid cells (fixed summation order, double accumulation), threshold = min(mean, density_thresh),
 * bitfield = packbits(density_grid, threshold) (:1709), mean_count = int(sum(step_counter[:total_step, 0]) / total_step)
 * (:1712-1714; step_counter i32 [16,2], total_step <= 16, 0: left unchanged).  No host synchronisation: state f32/i32 [8] =
 * {mean density, threshold, mean_count (i32), valid cells (u32), ...} stays on the device.  scratch: nb200_occ_scratch_bytes(). */
int nb200_occ_density(const float *cell_xyz, const float *noise, uint32_t G, uint32_t cascade, float bound, const float *table,
                      const int32_t *offsets, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                      uint32_t interp, const void *fwd_img, float *tmp_grid, void *stream);
/* the same query as encoder + density-only field launches over chunks of chunk_rows cells (default of update_extra_state:
 * measured faster than the one-kernel form, profiles/); xyz_buf f32 [chunk_rows,3], x_en_buf f16 [chunk_rows,32] scratch */
int nb200_occ_density_chunked(const float *cell_xyz, const float *noise, uint32_t G, uint32_t cascade, float bound,
                              const float *table, const int32_t *offsets, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                              int align_corners, uint32_t interp, const void *fwd_img, float *tmp_grid, float *xyz_buf,
                              void *x_en_buf, uint32_t chunk_rows, void *stream);
uint32_t nb200_occ_scratch_bytes(void);
int nb200_occ_finalize(float *density_grid, const float *tmp_grid, uint32_t n_cells, float decay, float density_thresh,
                       const int32_t *step_counter, uint32_t total_step, uint8_t *bitfield, float *state, void *scratch,
                       void *stream);

/* ============================================================================================
 * dense (non-cuda_ray) renderer: the sampler of NeRFRenderer.run (nerf/renderer.py:297-367) and sample_pdf (:21-55).
 * nb200_dense_coarse: z_c[n,i] = near + (far - near) * lin[i] (+ (noise[n,i] - 0.5) * (far - near) / S when noise != NULL),
 *   xyzs[n*S+i] = clamp(o + d * z, aabb) (:306-317); lin f32 [S] = torch.linspace(0, 1, S).
 * nb200_dense_importance (one CTA per ray): coarse weights alpha * cumprod(1 - alpha + 1e-15) (:330-336), inverse-CDF sampling of
 *   S_up depths from weights[1:-1] over the interval mid-points (:338-339, :21-55; u f32 [S_up] shared by all rays -- the
 *   deterministic linspace of eval mode -- or [N,S_up] when u_per_ray), all S + S_up depths sorted (:360-361) and written in
 *   the occupancy path's layout: z_all f32 [N,T], xyzs / dirs f32 [N*T,3], deltas f32 [N*T,2] = (z[k+1] - z[k] | last:
 *   (far - near) / S, ori_z[k] - ori_z[k-1]) with ori_z = clamp((z - near) / (far - near), 0, 1) (:431-432), rays i32 [N,3] =
 *   (n, n*T, T) -- so that nb200_composite_rays_train_* with T_thresh = 0 evaluates weights_sum_i (:420-439) and the field /
 *   LGIE kernels of the occupancy path serve the dense path unchanged.  3 <= S, S + S_up <= 256. */
int nb200_dense_coarse(const float *rays_o, const float *rays_d, const float *nears, const float *fars, const float *aabb,
                       const float *lin, const float *noise, uint32_t N, uint32_t S, float *z_c, float *xyzs, void *stream);
int nb200_dense_importance(const float *rays_o, const float *rays_d, const float *nears, const float *fars, const float *aabb,
                           const float *z_c, const float *sigma_c, const float *u, int u_per_ray, uint32_t N, uint32_t S,
                           uint32_t S_up, float *z_all, float *xyzs, float *dirs, float *deltas, int32_t *rays, void *stream);

/* ============================================================================================
 * L2 residency of the hash table (north_star kernel 1: "per-level tables staged so the working set stays L2-resident"; the
 * reference's intent at gridencoder.cu:387).  nb200_l2_persist_limit sets aside up to `bytes` of the L2 for persisting lines
 * (host pointers: *granted = what the device gave, *max_window = the largest access-policy window it accepts);
 * nb200_stream_access_window makes every later launch on `stream` (CUDA-graph captures included) keep a hit_ratio share of
 * the lines it touches in [base, base + bytes) in that carve-out and treat everything else as streaming; bytes == 0
 * removes the window.  nb200_l2_persist_reset demotes all persisting lines. */
int nb200_l2_persist_limit(uint64_t bytes, uint64_t *granted, uint64_t *max_window);
int nb200_stream_access_window(void *stream, const void *base, uint64_t bytes, float hit_ratio);
int nb200_l2_persist_reset(void);

/* get_embedder(4) of nerf/base.py:42-77 exactly as the field kernels evaluate it: dirs f32 [M,3] -> out f32 [M,27] =
 * [d, sin d, cos d, sin 2d, cos 2d, sin 4d, cos 4d, sin 8d, cos 8d] (one sincos + three angle doublings per component). */
int nb200_freq_embed(const float *dirs, float *out, uint32_t M, void *stream);

/* ============================================================================================
 * fused train step (no single reference counterpart: replaces the Python glue between the ops --
 * NeRFRenderer.run_cuda nerf/renderer.py:597-650, NeRFNetwork.forward nerf/network_grid.py:159-177, the autograd
 * Functions of gridencoder/grid.py:27-95 and raymarching/raymarching.py:239-292, and the loss / GradScaler / Adam
 * sequence of nerf/utils_init_nerf.py:224-234,612-629).  No host synchronisation, no allocation: capturable in a
 * CUDA graph.  The nb200_fs_* entry points are the individual stages; nb200_train_* run them from a plan.
 * ========================================================================================== */

/* count + scan (nb200_march_rays_train_count) and write (nb200_march_rays_train_write) into buffers of M_cap rows.
 * counter[0] += total samples, counter[1] += N as the reference; *m_eff (device i32) = number of leading rows covered
 * by complete ray segments (= min(total, offset of the first ray that does not fit)): the row bound of every later
 * stage.  _count initialises it, _write lowers it.  With aabb != NULL, _count also performs near_far_from_aabb
 * (raymarching.cu:108-144; same arithmetic, bit-identical) and WRITES nears / fars instead of reading them. */
int nb200_fs_march_count(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M_cap, float *nears,
                         float *fars, const float *noises, int32_t *rays, int32_t *counter, int32_t *m_eff,
                         int32_t *scratch, const float *aabb, float min_near, void *stream);
int nb200_fs_march_write(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M_cap, const float *nears,
                         const float *fars, const float *noises, const int32_t *rays, float *xyzs, float *dirs,
                         float *deltas, int32_t *m_eff, const int32_t *scratch, void *stream);
/* grid encode of raw positions xyz in [-bound, bound] (GridEncoder.forward's normalisation grid.py:156 fused in),
 * D = 3, C = 2, fp32 master table rounded to fp16 per load (autocast semantics of grid.py:45-46), x_en f16 [M_cap, 2L]. */
int nb200_fs_encode_forward(const float *xyz, float bound, const float *table, const int32_t *offsets, void *x_en,
                            uint32_t M_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                            uint32_t interp, const int32_t *count_dev, void *stream);
/* scatter of d_x_en f16 [M_cap, 2L] into grad_table f32 [rows, 2] (accumulated; warp-aggregated fp32 atomics). */
int nb200_fs_encode_backward(const void *d_x_en, const float *xyz, float bound, const int32_t *offsets, float *grad_table,
                             uint32_t M_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                             uint32_t interp, const int32_t *count_dev, void *stream);
/* the same scatter for levels [level_begin, level_end) only (the ray-sharded step scatters in two launches, NB200_PLAN_SPLIT) */
int nb200_fs_encode_backward_levels(const void *d_x_en, const float *xyz, float bound, const int32_t *offsets, float *grad_table,
                                    uint32_t M_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                    uint32_t interp, const int32_t *count_dev, uint32_t level_begin, uint32_t level_end,
                                    void *stream);
/* composite_rays_train forward / backward reading the field kernel's rgba f16 [M,4] rows directly (the reference
 * slices [..., :3] and casts to float, renderer.py:510,635) and writing grad_rgba as float4 rows [g_r, g_g, g_b, 0]. */
int nb200_fs_composite_forward(const float *sigmas, const void *rgba, const float *deltas, const int32_t *rays,
                               uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *depth, float *image,
                               const float *target, float inv_n, float loss_scale, float *loss, float *g_image,
                               const float *target_mask, float mask_weight, float *render_mask, float *g_render_mask,
                               const float *loss_scale_dev, void *stream);
/* loss_scale_dev (may be NULL): the loss scale is read from this device word instead of `loss_scale` (dynamic scaler).
 * target != NULL: nb200_mse_loss_grad fused in (per ray).  target_mask != NULL (the reference's train_conf term,
 * utils_init_nerf.py:231-233): render_mask[n] = sum_i w_i * mask_i (mask = 4th channel of the rgba rows; what
 * weights_sum_i renders on the dense path, renderer.py:460-463), loss += mask_weight * mean((render_mask - target_mask)^2),
 * g_render_mask = its gradient times loss_scale. */
int nb200_fs_composite_backward(const float *grad_weights_sum, const float *grad_image, const float *sigmas,
                                const void *rgba, const float *deltas, const int32_t *rays, const float *weights_sum,
                                const float *image, uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas,
                                float *grad_rgba, const float *grad_render_mask, const float *render_mask, void *stream);
/* both in ONE launch (target required): the warp that composited a ray walks its samples again for the backward -- same
 * outputs (weights_sum, depth, image, render_mask, loss, g_image, g_render_mask, grad_sigmas, grad_rgba) as the two calls.
 * grad_weights_sum may be NULL (0). */
int nb200_fs_composite_fused(const float *sigmas, const void *rgba, const float *deltas, const int32_t *rays, uint32_t M,
                             uint32_t N, float T_thresh, float *weights_sum, float *depth, float *image, const float *target,
                             float inv_n, float loss_scale, float *loss, float *g_image, const float *target_mask,
                             float mask_weight, float *render_mask, float *g_render_mask, const float *loss_scale_dev,
                             const float *grad_weights_sum, float *grad_sigmas, float *grad_rgba, void *stream);
/* grad_render_mask / render_mask (both or neither): the rendered mask as a 4th composited channel; its gradient lands in
 * grad_rgba[:, 3] and in grad_sigmas. */
/* loss[0] += sum((image - target)^2) * inv_n ;  g_image = 2 (image - target) * inv_n * loss_scale.
 * (F.mse_loss of utils_init_nerf.py:224 with inv_n = 1 / (3 * N_total), and its gradient times the loss scale.) */
int nb200_mse_loss_grad(const float *image, const float *target, uint32_t N, float inv_n, float loss_scale, float *loss,
                        float *g_image, void *stream);
/* One pass of Adam over a flat fp32 parameter vector with two hyper-parameter groups (elements [0, split) and
 * [split, n)): gradient unscale, moment update, parameter update and (zero_grad) gradient reset, 32 B per parameter.
 * hyper: device f32 [2][8] = {lr, beta1, beta2, eps, 1 - beta1^t, sqrt(1 - beta2^t), grad_scale, 0} per group.
 * The arithmetic of torch.optim.Adam (main.py:182); the last division m / denom is the 2-ulp fast form (the update is
 * bounded by lr, so its error is ~1e-10 absolute; tests/test_gpu_fused_step.py bounds the drift against torch at 1e-6).
 * All pointers 16-byte aligned, split % 4 == 0. */
int nb200_fused_adam(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, uint64_t split,
                     const float *hyper, int zero_grad, void *stream);

/* t = ++(*step) and the two hyper groups for optimiser step t, computed on the device (so a replayed graph never
 * depends on host memory): sched f32[8] = {lr0 group 0, lr0 group 1, beta1, beta2, eps, grad_scale, decay_base,
 * decay_iters}; lr = lr0 * decay_base^min((t-1)/decay_iters, 1) (LambdaLR of main.py:189; decay_iters <= 0: constant). */
int nb200_adam_hyper(int32_t *step, const float *sched, float *hyper, void *stream);
/* nb200_fused_adam on an explicit launch shape: `grid` CTAs of `threads` (256 or 512) threads, `unroll` (1, 2 or 4) float4
 * groups of every vector in flight per thread; grid == 0: the default wide sweep (8 CTAs of 256 threads per SM).  Few wide
 * CTAs with deep loads pull most of the HBM stream from a fraction of the SMs and leave the rest to a concurrent kernel. */
int nb200_fused_adam_cfg(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, uint64_t split,
                         const float *hyper, int zero_grad, uint32_t grid, uint32_t threads, uint32_t unroll, void *stream);

/* Dynamic loss scaling with skipped steps, on the device -- torch.cuda.amp.GradScaler as the reference trains with it
 * (nerf/utils_init_nerf.py:100,612-629: scale(loss).backward(); scaler.step(optimizer) skips the step when a gradient is
 * inf / NaN; scaler.update() halves the scale then, and doubles it after 2000 clean steps).  scaler = 8 device words:
 *   [0] scale f32   [1] growth tracker i32   [2] iteration u32   [3] skipped steps u32   [4], [5] status words of even / odd
 *   iterations (bit 31 = found-inf)   [6] growth interval i32   [7] unused.     Initialise to {scale, 0, 0, 0, 0, 0, 2000, 0}.
 * A step: the loss gradient is multiplied by [0] (nb200_fs_composite_forward: loss_scale_dev = scaler); the backward kernels
 * raise bit 31 of word [4 + (iteration & 1)] when a feature gradient or an MLP weight gradient is not finite
 * (nb200_field_backward: every gradient of the step flows through that kernel); nb200_adam_hyper_scaled evaluates the hyper-parameters for step *step + 1 WITHOUT
 * advancing it, with grad_scale = 1 / scale and (local_skip != 0) skip = the local flag; the Adam sweep / the peer-memory
 * update leave p, m, v untouched when skipping (the gradient is still reset); nb200_scaler_commit then advances *step only
 * if the step was taken, updates scale / tracker, clears the flag of the next iteration and increments the iteration.
 * peer_scalers (world > 1, peer-memory update): the scaler words of every rank -- the decision is the OR of all flags. */
int nb200_adam_hyper_scaled(const int32_t *step, const float *sched, float *hyper, uint32_t *scaler, int local_skip,
                            const int32_t *samples, void *stream);
int nb200_scaler_commit(int32_t *step, uint32_t *scaler, uint32_t *const *peer_scalers, uint32_t world, int32_t *max_samples,
                        void *stream);
/* The status words [4], [5] carry two things: bit 31 = found-inf, bits 0..30 = the step's sample count (samples, device i32, may
 * be NULL; nb200_adam_hyper_scaled ORs it in).  nb200_scaler_commit writes the maximum count over all ranks to *max_samples (may
 * be NULL) -- the same number on every rank, so that ray-sharded ranks grow their sample buffers at the same step.  In
 * nb200_train_update[_peer] samples = plan->counter + 6 (where the forward/backward parks the step's count) and
 * max_samples = plan->counter + 5: plan->counter must then be an 8-word block {count, rays, -, -, -, max over ranks, count copy, -}. */

#define NB200_PLAN_FUSED_FORWARD 1u   /* encode + field forward as one kernel (nb200_field_fused_forward): the stage timer
                                         then reports the pair under "field_forward" and ~0 under "grid_encode_forward" */
#define NB200_PLAN_SPLIT_COMPOSITE 4u /* compositing forward and backward as two launches (default: nb200_fs_composite_fused) */
#define NB200_PLAN_HYPER_DONE 2u      /* nb200_train_update[_peer] skips its first kernel (nb200_adam_hyper[_scaled]): the caller
                                         has launched it already -- a pipelined trainer does so BEFORE forking the update onto its
                                         side stream, so that the sweep and the next step's march become runnable together */
typedef struct nb200_train_plan {
    /* sizes and scalars */
    uint32_t N, M_cap, C, H, L, base_res, gridtype, max_steps;
    float bound, dt_gamma, S, T_thresh, min_near, loss_scale, inv_n_total;
    uint32_t flags;                                              /* NB200_PLAN_* */
    uint64_t n_params, n_table_params;
    /* inputs (device) */
    const float *rays_o, *rays_d, *target, *aabb, *noises;      /* noises may be NULL (perturb off) */
    const float *target_mask;                                    /* [N] or NULL: ground-truth mask of the train_conf loss term */
    float *render_mask, *g_render_mask;                          /* [N] each (used when target_mask != NULL) */
    float mask_weight, pad1;                                     /* train_conf */
    const uint8_t *bitfield;
    /* parameters: one flat fp32 vector [table | trunk | density | rgb]; the four pointers below alias into it */
    float *params_flat, *grads_flat, *exp_avg, *exp_avg_sq;
    float *hyper;                                                /* f32 [16], rewritten by every nb200_train_update */
    const float *sched;                                          /* f32 [8], see nb200_adam_hyper */
    int32_t *step;                                               /* optimiser step count (device) */
    const float *table, *trunk, *density, *rgb;
    const int32_t *offsets;
    float *g_table, *g_trunk, *g_density, *g_rgb;
    void *w_fwd, *w_bwd;                                         /* packed fp16 operand images of the MLP weights */
    /* per-ray work buffers [N ...] */
    float *nears, *fars, *weights_sum, *depth, *image, *g_weights_sum, *g_image, *loss;
    int32_t *rays, *counter, *m_eff, *scratch;
    /* per-sample work buffers [M_cap ...] */
    float *xyzs, *dirs, *deltas, *sigma, *sigma_arg, *d_sigma, *d_rgba;
    void *x_en, *rgba, *act, *d_x_en;
    float *wg_scratch;                                           /* nb200_field_wgrad_scratch_bytes() or NULL */
    void *timer;                                                 /* nb200_stage_timer or NULL */
    uint32_t *scaler;                                            /* device-side dynamic loss scaler (8 words, see
                                                                    nb200_scaler_commit) or NULL: constant loss_scale */
    uint32_t adam_grid, adam_threads, adam_unroll, split_level;         /* shape of the Adam sweep (nb200_fused_adam_cfg); all 0: the
                                                                    wide default.  A pipelined trainer runs a NARROW sweep
                                                                    (e.g. 64 CTAs x 512 threads x 4 groups in flight) so that the
                                                                    next step's ray march overlaps it */
    uint64_t split_elem;                                         /* split_level > 0 (ray-sharded step): the table gradient is
                                                                    scattered in two launches -- levels [split_level, L) first,
                                                                    then [0, split_level) -- and the peer-memory update in two
                                                                    parts: elements [split_elem, n) (= 2 * offsets[split_level]:
                                                                    the fine levels and the MLPs) while the second scatter
                                                                    launch still runs, then [0, split_elem) */
} nb200_train_plan;

/* ---- LGIE editing step (BASELINE.json configs[3]; reference: the fg / bg / all renders of NeRFRenderer.run,
 * nerf/renderer.py:383-474, that Trainer_Nerf.train_step_editing consumes, utils_init_nerf.py:243-265).  The same samples
 * are composited three times -- variant 0 "all": sigma; 1 "fg": sigma * e(m); 2 "bg": sigma * (1 - e(m)) -- with the
 * mask head's output m (4th rgba channel) as the edit mask: e = sigmoid((m - conf_thr) * 100) (soft_mask) or [m > 0.5];
 * every variant also renders its mask sum w * m.  detach_bg: samples with m < 0.5 give values but no gradient to the
 * "all" render (:409-418); detach_mask_from_field: the rendered masks' weights carry no gradient (:460-463).
 * Per-ray outputs / gradient inputs are laid out [3 variants][N ...].  The loss lives with the caller (the reference's is
 * the Stable-Diffusion guidance, out of scope): run nb200_train_lgie_forward, write the gradients of the loss with respect
 * to the rendered outputs into g_*, run nb200_train_lgie_backward, then nb200_train_update[_peer]. */
typedef struct nb200_lgie_plan {
    float conf_thr;
    int32_t soft_mask, detach_bg, detach_mask_from_field;
    float *weights_sum, *depth, *image, *render_mask;            /* [3,N], [3,N], [3,N,3], [3,N] */
    const float *g_weights_sum, *g_image, *g_render_mask;        /* [3,N], [3,N,3], [3,N] (depth has no gradient,
                                                                    raymarching.py:274-289) */
} nb200_lgie_plan;
int nb200_fs_composite_lgie_forward(int variant, const float *sigmas, const void *rgba, const float *deltas,
                                    const int32_t *rays, uint32_t M, uint32_t N, float T_thresh, float conf_thr,
                                    int soft_mask, float *weights_sum, float *depth, float *image, float *render_mask,
                                    void *stream);
/* variant 0 writes grad_sigmas / grad_rgba (float4 rows [g_r, g_g, g_b, g_m]); variants 1, 2 accumulate into them */
int nb200_fs_composite_lgie_backward(int variant, const float *grad_weights_sum, const float *grad_image,
                                     const float *grad_render_mask, const float *sigmas, const void *rgba,
                                     const float *deltas, const int32_t *rays, const float *weights_sum, const float *image,
                                     const float *render_mask, uint32_t M, uint32_t N, float T_thresh, float conf_thr,
                                     int soft_mask, int detach_bg, int detach_mask_from_field, float *grad_sigmas,
                                     float *grad_rgba, void *stream);
/* near/far -> march -> encode -> field -> the three composites */
int nb200_train_lgie_forward(const nb200_train_plan *plan, const nb200_lgie_plan *lgie, void *stream);
/* the three composites^T (summed per sample) -> field^T -> encode^T; gradients accumulate into grads_flat */
int nb200_train_lgie_backward(const nb200_train_plan *plan, const nb200_lgie_plan *lgie, void *stream);
uint32_t nb200_lgie_plan_bytes(void);

/* Per-stage device timing of a (non-captured) step: when plan->timer is set, nb200_train_forward_backward records an
 * event before its first kernel and after each of its NB200_FB_STAGES stages, nb200_train_update after each of its
 * NB200_UP_STAGES stages.  nb200_stage_timer_read synchronises on the last event and returns the stage durations in
 * microseconds: out_us[NB200_FB_STAGES + NB200_UP_STAGES] (host memory). */
#define NB200_FB_STAGES 8    /* march_count (+ near/far), march_write, encode, field, composite (+ MSE), composite^T, field^T, encode^T */
#define NB200_UP_STAGES 2    /* adam (hyper + sweep), weight pack */
int nb200_stage_timer_create(void **timer);
int nb200_stage_timer_destroy(void *timer);
int nb200_stage_timer_read(void *timer, float *out_us);

uint32_t nb200_train_plan_bytes(void);
/* near/far -> march -> encode -> field -> composite -> MSE -> composite^T -> field^T -> encode^T (gradients
 * accumulated into grads_flat).  Resets counter and loss first. */
int nb200_train_forward_backward(const nb200_train_plan *plan, void *stream);
/* The same in two phases.  NB200_PHASE_MARCH (near/far + march) reads only rays, noises and the occupancy bit field --
 * nothing the optimiser writes -- so a trainer may run it on a second stream concurrently with the PREVIOUS step's
 * nb200_train_update (a memory-bound sweep next to an issue-bound traversal); NB200_PHASE_REST is encode .. encode^T. */
#define NB200_PHASE_MARCH 1
#define NB200_PHASE_REST  2
#define NB200_PHASE_REST_A 4   /* split step: encode .. field^T and the scatter of levels [split_level, L) */
#define NB200_PHASE_REST_B 8   /* split step: the scatter of levels [0, split_level) */
int nb200_train_phase(const nb200_train_plan *plan, int phases, void *stream);
/* fused Adam over params_flat (zeroing grads_flat) and re-pack of the MLP operand images. */
int nb200_train_update(const nb200_train_plan *plan, void *stream);
/* the first kernel of nb200_train_update[_peer] on its own (see NB200_PLAN_HYPER_DONE); peer != 0: for nb200_train_update_peer */
int nb200_train_update_hyper(const nb200_train_plan *plan, int peer, void *stream);

/* ============================================================================================
 * Ray generation.  Replaces the direction / origin arithmetic of get_rays
 * (nerf/provider_utils.py:238-302; the index sampling of :263-284 stays with the caller):
 *   pixel p = inds[b, n] (or n when inds == NULL, which requires N == H * W) -> i = p % W + off_x, j = p / W + off_y
 *   dir = safe_normalize((i - cx) / fx, (j - cy) / fy, 1)  (:125-126, :289-293);  rays_d = R dir, rays_o = t  (:294-297)
 * poses f32 [B, 4, 4] row-major camera-to-world (device); inds i64 [B, N] (device) or NULL; outputs f32 [B, N, 3].
 * ========================================================================================== */
int nb200_get_rays(const float *poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, uint32_t B,
                   uint32_t N, const int64_t *inds, float off_x, float off_y, float *rays_o, float *rays_d, void *stream);

/* ============================================================================================
 * Multi-GPU optimiser update over NVLink peer memory ("one all-reduce(sum) per step over a flat
 * buffer ... then identical optimizer step on every rank"; reference anchors: the DDP wrap of nerf/utils_init_nerf.py:
 * 76-78 and torch.optim.Adam of main.py:182).  One kernel replaces all-reduce + Adam: rank r owns slice r of the flat
 * parameter vector, reads slice r of every rank's gradient out of the peers' memory, sums in rank order, runs Adam on
 * the slice (moments are only kept where they are owned) and stores the new parameters into every replica; the local
 * gradient is reset once the peers have read it.  csrc/peer_update.cu describes the flag barriers.
 *
 * Peer-visible memory: cudaIpc handles exist only for whole cudaMalloc allocations, so -- the one exception to "nothing
 * here allocates" -- nb200_peer_alloc returns such an allocation (zero-filled) on the current device; export its handle
 * (nb200_peer_handle_bytes() bytes of HOST memory), ship it to the other processes by any means, import there.
 * ========================================================================================== */
#define NB200_PEER_MAX 8
typedef struct nb200_peer_plan {
    uint32_t world, rank;
    uint32_t grid;                        /* CTAs: nb200_peer_grid(n, world, sms) or fewer -- identical on every rank */
    uint32_t threads;                     /* threads per CTA: 0 (= 256), 64, 128, 256 or 512 */
    uint64_t n, split;                    /* parameters; [0, split) hyper group 0, [split, n) group 1; both % 4 == 0 */
    float *params[NB200_PEER_MAX];        /* every rank's parameter vector  ([rank] local, the others imported) */
    float *grads[NB200_PEER_MAX];         /* every rank's gradient vector */
    uint32_t *signals[NB200_PEER_MAX];    /* every rank's flag words, nb200_peer_signal_bytes(grid) each, zero-filled */
    float *exp_avg, *exp_avg_sq;          /* local, indexed like params; only this rank's slice is touched */
    const float *hyper;                   /* as nb200_fused_adam */
    uint32_t *epoch;                      /* local u32 [grid + 1], zero-filled before the first call */
    uint32_t *status;                     /* local u32: set non-zero when a barrier timed out (a peer is gone) */
    float *mc_params, *mc_grads;          /* NVSwitch multicast mappings of the same two vectors (both or neither; NULL:
                                             plain peer loads / stores): multimem.ld_reduce sums the gradient inside the
                                             switch, multimem.st delivers the parameters to every replica */
    uint32_t unroll, pad;                 /* float4 groups in flight per thread and rank: 0 (default by world size), 1, 2 or 4 */
    uint32_t *scalers[NB200_PEER_MAX];    /* every rank's loss-scaler words (nb200_scaler_commit), or all NULL: when given, the
                                             update is skipped on EVERY rank if any rank's found-inf flag of this iteration is up */
} nb200_peer_plan;
uint32_t nb200_peer_plan_bytes(void);
uint32_t nb200_peer_handle_bytes(void);
uint64_t nb200_peer_signal_bytes(uint32_t grid);
uint32_t nb200_peer_grid(uint64_t n, uint32_t world, uint32_t sms);
/* elements [lo, hi) of the flat vector that `rank` owns (host arithmetic only) */
void nb200_peer_slice(uint64_t n, uint32_t world, uint32_t rank, uint64_t *lo, uint64_t *hi);
int nb200_peer_alloc(void **ptr, uint64_t bytes);
int nb200_peer_free(void *ptr);
int nb200_peer_export(void *ptr, void *handle);
int nb200_peer_import(const void *handle, void **ptr);
int nb200_peer_release(void *ptr);
/* Every rank must call this the same number of times (a CUDA-graph replay counts); world == 1 degenerates to
 * nb200_fused_adam with zero_grad. */
int nb200_peer_reduce_adam_bcast(const nb200_peer_plan *plan, void *stream);
/* One-CTA barrier over the same flag words: returns when every rank's stream has reached its own call (every rank must
 * call it the same number of times).  bench.py aligns the ranks with it between timed steps. */
int nb200_peer_rank_barrier(const nb200_peer_plan *plan, void *stream);
/* nb200_train_update with the Adam sweep replaced by nb200_peer_reduce_adam_bcast. */
int nb200_train_update_peer(const nb200_train_plan *plan, const nb200_peer_plan *peer, void *stream);
/* The same in two parts (plan->split_level > 0): part 1 = hyper kernel (unless NB200_PLAN_HYPER_DONE) + reduce / Adam /
 * broadcast of elements [split_elem, n); part 2 = elements [0, split_elem) + weight re-pack (+ scaler commit).  Within
 * each part rank r owns the r-th 1/world of that part's range.  Every rank must call both parts, in this order. */
int nb200_train_update_peer_part(const nb200_train_plan *plan, const nb200_peer_plan *peer, int part, void *stream);

/* L2 bandwidth probes (measurement aid, no reference counterpart): MEASURED_PEAKS.json carries no L2 figure, so bench.py
 * measures (a) a coalesced float4 stream over an L2-resident buffer and (b) random 8-byte gathers (one 32-byte sector
 * each, the access pattern of a hashed grid level) and reports the encoder's rates next to them. */
int nb200_l2_stream_probe(const void *buf, uint64_t bytes, uint32_t reps, float *sink, void *stream);
int nb200_l2_gather_probe(const void *buf, uint32_t words, uint32_t n_threads, uint32_t per_thread, float *sink, void *stream);
/* SM-issued float4 copy on `ctas` CTAs; with one side a peer mapping (nb200_peer_import) it measures the NVLink rate
 * available to loads (pull) / stores (push) from kernels: the roofline of nb200_peer_reduce_adam_bcast. */
int nb200_stream_copy_probe(void *dst, const void *src, uint64_t bytes, uint32_t ctas, void *stream);

/* ============================================================================================
 * tensor-core path self test (no reference counterpart): one-CTA tcgen05 GEMM D[128,N] = A[128,K] * B[N,K]^T with
 * fp16 operands / fp32 accumulation, used by the tests to pin the UMMA descriptor conventions the fused MLP
 * kernels rely on.  a_mn / b_mn select MN-major operand storage (A given as [K][128], B as [K][N]).
 * status (device u32) is set to 1 if the MMA never signalled completion.
 * ========================================================================================== */
int nb200_umma_selftest(const void *A, const void *B, float *D, uint32_t N, uint32_t K, uint32_t a_mn, uint32_t b_mn,
                        uint32_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
