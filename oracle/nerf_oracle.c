/*
 * oracle/nerf_oracle.c -- CPU restatement of the reference's native hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The
 * product package (customnerf_b200/) never imports, links or executes it.
 *
 * Each function restates one reference kernel, sequentially, one unit (point /
 * ray / byte) at a time, following the reference's operation order including
 * the FMA contraction nvcc applies to it, so integer
 * outputs (morton codes, bit-field, per-ray sample counts) are bit-exact.
 *
 * Parity pin: the reference ships no golden vectors, so
 * this file is pinned against the reference's OWN CUDA extensions compiled
 * unmodified into oracle/_ref/ (oracle/build_ref.py) and run on the GPU box:
 * tests/test_ref_ext_parity.py compares them live, and tests/golden/ holds
 * vectors minted from them (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/build.py).
 * -ffp-contract=off matters: every fused multiply-add below is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_MAX_D 5
#define ORC_MAX_C 8

/* ------------------------------------------------------------------ */
/* helpers: raymarching/src/raymarching.cu:19-81                       */
/* ------------------------------------------------------------------ */
static inline float orc_signf(float x) { return copysignf(1.0f, x); }             /* :30-32 */
static inline float orc_clamp(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); } /* :34-36 */

static inline int orc_mip_from_pos(float x, float y, float z, float max_cascade) { /* :42-47 */
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

static inline int orc_mip_from_dt(float dt, float H, float max_cascade) {         /* :49-54 */
    const float mx = (float)((double)(dt * H) * 0.5);
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

static inline uint32_t orc_expand_bits(uint32_t v) {                               /* :56-63 */
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

static inline uint32_t orc_m3d(uint32_t x, uint32_t y, uint32_t z) {          /* :65-71 */
    return orc_expand_bits(x) | (orc_expand_bits(y) << 1) | (orc_expand_bits(z) << 2);
}

static inline uint32_t orc_m3d_inv(uint32_t x) {                           /* :73-81 */
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* ------------------------------------------------------------------ */
/* near_far_from_aabb: raymarching.cu:91-145                           */
/* ------------------------------------------------------------------ */
void orc_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb,
                            uint32_t N, float min_near, float *nears, float *fars) {
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float rdx = 1.0f / dx, rdy = 1.0f / dy, rdz = 1.0f / dz;
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, tmp;
        if (near > far) { tmp = near; near = far; far = tmp; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* sph_from_ray: raymarching.cu:162-199 (float tolerance only; unused by the product) */
void orc_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float A = dx * dx + dy * dy + dz * dz;
        const float B = ox * dx + oy * dy + oz * dz;
        const float C = ox * ox + oy * oy + oz * oz - radius * radius;
        const float t = (-B + sqrtf(B * B - A * C)) / A;
        const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
        const float theta = atan2f(sqrtf(x * x + z * z), y);
        const float phi = atan2f(z, x);
        coords[n * 2] = 2 * theta * RPI - 1;
        coords[n * 2 + 1] = phi * RPI;
    }
}

/* morton3D / morton3D_invert: raymarching.cu:214-254 */
void orc_morton3D(const int32_t *coords, uint32_t N, int32_t *indices) {
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int32_t)orc_m3d((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}

void orc_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords) {
    for (uint32_t n = 0; n < N; n++) {
        const int32_t ind = indices[n];   /* arithmetic >> on int, as in the reference (:249-253) */
        coords[n * 3] = (int32_t)orc_m3d_inv((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int32_t)orc_m3d_inv((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int32_t)orc_m3d_inv((uint32_t)(ind >> 2));
    }
}

/* packbits: raymarching.cu:267-289.  N = number of output bytes. */
void orc_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield) {
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (uint8_t i = 0; i < 8; i++)
            bits |= (grid[(size_t)n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* ------------------------------------------------------------------ */
/* ray marching core: raymarching.cu:335-479 (train), :904-988 (infer) */
/* FMA placement per SURVEY.md Appendix A.3 (nvcc 12.9 -O3, sm_100a).  */
/* ------------------------------------------------------------------ */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float rH, H3, Hf, Cf, bound, dt_gamma, dt_min, dt_max;
    uint32_t H;
    const uint8_t *grid;
} orc_ray_t;

static void orc_ray_setup(orc_ray_t *r, const float *o, const float *d, const uint8_t *grid,
                          float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    r->ox = o[0]; r->oy = o[1]; r->oz = o[2];
    r->dx = d[0]; r->dy = d[1]; r->dz = d[2];
    r->rdx = 1.0f / r->dx; r->rdy = 1.0f / r->dy; r->rdz = 1.0f / r->dz;
    r->rH = 1.0f / (float)H;
    r->H3 = (float)(H * H * H);
    r->Hf = (float)H; r->Cf = (float)C; r->H = H;
    r->bound = bound; r->dt_gamma = dt_gamma;
    r->dt_min = 2 * 1.7320508075688772f / (float)max_steps;                  /* :345 */
    r->dt_max = 2 * 1.7320508075688772f * (float)(1 << (C - 1)) / (float)H;  /* :346 */
    r->grid = grid;
}

/* one loop iteration of :359-399.  Returns 1 if the cell is occupied (caller emits a sample at
 * (x,y,z) with step dt and advances t by dt); otherwise advances *t past the empty voxel. */
static inline int orc_ray_step(const orc_ray_t *r, float *t, float *x, float *y, float *z, float *dt_out) {
    const float tt0 = *t;
    *x = orc_clamp(fmaf(tt0, r->dx, r->ox), -r->bound, r->bound);
    *y = orc_clamp(fmaf(tt0, r->dy, r->oy), -r->bound, r->bound);
    *z = orc_clamp(fmaf(tt0, r->dz, r->oz), -r->bound, r->bound);
    const float dt = orc_clamp(tt0 * r->dt_gamma, r->dt_min, r->dt_max);
    *dt_out = dt;
    int level = orc_mip_from_pos(*x, *y, *z, r->Cf);
    const int l2 = orc_mip_from_dt(dt, r->Hf, r->Cf);
    if (l2 > level) level = l2;
    const float mip_bound = fminf(scalbnf(1.0f, level), r->bound);
    const float mip_rbound = 1.0f / mip_bound;
    /* 0.5 * (x * mip_rbound + 1) * H with a double literal: fp64 product, then clamp in fp32 (:374-376) */
    const int nx = (int)orc_clamp((float)(0.5 * (double)fmaf(*x, mip_rbound, 1.0f) * (double)r->H), 0.0f, (float)(r->H - 1));
    const int ny = (int)orc_clamp((float)(0.5 * (double)fmaf(*y, mip_rbound, 1.0f) * (double)r->H), 0.0f, (float)(r->H - 1));
    const int nz = (int)orc_clamp((float)(0.5 * (double)fmaf(*z, mip_rbound, 1.0f) * (double)r->H), 0.0f, (float)(r->H - 1));
    /* index evaluated in fp32: level * H3 + morton (:378) */
    const uint32_t index = (uint32_t)fmaf((float)level, r->H3, (float)orc_m3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const int occ = r->grid[index / 8] & (1 << (index % 8));
    if (occ) return 1;
    const float tx = (fmaf(fmaf(fmaf(0.5f, orc_signf(r->dx), (float)nx + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*x)) * r->rdx;
    const float ty = (fmaf(fmaf(fmaf(0.5f, orc_signf(r->dy), (float)ny + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*y)) * r->rdy;
    const float tz = (fmaf(fmaf(fmaf(0.5f, orc_signf(r->dz), (float)nz + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*z)) * r->rdz;
    const float tt = tt0 + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    float tc = tt0;
    do {
        tc += orc_clamp(tc * r->dt_gamma, r->dt_min, r->dt_max);
    } while (tc < tt);
    *t = tc;
    return 0;
}

/*
 * march_rays_train: raymarching.cu:311-480.  Rays are processed in ray-id order, so slot
 * reservation (the reference's atomicAdd order, :405-406) is the exclusive scan of the counts --
 * one valid member of the reference's non-deterministic output set.  counter is updated exactly
 * like the reference: counter[0] += sum(num_steps), counter[1] += N, starting from its entry value.
 */
void orc_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid,
                          float bound, float dt_gamma, uint32_t max_steps,
                          uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                          const float *nears, const float *fars,
                          float *xyzs, float *dirs, float *deltas,
                          int32_t *rays, int32_t *counter, const float *noises) {
    for (uint32_t n = 0; n < N; n++) {
        orc_ray_t r;
        orc_ray_setup(&r, rays_o + n * 3, rays_d + n * 3, grid, bound, dt_gamma, max_steps, C, H);
        const float far = fars[n];
        float t0 = nears[n];
        t0 = fmaf(orc_clamp(t0 * dt_gamma, r.dt_min, r.dt_max), noises[n], t0);   /* :351 (contracted to one FFMA) */
        float t = t0, x, y, z, dt;
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) {
            if (orc_ray_step(&r, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; }
        }
        const uint32_t point_index = (uint32_t)counter[0];
        const uint32_t ray_index = (uint32_t)counter[1];
        counter[0] += (int32_t)num_steps;
        counter[1] += 1;
        rays[ray_index * 3] = (int32_t)n;
        rays[ray_index * 3 + 1] = (int32_t)point_index;
        rays[ray_index * 3 + 2] = (int32_t)num_steps;
        if (num_steps == 0) continue;
        if (point_index + num_steps > M) continue;
        float *px = xyzs + (size_t)point_index * 3, *pd = dirs + (size_t)point_index * 3, *pl = deltas + (size_t)point_index * 2;
        t = t0;
        uint32_t step = 0;
        float last_t = t;
        while (t < far && step < num_steps) {
            if (orc_ray_step(&r, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* per-ray sample counts only (used by the tests to check the count kernel alone) */
void orc_march_rays_count(const float *rays_o, const float *rays_d, const uint8_t *grid,
                          float bound, float dt_gamma, uint32_t max_steps,
                          uint32_t N, uint32_t C, uint32_t H,
                          const float *nears, const float *fars, const float *noises, int32_t *counts) {
    for (uint32_t n = 0; n < N; n++) {
        orc_ray_t r;
        orc_ray_setup(&r, rays_o + n * 3, rays_d + n * 3, grid, bound, dt_gamma, max_steps, C, H);
        const float far = fars[n];
        float t = nears[n];
        t = fmaf(orc_clamp(t * dt_gamma, r.dt_min, r.dt_max), noises[n], t);
        float x, y, z, dt;
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) {
            if (orc_ray_step(&r, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; }
        }
        counts[n] = (int32_t)num_steps;
    }
}

/* march_rays (inference): raymarching.cu:884-989.  Output buffers must be zero-filled by the caller. */
void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                    const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                    uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                    float *xyzs, float *dirs, float *deltas, const float *noises) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        orc_ray_t r;
        orc_ray_setup(&r, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, dt_gamma, max_steps, C, H);
        float *px = xyzs + (size_t)n * n_step * 3, *pd = dirs + (size_t)n * n_step * 3, *pl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        const float far = fars[index];
        (void)nears;
        uint32_t step = 0;
        t = fmaf(orc_clamp(t * dt_gamma, r.dt_min, r.dt_max), noises[n], t);   /* :930 */
        float last_t = t, x, y, z, dt;
        while (t < far && step < n_step) {
            if (orc_ray_step(&r, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* compositing: raymarching.cu:500-577 (fwd), :691-772 (bwd), :1002-1089 (infer) */
/* __expf(x) == ex2.approx(x * log2(e)); restated with exp2f.          */
/* ------------------------------------------------------------------ */
static inline float orc_fast_expf(float x) { return exp2f(x * 1.4426950408889634f); }

void orc_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                      const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
                                      float *weights_sum, float *depth, float *image) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) {
            weights_sum[index] = 0; depth[index] = 0;
            image[index * 3] = image[index * 3 + 1] = image[index * 3 + 2] = 0;
            continue;
        }
        const float *s = sigmas + offset, *c = rgbs + (size_t)offset * 3, *dl = deltas + (size_t)offset * 2;
        uint32_t step = 0;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        while (step < num_steps) {
            const float alpha = 1.0f - orc_fast_expf(-s[0] * dl[0]);
            const float weight = alpha * T;
            r = fmaf(weight, c[0], r); g = fmaf(weight, c[1], g); b = fmaf(weight, c[2], b);
            t += dl[1];
            d = fmaf(weight, t, d);
            ws += weight;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; step++;
        }
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* grad_sigmas / grad_rgbs must be zero-filled by the caller (raymarching.py:284-285). */
void orc_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image,
                                       const float *sigmas, const float *rgbs, const float *deltas,
                                       const int32_t *rays, const float *weights_sum, const float *image,
                                       uint32_t M, uint32_t N, float T_thresh,
                                       float *grad_sigmas, float *grad_rgbs) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float gws = grad_weights_sum[index];
        const float *gi = grad_image + (size_t)index * 3;
        const float r_final = image[index * 3], g_final = image[index * 3 + 1], b_final = image[index * 3 + 2];
        const float ws_final = weights_sum[index];
        const float *s = sigmas + offset, *c = rgbs + (size_t)offset * 3, *dl = deltas + (size_t)offset * 2;
        float *gs = grad_sigmas + offset, *gc = grad_rgbs + (size_t)offset * 3;
        uint32_t step = 0;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        while (step < num_steps) {
            const float alpha = 1.0f - orc_fast_expf(-s[0] * dl[0]);
            const float weight = alpha * T;
            r = fmaf(weight, c[0], r); g = fmaf(weight, c[1], g); b = fmaf(weight, c[2], b);
            ws += weight;
            T *= 1.0f - alpha;
            gc[0] = gi[0] * weight; gc[1] = gi[1] * weight; gc[2] = gi[2] * weight;
            gs[0] = dl[0] * (gi[0] * (T * c[0] - (r_final - r)) +
                             gi[1] * (T * c[1] - (g_final - g)) +
                             gi[2] * (T * c[2] - (b_final - b)) +
                             gws * (1 - ws_final));
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; gs++; gc += 3; step++;
        }
    }
}

void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t,
                        const float *sigmas, const float *rgbs, const float *deltas,
                        float *weights_sum, float *depth, float *image) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float *s = sigmas + (size_t)n * n_step, *c = rgbs + (size_t)n * n_step * 3, *dl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        float weight_sum = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = 1.0f - orc_fast_expf(-s[0] * dl[0]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t += dl[1];
            d = fmaf(weight, t, d);
            r = fmaf(weight, c[0], r); g = fmaf(weight, c[1], g); b = fmaf(weight, c[2], b);
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; step++;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = weight_sum; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* ------------------------------------------------------------------ */
/* grid encoder: gridencoder/src/gridencoder.cu                        */
/* ------------------------------------------------------------------ */
static inline uint32_t orc_fast_hash(uint32_t D, const uint32_t *pos_grid) {       /* :50-63 */
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t result = 0;
    for (uint32_t i = 0; i < D; ++i) result ^= pos_grid[i] * primes[i];
    return result;
}

static inline uint32_t orc_grid_index(uint32_t D, uint32_t C, uint32_t gridtype, int align_corners, uint32_t ch,
                                      uint32_t hashmap_size, uint32_t resolution, const uint32_t *pos_grid) { /* :66-84 */
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = orc_fast_hash(D, pos_grid);
    return (index % hashmap_size) * C + ch;
}

static inline float orc_smoothstep(float v) { return v * v * (3.0f - 2.0f * v); }  /* :39-42 */
static inline float orc_smoothstep_d(float v) { return 6 * v * (1.0f - v); }       /* :44-47 */

static inline float orc_round_half(float v) { return (float)(_Float16)v; }

/* per (point, level) locate step shared by forward / backward / tv (:137-158).  Returns 0 if oob. */
static inline int orc_locate(const float *x, uint32_t D, uint32_t level, float S, uint32_t H, const float *scales,
                             int align_corners, uint32_t interp, float *scale_out, uint32_t *resolution, float *pos, float *pos_deriv,
                             uint32_t *pos_grid) {
    for (uint32_t d = 0; d < D; d++)
        if (x[d] < 0 || x[d] > 1) return 0;
    /* scale = exp2f(level * S) * H - 1.0f (:138).  exp2f is the one libm-dependent value on this path (device
     * ex2.approx vs glibc differ by an ulp at some levels, which moves fine-level outputs by ~1e-4), so callers
     * may pass the per-level scales the device computed; everything downstream is then identical arithmetic. */
    const float scale = scales ? scales[level] : fmaf(exp2f((float)level * S), (float)H, -1.0f);
    *scale_out = scale;
    *resolution = (uint32_t)ceilf(scale) + 1;
    for (uint32_t d = 0; d < D; d++) {
        pos[d] = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
        pos_grid[d] = (uint32_t)floorf(pos[d]);
        pos[d] -= (float)pos_grid[d];
        if (interp == 1) {
            pos_deriv[d] = orc_smoothstep_d(pos[d]);
            pos[d] = orc_smoothstep(pos[d]);
        } else {
            pos_deriv[d] = 1.0f;
        }
    }
    return 1;
}

/*
 * grid_encode_forward: gridencoder.cu:87-244.  embeddings / outputs / dy_dx are float arrays; with
 * half_mode != 0 the table values are expected to be fp16-representable and the accumulator follows
 * the reference's Half arithmetic (product rounded to half, then half += half).
 * outputs layout [L, B, C] (the reference's native layout; grid.py:49,63 permutes afterwards).
 */
void orc_grid_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets, float *outputs,
                             uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                             float *dy_dx, uint32_t gridtype, int align_corners, uint32_t interp, int half_mode,
                             const float *scales) {
    for (uint32_t level = 0; level < max_level; level++) {
        const float *grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        for (uint32_t b = 0; b < B; b++) {
            const float *x = inputs + (size_t)b * D;
            float *out = outputs + ((size_t)level * B + b) * C;
            float *dd = dy_dx ? dy_dx + (size_t)b * D * L * C + (size_t)level * D * C : NULL;
            float scale, pos[ORC_MAX_D], pos_deriv[ORC_MAX_D];
            uint32_t resolution, pos_grid[ORC_MAX_D];
            if (!orc_locate(x, D, level, S, H, scales, align_corners, interp, &scale, &resolution, pos, pos_deriv, pos_grid)) {
                for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
                if (dd) for (uint32_t i = 0; i < D * C; i++) dd[i] = 0;
                continue;
            }
            float results[ORC_MAX_C] = {0};
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pgl[ORC_MAX_D];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                }
                const uint32_t index = orc_grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pgl);
                for (uint32_t ch = 0; ch < C; ch++) {
                    if (half_mode) results[ch] = orc_round_half(results[ch] + orc_round_half(w * grid[index + ch]));
                    else results[ch] = fmaf(w, grid[index + ch], results[ch]);
                }
            }
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = results[ch];
            if (dd) {
                for (uint32_t gd = 0; gd < D; gd++) {
                    float rg[ORC_MAX_C] = {0};
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                        float w = scale;
                        uint32_t pgl[ORC_MAX_D];
                        for (uint32_t nd = 0; nd < D - 1; nd++) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                            else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                        }
                        pgl[gd] = pos_grid[gd];
                        const uint32_t il = orc_grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pgl);
                        pgl[gd] = pos_grid[gd] + 1;
                        const uint32_t ir = orc_grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pgl);
                        for (uint32_t ch = 0; ch < C; ch++) {
                            const float v = w * (grid[ir + ch] - grid[il + ch]) * pos_deriv[gd];
                            if (half_mode) rg[ch] = orc_round_half(rg[ch] + orc_round_half(v));
                            else rg[ch] += v;
                        }
                    }
                    for (uint32_t ch = 0; ch < C; ch++) dd[gd * C + ch] = rg[ch];
                }
            }
        }
    }
}

/*
 * grid_encode_backward: gridencoder.cu:247-368.  grad layout [L, B, C].  The reference scatters with
 * float (or half2) atomics in arbitrary order; the restatement accumulates in double and rounds once,
 * which is the value every atomic ordering approximates.  grad_embeddings must be zero-filled.
 * grad_inputs (optional, with dy_dx): gridencoder.cu:342-368.
 */
void orc_grid_encode_backward(const float *grad, const float *inputs, const int32_t *offsets, float *grad_embeddings,
                              uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                              const float *dy_dx, float *grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                              const float *scales) {
    const size_t total = (size_t)(uint32_t)offsets[L] * C;
    double *acc = (double *)calloc(total, sizeof(double));
    for (uint32_t level = 0; level < max_level; level++) {
        double *gg = acc + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        for (uint32_t b = 0; b < B; b++) {
            const float *x = inputs + (size_t)b * D;
            const float *g = grad + ((size_t)level * B + b) * C;
            float scale, pos[ORC_MAX_D], pos_deriv[ORC_MAX_D];
            uint32_t resolution, pos_grid[ORC_MAX_D];
            if (!orc_locate(x, D, level, S, H, scales, align_corners, interp, &scale, &resolution, pos, pos_deriv, pos_grid)) continue;
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pgl[ORC_MAX_D];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                }
                const uint32_t index = orc_grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pgl);
                for (uint32_t ch = 0; ch < C; ch++) gg[index + ch] += (double)w * (double)g[ch];
            }
        }
    }
    for (size_t i = 0; i < total; i++) grad_embeddings[i] = (float)acc[i];
    free(acc);
    if (dy_dx && grad_inputs) {
        for (uint32_t b = 0; b < B; b++)
            for (uint32_t d = 0; d < D; d++) {
                double result = 0;
                const float *dd = dy_dx + (size_t)b * L * D * C;
                for (uint32_t l = 0; l < L; l++)
                    for (uint32_t ch = 0; ch < C; ch++)
                        result += (double)grad[((size_t)l * B + b) * C + ch] * (double)dd[l * D * C + d * C + ch];
                grad_inputs[(size_t)b * D + d] = (float)result;
            }
    }
}

/* grad_total_variation: gridencoder.cu:505-609.  Adds into grad (double accumulate, rounded once). */
void orc_grad_total_variation(const float *inputs, const float *embeddings, float *grad, const int32_t *offsets,
                              float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                              uint32_t gridtype, int align_corners, const float *scales) {
    const size_t total = (size_t)(uint32_t)offsets[L] * C;
    double *acc = (double *)calloc(total, sizeof(double));
    for (uint32_t level = 0; level < L; level++) {
        const float *grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
        double *gg = acc + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        for (uint32_t b = 0; b < B; b++) {
            const float *x = inputs + (size_t)b * D;
            float scale, pos[ORC_MAX_D], pos_deriv[ORC_MAX_D];
            uint32_t resolution, pos_grid[ORC_MAX_D];
            if (!orc_locate(x, D, level, S, H, scales, align_corners, 0, &scale, &resolution, pos, pos_deriv, pos_grid)) continue;
            float results[ORC_MAX_C] = {0}, idelta[ORC_MAX_C] = {0};
            const uint32_t index = orc_grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pos_grid);
            const float w = weight / (2 * D);
            for (uint32_t d = 0; d < D; d++) {
                const uint32_t cur_d = pos_grid[d];
                if (cur_d < resolution) {
                    pos_grid[d] = cur_d + 1;
                    const uint32_t ir = orc_grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pos_grid);
                    for (uint32_t ch = 0; ch < C; ch++) {
                        const float gv = grid[index + ch] - grid[ir + ch];
                        results[ch] += gv; idelta[ch] += gv * gv;
                    }
                }
                if (cur_d > 0) {
                    pos_grid[d] = cur_d - 1;
                    const uint32_t il = orc_grid_index(D, C, gridtype, align_corners, 0, hashmap_size, resolution, pos_grid);
                    for (uint32_t ch = 0; ch < C; ch++) {
                        const float gv = grid[index + ch] - grid[il + ch];
                        results[ch] += gv; idelta[ch] += gv * gv;
                    }
                }
                pos_grid[d] = cur_d;
            }
            for (uint32_t ch = 0; ch < C; ch++)
                gg[index + ch] += (double)(w * results[ch] * (1.0f / sqrtf(idelta[ch] + 1e-9f)));
        }
    }
    for (size_t i = 0; i < total; i++) grad[i] = (float)((double)grad[i] + acc[i]);
    free(acc);
}
