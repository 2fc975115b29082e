// raymarching.cu -- occupancy-grid ray marching and alpha compositing for sm_100a.
//
// Replaces raymarching/src/raymarching.cu of the reference (see include/nerf_b200.h for the per-entry
// file:line map).  Design differences from the reference's one-thread-per-ray kernels:
//   * march_rays_train: count -> scan -> write.  Slot offsets are the exclusive scan of the per-ray
//     counts in ray-id order (warp shuffles + one block-sum pass) instead of two global atomics per ray,
//     so outputs are deterministic and the caller can size them exactly.
//   * composite_rays_train fwd/bwd: one WARP per ray, 32 samples per trip, transmittance / colour
//     prefixes through warp-shuffle scans; loads of a ray's segment are coalesced.
// The per-step arithmetic of the marcher is pinned with explicit round-to-nearest intrinsics in the
// order nvcc emits for the reference source, because per-ray sample COUNTS
// are integers that must match the reference bit-exactly.
#include "common.cuh"
#include <float.h>
#include <stdlib.h>

namespace {

constexpr float kSqrt3 = 1.7320508075688772f;

__device__ __forceinline__ float rm_clamp(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// frexpf exponent of a finite non-negative float, clamped to [0, C-1] (raymarching.cu:42-54).
__device__ __forceinline__ int rm_mip_level(float mx, int Cm1) {
    const int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 126;
    return min(Cm1, max(0, e));
}

struct RayCtx {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float sx, sy, sz;          // 0.5f * sign(d)
    float rH, H3, Hf, Hm1f, bound, rbound, dt_gamma, dt_min, dt_max;
    float halfH;               // 0.5f * H when H is a power of two (then the fp64 product of :374 is an exact fp32 one), else 0
    double Hd;
    int Cm1;
    int level_dt;              // mip_from_dt of the constant step when dt_gamma == 0, else -1
    float dt_min_c;            // that constant step
    const uint8_t *grid;
};

__device__ __forceinline__ void rm_setup(RayCtx &r, const float *__restrict__ o, const float *__restrict__ d,
                                         const uint8_t *grid, float bound, float dt_gamma, uint32_t max_steps,
                                         uint32_t C, uint32_t H) {
    r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
    r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
    r.rdx = __fdiv_rn(1.0f, r.dx); r.rdy = __fdiv_rn(1.0f, r.dy); r.rdz = __fdiv_rn(1.0f, r.dz);
    r.sx = copysignf(0.5f, r.dx); r.sy = copysignf(0.5f, r.dy); r.sz = copysignf(0.5f, r.dz);
    r.rH = __fdiv_rn(1.0f, (float)H);
    r.H3 = (float)(H * H * H);
    r.Hf = (float)H; r.Hm1f = (float)(H - 1); r.Hd = (double)H;
    r.bound = bound; r.dt_gamma = dt_gamma;
    r.rbound = __fdiv_rn(1.0f, bound);
    r.halfH = ((H & (H - 1)) == 0) ? 0.5f * (float)H : 0.0f;
    r.dt_min = __fdiv_rn(2 * kSqrt3, (float)max_steps);                                   // :345
    r.dt_max = __fdiv_rn(__fmul_rn(2 * kSqrt3, (float)(1 << (C - 1))), (float)H);          // :346
    r.Cm1 = (int)C - 1;
    r.level_dt = -1;
    r.dt_min_c = 0.0f;
    if (dt_gamma == 0.0f) {
        const float dt = rm_clamp(0.0f, r.dt_min, r.dt_max);                               // t * 0 == 0 for every finite t
        r.level_dt = rm_mip_level(__fmul_rn(__fmul_rn(dt, r.Hf), 0.5f), r.Cm1);
        r.dt_min_c = dt;
    }
    r.grid = grid;
}

__device__ __forceinline__ float rm_dt(const RayCtx &r, float t) {
    return rm_clamp(__fmul_rn(t, r.dt_gamma), r.dt_min, r.dt_max);
}

// (int) clamp(0.5 * (p * mip_rbound + 1) * H, 0, H-1): fp64 product because the literal is a double (:374-376).
// When H is a power of two, 0.5 * f * H only changes f's exponent, so the fp64 product rounded to fp32 equals the
// single fp32 product f * (0.5 * H) bit for bit (no overflow / underflow is possible for f in [0, 2]).
// kFast (here and below): the caller guarantees dt_gamma == 0 and H a power of two, so the constant-step and the
// single-fp32-product paths are selected at compile time instead of being predicated into every probe.
template <bool kFast = false>
__device__ __forceinline__ int rm_cell(const RayCtx &r, float p, float mip_rbound) {
    const float f = __fmaf_rn(p, mip_rbound, 1.0f);
    float v;
    if (kFast || r.halfH != 0.0f) v = __fmul_rn(f, r.halfH);
    else v = __double2float_rn(__dmul_rn(__dmul_rn(0.5, (double)f), r.Hd));
    return __float2int_rz(rm_clamp(v, 0.0f, r.Hm1f));
}

// do { t += dt; } while (t < tt);   (:396-398)
//
// With dt_gamma == 0 the step is the constant dt_min, and while t stays inside one binade [2^e, 2^(e+1)) every
// rounded addition t + dt advances t by the SAME multiple c of the binade's ulp u (t = T u, dt = q u + r with
// 0 <= r < u; the sum rounds to (T + q) u or (T + q + 1) u depending only on r, unless r == u / 2 where ties-to-even
// looks at T's parity).  So the loop's result is t + n c for the smallest n >= 1 with t + n c >= tt, evaluated exactly
// by one FMA (the value is a multiple of u below 2^24 u) -- bit-identical to the serial loop, without the loop.
// Ties, a binade crossing, or t outside the normal range fall back to the serial loop.
// constant step c of the binade of t (see above); false when the closed form does not apply (tie, range)
__device__ __forceinline__ bool rm_binade_step(float t, float dt, float &c, float &top) {
    const uint32_t eb = __float_as_uint(t) & 0x7f800000u;
    if (!(eb > (24u << 23) && eb < (253u << 23) && t >= dt)) return false;   // t >= dt: the differences are exact (Sterbenz)
    const float u = __uint_as_float(eb - (23u << 23));
    top = __uint_as_float(eb + (1u << 23));
    c = __fsub_rn(__fadd_rn(t, dt), t);
    const float tn = __fadd_rn(t, u);
    const float c2 = __fsub_rn(__fadd_rn(tn, dt), tn);          // the other parity of T: differs from c only on a tie
    return c == c2 && c > 0.0f;
}

// returns the advanced t; nsteps = number of additions the loop performed
template <bool kFast = false>
__device__ __forceinline__ float rm_advance(const RayCtx &r, float t, float tt, uint32_t &nsteps) {
    uint32_t n = 0;
    if (kFast || r.level_dt >= 0) {                                 // dt_gamma == 0: dt is a constant
        const float dt = r.dt_min_c;
        float c, top;
        if (rm_binade_step(t, dt, c, top)) {
            float nf = fmaxf(ceilf(__fdividef(__fsub_rn(tt, t), c)), 1.0f);
            if (nf < 65536.0f) {
                float t1 = __fmaf_rn(nf, c, t), t0 = __fmaf_rn(nf - 1.0f, c, t);
                if (t1 < tt) { nf += 1.0f; t0 = t1; t1 = __fmaf_rn(nf, c, t); }              // estimate one short
                else if (nf > 1.0f && t0 >= tt) { nf -= 1.0f; t1 = t0; t0 = __fmaf_rn(nf - 1.0f, c, t); }   // one long
                // accept only a verified answer: n minimal, and every point before the last inside the binade
                // (t0 = t + (n-1) c < top makes all of them exact multiples of u reached by constant steps)
                if (t1 >= tt && (nf == 1.0f || t0 < tt) && t0 < top) {
                    n = (uint32_t)nf;
                    if (t1 < top) { nsteps = n; return t1; }
                    t = __fadd_rn(t0, dt);          // the last step crosses the binade: rounded with the new ulp
                    if (t >= tt) { nsteps = n; return t; }          // (else keep stepping serially from here)
                }
            }
        }
    }
    do { t = __fadd_rn(t, kFast ? r.dt_min_c : rm_dt(r, t)); n++; } while (t < tt);
    nsteps = n;
    return t;
}

// t after n additions of the constant step dt (dt_gamma == 0): binade by binade with the constant-step closed form,
// the crossing addition done for real.  Bit-identical to n serial additions.
__device__ __forceinline__ float rm_lattice_jump(float t, float dt, uint32_t n) {
    while (n > 0) {
        float c, top;
        if (rm_binade_step(t, dt, c, top)) {
            // largest j with t + j c < top
            float jf = fmaxf(ceilf(__fdividef(__fsub_rn(top, t), c)) - 1.0f, 0.0f);
            if (jf < 16777216.0f) {
                if (__fmaf_rn(jf, c, t) >= top) jf -= 1.0f;                              // estimate one long
                else if (__fmaf_rn(jf + 1.0f, c, t) < top) jf += 1.0f;                   // one short
                if (jf >= 0.0f && __fmaf_rn(jf, c, t) < top && __fmaf_rn(jf + 1.0f, c, t) >= top) {
                    const uint32_t j = min(n, (uint32_t)jf);
                    t = __fmaf_rn((float)j, c, t);
                    n -= j;
                    if (n == 0) break;
                }
            }
        }
        t = __fadd_rn(t, dt);       // the crossing step (or a step the closed form does not cover)
        n -= 1;
    }
    return t;
}

__device__ __forceinline__ float rm_exit(float n, float s, float rH, float mip_bound, float p, float rd) {
    // (((n + 0.5f + 0.5f * sign(d)) * rH * 2 - 1) * mip_bound - p) * rd   (:390-392).  s = 0.5f*sign(d) is exact,
    // so the FFMA(0.5, sign, n + 0.5) nvcc emits for the reference equals this second add.
    const float b = __fmul_rn(__fadd_rn(__fadd_rn(n, 0.5f), s), rH);
    return __fmul_rn(__fmaf_rn(__fmaf_rn(b, 2.0f, -1.0f), mip_bound, -p), rd);
}

// Occupancy probe of the marching loop (:359-387) at parameter t: x,y,z,dt are the would-be sample; when the cell is
// empty, tt is the parameter at which the ray leaves the voxel (:388-394).
template <bool kFast = false>
__device__ __forceinline__ bool rm_probe(const RayCtx &r, float t, float &x, float &y, float &z, float &dt, float &tt) {
    x = rm_clamp(__fmaf_rn(t, r.dx, r.ox), -r.bound, r.bound);
    y = rm_clamp(__fmaf_rn(t, r.dy, r.oy), -r.bound, r.bound);
    z = rm_clamp(__fmaf_rn(t, r.dz, r.oz), -r.bound, r.bound);
    dt = kFast ? r.dt_min_c : rm_dt(r, t);
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    const int level_dt = (kFast || r.level_dt >= 0) ? r.level_dt : rm_mip_level(__fmul_rn(__fmul_rn(dt, r.Hf), 0.5f), r.Cm1);
    const int level = max(rm_mip_level(mx, r.Cm1), level_dt);
    const float pw = __int_as_float((127 + level) << 23);                                   // scalbnf(1, level)
    const bool capped = pw > r.bound;
    const float mip_bound = capped ? r.bound : pw;                                          // fminf(2^level, bound)
    const float mip_rbound = capped ? r.rbound : __int_as_float((127 - level) << 23);       // 1 / mip_bound, exact
    const int nx = rm_cell<kFast>(r, x, mip_rbound);
    const int ny = rm_cell<kFast>(r, y, mip_rbound);
    const int nz = rm_cell<kFast>(r, z, mip_rbound);
    const uint32_t index = __float2uint_rz(__fmaf_rn((float)level, r.H3, (float)nb_morton3D(nx, ny, nz)));   // :378
    const bool occ = (__ldg(r.grid + (index >> 3)) >> (index & 7u)) & 1u;
    if (occ) return true;
    const float tx = rm_exit((float)nx, r.sx, r.rH, mip_bound, x, r.rdx);
    const float ty = rm_exit((float)ny, r.sy, r.rH, mip_bound, y, r.rdy);
    const float tz = rm_exit((float)nz, r.sz, r.rH, mip_bound, z, r.rdz);
    tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    return false;
}

// One iteration of the marching loop (:359-399).  Returns true when the cell at t is occupied: x,y,z,dt are the
// sample; the caller emits it and advances t by dt.  Otherwise t has been advanced past the empty voxel.
__device__ __forceinline__ bool rm_step(const RayCtx &r, float &t, float &x, float &y, float &z, float &dt) {
    float tt;
    if (rm_probe(r, t, x, y, z, dt, tt)) return true;
    uint32_t n;
    t = rm_advance(r, t, tt, n);
    return false;
}

// ------------------------------------------------------------------------------------------------
// utils
// ------------------------------------------------------------------------------------------------
// slab test of raymarching.cu:108-144 on precomputed reciprocal directions; a miss gives near = far = FLT_MAX
__device__ __forceinline__ void rm_near_far(float ox, float oy, float oz, float rdx, float rdy, float rdz,
                                            const float *__restrict__ aabb, float min_near, float &near, float &far) {
    float tmp;
    near = __fmul_rn(__fsub_rn(aabb[0], ox), rdx); far = __fmul_rn(__fsub_rn(aabb[3], ox), rdx);
    if (near > far) { tmp = near; near = far; far = tmp; }
    float near_y = __fmul_rn(__fsub_rn(aabb[1], oy), rdy), far_y = __fmul_rn(__fsub_rn(aabb[4], oy), rdy);
    if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
    if (near > far_y || near_y > far) { near = far = FLT_MAX; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = __fmul_rn(__fsub_rn(aabb[2], oz), rdz), far_z = __fmul_rn(__fsub_rn(aabb[5], oz), rdz);
    if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
    if (near > far_z || near_z > far) { near = far = FLT_MAX; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
}

__global__ void k_near_far_from_aabb(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                     const float *__restrict__ aabb, uint32_t N, float min_near,
                                     float *__restrict__ nears, float *__restrict__ fars) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float rdx = __fdiv_rn(1.0f, rays_d[n * 3]), rdy = __fdiv_rn(1.0f, rays_d[n * 3 + 1]),
                rdz = __fdiv_rn(1.0f, rays_d[n * 3 + 2]);
    float near, far;
    rm_near_far(ox, oy, oz, rdx, rdy, rdz, aabb, min_near, near, far);
    nears[n] = near;
    fars[n] = far;
}

__global__ void k_sph_from_ray(const float *__restrict__ rays_o, const float *__restrict__ rays_d, float radius,
                               uint32_t N, float *__restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    constexpr float RPI = 0.3183098861837907f;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float Cq = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * Cq)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    coords[n * 2] = 2 * atan2f(sqrtf(x * x + z * z), y) * RPI - 1;
    coords[n * 2 + 1] = atan2f(z, x) * RPI;
}

__global__ void k_morton3D(const int32_t *__restrict__ coords, uint32_t N, int32_t *__restrict__ indices) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    indices[n] = (int32_t)nb_morton3D((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}

__global__ void k_morton3D_invert(const int32_t *__restrict__ indices, uint32_t N, int32_t *__restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const int32_t ind = indices[n];
    coords[n * 3] = (int32_t)nb_morton3D_invert((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int32_t)nb_morton3D_invert((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int32_t)nb_morton3D_invert((uint32_t)(ind >> 2));
}

// one thread per output byte, 2 x float4 loads
__global__ void k_packbits(const float *__restrict__ grid, uint32_t N, float thresh, uint8_t *__restrict__ bitfield) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float4 a = __ldg(reinterpret_cast<const float4 *>(grid) + (size_t)n * 2);
    const float4 b = __ldg(reinterpret_cast<const float4 *>(grid) + (size_t)n * 2 + 1);
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;   bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;   bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;  bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;  bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// ------------------------------------------------------------------------------------------------
// march_rays_train: count (+ record) -> scan -> expand
// ------------------------------------------------------------------------------------------------
// The traversal is a serial dependent chain per ray (~300 voxel visits of ~100 dependent instructions each) and an
// image is only ~15 k rays, i.e. < 1 warp per SM scheduler when 32 rays share a warp: the kernel is latency bound with
// most issue slots empty.  So (1) a warp carries only `rpw` rays (lanes >= rpw idle): 32 / rpw times more warps in
// flight to hide the chain's latency, and less divergence inside a warp; (2) every ray is traversed ONCE: the count
// pass records the parameter t of each sample it would emit (up to `tcap` per ray) and the second pass only expands
// those records into xyzs / dirs / deltas with one warp per ray and coalesced stores (a ray with more than `tcap`
// samples is re-marched by one lane, as the reference does for every ray, raymarching.cu:418-479).
constexpr int kMarchBlock = 32;    // one warp per block

// sample records kept per ray: all of them (max_steps <= 1024) while that stays under ~1 GB
__host__ __device__ __forceinline__ uint32_t march_tcap(uint32_t N) { return N <= (1u << 18) ? 1024u : (N <= (1u << 21) ? 256u : 64u); }
static uint32_t march_rpw(uint32_t N) {
    static int env = -1;
    if (env < 0) { const char *e = getenv("NB200_MARCH_RPW"); env = e ? atoi(e) : 0; }
    if (env == 1 || env == 2 || env == 4 || env == 8 || env == 16 || env == 32) return (uint32_t)env;
    uint32_t r = 32;                                   // aim at >= ~6 warps per SM scheduler (148 x 4 of them)
    while (r > 4 && (uint64_t)N / r < 148ull * 4 * 6) r >>= 1;
    return r;
}

// pass 1 (:353-400): counts, sample records, block-local exclusive offsets (warp shuffles), per-block sums
__global__ void __launch_bounds__(kMarchBlock)
k_march_count(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
              float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
              const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
              int32_t *__restrict__ rays, int32_t *__restrict__ block_sums, uint32_t rpw, float *__restrict__ trec,
              uint32_t tcap) {
    const uint32_t lane = threadIdx.x;
    const uint32_t n = blockIdx.x * rpw + lane;
    const bool live = lane < rpw && n < N;
    int num_steps = 0;
    if (live) {
        RayCtx r;
        rm_setup(r, rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, grid, bound, dt_gamma, max_steps, C, H);
        const float far = fars[n];
        float t = nears[n];
        t = __fmaf_rn(rm_dt(r, t), noises ? noises[n] : 0.0f, t);      // :351
        float *rec = trec ? trec + (size_t)n * tcap : nullptr;
        float x, y, z, dt;
        while (t < far && (uint32_t)num_steps < max_steps) {
            if (rm_step(r, t, x, y, z, dt)) {
                if (rec && (uint32_t)num_steps < tcap) rec[num_steps] = t;
                num_steps++;
                t = __fadd_rn(t, dt);
            }
        }
    }
    const int incl = nb_warp_incl_scan(num_steps);
    if (live) {
        rays[n * 3] = (int32_t)n;
        rays[n * 3 + 1] = incl - num_steps;             // block-local exclusive offset, globalised by k_march_fixup
        rays[n * 3 + 2] = num_steps;
    }
    if (lane == 31) block_sums[blockIdx.x] = incl;
}

// ---- occupied-region bounds: rays that cannot meet an occupied cell are not marched at all ------------------------
// obounds[level][6] = {max(x + 1), max(y + 1), max(z + 1), max(H - x), max(H - y), max(H - z)} over the set cells of a
// cascade level (all zero = level empty; the buffer is zero-filled before the launch).
__global__ void __launch_bounds__(256)
k_occ_bounds(const uint8_t *__restrict__ grid, uint32_t C, uint32_t H, int32_t *__restrict__ obounds) {
    const uint32_t H3 = H * H * H, wpl = H3 / 32;                    // 32-cell words per level
    // one level at a time so that the six running maxima stay in registers; one warp reduction + 6 atomics per level
    for (uint32_t level = 0; level < C; level++) {
        int m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0;
        for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < wpl; wi += gridDim.x * blockDim.x) {
            uint32_t bits = __ldg(reinterpret_cast<const uint32_t *>(grid) + level * wpl + wi);
            while (bits) {
                const uint32_t b = (uint32_t)__ffs((int)bits) - 1u;
                bits &= bits - 1u;
                const uint32_t m = wi * 32 + b;
                const int x = (int)nb_morton3D_invert(m), y = (int)nb_morton3D_invert(m >> 1), z = (int)nb_morton3D_invert(m >> 2);
                m0 = max(m0, x + 1); m1 = max(m1, y + 1); m2 = max(m2, z + 1);
                m3 = max(m3, (int)H - x); m4 = max(m4, (int)H - y); m5 = max(m5, (int)H - z);
            }
        }
        m0 = __reduce_max_sync(0xffffffffu, m0); m1 = __reduce_max_sync(0xffffffffu, m1);
        m2 = __reduce_max_sync(0xffffffffu, m2); m3 = __reduce_max_sync(0xffffffffu, m3);
        m4 = __reduce_max_sync(0xffffffffu, m4); m5 = __reduce_max_sync(0xffffffffu, m5);
        if (nb_lane() == 0 && m0 > 0) {
            int32_t *o = obounds + level * 6;
            atomicMax(o + 0, m0); atomicMax(o + 1, m1); atomicMax(o + 2, m2);
            atomicMax(o + 3, m3); atomicMax(o + 4, m4); atomicMax(o + 5, m5);
        }
    }
}

// Can the ray segment t in [t0, far] come within one cell of an occupied cell of any level?  Conservative by
// construction: each level's box of occupied cells is dilated by a full cell (>= 1/64 of the scene, five orders of
// magnitude above the rounding of position and cell index, raymarching.cu:361-376) and the slab intervals by 1e-3, so a
// "no" proves that every cell the marching loop would probe is empty and the ray emits nothing.  All lanes of a warp
// evaluate the same ray: the result is warp-uniform.
__device__ __forceinline__ bool rm_may_hit(const RayCtx &r, const int32_t *__restrict__ obounds, uint32_t C, uint32_t H,
                                           float t0, float far) {
    if (!obounds) return true;
    // positions are clamped to the scene box (:361-363): only argue about rays whose whole range lies inside it
    {
        const float e0 = fmaf(t0, r.dx, r.ox), e1 = fmaf(t0, r.dy, r.oy), e2 = fmaf(t0, r.dz, r.oz);
        const float f0 = fmaf(far, r.dx, r.ox), f1 = fmaf(far, r.dy, r.oy), f2 = fmaf(far, r.dz, r.oz);
        const float lim = r.bound + 1e-3f;
        const float m = fmaxf(fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fabsf(e2)), fmaxf(fmaxf(fabsf(f0), fabsf(f1)), fabsf(f2)));
        if (!(m <= lim)) return true;           // (also catches NaN / inf)
    }
    for (uint32_t level = 0; level < C; level++) {
        const int32_t *o = obounds + level * 6;
        const int hx = __ldg(o + 0), hy = __ldg(o + 1), hz = __ldg(o + 2);
        if (hx <= 0) continue;                                       // no occupied cell at this level
        const int lx = (int)H - __ldg(o + 3), ly = (int)H - __ldg(o + 4), lz = (int)H - __ldg(o + 5);
        const float mb = fminf(__int_as_float((127 + (int)level) << 23), r.bound);
        const float cs = 2.0f * mb / (float)H;
        // occupied cells [l, h) per axis, dilated by one cell
        const float bx0 = (float)(lx - 1) * cs - mb, bx1 = (float)(hx + 1) * cs - mb;
        const float by0 = (float)(ly - 1) * cs - mb, by1 = (float)(hy + 1) * cs - mb;
        const float bz0 = (float)(lz - 1) * cs - mb, bz1 = (float)(hz + 1) * cs - mb;
        float tn = t0 - 1e-3f, tf = far + 1e-3f;
        bool miss = false;
        const float oo[3] = {r.ox, r.oy, r.oz}, dd[3] = {r.dx, r.dy, r.dz};
        const float b0[3] = {bx0, by0, bz0}, b1[3] = {bx1, by1, bz1};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (fabsf(dd[a]) < 1e-12f) {
                if (oo[a] < b0[a] || oo[a] > b1[a]) miss = true;     // parallel to the slab and outside it
            } else {
                const float inv = 1.0f / dd[a];
                float ta = (b0[a] - oo[a]) * inv, tb = (b1[a] - oo[a]) * inv;
                if (ta > tb) { const float tmp = ta; ta = tb; tb = tmp; }
                tn = fmaxf(tn, ta - 1e-3f);
                tf = fminf(tf, tb + 1e-3f);
            }
        }
        if (!miss && tn <= tf) return true;
    }
    return false;
}

// ---- pass 1 for the constant-step case (dt_gamma == 0): one WARP per ray, speculative segments -----------------
// With a constant step every parameter the loop can visit lies on the lattice T(0) = t0, T(k+1) = fl(T(k) + dt), and
// T(k) is available in closed form (rm_lattice_jump).  The loop is a chain over lattice indices:
//     k occupied -> emit, k + 1;      k empty -> the first j > k with T(j) >= voxel exit.
// The chain is serial, but its restriction to a window of indices depends only on where it ENTERS the window.  So the
// warp cuts a window of 32 * seg indices into 32 segments; lane s marches segment s speculating that its first index
// a_s is visited, recording the samples it would emit and where it lands beyond the segment.  A cheap sequential pass
// then threads the true chain through the segments: the entry e of segment s (the landing of the last valid segment)
// must be a_s itself or the first point lane s moved to (v1) -- both lie on lane s's chain, so everything lane s did
// from e on IS the true chain (a sample speculatively emitted AT a_s is dropped when e = v1).  In the rare case where
// e is neither, lane s re-marches its segment from e.  Samples are therefore identical to the serial loop's by
// construction, never by tolerance; the critical path drops from the whole ray (~500 dependent voxel visits) to one
// segment (<= 64 lattice points).
constexpr int kSegWarps = 4;            // rays per block
constexpr int kSegMax = 64;             // lattice indices per lane and round
constexpr int kSegMin = 4;
constexpr int kRecStride = kSegMax + 1; // +1: lanes write their own rows, keep them on different banks

struct SegResult { uint32_t cnt, v1, land; float tland; bool occ_a; };

__device__ __forceinline__ SegResult rm_march_segment(const RayCtx &r, uint32_t k, float t, uint32_t b, float far,
                                                      float *__restrict__ rec) {
    SegResult o;
    o.cnt = 0; o.v1 = 0xffffffffu; o.occ_a = false;
    const uint32_t a = k;
    float x, y, z, dt, tt;
    while (k < b && t < far) {
        if (rm_probe<true>(r, t, x, y, z, dt, tt)) {
            rec[o.cnt++] = t;
            if (k == a) o.occ_a = true;
            k += 1;
            t = __fadd_rn(t, dt);
        } else {
            uint32_t n;
            t = rm_advance<true>(r, t, tt, n);
            k += n;
        }
        if (o.v1 == 0xffffffffu) o.v1 = k;
    }
    o.land = k; o.tland = t;
    return o;
}

// kLanes lanes cooperate on one ray (32 / kLanes rays per warp), and every lane marches kChunks segments that lie a
// quarter (1 / kChunks) of the window apart BEFORE the warp synchronises to thread the chain through them.  A warp
// issues for its slowest lane: one contiguous segment per lane leaves the lanes inside the object marching several
// times longer than the lanes in empty space (measured: 3-4x more probes issued than a balanced warp would need), while
// the sum of segments spread over the ray evens out.
template <int kLanes, int kChunks>
__global__ void __launch_bounds__(kSegWarps * 32)
k_march_count_seg(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
                  float bound, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                  float *__restrict__ nears, float *__restrict__ fars, const float *__restrict__ noises,
                  int32_t *__restrict__ counts, int32_t *__restrict__ group_sums, float *__restrict__ trec, uint32_t tcap,
                  const int32_t *__restrict__ obounds, const float *__restrict__ aabb, float min_near) {
    __shared__ float rec_s[kSegWarps][32][kRecStride];
    constexpr uint32_t kFull = 0xffffffffu, kRaysPerWarp = 32 / kLanes, kGroupMask = kLanes == 32 ? kFull : ((1u << kLanes) - 1u);
    constexpr int kSegCap = kSegMax / kChunks;                      // a lane's kChunks segments share its record row
    const uint32_t lane = nb_lane(), w = threadIdx.x >> 5;
    const uint32_t li = lane % kLanes, gshift = lane - li;          // index inside the ray's lane group, group's first lane
    const uint32_t n = (blockIdx.x * kSegWarps + w) * kRaysPerWarp + lane / kLanes;
    uint32_t total = 0;
    RayCtx r;
    float far = 0.0f, t0 = 0.0f, dt = 0.0f;
    bool ended = true;
    float *rec = rec_s[w][lane];
    float *out = nullptr;
    if (n < N) {
        rm_setup(r, rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, grid, bound, 0.0f, max_steps, C, H);
        if (aabb) {         // near_far_from_aabb fused in (same reciprocals, same operation order: bit-identical)
            rm_near_far(r.ox, r.oy, r.oz, r.rdx, r.rdy, r.rdz, aabb, min_near, t0, far);
            if (li == 0) { nears[n] = t0; fars[n] = far; }
        } else {
            far = fars[n]; t0 = nears[n];
        }
        dt = r.dt_min_c;
        t0 = __fmaf_rn(rm_dt(r, t0), noises ? noises[n] : 0.0f, t0);      // :351
        out = trec ? trec + (size_t)n * tcap : nullptr;
        ended = !(t0 < far && rm_may_hit(r, obounds, C, H, t0, far));
    }
    const float kest = ended ? 0.0f : fminf(__fdividef(far - t0, dt) + 2.0f, 1.0e9f);
    const uint32_t seg = (uint32_t)min(max((int)ceilf(kest * (1.0f / (kLanes * kChunks))), kSegMin), kSegCap);
    uint32_t K0 = 0;            // first index of the window == the chain's entry into it (exact for the group's lane 0)
    float tK0 = t0;
    // groups of a warp finish at different times: the collectives below always run warp-wide, a finished group's lanes
    // contribute values nobody uses
    while (__any_sync(kFull, !ended)) {
        // ---- march: chunk c of the window is lattice [K0 + c kLanes seg, K0 + (c + 1) kLanes seg), this lane's piece of
        //      it starts li seg further; no warp-level synchronisation between a lane's pieces
        uint32_t a[kChunks];
        SegResult sr[kChunks];
        {
            uint32_t kprev = K0;
            float tprev = tK0;
#pragma unroll
            for (int c = 0; c < kChunks; c++) {
                a[c] = 0;
                sr[c].cnt = 0; sr[c].v1 = 0; sr[c].land = 0; sr[c].tland = 0.0f; sr[c].occ_a = false;
                if (!ended) {
                    a[c] = K0 + ((uint32_t)c * kLanes + li) * seg;
                    const float ta = rm_lattice_jump(tprev, dt, a[c] - kprev);
                    sr[c] = rm_march_segment(r, a[c], ta, a[c] + seg, far, rec + c * seg);
                    kprev = a[c]; tprev = ta;
                }
            }
        }
        // ---- thread the true chain through the chunks, one after the other
        uint32_t e = K0;
        float te = tK0;
        bool now_ended = false;
#pragma unroll
        for (int c = 0; c < kChunks; c++) {
            const bool live = !ended && !now_ended;         // this group still follows the chain into chunk c
            uint32_t drop = 0, valid = 0;
            // common case, checked in parallel: every segment up to the one where the ray ends is entered at its own
            // first index or at the first point its lane moved to, i.e. the previous segment landed on a_s or on v1_s
            uint32_t pl = __shfl_up_sync(kFull, sr[c].land, 1, kLanes);
            if (li == 0) pl = e;
            const bool ok = pl == a[c] || pl == sr[c].v1;
            const uint32_t endm = (__ballot_sync(kFull, !(sr[c].tland < far)) >> gshift) & kGroupMask;
            const uint32_t last = endm ? (uint32_t)__ffs((int)endm) - 1u : (uint32_t)kLanes - 1u;      // segment in which the chain ends
            const uint32_t upto = last == 31u ? kFull : ((2u << last) - 1u);
            const uint32_t bad = (__ballot_sync(kFull, !ok) >> gshift) & kGroupMask & upto;
            const uint32_t land_l = __shfl_sync(kFull, sr[c].land, last, kLanes);
            const float tland_l = __shfl_sync(kFull, sr[c].tland, last, kLanes);
            uint32_t e2 = e;
            float te2 = te;
            bool end2 = false;
            if (bad == 0) {
                if (li <= last) {
                    drop = (pl != a[c] && sr[c].occ_a) ? 1u : 0u;
                    valid = sr[c].cnt - drop;
                }
                e2 = land_l; te2 = tland_l;
                end2 = endm != 0;
            }
            if (__any_sync(kFull, bad != 0 && live)) {        // rare: some group must be threaded sequentially
                const bool mine = bad != 0 && live;
                bool stop = !mine;
                for (uint32_t sgm = 0; sgm < (uint32_t)kLanes; sgm++) {
                    const uint32_t as = K0 + ((uint32_t)c * kLanes + sgm) * seg, bs = as + seg;
                    const uint32_t v1s = __shfl_sync(kFull, sr[c].v1, sgm, kLanes);
                    const bool occs = __shfl_sync(kFull, (int)sr[c].occ_a, sgm, kLanes) != 0;
                    const bool visit = !stop && e2 < bs;                    // (else the chain jumps over this segment)
                    uint32_t dr = 0;
                    if (visit && e2 != as) {
                        if (e2 == v1s) dr = occs ? 1u : 0u;
                        else if (li == sgm) sr[c] = rm_march_segment(r, e2, te2, bs, far, rec + c * seg);   // mis-speculation: re-march
                    }
                    if (visit && li == sgm) { drop = dr; valid = sr[c].cnt - dr; }
                    const uint32_t land_s = __shfl_sync(kFull, sr[c].land, sgm, kLanes);
                    const float tland_s = __shfl_sync(kFull, sr[c].tland, sgm, kLanes);
                    if (visit) {
                        e2 = land_s; te2 = tland_s;
                        if (!(te2 < far)) { end2 = true; stop = true; }
                    }
                }
            }
            if (!live) { valid = 0; drop = 0; }
            // ---- compact this chunk's samples behind the ray's earlier ones (max_steps caps the ray, :359)
            uint32_t incl = valid;
#pragma unroll
            for (int o = 1; o < kLanes; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, incl, o, kLanes);
                if ((int)li >= o) incl += t;
            }
            const uint32_t chunk_total = __shfl_sync(kFull, incl, kLanes - 1, kLanes);
            if (live) {
                const uint32_t first = total + incl - valid;
                if (out) {
                    const float *src = rec + c * seg + drop;
                    for (uint32_t i = 0; i < valid; i++) {
                        const uint32_t idx = first + i;
                        if (idx < max_steps && idx < tcap) out[idx] = src[i];
                    }
                }
                total += chunk_total;
                e = e2; te = te2;
                if (end2) now_ended = true;
                if (total >= max_steps) { total = max_steps; now_ended = true; }
            }
        }
        if (!ended) {
            K0 = e; tK0 = te;
            ended = now_ended;
        }
    }
    if (li == 0 && n < N) {
        counts[n] = (int32_t)total;
        if (total) atomicAdd(group_sums + (n >> 5), (int32_t)total);        // 32-ray groups: scanned by k_march_scan
    }
}

// rays[n] = (n, prefix of its 32-ray group + exclusive prefix inside the group, count): one thread per ray, coalesced
__global__ void __launch_bounds__(256)
k_march_finalize(const int32_t *__restrict__ counts, const int32_t *__restrict__ group_prefix, int32_t *__restrict__ rays,
                 uint32_t N) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    const int c = n < N ? counts[n] : 0;
    const int incl = nb_warp_incl_scan(c);
    if (n >= N) return;
    rays[n * 3] = (int32_t)n;
    rays[n * 3 + 1] = group_prefix[n >> 5] + incl - c;
    rays[n * 3 + 2] = c;
}

// single block: exclusive scan of the block sums starting at counter[0]; counter += (sum, N)  (:405-406)
__global__ void __launch_bounds__(1024)
k_march_scan(const int32_t *__restrict__ block_sums, int32_t *__restrict__ block_prefix, uint32_t nb, uint32_t N,
             int32_t *__restrict__ counter, uint32_t M_cap, int32_t *__restrict__ m_eff) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s, chunk_total_s;
    if (threadIdx.x == 0) carry_s = counter[0];
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const int v = (i < nb) ? block_sums[i] : 0;
        const int incl = nb_warp_incl_scan(v);
        if (nb_lane() == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            const int wt = warp_tot[threadIdx.x];
            const int wi = nb_warp_incl_scan(wt);
            warp_tot[threadIdx.x] = wi - wt;            // exclusive warp offsets
            if (threadIdx.x == 31) chunk_total_s = wi;
        }
        __syncthreads();
        const int carry = carry_s;
        if (i < nb) block_prefix[i] = carry + warp_tot[threadIdx.x >> 5] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + chunk_total_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counter[0] = carry_s;
        counter[1] += (int32_t)N;
        // rows the downstream kernels of the fused step may touch: every segment below this row is complete
        // (k_march_expand lowers it to the offset of the first ray that does not fit in M_cap rows)
        if (m_eff) *m_eff = (int32_t)min((uint32_t)max(carry_s, 0), M_cap);
    }
}

__global__ void k_march_fixup(int32_t *__restrict__ rays, const int32_t *__restrict__ block_prefix, uint32_t N,
                              uint32_t rpw) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    rays[n * 3 + 1] += block_prefix[n / rpw];
}

// pass 2 (:418-479): one warp per ray.  Lane i of trip k owns sample 32 k + i of the ray: the sample position is
// recomputed from its recorded t with the same FMA / clamp as the traversal, deltas[:,1] from the previous record.
constexpr int kExpandBlock = 256;
__global__ void __launch_bounds__(kExpandBlock)
k_march_expand(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
               float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
               const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
               const int32_t *__restrict__ rays, float *__restrict__ xyzs, float *__restrict__ dirs,
               float *__restrict__ deltas, int32_t *__restrict__ m_eff, const float *__restrict__ trec, uint32_t tcap) {
    const uint32_t n = (threadIdx.x + blockIdx.x * blockDim.x) >> 5;
    if (n >= N) return;
    const uint32_t lane = nb_lane();
    const uint32_t point_index = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0) return;
    if (point_index + num_steps > M) {
        if (m_eff && lane == 0) atomicMin(m_eff, (int32_t)min(point_index, M));
        return;
    }
    RayCtx r;
    rm_setup(r, rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, grid, bound, dt_gamma, max_steps, C, H);
    float t = nears[n];
    t = __fmaf_rn(rm_dt(r, t), noises ? noises[n] : 0.0f, t);
    float *px = xyzs + (size_t)point_index * 3, *pd = dirs + (size_t)point_index * 3, *pl = deltas + (size_t)point_index * 2;
    if (trec && num_steps <= tcap) {
        const float *rec = trec + (size_t)n * tcap;
        for (uint32_t base = 0; base < num_steps; base += 32) {
            const uint32_t i = base + lane;
            if (i < num_steps) {
                const float ti = rec[i];
                float last_t = t;                                   // the ray's (perturbed) start for its first sample
                if (i > 0) { const float tp = rec[i - 1]; last_t = __fadd_rn(tp, rm_dt(r, tp)); }
                const float dt = rm_dt(r, ti);
                px[i * 3] = rm_clamp(__fmaf_rn(ti, r.dx, r.ox), -r.bound, r.bound);
                px[i * 3 + 1] = rm_clamp(__fmaf_rn(ti, r.dy, r.oy), -r.bound, r.bound);
                px[i * 3 + 2] = rm_clamp(__fmaf_rn(ti, r.dz, r.oz), -r.bound, r.bound);
                pd[i * 3] = r.dx; pd[i * 3 + 1] = r.dy; pd[i * 3 + 2] = r.dz;
                *reinterpret_cast<float2 *>(pl + i * 2) = make_float2(dt, __fsub_rn(__fadd_rn(ti, dt), last_t));
            }
        }
        return;
    }
    if (lane != 0) return;                  // more samples than records: re-march this ray serially
    const float far = fars[n];
    float last_t = t, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < num_steps) {
        if (rm_step(r, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t = __fadd_rn(t, dt);
            pl[0] = dt; pl[1] = __fsub_rn(t, last_t);
            last_t = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// compositing (training): one warp per ray
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_incl_prod(float p) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float q = __shfl_up_sync(0xffffffffu, p, o);
        if ((int)nb_lane() >= o) p *= q;
    }
    return p;
}
__device__ __forceinline__ float warp_incl_sum(float p) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float q = __shfl_up_sync(0xffffffffu, p, o);
        if ((int)nb_lane() >= o) p += q;
    }
    return p;
}

constexpr int kCompBlock = 256;   // 8 rays per block

// colour of sample s: float [M,3] (the reference layout) or half [M,4] (rgb + mask, straight from the field kernel)
template <typename TC>
__device__ __forceinline__ void ld_rgba(const TC *__restrict__ rgbs, size_t s, float &c0, float &c1, float &c2, float &c3) {
    if constexpr (sizeof(TC) == 2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(rgbs) + s);
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
        c0 = a.x; c1 = a.y; c2 = b.x; c3 = b.y;                 // 4th channel: the mask head (train_conf)
    } else {
        c0 = __ldg(rgbs + s * 3); c1 = __ldg(rgbs + s * 3 + 1); c2 = __ldg(rgbs + s * 3 + 2); c3 = 0.0f;
    }
}
template <typename TC>
__device__ __forceinline__ void ld_rgb(const TC *__restrict__ rgbs, size_t s, float &c0, float &c1, float &c2) {
    float c3;
    ld_rgba<TC>(rgbs, s, c0, c1, c2, c3);
}

// optional fused MSE (fused train step): loss[0] += sum (image - target)^2 * inv_n, g_image = 2 (image - target) inv_n scale
// mask part (train_conf, utils_init_nerf.py:231-233): render_mask = sum w * mask (the 4th channel of the half rgba rows),
// loss += mask_weight * mean (render_mask - target_mask)^2, g_render_mask its gradient
struct MseArgs {
    const float *target; float *loss, *g_image; float inv_n, scale;
    const float *target_mask; float *render_mask, *g_render_mask; float mask_weight;
    const float *scale_dev = nullptr;       // device-side loss scale (the scaler words of adam.cuh): overrides `scale`
};

// LGIE composites (the editing render of nerf/renderer.py:383-474, on the occupancy path: rendering._lgie_composites):
// the same samples composited three times with the density gated by the mask head's output m (the 4th rgba channel):
//   V = 0 "all": sigma;   V = 1 "fg": sigma * e(m);   V = 2 "bg": sigma * (1 - e(m))
//   e = sigmoid((m - conf_thr) * 100) (soft_mask, :421-426) or [m > 0.5] (hard)
// V = -1 is the plain composite (no gate; identical code to before the LGIE variants existed).
struct LgieArgs { float thr; int soft, detach_bg, detach_mask, accumulate; };
template <int V>
__device__ __forceinline__ float lgie_gate(float m, const LgieArgs &a, float &dgate_dm) {
    dgate_dm = 0.0f;
    if constexpr (V <= 0) return 1.0f;
    float e;
    if (a.soft) {
        e = 1.0f / (1.0f + expf(-(m - a.thr) * 100.0f));
        dgate_dm = 100.0f * e * (1.0f - e);
    } else {
        e = m > 0.5f ? 1.0f : 0.0f;
    }
    if constexpr (V == 2) { dgate_dm = -dgate_dm; return 1.0f - e; }
    return e;
}

// forward compositing of one ray by one warp: 32 samples per trip, prefixes by shuffle scans; every lane returns the sums
template <typename TC, int V>
__device__ __forceinline__ void comp_fwd_ray(const float *__restrict__ sigmas, const TC *__restrict__ rgbs,
                                             const float *__restrict__ deltas, uint32_t offset, uint32_t num_steps, uint32_t M,
                                             float T_thresh, uint32_t lane, const LgieArgs &lg, bool need_mask, float &r, float &g,
                                             float &b, float &ws, float &d, float &mk) {
    if (num_steps != 0 && offset + num_steps <= M) {
        float T_carry = 1.0f, t_carry = 0.0f;
        for (uint32_t base = 0; base < num_steps; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < num_steps;
            const size_t s = (size_t)offset + i;
            float sigma = 0, d0 = 0, d1 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            if (valid) {
                sigma = __ldg(sigmas + s);
                const float2 dl = __ldg(reinterpret_cast<const float2 *>(deltas) + s);
                d0 = dl.x; d1 = dl.y;
                ld_rgba<TC>(rgbs, s, c0, c1, c2, c3);
                if constexpr (V > 0) { float dg; sigma *= lgie_gate<V>(c3, lg, dg); }
            }
            const float alpha = valid ? 1.0f - __expf(-sigma * d0) : 0.0f;
            const float pin = warp_incl_prod(1.0f - alpha);
            float pex = __shfl_up_sync(0xffffffffu, pin, 1);
            if (lane == 0) pex = 1.0f;
            const float T_after = T_carry * pin;
            const float t_i = t_carry + warp_incl_sum(d1);
            const uint32_t term = __ballot_sync(0xffffffffu, valid && (T_after < T_thresh));
            const uint32_t last = term ? (uint32_t)(__ffs(term) - 1) : 31u;   // the breaking sample is included (:554-557)
            const float weight = (valid && lane <= last) ? alpha * (T_carry * pex) : 0.0f;
            r = fmaf(weight, c0, r); g = fmaf(weight, c1, g); b = fmaf(weight, c2, b);
            mk = fmaf(weight, c3, mk);
            d = fmaf(weight, t_i, d);
            ws += weight;
            if (term) break;
            T_carry = __shfl_sync(0xffffffffu, T_after, 31);
            t_carry = __shfl_sync(0xffffffffu, t_i, 31);
        }
        r = nb_warp_sum(r); g = nb_warp_sum(g); b = nb_warp_sum(b); ws = nb_warp_sum(ws); d = nb_warp_sum(d);
        if (need_mask) mk = nb_warp_sum(mk);
    }
}

template <typename TC, int V = -1>
__global__ void __launch_bounds__(kCompBlock)
k_composite_train_fwd(const float *__restrict__ sigmas, const TC *__restrict__ rgbs, const float *__restrict__ deltas,
                      const int32_t *__restrict__ rays, uint32_t M, uint32_t N, float T_thresh,
                      float *__restrict__ weights_sum, float *__restrict__ depth, float *__restrict__ image, MseArgs mse,
                      LgieArgs lg = LgieArgs{}) {
    __shared__ float loss_part[kCompBlock / 32];
    const uint32_t n = (threadIdx.x + blockIdx.x * blockDim.x) >> 5;
    const uint32_t lane = nb_lane();
    if (mse.target) {                       // whole block stays alive for the block-level loss reduction
        if (lane == 0) loss_part[threadIdx.x >> 5] = 0.0f;
        if (n >= N) {
            __syncthreads();
            if (threadIdx.x == 0) {
                float v = 0.0f;
                for (int i = 0; i < kCompBlock / 32; i++) v += loss_part[i];
                if (v != 0.0f) atomicAdd(mse.loss, v * mse.inv_n);
            }
            return;
        }
    } else if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    float r = 0, g = 0, b = 0, ws = 0, d = 0, mk = 0;
    comp_fwd_ray<TC, V>(sigmas, rgbs, deltas, offset, num_steps, M, T_thresh, lane, lg, mse.render_mask != nullptr, r, g, b, ws, d, mk);
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
        if (mse.render_mask) mse.render_mask[index] = mk;
    }
    if (mse.target) {
        if (lane == 0) {
            const float d0 = r - mse.target[index * 3], d1 = g - mse.target[index * 3 + 1], d2 = b - mse.target[index * 3 + 2];
            const float lscale = mse.scale_dev ? __ldg(mse.scale_dev) : mse.scale;
            const float k = 2.0f * mse.inv_n * lscale;
            mse.g_image[index * 3] = d0 * k; mse.g_image[index * 3 + 1] = d1 * k; mse.g_image[index * 3 + 2] = d2 * k;
            float part = d0 * d0 + d1 * d1 + d2 * d2;
            if (mse.target_mask) {          // mean over N x 1 entries = 3 x the per-element weight of the N x 3 image
                const float dm = mk - mse.target_mask[index];
                mse.g_render_mask[index] = 2.0f * dm * (3.0f * mse.inv_n) * mse.mask_weight * lscale;
                part += 3.0f * mse.mask_weight * dm * dm;
            }
            loss_part[threadIdx.x >> 5] = part;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = 0.0f;
            for (int i = 0; i < kCompBlock / 32; i++) v += loss_part[i];
            atomicAdd(mse.loss, v * mse.inv_n);
        }
    }
}

// backward compositing of one ray by one warp (the loop of k_composite_train_bwd; also the second half of the fused kernel)
template <typename TC, int GS, int V>
__device__ __forceinline__ void comp_bwd_ray(const float *__restrict__ sigmas, const TC *__restrict__ rgbs,
                                             const float *__restrict__ deltas, uint32_t offset, uint32_t num_steps, float T_thresh,
                                             uint32_t lane, const LgieArgs &lg, bool has_m, float gm, float m_final, float gws,
                                             float gi0, float gi1, float gi2, float r_final, float g_final, float b_final,
                                             float ws_final, float *__restrict__ grad_sigmas, float *__restrict__ grad_rgbs) {
    float m_c = 0;
    const float ws_term = gws * (1.0f - ws_final);
    float T_carry = 1.0f, r_c = 0, g_c = 0, b_c = 0;
    bool done = false;
    for (uint32_t base = 0; base < num_steps; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < num_steps;
        const size_t s = (size_t)offset + i;
        float gs = 0, gr0 = 0, gr1 = 0, gr2 = 0, gr3 = 0;
        if (!done) {
            float sigma = 0, d0 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            [[maybe_unused]] float sigma_raw = 0, gate = 1.0f, dgate = 0.0f;
            if (valid) {
                sigma = __ldg(sigmas + s);
                d0 = __ldg(deltas + s * 2);
                ld_rgba<TC>(rgbs, s, c0, c1, c2, c3);
                if constexpr (V >= 0) { sigma_raw = sigma; gate = lgie_gate<V>(c3, lg, dgate); sigma *= gate; }
            }
            const float alpha = valid ? 1.0f - __expf(-sigma * d0) : 0.0f;
            const float pin = warp_incl_prod(1.0f - alpha);
            float pex = __shfl_up_sync(0xffffffffu, pin, 1);
            if (lane == 0) pex = 1.0f;
            const float T_after = T_carry * pin;
            const uint32_t term = __ballot_sync(0xffffffffu, valid && (T_after < T_thresh));
            const uint32_t last = term ? (uint32_t)(__ffs(term) - 1) : 31u;
            const bool act = valid && lane <= last;
            const float weight = act ? alpha * (T_carry * pex) : 0.0f;
            const float r_i = r_c + warp_incl_sum(weight * c0);
            const float g_i = g_c + warp_incl_sum(weight * c1);
            const float b_i = b_c + warp_incl_sum(weight * c2);
            float m_i = 0.0f;
            if (has_m) m_i = m_c + warp_incl_sum(weight * c3);      // (warp-uniform branch)
            if (act) {
                gr0 = gi0 * weight; gr1 = gi1 * weight; gr2 = gi2 * weight; gr3 = gm * weight;
                float mask_term = gm * (T_after * c3 - (m_final - m_i));
                if constexpr (V >= 0) { if (lg.detach_mask) mask_term = 0.0f; }
                gs = d0 * (gi0 * (T_after * c0 - (r_final - r_i)) + gi1 * (T_after * c1 - (g_final - g_i)) +
                           gi2 * (T_after * c2 - (b_final - b_i)) + mask_term + ws_term);   // :752-757
                if constexpr (V >= 0) {
                    const float a = (V == 0 && lg.detach_bg && !(c3 >= 0.5f)) ? 0.0f : 1.0f;
                    gr0 *= a; gr1 *= a; gr2 *= a;
                    gr3 += gs * sigma_raw * dgate;
                    gs *= gate * a;
                }
            }
            if (term) done = true;
            T_carry = __shfl_sync(0xffffffffu, T_after, 31);
            r_c = __shfl_sync(0xffffffffu, r_i, 31);
            g_c = __shfl_sync(0xffffffffu, g_i, 31);
            b_c = __shfl_sync(0xffffffffu, b_i, 31);
            if (has_m) m_c = __shfl_sync(0xffffffffu, m_i, 31);
        }
        if (valid) {   // rows after the early-out get explicit zeros (raymarching.py:284-285 zero-fills instead)
            if constexpr (V >= 0 && GS == 4) {
                if (lg.accumulate) {      // the same lane owns row s in every variant's launch: a plain read-modify-write
                    const float4 q = reinterpret_cast<const float4 *>(grad_rgbs)[s];
                    gs += grad_sigmas[s]; gr0 += q.x; gr1 += q.y; gr2 += q.z; gr3 += q.w;
                }
            }
            grad_sigmas[s] = gs;
            if constexpr (GS == 4) {
                reinterpret_cast<float4 *>(grad_rgbs)[s] = make_float4(gr0, gr1, gr2, gr3);
            } else {
                grad_rgbs[s * 3] = gr0; grad_rgbs[s * 3 + 1] = gr1; grad_rgbs[s * 3 + 2] = gr2;
            }
        }
    }
}

// GS = row stride of grad_rgbs: 3 (reference layout) or 4 (float4 rows [g_r, g_g, g_b, 0] for the field backward)
// V >= 0 (LGIE, GS == 4): contributions of variant V to the per-sample gradients -- d sigma = gs * gate * a,
// d m = g_mask * w + gs * sigma * d gate / d m, d rgb = a * g_rgb, with a = [m >= 0.5] for the "all" variant under
// detach_bg (background samples give values but no gradient to the global image, :409-418) and 1 otherwise;
// detach_mask: the rendered mask's weights are detached (:460-463), so its term leaves gs.  accumulate: add to the rows.
template <typename TC, int GS, int V = -1>
__global__ void __launch_bounds__(kCompBlock)
k_composite_train_bwd(const float *__restrict__ grad_weights_sum, const float *__restrict__ grad_image,
                      const float *__restrict__ sigmas, const TC *__restrict__ rgbs, const float *__restrict__ deltas,
                      const int32_t *__restrict__ rays, const float *__restrict__ weights_sum,
                      const float *__restrict__ image, uint32_t M, uint32_t N, float T_thresh,
                      float *__restrict__ grad_sigmas, float *__restrict__ grad_rgbs,
                      const float *__restrict__ grad_render_mask, const float *__restrict__ render_mask,
                      LgieArgs lg = LgieArgs{}) {
    const uint32_t n = (threadIdx.x + blockIdx.x * blockDim.x) >> 5;
    if (n >= N) return;
    const uint32_t lane = nb_lane();
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;
    // optional 4th composited channel (the rendered mask): same recurrences as a colour channel
    const bool has_m = grad_render_mask != nullptr;
    const float gm = has_m ? grad_render_mask[index] : 0.0f, m_final = has_m ? render_mask[index] : 0.0f;
    const float gws = grad_weights_sum[index];
    comp_bwd_ray<TC, GS, V>(sigmas, rgbs, deltas, offset, num_steps, T_thresh, lane, lg, has_m, gm, m_final, gws,
                            grad_image[index * 3], grad_image[index * 3 + 1], grad_image[index * 3 + 2], image[index * 3],
                            image[index * 3 + 1], image[index * 3 + 2], weights_sum[index], grad_sigmas, grad_rgbs);
}

// Fused train step: compositing forward + MSE (+ mask term) + compositing backward of a ray in ONE launch.  Everything the
// backward needs from the forward is per ray (the final sums and the loss gradient of that ray's pixel), so the warp that
// composited a ray turns round and walks its samples again -- they are still in L1 / L2 -- instead of a second kernel
// re-reading the per-ray results (two launch-latency-sized kernels of 12 + 10 us at 270 k samples).  Same loops, same
// arithmetic and the same outputs as k_composite_train_fwd<TC> followed by k_composite_train_bwd<TC, 4>.
template <typename TC>
__global__ void __launch_bounds__(kCompBlock)
k_composite_train_fused(const float *__restrict__ sigmas, const TC *__restrict__ rgbs, const float *__restrict__ deltas,
                        const int32_t *__restrict__ rays, uint32_t M, uint32_t N, float T_thresh,
                        float *__restrict__ weights_sum, float *__restrict__ depth, float *__restrict__ image, MseArgs mse,
                        const float *__restrict__ grad_weights_sum, float *__restrict__ grad_sigmas,
                        float *__restrict__ grad_rgbs) {
    __shared__ float loss_part[kCompBlock / 32];
    const uint32_t n = (threadIdx.x + blockIdx.x * blockDim.x) >> 5;
    const uint32_t lane = nb_lane();
    if (lane == 0) loss_part[threadIdx.x >> 5] = 0.0f;
    if (n < N) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        const LgieArgs lg{};
        const bool has_m = mse.target_mask != nullptr;
        float r = 0, g = 0, b = 0, ws = 0, d = 0, mk = 0;
        comp_fwd_ray<TC, -1>(sigmas, rgbs, deltas, offset, num_steps, M, T_thresh, lane, lg, mse.render_mask != nullptr, r, g, b, ws,
                             d, mk);
        // the pixel's loss gradient, in every lane (the loads are warp-uniform)
        const float e0 = r - mse.target[index * 3], e1 = g - mse.target[index * 3 + 1], e2 = b - mse.target[index * 3 + 2];
        const float lscale = mse.scale_dev ? __ldg(mse.scale_dev) : mse.scale;
        const float k = 2.0f * mse.inv_n * lscale;
        const float gi0 = e0 * k, gi1 = e1 * k, gi2 = e2 * k;
        float gm = 0.0f, part = e0 * e0 + e1 * e1 + e2 * e2;
        if (has_m) {                        // mean over N x 1 entries = 3 x the per-element weight of the N x 3 image
            const float dm = mk - mse.target_mask[index];
            gm = 2.0f * dm * (3.0f * mse.inv_n) * mse.mask_weight * lscale;
            part += 3.0f * mse.mask_weight * dm * dm;
        }
        if (lane == 0) {
            weights_sum[index] = ws;
            depth[index] = d;
            image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
            if (mse.render_mask) mse.render_mask[index] = mk;
            mse.g_image[index * 3] = gi0; mse.g_image[index * 3 + 1] = gi1; mse.g_image[index * 3 + 2] = gi2;
            if (has_m) mse.g_render_mask[index] = gm;
            loss_part[threadIdx.x >> 5] = part;
        }
        if (num_steps != 0 && offset + num_steps <= M)
            comp_bwd_ray<TC, 4, -1>(sigmas, rgbs, deltas, offset, num_steps, T_thresh, lane, lg, has_m, gm, mk,
                                    grad_weights_sum ? grad_weights_sum[index] : 0.0f, gi0, gi1, gi2, r, g, b, ws, grad_sigmas,
                                    grad_rgbs);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.0f;
        for (int i = 0; i < kCompBlock / 32; i++) v += loss_part[i];
        if (v != 0.0f) atomicAdd(mse.loss, v * mse.inv_n);
    }
}

// ------------------------------------------------------------------------------------------------
// inference
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *__restrict__ rays_alive, const float *__restrict__ rays_t,
             const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, float dt_gamma,
             uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
             const float *__restrict__ fars, float *__restrict__ xyzs, float *__restrict__ dirs,
             float *__restrict__ deltas, const float *__restrict__ noises) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    RayCtx r;
    rm_setup(r, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, dt_gamma, max_steps, C, H);
    float *px = xyzs + (size_t)n * n_step * 3, *pd = dirs + (size_t)n * n_step * 3, *pl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    const float far = fars[index];
    t = __fmaf_rn(rm_dt(r, t), noises ? noises[n] : 0.0f, t);       // :930
    float last_t = t, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < n_step) {
        if (rm_step(r, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t = __fadd_rn(t, dt);
            pl[0] = dt; pl[1] = __fsub_rn(t, last_t);
            last_t = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
}

__global__ void __launch_bounds__(128)
k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *__restrict__ rays_alive,
                 float *__restrict__ rays_t, const float *__restrict__ sigmas, const float *__restrict__ rgbs,
                 const float *__restrict__ deltas, float *__restrict__ weights_sum, float *__restrict__ depth,
                 float *__restrict__ image) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    const float *s = sigmas + (size_t)n * n_step, *c = rgbs + (size_t)n * n_step * 3, *dl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth[index];
    float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        if (dl[0] == 0) break;                                   // :1042
        const float alpha = 1.0f - __expf(-s[0] * dl[0]);
        const float T = 1 - weight_sum;                          // :1052
        const float weight = alpha * T;
        weight_sum += weight;
        t += dl[1];
        d = fmaf(weight, t, d);
        r = fmaf(weight, c[0], r); g = fmaf(weight, c[1], g); b = fmaf(weight, c[2], b);
        if (T < T_thresh) break;                                 // :1066
        s++; c += 3; dl += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;   // :1078-1082
    weights_sum[index] = weight_sum; depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

// ------------------------------------------------------------------------------------------------
// inference rounds driven from the device (no per-round host synchronisation)
// ------------------------------------------------------------------------------------------------
// The reference's eval loop (nerf/renderer.py:651-688) reads n_alive back every round to pick n_step and to compact
// rays_alive with a boolean mask.  Here the round state lives on the device: state = {n_alive, n_step, step, samples}.
// A round = plan -> march -> (encode, field) -> composite -> compact, every kernel launched for the worst case (N rays /
// N + pad samples) and bounded by the state, so a round can be captured once in a CUDA graph and replayed; the host
// only looks at n_alive every few rounds.  Per ray the samples and their order are those of the reference loop, so the
// composited image is the same.
__global__ void k_infer_plan(int32_t *__restrict__ state, uint32_t N, uint32_t max_steps) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n_alive = state[0];
    const int step = state[2];
    if (step >= (int)max_steps) n_alive = 0;                               // while step < max_steps   (:667)
    const int n_step = n_alive > 0 ? max(min((int)N / n_alive, 8), 1) : 0;   // :676
    state[0] = n_alive;
    state[1] = n_step;
    state[3] = n_alive * n_step;                                           // sample rows of this round
}

__global__ void __launch_bounds__(128)
k_march_rays_dev(const int32_t *__restrict__ state, const int32_t *__restrict__ rays_alive, const float *__restrict__ rays_t,
                 const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, float dt_gamma,
                 uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
                 const float *__restrict__ fars, float *__restrict__ xyzs, float *__restrict__ dirs,
                 float *__restrict__ deltas, const float *__restrict__ noises) {
    const uint32_t n_alive = (uint32_t)state[0], n_step = (uint32_t)state[1];
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    RayCtx r;
    rm_setup(r, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, dt_gamma, max_steps, C, H);
    float *px = xyzs + (size_t)n * n_step * 3, *pd = dirs + (size_t)n * n_step * 3, *pl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    const float far = fars[index];
    t = __fmaf_rn(rm_dt(r, t), (noises && state[2] == 0) ? noises[n] : 0.0f, t);      // perturb in the first round only (:677)
    float last_t = t, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < n_step) {
        if (rm_step(r, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t = __fadd_rn(t, dt);
            pl[0] = dt; pl[1] = __fsub_rn(t, last_t);
            last_t = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
    for (; step < n_step; step++) {          // unused slots read as "ray finished" (the reference zero-fills, raymarching.py:391-393)
        px[0] = px[1] = px[2] = 0.0f; pd[0] = pd[1] = pd[2] = 0.0f; pl[0] = pl[1] = 0.0f;
        px += 3; pd += 3; pl += 2;
    }
}

// The same round from the RECORDS of one whole-ray traversal (nb200_march_rays_train_count's speculative-segment march, once per
// frame): a round then only looks up the ray's next n_step recorded sample parameters instead of walking the grid from
// rays_t again -- the per-round DDA (one thread per ray, hundreds of dependent voxel visits in the first rounds) was 58 %
// of an inference frame.  Bit-identical by construction, with a check: the records are the chain started at the ray's
// (perturbed) near point; a round may consume them only if rays_t -- which compositing re-accumulates from the deltas,
// t += delta (raymarching.cu:1058), normally exactly -- still equals the parameter the chain reached after the samples
// consumed so far.  A ray for which it does not (or whose record list was truncated at max_steps / tcap) walks the grid
// from rays_t like k_march_rays_dev, from then on (consumed = -1; state[5] counts such rays per frame).
__global__ void __launch_bounds__(128)
k_march_rays_rec(int32_t *__restrict__ state, const int32_t *__restrict__ rays_alive, const float *__restrict__ rays_t,
                 const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, float dt_gamma,
                 uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
                 const float *__restrict__ fars, float *__restrict__ xyzs, float *__restrict__ dirs,
                 float *__restrict__ deltas, const float *__restrict__ noises, const int32_t *__restrict__ rays,
                 const float *__restrict__ trec, uint32_t tcap, int32_t *__restrict__ consumed) {
    const uint32_t n_alive = (uint32_t)state[0], n_step = (uint32_t)state[1];
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    RayCtx r;
    rm_setup(r, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, dt_gamma, max_steps, C, H);
    float *px = xyzs + (size_t)n * n_step * 3, *pd = dirs + (size_t)n * n_step * 3, *pl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    const bool first_round = state[2] == 0;
    t = __fmaf_rn(rm_dt(r, t), (noises && first_round) ? noises[n] : 0.0f, t);      // perturb in the first round only (:677)
    const uint32_t num = (uint32_t)rays[index * 3 + 2];
    const int32_t c = consumed[index];
    const float *rec = trec + (size_t)index * tcap;
    bool fast = c >= 0 && num <= tcap && num < max_steps;
    if (fast && c > 0) { const float tp = rec[c - 1]; fast = t == __fadd_rn(tp, rm_dt(r, tp)); }
    if (fast && c == 0) fast = first_round;
    uint32_t step = 0;
    if (fast) {
        float last_t = t;
        for (; step < n_step && (uint32_t)c + step < num; step++) {
            const float ti = rec[c + step], dt = rm_dt(r, ti), tn = __fadd_rn(ti, dt);
            px[0] = rm_clamp(__fmaf_rn(ti, r.dx, r.ox), -r.bound, r.bound);
            px[1] = rm_clamp(__fmaf_rn(ti, r.dy, r.oy), -r.bound, r.bound);
            px[2] = rm_clamp(__fmaf_rn(ti, r.dz, r.oz), -r.bound, r.bound);
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            pl[0] = dt; pl[1] = __fsub_rn(tn, last_t);
            last_t = tn;
            px += 3; pd += 3; pl += 2;
        }
        consumed[index] = c + (int32_t)step;
    } else {
        if (c >= 0) { consumed[index] = -1; atomicAdd(state + 5, 1); }
        const float far = fars[index];
        float last_t = t, x, y, z, dt;
        while (t < far && step < n_step) {
            if (rm_step(r, t, x, y, z, dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t = __fadd_rn(t, dt);
                pl[0] = dt; pl[1] = __fsub_rn(t, last_t);
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
    for (; step < n_step; step++) {          // unused slots read as "ray finished" (the reference zero-fills, raymarching.py:391-393)
        px[0] = px[1] = px[2] = 0.0f; pd[0] = pd[1] = pd[2] = 0.0f; pl[0] = pl[1] = 0.0f;
        px += 3; pd += 3; pl += 2;
    }
}

template <typename TC>
__global__ void __launch_bounds__(128)
k_composite_rays_dev(const int32_t *__restrict__ state, float T_thresh, int32_t *__restrict__ rays_alive,
                     float *__restrict__ rays_t, const float *__restrict__ sigmas, const TC *__restrict__ rgbs,
                     const float *__restrict__ deltas, float *__restrict__ weights_sum, float *__restrict__ depth,
                     float *__restrict__ image) {
    const uint32_t n_alive = (uint32_t)state[0], n_step = (uint32_t)state[1];
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    size_t s = (size_t)n * n_step;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth[index];
    float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        const float d0 = deltas[s * 2], d1 = deltas[s * 2 + 1];
        if (d0 == 0) break;                                      // :1042
        const float alpha = 1.0f - __expf(-sigmas[s] * d0);
        const float T = 1 - weight_sum;                          // :1052
        const float weight = alpha * T;
        float c0, c1, c2;
        ld_rgb<TC>(rgbs, s, c0, c1, c2);
        weight_sum += weight;
        t += d1;
        d = fmaf(weight, t, d);
        r = fmaf(weight, c0, r); g = fmaf(weight, c1, g); b = fmaf(weight, c2, b);
        if (T < T_thresh) break;                                 // :1066
        s++; step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;   // :1078-1082
    weights_sum[index] = weight_sum; depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

// compaction rays_alive[rays_alive >= 0] (:685), all SMs: warp ballots + one atomic per warp.  The survivors' ORDER is
// not the reference's (it is whatever order the warps' atomics land in) -- it only decides which slot of the next
// round's sample buffers a ray uses, never what is accumulated for the ray.  state[4] counts the survivors.
__global__ void __launch_bounds__(256)
k_compact_alive(int32_t *__restrict__ state, const int32_t *__restrict__ alive_in, int32_t *__restrict__ alive_out) {
    const uint32_t n_alive = (uint32_t)state[0];
    const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
    const int32_t v = (i < n_alive) ? alive_in[i] : -1;
    const bool keep = v >= 0;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (m == 0) return;
    int base = 0;
    if (nb_lane() == 0) base = atomicAdd(state + 4, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) alive_out[base + __popc(m & ((1u << nb_lane()) - 1u))] = v;
}

__global__ void k_copy_alive(const int32_t *__restrict__ state, const int32_t *__restrict__ src, int32_t *__restrict__ dst) {
    const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < (uint32_t)state[4]) dst[i] = src[i];
}

__global__ void k_infer_commit(int32_t *__restrict__ state) {        // n_alive = survivors, step += n_step
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    state[0] = state[4];
    state[2] += state[1];
    state[4] = 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int nb200_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N, float min_near,
                             float *nears, float *fars, void *stream) {
    if (N == 0) return 0;
    k_near_far_from_aabb<<<nb_div_up(N, 128), 128, 0, nb_stream(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords, void *stream) {
    if (N == 0) return 0;
    k_sph_from_ray<<<nb_div_up(N, 128), 128, 0, nb_stream(stream)>>>(rays_o, rays_d, radius, N, coords);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_morton3D(const int32_t *coords, uint32_t N, int32_t *indices, void *stream) {
    if (N == 0) return 0;
    k_morton3D<<<nb_div_up(N, 256), 256, 0, nb_stream(stream)>>>(coords, N, indices);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords, void *stream) {
    if (N == 0) return 0;
    k_morton3D_invert<<<nb_div_up(N, 256), 256, 0, nb_stream(stream)>>>(indices, N, coords);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream) {
    if (N == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(grid) & 15u) != 0) return NB200_E_BAD_ARG;
    k_packbits<<<nb_div_up(N, 256), 256, 0, nb_stream(stream)>>>(grid, N, density_thresh, bitfield);
    NB_LAUNCH_CHECK();
    return 0;
}

// scratch layout (int32 units): [per-ray counts: N + 8  (serial kernel: block sums nb4 | block prefixes nb4)]
// [occupied bounds: 8 levels x 6 + 16 spare][sample records: N * tcap floats], nb4 = ceil(N / 4)
constexpr uint32_t kScratchFixed = 64;
static inline uint32_t march_nb_max(uint32_t N) { return nb_div_up(N, 4); }
// per-ray counts [N + 8] then 32-ray group sums and prefixes [2 x (N / 32 + 8)]  (serial kernel: 2 x nb4 block sums / prefixes)
static inline uint32_t march_groups(uint32_t N) { return nb_div_up(N, 32) + 8; }
static inline uint32_t march_hdr(uint32_t N) { return N + 8 + 2 * march_groups(N); }
uint32_t nb200_march_scratch_ints(uint32_t N) { return march_hdr(N) + kScratchFixed + N * march_tcap(N); }
static inline int32_t *march_obounds(int32_t *scratch, uint32_t N) { return scratch + march_hdr(N); }
static inline float *march_trec(int32_t *scratch, uint32_t N) {
    return reinterpret_cast<float *>(scratch + march_hdr(N) + kScratchFixed);
}

static int march_count_impl(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                            float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                            float *nears, float *fars, const float *noises,
                            int32_t *rays, int32_t *counter, int32_t *scratch, uint32_t M_cap, int32_t *m_eff,
                            const float *aabb, float min_near, void *stream) {
    if (N == 0) return 0;
    if (!scratch || !rays || !counter) return NB200_E_BAD_ARG;
    static int serial_env = -1;
    if (serial_env < 0) { const char *e = getenv("NB200_MARCH_SERIAL"); serial_env = (e && atoi(e)) ? 1 : 0; }
    const bool seg = dt_gamma == 0.0f && (H & (H - 1)) == 0 && !serial_env;   // k_march_count_seg's compile-time fast paths
    const uint32_t rpw = seg ? (uint32_t)kSegWarps : march_rpw(N);
    const uint32_t nb = nb_div_up(N, rpw);
    int32_t *block_sums = scratch, *block_prefix = scratch + march_nb_max(N);
    cudaStream_t st = nb_stream(stream);
    if (seg) {
        static int nocull_env = -1;
        if (nocull_env < 0) { const char *e = getenv("NB200_MARCH_CULL"); nocull_env = (e && atoi(e)) ? 0 : 1; }
        int32_t *obounds = nullptr;
        if (!nocull_env && C <= 8 && (H * H * H) % 32 == 0 && (reinterpret_cast<uintptr_t>(grid) & 3u) == 0) {
            obounds = march_obounds(scratch, N);
            cudaError_t e = cudaMemsetAsync(obounds, 0, 48 * sizeof(int32_t), st);
            if (e != cudaSuccess) return (int)e;
            k_occ_bounds<<<148, 256, 0, st>>>(grid, C, H, obounds);
            NB_LAUNCH_CHECK();
        }
        static int lanes_env = -1;
        if (lanes_env < 0) { const char *e = getenv("NB200_MARCH_LANES"); lanes_env = e ? atoi(e) : 0; }
        uint32_t lanes = (lanes_env == 4 || lanes_env == 8 || lanes_env == 16 || lanes_env == 32) ? (uint32_t)lanes_env : 8u;
        while (lanes < 32 && (uint64_t)N * lanes < 148ull * 4 * 32 * 4) lanes <<= 1;     // few rays: keep >= ~4 warps per scheduler
        const uint32_t rays_per_block = kSegWarps * 32 / lanes, nbk = nb_div_up(N, rays_per_block);
        int32_t *gsum = scratch + N + 8, *gprefix = gsum + march_groups(N);
        {
            cudaError_t e = cudaMemsetAsync(gsum, 0, (size_t)march_groups(N) * sizeof(int32_t), st);
            if (e != cudaSuccess) return (int)e;
        }
        static int chunks_env = -1;
        if (chunks_env < 0) { const char *e = getenv("NB200_MARCH_CHUNKS"); chunks_env = e ? atoi(e) : 0; }
        const int chunks = (chunks_env == 1 || chunks_env == 2 || chunks_env == 4) ? chunks_env : 2;   // 2 measured best (profiles/)
#define NB_SEG_LAUNCH(L, R) k_march_count_seg<L, R><<<nbk, kSegWarps * 32, 0, st>>>(rays_o, rays_d, grid, bound, max_steps, N, C, H, \
        nears, fars, noises, scratch, gsum, march_trec(scratch, N), march_tcap(N), obounds, aabb, min_near)
#define NB_SEG_LANES(R) do { if (lanes == 8) NB_SEG_LAUNCH(8, R); else if (lanes == 16) NB_SEG_LAUNCH(16, R); else NB_SEG_LAUNCH(32, R); } while (0)
        if (lanes == 4) lanes = 8;
        if (chunks == 1) NB_SEG_LANES(1);
        else if (chunks == 2) NB_SEG_LANES(2);
        else NB_SEG_LANES(4);
#undef NB_SEG_LANES
#undef NB_SEG_LAUNCH
        NB_LAUNCH_CHECK();
        k_march_scan<<<1, 1024, 0, st>>>(gsum, gprefix, nb_div_up(N, 32), N, counter, M_cap, m_eff);
        NB_LAUNCH_CHECK();
        k_march_finalize<<<nb_div_up(N, 256), 256, 0, st>>>(scratch, gprefix, rays, N);
        NB_LAUNCH_CHECK();
        return 0;
    }
    if (aabb) {             // the serial kernel reads nears / fars: produce them first
        k_near_far_from_aabb<<<nb_div_up(N, 128), 128, 0, st>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
        NB_LAUNCH_CHECK();
    }
    k_march_count<<<nb, kMarchBlock, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars,
                                              noises, rays, block_sums, rpw, march_trec(scratch, N), march_tcap(N));
    NB_LAUNCH_CHECK();
    k_march_scan<<<1, 1024, 0, st>>>(block_sums, block_prefix, nb, N, counter, M_cap, m_eff);
    NB_LAUNCH_CHECK();
    k_march_fixup<<<nb_div_up(N, 256), 256, 0, st>>>(rays, block_prefix, N, rpw);
    NB_LAUNCH_CHECK();
    return 0;
}

static int march_expand_impl(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                             const float *fars, const float *noises, const int32_t *rays, float *xyzs, float *dirs,
                             float *deltas, int32_t *m_eff, const int32_t *scratch, void *stream) {
    if (N == 0) return 0;
    const float *trec = scratch ? march_trec(const_cast<int32_t *>(scratch), N) : nullptr;
    k_march_expand<<<nb_div_up((uint64_t)N * 32, kExpandBlock), kExpandBlock, 0, nb_stream(stream)>>>(
        rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, noises, rays, xyzs, dirs, deltas,
        m_eff, trec, march_tcap(N));
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_march_rays_train_count(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                 const float *nears, const float *fars, const float *noises,
                                 int32_t *rays, int32_t *counter, int32_t *scratch, void *stream) {
    return march_count_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, const_cast<float *>(nears),
                            const_cast<float *>(fars), noises, rays, counter, scratch, 0, nullptr, nullptr, 0.0f, stream);
}

// fused train step: count + scan, then expand, into buffers of M_cap rows; *m_eff = rows covered by complete segments
int nb200_fs_march_count(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M_cap, float *nears,
                         float *fars, const float *noises, int32_t *rays, int32_t *counter, int32_t *m_eff,
                         int32_t *scratch, const float *aabb, float min_near, void *stream) {
    if (N == 0) return 0;
    if (!m_eff || !nears || !fars) return NB200_E_BAD_ARG;
    return march_count_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays,
                            counter, scratch, M_cap, m_eff, aabb, min_near, stream);
}

int nb200_fs_march_write(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M_cap, const float *nears,
                         const float *fars, const float *noises, const int32_t *rays, float *xyzs, float *dirs,
                         float *deltas, int32_t *m_eff, const int32_t *scratch, void *stream) {
    if (N == 0) return 0;
    if (!m_eff) return NB200_E_BAD_ARG;
    return march_expand_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M_cap, nears, fars, noises, rays,
                             xyzs, dirs, deltas, m_eff, scratch, stream);
}

// scratch: the buffer nb200_march_rays_train_count filled for the same rays (its sample records), or NULL to re-march
int nb200_march_rays_train_write(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                 const float *nears, const float *fars, const float *noises, const int32_t *rays,
                                 float *xyzs, float *dirs, float *deltas, const int32_t *scratch, void *stream) {
    return march_expand_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, noises, rays,
                             xyzs, dirs, deltas, nullptr, scratch, stream);
}

int nb200_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                           uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                           const float *fars, float *xyzs, float *dirs, float *deltas, int32_t *rays, int32_t *counter,
                           const float *noises, int32_t *scratch, void *stream) {
    int rc = nb200_march_rays_train_count(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises,
                                          rays, counter, scratch, stream);
    if (rc) return rc;
    return nb200_march_rays_train_write(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars,
                                        noises, rays, xyzs, dirs, deltas, scratch, stream);
}

int nb200_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                                       uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *depth,
                                       float *image, void *stream) {
    if (N == 0) return 0;
    k_composite_train_fwd<float><<<nb_div_up((uint64_t)N * 32, kCompBlock), kCompBlock, 0, nb_stream(stream)>>>(
        sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image, MseArgs{nullptr, nullptr, nullptr, 0.0f, 0.0f, nullptr, nullptr, nullptr, 0.0f});
    NB_LAUNCH_CHECK();
    return 0;
}

// fused train step: colours are the field kernel's half [M,4] rows; grad_rgba rows are float4 [g_r, g_g, g_b, 0]
int nb200_fs_composite_forward(const float *sigmas, const void *rgba, const float *deltas, const int32_t *rays,
                               uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *depth, float *image,
                               const float *target, float inv_n, float loss_scale, float *loss, float *g_image,
                               const float *target_mask, float mask_weight, float *render_mask, float *g_render_mask,
                               const float *loss_scale_dev, void *stream) {
    if (N == 0) return 0;
    if (target && (!loss || !g_image)) return NB200_E_BAD_ARG;
    if (target_mask && (!target || !render_mask || !g_render_mask)) return NB200_E_BAD_ARG;
    k_composite_train_fwd<__half><<<nb_div_up((uint64_t)N * 32, kCompBlock), kCompBlock, 0, nb_stream(stream)>>>(
        sigmas, (const __half *)rgba, deltas, rays, M, N, T_thresh, weights_sum, depth, image,
        MseArgs{target, loss, g_image, inv_n, loss_scale, target_mask, render_mask, g_render_mask, mask_weight, loss_scale_dev});
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_fs_composite_backward(const float *grad_weights_sum, const float *grad_image, const float *sigmas,
                                const void *rgba, const float *deltas, const int32_t *rays, const float *weights_sum,
                                const float *image, uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas,
                                float *grad_rgba, const float *grad_render_mask, const float *render_mask, void *stream) {
    if (N == 0) return 0;
    if (grad_render_mask && !render_mask) return NB200_E_BAD_ARG;
    k_composite_train_bwd<__half, 4><<<nb_div_up((uint64_t)N * 32, kCompBlock), kCompBlock, 0, nb_stream(stream)>>>(
        grad_weights_sum, grad_image, sigmas, (const __half *)rgba, deltas, rays, weights_sum, image, M, N, T_thresh,
        grad_sigmas, grad_rgba, grad_render_mask, render_mask);
    NB_LAUNCH_CHECK();
    return 0;
}

// compositing forward + MSE + compositing backward of the fused train step in one launch (same outputs as the two calls)
int nb200_fs_composite_fused(const float *sigmas, const void *rgba, const float *deltas, const int32_t *rays, uint32_t M,
                             uint32_t N, float T_thresh, float *weights_sum, float *depth, float *image, const float *target,
                             float inv_n, float loss_scale, float *loss, float *g_image, const float *target_mask,
                             float mask_weight, float *render_mask, float *g_render_mask, const float *loss_scale_dev,
                             const float *grad_weights_sum, float *grad_sigmas, float *grad_rgba, void *stream) {
    if (N == 0) return 0;
    if (!target || !loss || !g_image || !grad_sigmas || !grad_rgba) return NB200_E_BAD_ARG;
    if (target_mask && (!render_mask || !g_render_mask)) return NB200_E_BAD_ARG;
    k_composite_train_fused<__half><<<nb_div_up((uint64_t)N * 32, kCompBlock), kCompBlock, 0, nb_stream(stream)>>>(
        sigmas, (const __half *)rgba, deltas, rays, M, N, T_thresh, weights_sum, depth, image,
        MseArgs{target, loss, g_image, inv_n, loss_scale, target_mask, render_mask, g_render_mask, mask_weight, loss_scale_dev},
        grad_weights_sum, grad_sigmas, grad_rgba);
    NB_LAUNCH_CHECK();
    return 0;
}

// LGIE composites of the fused editing step: variant 0 = all, 1 = fg, 2 = bg (see LgieArgs).  Per-ray outputs of the
// forward and per-ray gradient inputs of the backward are per variant; the backward of variant 0 writes the per-sample rows,
// variants 1 and 2 accumulate into them (launch them in that order on one stream).
int nb200_fs_composite_lgie_forward(int variant, const float *sigmas, const void *rgba, const float *deltas,
                                    const int32_t *rays, uint32_t M, uint32_t N, float T_thresh, float conf_thr,
                                    int soft_mask, float *weights_sum, float *depth, float *image, float *render_mask,
                                    void *stream) {
    if (N == 0) return 0;
    if (!render_mask || variant < 0 || variant > 2) return NB200_E_BAD_ARG;
    const MseArgs mse{nullptr, nullptr, nullptr, 0.0f, 0.0f, nullptr, render_mask, nullptr, 0.0f};
    const LgieArgs lg{conf_thr, soft_mask, 0, 0, 0};
    const uint32_t grid = nb_div_up((uint64_t)N * 32, kCompBlock);
    cudaStream_t st = nb_stream(stream);
#define NB_LGIE_FWD(V) k_composite_train_fwd<__half, V><<<grid, kCompBlock, 0, st>>>(sigmas, (const __half *)rgba, deltas, rays, \
        M, N, T_thresh, weights_sum, depth, image, mse, lg)
    if (variant == 0) NB_LGIE_FWD(0); else if (variant == 1) NB_LGIE_FWD(1); else NB_LGIE_FWD(2);
#undef NB_LGIE_FWD
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_fs_composite_lgie_backward(int variant, const float *grad_weights_sum, const float *grad_image,
                                     const float *grad_render_mask, const float *sigmas, const void *rgba,
                                     const float *deltas, const int32_t *rays, const float *weights_sum, const float *image,
                                     const float *render_mask, uint32_t M, uint32_t N, float T_thresh, float conf_thr,
                                     int soft_mask, int detach_bg, int detach_mask_from_field, float *grad_sigmas,
                                     float *grad_rgba, void *stream) {
    if (N == 0) return 0;
    if (!grad_render_mask || !render_mask || variant < 0 || variant > 2) return NB200_E_BAD_ARG;
    const LgieArgs lg{conf_thr, soft_mask, detach_bg, detach_mask_from_field, variant != 0};
    const uint32_t grid = nb_div_up((uint64_t)N * 32, kCompBlock);
    cudaStream_t st = nb_stream(stream);
#define NB_LGIE_BWD(V) k_composite_train_bwd<__half, 4, V><<<grid, kCompBlock, 0, st>>>(grad_weights_sum, grad_image, sigmas, \
        (const __half *)rgba, deltas, rays, weights_sum, image, M, N, T_thresh, grad_sigmas, grad_rgba, grad_render_mask,    \
        render_mask, lg)
    if (variant == 0) NB_LGIE_BWD(0); else if (variant == 1) NB_LGIE_BWD(1); else NB_LGIE_BWD(2);
#undef NB_LGIE_BWD
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image, const float *sigmas,
                                        const float *rgbs, const float *deltas, const int32_t *rays,
                                        const float *weights_sum, const float *image, uint32_t M, uint32_t N,
                                        float T_thresh, float *grad_sigmas, float *grad_rgbs, void *stream) {
    if (N == 0) return 0;
    k_composite_train_bwd<float, 3><<<nb_div_up((uint64_t)N * 32, kCompBlock), kCompBlock, 0, nb_stream(stream)>>>(
        grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh, grad_sigmas,
        grad_rgbs, nullptr, nullptr);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                     const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                     uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                     float *xyzs, float *dirs, float *deltas, const float *noises, void *stream) {
    (void)nears;
    if (n_alive == 0) return 0;
    k_march_rays<<<nb_div_up(n_alive, 128), 128, 0, nb_stream(stream)>>>(n_alive, n_step, rays_alive, rays_t, rays_o,
                                                                        rays_d, bound, dt_gamma, max_steps, C, H, grid,
                                                                        fars, xyzs, dirs, deltas, noises);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t,
                         const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum, float *depth,
                         float *image, void *stream) {
    if (n_alive == 0) return 0;
    k_composite_rays<<<nb_div_up(n_alive, 128), 128, 0, nb_stream(stream)>>>(n_alive, n_step, T_thresh, rays_alive,
                                                                            rays_t, sigmas, rgbs, deltas, weights_sum,
                                                                            depth, image);
    NB_LAUNCH_CHECK();
    return 0;
}

// ---- device-driven inference rounds (see k_infer_plan) --------------------------------------------------------------
int nb200_infer_plan(int32_t *state, uint32_t N, uint32_t max_steps, void *stream) {
    if (!state) return NB200_E_BAD_ARG;
    k_infer_plan<<<1, 32, 0, nb_stream(stream)>>>(state, N, max_steps);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_march_rays_dev(const int32_t *state, uint32_t N, const int32_t *rays_alive, const float *rays_t,
                         const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                         uint32_t C, uint32_t H, const uint8_t *grid, const float *fars, float *xyzs, float *dirs,
                         float *deltas, const float *noises, void *stream) {
    if (N == 0) return 0;
    if (!state) return NB200_E_BAD_ARG;
    k_march_rays_dev<<<nb_div_up(N, 128), 128, 0, nb_stream(stream)>>>(state, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma,
                                                                       max_steps, C, H, grid, fars, xyzs, dirs, deltas, noises);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_march_rays_rec(int32_t *state, uint32_t N, const int32_t *rays_alive, const float *rays_t, const float *rays_o,
                         const float *rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                         const uint8_t *grid, const float *fars, float *xyzs, float *dirs, float *deltas, const float *noises,
                         const int32_t *rays, const int32_t *scratch, int32_t *consumed, void *stream) {
    if (N == 0) return 0;
    if (!state || !rays || !scratch || !consumed) return NB200_E_BAD_ARG;
    k_march_rays_rec<<<nb_div_up(N, 128), 128, 0, nb_stream(stream)>>>(state, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma,
                                                                       max_steps, C, H, grid, fars, xyzs, dirs, deltas, noises, rays,
                                                                       march_trec(const_cast<int32_t *>(scratch), N), march_tcap(N),
                                                                       consumed);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_composite_rays_dev(const int32_t *state, uint32_t N, float T_thresh, int32_t *rays_alive, float *rays_t,
                             const float *sigmas, const void *rgba, const float *deltas, float *weights_sum, float *depth,
                             float *image, void *stream) {
    if (N == 0) return 0;
    if (!state) return NB200_E_BAD_ARG;
    k_composite_rays_dev<__half><<<nb_div_up(N, 128), 128, 0, nb_stream(stream)>>>(state, T_thresh, rays_alive, rays_t, sigmas,
                                                                                  (const __half *)rgba, deltas, weights_sum,
                                                                                  depth, image);
    NB_LAUNCH_CHECK();
    return 0;
}

int nb200_compact_alive(int32_t *state, uint32_t N, int32_t *rays_alive, int32_t *tmp, void *stream) {
    if (N == 0) return 0;
    if (!state || !rays_alive || !tmp) return NB200_E_BAD_ARG;
    k_compact_alive<<<nb_div_up(N, 256), 256, 0, nb_stream(stream)>>>(state, rays_alive, tmp);
    NB_LAUNCH_CHECK();
    k_copy_alive<<<nb_div_up(N, 256), 256, 0, nb_stream(stream)>>>(state, tmp, rays_alive);
    NB_LAUNCH_CHECK();
    k_infer_commit<<<1, 32, 0, nb_stream(stream)>>>(state);
    NB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
