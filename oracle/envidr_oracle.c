/*
 * envidr_oracle.c -- CPU restatement of the ENVIDR volumetric-render hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under envidr_b200/ (the product) may link,
 * import or call this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Every function restates, in scalar fp32 C, what the reference's CUDA kernel
 * computes, citing the reference file:line it follows (paths relative to the
 * reference checkout).  Parity status: the reference ships no tests or golden
 * vectors for these kernels, so the oracle is pinned against the
 * reference kernels themselves, rebuilt unmodified for sm_100a into oracle/_ref/
 * (oracle/build_ref.py) and compared on the GPU box by tests/test_oracle_c.py / tests/test_gpu_ops.py.
 *
 * Floating-point contract.  nvcc contracts a*b+c into one FMA (--fmad=true); gcc is
 * run with -ffp-contract=off and the contractions nvcc performs on the march path
 * are written out as fmaf() so that sample positions, counts and occupancy decisions
 * are bit-identical.  __expf/__sinf (GPU fast intrinsics) have no CPU twin: composite
 * and freq outputs are therefore compared with a tolerance, never bit-wise.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC envidr_oracle.c -o liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* small helpers (raymarching/src/raymarching.cu:30-81)                      */
/* ------------------------------------------------------------------------- */
static inline float sgn1(float x) { return copysignf(1.0f, x); }
static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

/* raymarching.cu:56-71: 10-bit-per-axis bit interleave */
static inline uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
/* raymarching.cu:73-81 */
static inline uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* raymarching.cu:42-54: cascade level from position / from step size.
 * The float fminf/fmaxf round trip of the reference is exact for these small ints. */
static inline int level_from_pos(float x, float y, float z, int cascades) {
    float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e; frexpf(mx, &e);
    return (int)fminf((float)(cascades - 1), fmaxf(0.0f, (float)e));
}
static inline int level_from_dt(float dt, float H, int cascades) {
    float mx = dt * H * 0.5f;               /* double literal in the reference; exact for any H */
    int e; frexpf(mx, &e);
    return (int)fminf((float)(cascades - 1), fmaxf(0.0f, (float)e));
}

ORC_API void orc_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
    for (uint32_t n = 0; n < N; n++)   /* raymarching.cu:214-226 */
        indices[n] = (int32_t)morton3((uint32_t)coords[3*n], (uint32_t)coords[3*n+1], (uint32_t)coords[3*n+2]);
}
ORC_API void orc_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
    for (uint32_t n = 0; n < N; n++) { /* raymarching.cu:237-254: note arithmetic >> on int */
        int32_t ind = indices[n];
        coords[3*n+0] = (int32_t)compact3((uint32_t)(ind >> 0));
        coords[3*n+1] = (int32_t)compact3((uint32_t)(ind >> 1));
        coords[3*n+2] = (int32_t)compact3((uint32_t)(ind >> 2));
    }
}

/* raymarching.cu:91-145 -- slab test against aabb[6] */
ORC_API void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                                    uint32_t N, float min_near, float* nears, float* fars) {
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[3*n], oy = rays_o[3*n+1], oz = rays_o[3*n+2];
        const float rdx = 1.0f / rays_d[3*n], rdy = 1.0f / rays_d[3*n+1], rdz = 1.0f / rays_d[3*n+2];
        float tn = (aabb[0] - ox) * rdx, tf = (aabb[3] - ox) * rdx, s;
        if (tn > tf) { s = tn; tn = tf; tf = s; }
        float yn = (aabb[1] - oy) * rdy, yf = (aabb[4] - oy) * rdy;
        if (yn > yf) { s = yn; yn = yf; yf = s; }
        if (tn > yf || yn > tf) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (yn > tn) tn = yn;
        if (yf < tf) tf = yf;
        float zn = (aabb[2] - oz) * rdz, zf = (aabb[5] - oz) * rdz;
        if (zn > zf) { s = zn; zn = zf; zf = s; }
        if (tn > zf || zn > tf) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (zn > tn) tn = zn;
        if (zf < tf) tf = zf;
        if (tn < min_near) tn = min_near;
        nears[n] = tn; fars[n] = tf;
    }
}

/* raymarching.cu:163-198 -- far intersection with the background sphere -> (theta, phi) in [-1,1]^2.
 * Transcendentals differ in the last ulp between libm and CUDA: tolerance-compared. */
ORC_API void orc_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[3*n], oy = rays_o[3*n+1], oz = rays_o[3*n+2];
        const float dx = rays_d[3*n], dy = rays_d[3*n+1], dz = rays_d[3*n+2];
        const float A = dx*dx + dy*dy + dz*dz;
        const float B = ox*dx + oy*dy + oz*dz;
        const float C = ox*ox + oy*oy + oz*oz - radius*radius;
        const float t = (-B + sqrtf(B*B - A*C)) / A;
        const float x = ox + t*dx, y = oy + t*dy, z = oz + t*dz;
        coords[2*n]   = 2 * atan2f(sqrtf(x*x + z*z), y) * RPI - 1;
        coords[2*n+1] = atan2f(z, x) * RPI;
    }
}

/* raymarching.cu:267-289 -- one output byte per 8 consecutive cells, bit i = grid[8n+i] > thresh */
ORC_API void orc_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; n++) {
        uint8_t b = 0;
        for (int i = 0; i < 8; i++) if (grid[8u*n + i] > thresh) b |= (uint8_t)(1u << i);
        bitfield[n] = b;
    }
}

/* raymarching.cu:302-322 */
ORC_API void orc_get_scatter_idx(const int32_t* rays, uint32_t N, int32_t* idx_map) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t id = rays[3*n], off = rays[3*n+1], cnt = rays[3*n+2];
        for (uint32_t s = 0; s < cnt; s++) idx_map[off + s] = (int32_t)id;
    }
}

/* ------------------------------------------------------------------------- */
/* occupancy-grid DDA shared by the train and inference march                */
/* ------------------------------------------------------------------------- */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH, Hf, far;
    uint32_t C, H, H3;
    const uint8_t* grid;
} dda_t;

static void dda_init(dda_t* s, const float* o, const float* d, float bound, float dt_gamma,
                     uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, float far) {
    s->ox = o[0]; s->oy = o[1]; s->oz = o[2];
    s->dx = d[0]; s->dy = d[1]; s->dz = d[2];
    s->rdx = 1.0f / s->dx; s->rdy = 1.0f / s->dy; s->rdz = 1.0f / s->dz;
    s->bound = bound; s->dt_gamma = dt_gamma;
    s->Hf = (float)H; s->rH = 1.0f / (float)H;
    s->H3 = (uint32_t)((float)(H * H * H));         /* raymarching.cu:368,873: float H3 */
    s->dt_min = 2 * 1.7320508075688772f / (float)max_steps;                 /* :374,878 */
    s->dt_max = 2 * 1.7320508075688772f * (float)(1 << (C - 1)) / (float)H; /* :375,879 */
    s->C = C; s->H = H; s->grid = grid; s->far = far;
}

/* One DDA decision at parameter *t (raymarching.cu:889-942 / 388-428 / 456-507).
 * Returns 1 and fills xyz/dt when the cell is occupied (caller advances t by dt);
 * returns 0 after having advanced *t past the empty cell. */
static inline int dda_step(const dda_t* s, float* t, float xyz[3], float* dt_out) {
    const float tt0 = *t;
    const float x = clampf(fmaf(tt0, s->dx, s->ox), -s->bound, s->bound);
    const float y = clampf(fmaf(tt0, s->dy, s->oy), -s->bound, s->bound);
    const float z = clampf(fmaf(tt0, s->dz, s->oz), -s->bound, s->bound);
    const float dt = clampf(tt0 * s->dt_gamma, s->dt_min, s->dt_max);
    int lv_p = level_from_pos(x, y, z, (int)s->C), lv_d = level_from_dt(dt, s->Hf, (int)s->C);
    const int level = lv_p > lv_d ? lv_p : lv_d;
    const float mip_bound = fminf(scalbnf(1.0f, level), s->bound);
    const float mip_rbound = 1.0f / mip_bound;
    /* 0.5 * (x * r + 1) * H: the double promotion of the reference is exact in fp32 (power-of-two factors) */
    const int nx = (int)clampf(0.5f * fmaf(x, mip_rbound, 1.0f) * s->Hf, 0.0f, (float)(s->H - 1));
    const int ny = (int)clampf(0.5f * fmaf(y, mip_rbound, 1.0f) * s->Hf, 0.0f, (float)(s->H - 1));
    const int nz = (int)clampf(0.5f * fmaf(z, mip_rbound, 1.0f) * s->Hf, 0.0f, (float)(s->H - 1));
    const uint32_t index = (uint32_t)((float)level * (float)s->H3) + morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const int occ = s->grid[index / 8] & (1 << (index % 8));
    if (occ) {
        xyz[0] = x; xyz[1] = y; xyz[2] = z; *dt_out = dt;
        return 1;
    }
    /* distance to the exit face of this voxel (raymarching.cu:934-937) */
    const float tx = fmaf(((nx + 0.5f + 0.5f * sgn1(s->dx)) * s->rH * 2 - 1), mip_bound, -x) * s->rdx;
    const float ty = fmaf(((ny + 0.5f + 0.5f * sgn1(s->dy)) * s->rH * 2 - 1), mip_bound, -y) * s->rdy;
    const float tz = fmaf(((nz + 0.5f + 0.5f * sgn1(s->dz)) * s->rH * 2 - 1), mip_bound, -z) * s->rdz;
    const float tt = tt0 + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    float tc = tt0;
    do { tc += clampf(tc * s->dt_gamma, s->dt_min, s->dt_max); } while (tc < tt);
    *t = tc;
    return 0;
}

/* raymarching.cu:340-509.  The reference allocates output slots with two atomicAdds, so
 * its rays[:,1] offsets depend on thread scheduling.  This restatement visits rays in
 * index order (slot r = ray r, offsets = running sum); parity is defined on
 * ray id -> (count, per-ray sample sequence), see SURVEY.md 8a-3. */
ORC_API void orc_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
        float bound, float dt_gamma, uint32_t max_steps, uint32_t early_stop_steps,
        uint32_t N, uint32_t C, uint32_t H, uint32_t M,
        const float* nears, const float* fars,
        float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter, const float* noises) {
    for (uint32_t n = 0; n < N; n++) {
        dda_t s; dda_init(&s, rays_o + 3*n, rays_d + 3*n, bound, dt_gamma, max_steps, C, H, grid, fars[n]);
        const float near = nears[n];
        float t0 = near;
        t0 = fmaf(clampf(t0 * dt_gamma, s.dt_min, s.dt_max), noises[n], t0);  /* :380 */
        float t = t0, xyz[3], dt; uint32_t num_steps = 0;
        while (t < s.far && num_steps < early_stop_steps) {                  /* :388 count pass */
            if (dda_step(&s, &t, xyz, &dt)) { num_steps++; t += dt; }
        }
        const uint32_t point_index = (uint32_t)counter[0]; counter[0] += (int32_t)num_steps;
        const uint32_t ray_index = (uint32_t)counter[1];   counter[1] += 1;
        rays[3*ray_index] = (int32_t)n; rays[3*ray_index+1] = (int32_t)point_index; rays[3*ray_index+2] = (int32_t)num_steps;
        if (num_steps == 0 || point_index + num_steps > M) continue;          /* :444-445 */
        t = t0; uint32_t step = 0; float last_t = near;                      /* :451-454 */
        float* px = xyzs + 3*(size_t)point_index; float* pd = dirs + 3*(size_t)point_index; float* pl = deltas + 2*(size_t)point_index;
        while (t < s.far && step < num_steps) {
            if (dda_step(&s, &t, xyz, &dt)) {
                px[0] = xyz[0]; px[1] = xyz[1]; px[2] = xyz[2];
                pd[0] = s.dx; pd[1] = s.dy; pd[2] = s.dz;
                t += dt; pl[0] = dt; pl[1] = t - last_t; last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* raymarching.cu:839-944.  Fixed slot layout: ray slot n owns samples [n*n_step, (n+1)*n_step);
 * slots past the end of the ray keep the caller's zero fill.  counts (optional) = samples per slot. */
ORC_API void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
        const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
        uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars,
        float* xyzs, float* dirs, float* deltas, const float* noises, int32_t* counts) {
    (void)nears;
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        dda_t s; dda_init(&s, rays_o + 3*(size_t)index, rays_d + 3*(size_t)index, bound, dt_gamma, max_steps, C, H, grid, fars[index]);
        float t = rays_t[index], last_t = t, xyz[3], dt;
        t = fmaf(clampf(t * dt_gamma, s.dt_min, s.dt_max), noises[n], t);     /* :887 */
        float* px = xyzs + 3*(size_t)n*n_step; float* pd = dirs + 3*(size_t)n*n_step; float* pl = deltas + 2*(size_t)n*n_step;
        uint32_t step = 0;
        while (t < s.far && step < n_step) {
            if (dda_step(&s, &t, xyz, &dt)) {
                px[0] = xyz[0]; px[1] = xyz[1]; px[2] = xyz[2];
                pd[0] = s.dx; pd[1] = s.dy; pd[2] = s.dz;
                t += dt; pl[0] = dt; pl[1] = t - last_t; last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
        if (counts) counts[n] = (int32_t)step;
    }
}

/* ------------------------------------------------------------------------- */
/* alpha compositing                                                          */
/* ------------------------------------------------------------------------- */
/* raymarching.cu:529-608 / 618-701.  weights may be NULL (no-weights kernel). */
ORC_API void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
        const int32_t* rays, uint32_t M, uint32_t N, float T_thresh, uint32_t accum_deltas, uint32_t input_alpha,
        float* weights_sum, float* depth, float* image, float* weights) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = rays[3*n], offset = rays[3*n+1], num_steps = rays[3*n+2];
        if (num_steps == 0 || offset + num_steps > M) {
            weights_sum[index] = 0; depth[index] = 0;
            image[3*index] = image[3*index+1] = image[3*index+2] = 0;
            continue;
        }
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        for (uint32_t s = 0; s < num_steps; s++) {
            const uint32_t i = offset + s;
            const float alpha = input_alpha ? 0.0f + sigmas[i] : 1.0f - expf(-sigmas[i] * deltas[2*i]);
            const float w = alpha * T;
            if (weights) weights[i] = w;
            r += w * rgbs[3*i]; g += w * rgbs[3*i+1]; b += w * rgbs[3*i+2];
            t = accum_deltas ? t + deltas[2*i+1] : deltas[2*i+1];
            d += w * t;
            ws += w;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;                                     /* :585-588: test AFTER the update */
        }
        weights_sum[index] = ws; depth[index] = d;
        image[3*index] = r; image[3*index+1] = g; image[3*index+2] = b;
    }
}

/* raymarching.cu:731-821.  Reference quirk kept: depth / grad_depth are NOT offset by the
 * ray index (lines 760-768, 774, 804), i.e. every ray reads element 0. */
ORC_API void orc_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* grad_depth,
        const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
        const float* weights_sum, const float* image, const float* depth,
        uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas, float* grad_rgbs,
        uint32_t accum_deltas, uint32_t input_alpha) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = rays[3*n], offset = rays[3*n+1], num_steps = rays[3*n+2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float* gi = grad_image + 3*index;
        const float gws = grad_weights_sum[index];
        const float r_final = image[3*index], g_final = image[3*index+1], b_final = image[3*index+2];
        const float ws_final = weights_sum[index], d_final = depth[0];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        for (uint32_t s = 0; s < num_steps; s++) {
            const uint32_t i = offset + s;
            const float alpha = input_alpha ? 0.0f + sigmas[i] : 1.0f - expf(-sigmas[i] * deltas[2*i]);
            const float w = alpha * T;
            const float gscale = input_alpha ? (1.0f / (1.0f - alpha + 1e-4f)) : deltas[2*i];
            r += w * rgbs[3*i]; g += w * rgbs[3*i+1]; b += w * rgbs[3*i+2];
            ws += w;
            t = accum_deltas ? t + deltas[2*i+1] : deltas[2*i+1];
            d += w * t;
            T *= 1.0f - alpha;
            grad_rgbs[3*i] = gi[0] * w; grad_rgbs[3*i+1] = gi[1] * w; grad_rgbs[3*i+2] = gi[2] * w;
            grad_sigmas[i] = gscale * (
                gi[0] * (T * rgbs[3*i]   - (r_final - r)) +
                gi[1] * (T * rgbs[3*i+1] - (g_final - g)) +
                gi[2] * (T * rgbs[3*i+2] - (b_final - b)) +
                grad_depth[0] * (T * t - (d_final - d)) +
                gws * (1 - ws_final));
            if (T < T_thresh) break;
        }
    }
}

/* raymarching.cu:957-1046.  In-place update of per-ray accumulators; a ray dies
 * (rays_alive[n] = -1) when it meets a zero-delta slot or when the transmittance
 * it had BEFORE the sample just accumulated was already below T_thresh. */
ORC_API void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, uint32_t accum_deltas, uint32_t input_alpha,
        int32_t* rays_alive, float* rays_t, const float* sigmas, const float* rgbs, const float* deltas,
        float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float* sg = sigmas + (size_t)n*n_step; const float* cl = rgbs + 3*(size_t)n*n_step; const float* dl = deltas + 2*(size_t)n*n_step;
        float t = rays_t[index], ws = weights_sum[index], d = depth[index];
        float r = image[3*index], g = image[3*index+1], b = image[3*index+2];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = input_alpha ? 0.0f + sg[0] : 1.0f - expf(-sg[0] * dl[0]);
            const float T = 1 - ws;
            const float w = alpha * T;
            ws += w;
            t = accum_deltas ? t + dl[1] : dl[1];
            d += w * t; r += w * cl[0]; g += w * cl[1]; b += w * cl[2];
            if (T < T_thresh) break;
            sg++; cl += 3; dl += 2; step++;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = ws; depth[index] = d;
        image[3*index] = r; image[3*index+1] = g; image[3*index+2] = b;
    }
}

/* ------------------------------------------------------------------------- */
/* multi-resolution grid encoders                                             */
/* ------------------------------------------------------------------------- */
#define MAXD 5
#define MAXC 8

typedef struct {
    int D, C, smooth;        /* smooth=1: hashencoder (smoothstep, no +0.5, stride res);
                                smooth=0: gridencoder (linear, +0.5 unless align_corners, stride res+1) */
    int gridtype, align_corners;
} enc_cfg_t;

/* hashencoder.cu:36-72 / gridencoder.cu:36-72.  uint32 wrap-around is part of the hash. */
static inline uint32_t cell_index(const enc_cfg_t* cf, uint32_t hashmap_size, uint32_t resolution, const uint32_t* pg) {
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t stride = 1, index = 0;
    for (int d = 0; d < cf->D && stride <= hashmap_size; d++) {
        index += pg[d] * stride;
        stride *= cf->smooth ? resolution : (cf->align_corners ? resolution : resolution + 1);
    }
    if (stride > hashmap_size && (cf->smooth || cf->gridtype == 0)) {
        uint32_t h = 0;
        for (int d = 0; d < cf->D; d++) h ^= pg[d] * primes[d];
        index = h;
    }
    return index % hashmap_size;
}

/* Optional override of the per-level `scale`.  The reference evaluates exp2f(level*S) with the GPU's ex2.approx
 * instruction (<= 2 ulp from libm's exp2f); one ulp of scale moves the fractional cell position by up to 1.2e-4
 * at the finest level.  Tests that compare at 1e-6 read the device's values (envidr_debug_level_scales) and
 * install them here; without an override libm's exp2f is used. */
static float g_level_scales[64];
static int g_n_level_scales = 0;
ORC_API void orc_set_level_scales(const float* scales, int n) {
    g_n_level_scales = (scales && n > 0 && n <= 64) ? n : 0;
    for (int i = 0; i < g_n_level_scales; i++) g_level_scales[i] = scales[i];
}

/* per-level geometry (hashencoder.cu:151-167, gridencoder.cu:124-138): returns 0 when x is out of [0,1]^D */
static inline int level_setup(const enc_cfg_t* cf, const float* x, int level, float S, uint32_t H,
                              float* scale_out, uint32_t* res_out, float* w, float* dw, uint32_t* pg) {
    for (int d = 0; d < cf->D; d++) if (x[d] < 0 || x[d] > 1) return 0;
    const float scale = (level < g_n_level_scales) ? g_level_scales[level] : exp2f((float)level * S) * (float)H - 1.0f;
    *scale_out = scale;
    *res_out = (uint32_t)ceilf(scale) + 1;
    for (int d = 0; d < cf->D; d++) {
        float p;
        if (cf->smooth) p = x[d] * scale;
        else            p = fmaf(x[d], scale, cf->align_corners ? 0.0f : 0.5f);
        const float fl = floorf(p);
        pg[d] = (uint32_t)fl;
        float f = p - (float)pg[d];
        if (cf->smooth) { dw[d] = 6 * f * (1.0f - f); w[d] = f * f * (3.0f - 2.0f * f); }
        else            { dw[d] = 1.0f; w[d] = f; }
    }
    return 1;
}

/* forward: outputs [L,B,C]; dy_dx [B, L*D*C] or NULL (hashencoder.cu:103-254, gridencoder.cu:75-223) */
static void encode_forward(const enc_cfg_t* cf, const float* inputs, const float* emb, const int32_t* offsets,
                           float* outputs, uint32_t B, uint32_t L, float S, uint32_t H, float* dy_dx) {
    const int D = cf->D, C = cf->C;
    #pragma omp parallel for schedule(static)
    for (uint32_t b = 0; b < B; b++) {
        for (uint32_t l = 0; l < L; l++) {
            const float* grid = emb + (size_t)(uint32_t)offsets[l] * C;
            float* out = outputs + ((size_t)l * B + b) * C;
            float* jac = dy_dx ? dy_dx + (size_t)b * D * L * C + (size_t)l * D * C : NULL;
            float w[MAXD], dw[MAXD], scale; uint32_t pg[MAXD], res;
            const uint32_t hsz = (uint32_t)(offsets[l+1] - offsets[l]);
            if (!level_setup(cf, inputs + (size_t)b * D, (int)l, S, H, &scale, &res, w, dw, pg)) {
                for (int c = 0; c < C; c++) out[c] = 0;
                if (jac) for (int i = 0; i < D * C; i++) jac[i] = 0;
                continue;
            }
            float acc[MAXC] = {0};
            for (uint32_t corner = 0; corner < (1u << D); corner++) {
                float wt = 1; uint32_t pl[MAXD];
                for (int d = 0; d < D; d++) {
                    if (corner & (1u << d)) { wt *= w[d]; pl[d] = pg[d] + 1; }
                    else                    { wt *= 1 - w[d]; pl[d] = pg[d]; }
                }
                const uint32_t idx = cell_index(cf, hsz, res, pl) * C;
                for (int c = 0; c < C; c++) acc[c] = fmaf(wt, grid[idx + c], acc[c]);
            }
            for (int c = 0; c < C; c++) out[c] = acc[c];
            if (!jac) continue;
            for (int gd = 0; gd < D; gd++) {
                float g[MAXC] = {0};
                for (uint32_t corner = 0; corner < (1u << (D - 1)); corner++) {
                    float wt = scale; uint32_t pl[MAXD];
                    for (int nd = 0; nd < D - 1; nd++) {
                        const int d = nd >= gd ? nd + 1 : nd;
                        if (corner & (1u << nd)) { wt *= w[d]; pl[d] = pg[d] + 1; }
                        else                     { wt *= 1 - w[d]; pl[d] = pg[d]; }
                    }
                    pl[gd] = pg[gd];     const uint32_t il = cell_index(cf, hsz, res, pl) * C;
                    pl[gd] = pg[gd] + 1; const uint32_t ir = cell_index(cf, hsz, res, pl) * C;
                    for (int c = 0; c < C; c++) {
                        if (cf->smooth) g[c] = fmaf(wt * (grid[ir + c] - grid[il + c]), dw[gd], g[c]);
                        else            g[c] = fmaf(wt, grid[ir + c] - grid[il + c], g[c]);
                    }
                }
                for (int c = 0; c < C; c++) jac[gd * C + c] = g[c];
            }
        }
    }
}

/* backward into the table (+ optional input gradient) -- hashencoder.cu:257-372, gridencoder.cu:226-342.
 * Sequential accumulation: summation order differs from the reference's float atomics. */
static void encode_backward(const enc_cfg_t* cf, const float* grad, const float* inputs, const int32_t* offsets,
                            float* grad_emb, uint32_t B, uint32_t L, float S, uint32_t H,
                            const float* dy_dx, float* grad_inputs) {
    const int D = cf->D, C = cf->C;
    for (uint32_t l = 0; l < L; l++) {
        float* gg = grad_emb + (size_t)(uint32_t)offsets[l] * C;
        const uint32_t hsz = (uint32_t)(offsets[l+1] - offsets[l]);
        for (uint32_t b = 0; b < B; b++) {
            float w[MAXD], dw[MAXD], scale; uint32_t pg[MAXD], res;
            if (!level_setup(cf, inputs + (size_t)b * D, (int)l, S, H, &scale, &res, w, dw, pg)) continue;
            const float* g = grad + ((size_t)l * B + b) * C;
            for (uint32_t corner = 0; corner < (1u << D); corner++) {
                float wt = 1; uint32_t pl[MAXD];
                for (int d = 0; d < D; d++) {
                    if (corner & (1u << d)) { wt *= w[d]; pl[d] = pg[d] + 1; }
                    else                    { wt *= 1 - w[d]; pl[d] = pg[d]; }
                }
                const uint32_t idx = cell_index(cf, hsz, res, pl) * C;
                for (int c = 0; c < C; c++) gg[idx + c] += wt * g[c];
            }
        }
    }
    if (dy_dx && grad_inputs) {
        for (uint32_t b = 0; b < B; b++) for (int d = 0; d < D; d++) {
            float r = 0;
            for (uint32_t l = 0; l < L; l++) for (int c = 0; c < C; c++)
                r += grad[((size_t)l * B + b) * C + c] * dy_dx[(size_t)b * L * D * C + (size_t)l * D * C + d * C + c];
            grad_inputs[(size_t)b * D + d] = r;
        }
    }
}

ORC_API int orc_hash_encode_forward(const float* inputs, const float* emb, const int32_t* offsets, float* outputs,
        uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, float* dy_dx) {
    if (D > MAXD || C > MAXC) return 1;
    enc_cfg_t cf = {(int)D, (int)C, 1, 0, 0};
    encode_forward(&cf, inputs, emb, offsets, outputs, B, L, S, H, calc_grad_inputs ? dy_dx : NULL);
    return 0;
}
ORC_API int orc_hash_encode_backward(const float* grad, const float* inputs, const float* emb, const int32_t* offsets,
        float* grad_emb, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
        int calc_grad_inputs, const float* dy_dx, float* grad_inputs) {
    (void)emb;
    if (D > MAXD || C > MAXC) return 1;
    enc_cfg_t cf = {(int)D, (int)C, 1, 0, 0};
    encode_backward(&cf, grad, inputs, offsets, grad_emb, B, L, S, H, calc_grad_inputs ? dy_dx : NULL, grad_inputs);
    return 0;
}

/* hashencoder.cu:375-595: derivative of grad_inputs = sum_{l,c} grad[l,b,c] * dy_dx[b,l,:,c]
 * w.r.t. `grad` (-> grad_grad) and w.r.t. the table (-> grad2_embeddings). */
ORC_API int orc_hash_encode_second_backward(const float* grad, const float* inputs, const float* emb, const int32_t* offsets,
        uint32_t B, uint32_t D_, uint32_t C_, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
        const float* dy_dx, const float* grad_grad_inputs, float* grad_grad, float* grad2_emb) {
    (void)emb; (void)calc_grad_inputs;
    if (D_ > MAXD || C_ > MAXC) return 1;
    const int D = (int)D_, C = (int)C_;
    enc_cfg_t cf = {D, C, 1, 0, 0};
    for (uint32_t l = 0; l < L; l++) {
        float* g2 = grad2_emb + (size_t)(uint32_t)offsets[l] * C;
        const uint32_t hsz = (uint32_t)(offsets[l+1] - offsets[l]);
        for (uint32_t b = 0; b < B; b++) {
            const float* ggx = grad_grad_inputs + (size_t)b * D;
            const float* jac = dy_dx + (size_t)b * L * D * C + (size_t)l * D * C;
            float* og = grad_grad + ((size_t)l * B + b) * C;
            for (int c = 0; c < C; c++) {           /* :375-413 -- computed even for out-of-range inputs */
                float r = 0;
                for (int d = 0; d < D; d++) r += ggx[d] * jac[d * C + c];
                og[c] = r;
            }
            float w[MAXD], dw[MAXD], scale; uint32_t pg[MAXD], res;
            if (!level_setup(&cf, inputs + (size_t)b * D, (int)l, S, H, &scale, &res, w, dw, pg)) continue;
            const float* g = grad + ((size_t)l * B + b) * C;
            float cache[(1 << MAXD) * MAXC]; memset(cache, 0, sizeof(cache));
            for (int gd = 0; gd < D; gd++) {
                for (uint32_t corner = 0; corner < (1u << (D - 1)); corner++) {
                    float wt = scale; uint32_t bits = 0;
                    for (int nd = 0; nd < D - 1; nd++) {
                        const int d = nd >= gd ? nd + 1 : nd;
                        if (corner & (1u << nd)) { wt *= w[d]; bits |= 1u << d; }
                        else                     { wt *= 1 - w[d]; }
                    }
                    const uint32_t left = bits, right = bits | (1u << gd);
                    for (int c = 0; c < C; c++) {
                        const float v = wt * g[c] * ggx[gd] * dw[gd];
                        cache[right * C + c] += v;
                        cache[left * C + c] -= v;
                    }
                }
            }
            for (uint32_t corner = 0; corner < (1u << D); corner++) {
                uint32_t pl[MAXD];
                for (int d = 0; d < D; d++) pl[d] = pg[d] + ((corner >> d) & 1u);
                const uint32_t idx = cell_index(&cf, hsz, res, pl) * C;
                for (int c = 0; c < C; c++) g2[idx + c] += cache[corner * C + c];
            }
        }
    }
    return 0;
}

ORC_API int orc_grid_encode_forward(const float* inputs, const float* emb, const int32_t* offsets, float* outputs,
        uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, float* dy_dx,
        uint32_t gridtype, int align_corners) {
    if (D > MAXD || C > MAXC) return 1;
    enc_cfg_t cf = {(int)D, (int)C, 0, (int)gridtype, align_corners};
    encode_forward(&cf, inputs, emb, offsets, outputs, B, L, S, H, dy_dx);
    return 0;
}
ORC_API int orc_grid_encode_backward(const float* grad, const float* inputs, const float* emb, const int32_t* offsets,
        float* grad_emb, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
        const float* dy_dx, float* grad_inputs, uint32_t gridtype, int align_corners) {
    (void)emb;
    if (D > MAXD || C > MAXC) return 1;
    enc_cfg_t cf = {(int)D, (int)C, 0, (int)gridtype, align_corners};
    encode_backward(&cf, grad, inputs, offsets, grad_emb, B, L, S, H, dy_dx, grad_inputs);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* frequency encoding (freqencoder/src/freqencoder.cu:30-94)                  */
/* ------------------------------------------------------------------------- */
ORC_API void orc_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs) {
    (void)deg;
    const float HALF_PI = 3.141592653589793f / 2;
    for (uint32_t b = 0; b < B; b++) for (uint32_t c = 0; c < C; c++) {
        float v;
        if (c < D) v = inputs[(size_t)b * D + c];
        else {
            const uint32_t col = c / D - 1, d = c % D, f = col / 2;
            v = sinf(scalbnf(inputs[(size_t)b * D + d], (int)f) + (float)(col % 2) * HALF_PI);
        }
        outputs[(size_t)b * C + c] = v;
    }
}
ORC_API void orc_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* grad_inputs) {
    for (uint32_t b = 0; b < B; b++) for (uint32_t d = 0; d < D; d++) {
        const float* g = grad + (size_t)b * C; const float* o = outputs + (size_t)b * C;
        float r = g[d];
        g += D; o += D;
        for (uint32_t f = 0; f < deg; f++) {
            r += scalbnf(1.0f, (int)f) * (g[d] * o[D + d] - g[D + d] * o[d]);
            g += 2 * D; o += 2 * D;
        }
        grad_inputs[(size_t)b * D + d] = r;
    }
}

/* ------------------------------------------------------------------------- */
/* real spherical harmonics (shencoder/src/shencoder.cu:27-383)               */
/* ------------------------------------------------------------------------- */
/* The reference hard-codes 64 polynomials (+3x64 derivatives).  Each is
 *   Y_l^m(x,y,z) = N_l^|m| * Q_l^|m|(z) * { Re, Im }[(x+iy)^|m|],  index l*l + l + m,
 * with Q_l^m(z) = (-1)^m d^m/dz^m P_l(z) (Condon-Shortley phase, no (1-z^2)^{m/2} factor), and the
 * polynomials are used as-is for non-unit inputs.  This restatement evaluates the same polynomials
 * by recurrence in double precision (identical values up to fp32 rounding of the reference). */
static void sh_eval(double x, double y, double z, int deg, double* Y, double* dYx, double* dYy, double* dYz) {
    double A[9], Bm[9];           /* Re/Im (x+iy)^m */
    A[0] = 1; Bm[0] = 0;
    for (int m = 1; m <= deg; m++) { A[m] = x * A[m-1] - y * Bm[m-1]; Bm[m] = x * Bm[m-1] + y * A[m-1]; }
    double Ql[10][10];            /* Q[l][m] */
    memset(Ql, 0, sizeof(Ql));
    for (int m = 0; m <= deg; m++) {
        double qmm = 1;           /* (-1)^m (2m-1)!! */
        for (int k = 1; k <= m; k++) qmm *= -(2.0 * k - 1.0);
        Ql[m][m] = qmm;
        if (m + 1 <= deg) Ql[m+1][m] = (2.0 * m + 1.0) * z * qmm;
        for (int l = m + 2; l <= deg; l++)
            Ql[l][m] = ((2.0 * l - 1.0) * z * Ql[l-1][m] - (l + m - 1.0) * Ql[l-2][m]) / (double)(l - m);
    }
    for (int l = 0; l < deg; l++) {
        for (int m = 0; m <= l; m++) {
            double fac = 1;       /* (l-m)!/(l+m)! */
            for (int k = l - m + 1; k <= l + m; k++) fac /= (double)k;
            double Nlm = sqrt((2.0 * l + 1.0) / (4.0 * M_PI) * fac) * (m ? sqrt(2.0) : 1.0);
            const double q = Ql[l][m];
            const double dq = (m + 1 <= l) ? -Ql[l][m+1] : 0.0;     /* d/dz Q_l^m = -Q_l^{m+1} (phase) */
            const int ip = l * l + l + m, in = l * l + l - m;
            Y[ip] = Nlm * q * A[m];
            if (dYx) {
                dYx[ip] = m ? Nlm * q * m * A[m-1] : 0.0;
                dYy[ip] = m ? -Nlm * q * m * Bm[m-1] : 0.0;
                dYz[ip] = Nlm * dq * A[m];
            }
            if (m) {
                Y[in] = Nlm * q * Bm[m];
                if (dYx) {
                    dYx[in] = Nlm * q * m * Bm[m-1];
                    dYy[in] = Nlm * q * m * A[m-1];
                    dYz[in] = Nlm * dq * Bm[m];
                }
            }
        }
    }
}
/* outputs [B, deg^2]; dy_dx [B, 3*deg^2] (x block, y block, z block) or NULL */
ORC_API void orc_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t deg, float* dy_dx) {
    const uint32_t C2 = deg * deg;
    for (uint32_t b = 0; b < B; b++) {
        double Y[64], gx[64], gy[64], gz[64];
        sh_eval(inputs[(size_t)b*D], inputs[(size_t)b*D+1], inputs[(size_t)b*D+2], (int)deg, Y, dy_dx ? gx : NULL, gy, gz);
        for (uint32_t i = 0; i < C2; i++) outputs[(size_t)b*C2 + i] = (float)Y[i];
        if (dy_dx) for (uint32_t i = 0; i < C2; i++) {
            dy_dx[(size_t)b*3*C2 + i] = (float)gx[i];
            dy_dx[(size_t)b*3*C2 + C2 + i] = (float)gy[i];
            dy_dx[(size_t)b*3*C2 + 2*C2 + i] = (float)gz[i];
        }
    }
}
/* shencoder.cu:358-383: grad_inputs[b,d] += sum_k grad[b,k] * dy_dx[b,d,k]  (accumulates) */
ORC_API void orc_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t deg,
                                    const float* dy_dx, float* grad_inputs) {
    (void)inputs;
    const uint32_t C2 = deg * deg;
    for (uint32_t b = 0; b < B; b++) for (uint32_t d = 0; d < D; d++) {
        float r = 0;
        for (uint32_t k = 0; k < C2; k++) r += grad[(size_t)b*C2 + k] * dy_dx[(size_t)b*D*C2 + (size_t)d*C2 + k];
        grad_inputs[(size_t)b*D + d] += r;
    }
}
