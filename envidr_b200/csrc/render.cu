// render.cu -- fused inference render loop for sm_100a.
//
// Replaces the host-driven while-loop of the reference (nerf/render_func/cuda_ray.py:238-359: per
// iteration march_rays -> forward_sigma -> get_color_mlp_extra_params -> forward_color -> up to four
// composite_rays launches -> boolean-mask compaction with a device->host sync) by a device-driven
// wavefront: per iteration THREE launches (march+compact, fused field, composite+compact+advance), no
// host synchronisation, no zero-padded sample slots, and the alive list / step schedule kept on the device.
//
// The reference's schedule is preserved exactly: n_step = max(min(N / n_alive, 8), 1) samples per alive ray
// and iteration, rays_t re-synchronised from the accumulated deltas at iteration boundaries, a ray dies when
// it produced fewer than n_step samples or when the transmittance it had before its last sample is below
// T_thresh (raymarching/src/raymarching.cu:996-1039).  Sample positions are therefore bit-identical to the
// reference loop; only the placement of samples in the batch differs (compact, warp-aggregated allocation).
#include <stdlib.h>
#include "common.cuh"

namespace envidr {

struct RecCapture { float* rec; const uint32_t* base_dev; uint32_t cap; };        // same as field_tc.cuh
int field_forward_launch(const envidr_field* field, const float* xyzs, const float* dirs, const float* r_images,
                         const uint32_t* M_dev, uint32_t M_host, int mode, const envidr_field_out* out, cudaStream_t st,
                         cudaEvent_t* ev, int* ev_recorded, const RecCapture* cap);

constexpr int kMarchBlock = 128;
constexpr int kMaxNStep = 16;          // hard upper bound (staging array of k_march_compact); the reference schedule caps at 8
constexpr int kRefNStepCap = 8;

struct Counters {
    uint32_t n_alive;        // alive rays entering the current iteration
    uint32_t n_alive_next;   // filled by the compositor
    uint32_t M;              // samples produced by the current march
    uint32_t n_step;
    uint32_t step_total;     // sum of n_step so far (the reference's `step`)
    uint32_t iters;
    uint32_t total_samples_lo, total_samples_hi;
    uint32_t done_blocks;    // ticket for the last-block advance
    uint32_t pad[7];
};

struct RenderBuffers {
    float *nears, *fars, *rays_t;
    int32_t* alive[2];
    int2* slot;              // (offset, count) per alive slot
    float *s_xyz, *s_dir, *s_delta, *s_rimg;
    float *s_sigma, *s_rgb, *s_normal, *s_cd, *s_cs, *s_rough;
    Counters* ctr;
    int* occ_box;            // {x0, x1, y0, y1, z0, z1}: bounding box of the occupied level-0 cells (single-cascade fast path)
};

struct RenderOutDev {
    float *image, *depth, *weights_sum, *normal_image, *diffuse_image, *specular_image, *roughness_image;
    int32_t* sample_count;
    // sample log (envidr_sample_log): per marched sample, at index total_samples_before_this_iteration + batch slot
    float *log_sigma, *log_delta;
    int32_t *log_ray, *log_seq;
    uint32_t log_cap;
};

__global__ void __launch_bounds__(256) k_render_init(const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t N,
                                                    uint32_t n_step_floor, float min_near, float a0, float a1, float a2, float a3, float a4, float a5,
                                                    RenderBuffers B, RenderOutDev O) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) {
        Counters c{};
        c.n_alive = N; c.n_step = n_step_floor;
        *B.ctr = c;
        for (int a = 0; a < 3; a++) { B.occ_box[2 * a] = 0x7fffffff; B.occ_box[2 * a + 1] = -1; }
    }
    if (n >= N) return;
    // slab test (same arithmetic as k_near_far / reference raymarching.cu:91-145)
    const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
    const float rdx = 1 / rays_d[3 * n], rdy = 1 / rays_d[3 * n + 1], rdz = 1 / rays_d[3 * n + 2];
    float lo = (a0 - ox) * rdx, hi = (a3 - ox) * rdx;
    if (lo > hi) { float s = lo; lo = hi; hi = s; }
    float lo2 = (a1 - oy) * rdy, hi2 = (a4 - oy) * rdy;
    if (lo2 > hi2) { float s = lo2; lo2 = hi2; hi2 = s; }
    bool miss = (lo > hi2) || (lo2 > hi);
    if (!miss) {
        if (lo2 > lo) lo = lo2;
        if (hi2 < hi) hi = hi2;
        lo2 = (a2 - oz) * rdz; hi2 = (a5 - oz) * rdz;
        if (lo2 > hi2) { float s = lo2; lo2 = hi2; hi2 = s; }
        miss = (lo > hi2) || (lo2 > hi);
        if (!miss) {
            if (lo2 > lo) lo = lo2;
            if (hi2 < hi) hi = hi2;
            if (lo < min_near) lo = min_near;
        }
    }
    const float near = miss ? 3.402823466e+38f : lo, far = miss ? 3.402823466e+38f : hi;
    B.nears[n] = near; B.fars[n] = far; B.rays_t[n] = near;
    B.alive[0][n] = (int32_t)n;
    O.weights_sum[n] = 0; O.depth[n] = 0;
    O.image[3 * n] = 0; O.image[3 * n + 1] = 0; O.image[3 * n + 2] = 0;
    if (O.normal_image) { O.normal_image[3 * n] = 0; O.normal_image[3 * n + 1] = 0; O.normal_image[3 * n + 2] = 0; }
    if (O.diffuse_image) { O.diffuse_image[3 * n] = 0; O.diffuse_image[3 * n + 1] = 0; O.diffuse_image[3 * n + 2] = 0; }
    if (O.specular_image) { O.specular_image[3 * n] = 0; O.specular_image[3 * n + 1] = 0; O.specular_image[3 * n + 2] = 0; }
    if (O.roughness_image) O.roughness_image[n] = 0;
    if (O.sample_count) O.sample_count[n] = 0;
}

// bounding box of the occupied cells of cascade level 0 (bit = Morton(x, y, z), reference raymarching.cu:56-81)
__global__ void __launch_bounds__(256) k_occ_box(const uint8_t* __restrict__ grid, uint32_t n_words, int* __restrict__ box) {
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
    const uint32_t* __restrict__ g32 = reinterpret_cast<const uint32_t*>(grid);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) {
        uint32_t w = g32[i];
        while (w) {
            const uint32_t b = __ffs(w) - 1;
            w &= w - 1;
            const uint32_t idx = i * 32 + b;
            const int c[3] = {(int)compact3(idx), (int)compact3(idx >> 1), (int)compact3(idx >> 2)};
            #pragma unroll
            for (int a = 0; a < 3; a++) { lo[a] = min(lo[a], c[a]); hi[a] = max(hi[a], c[a]); }
        }
    }
    #pragma unroll
    for (int a = 0; a < 3; a++) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0 && hi[a] >= 0) { atomicMin(box + 2 * a, lo[a]); atomicMax(box + 2 * a + 1, hi[a]); }
    }
}

// march n_step samples for every alive ray and append them compactly to the sample batch
__global__ void __launch_bounds__(kMarchBlock) k_march_compact(
        const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ r_images,
        const uint8_t* __restrict__ grid, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
        const float* __restrict__ noises, int fast, RenderBuffers B) {
    __shared__ float stage[kMarchBlock][kMaxNStep][5];
    Counters* ctr = B.ctr;
    const uint32_t n_alive = ctr->n_alive, n_step = ctr->n_step;
    if (n_alive == 0 || ctr->step_total >= max_steps) return;
    const int32_t* __restrict__ alive = B.alive[ctr->iters & 1];
    const bool first = ctr->iters == 0;
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * kMarchBlock; base < n_alive; base += gridDim.x * kMarchBlock) {
        const uint32_t n = base + threadIdx.x;
        uint32_t count = 0;
        int index = -1;
        Dda s;
        if (n < n_alive) {
            index = alive[n];
            s.init(rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, dt_gamma, max_steps, C, H, grid);
            const float far = B.fars[index];
            float t = B.rays_t[index];
            float last_t = t;
            const float noise = (first && noises) ? noises[n] : 0.0f;
            t += s.step_size(t) * noise;
            float x, y, z, dt;
            if (fast) {
                s.init_fast(B.occ_box);
                if (first && s.misses_box()) t = far;
                while (t < far && count < n_step) {
                    if (s.probe_fast(t, x, y, z, dt, far)) {
                        float* q = stage[threadIdx.x][count];
                        q[0] = x; q[1] = y; q[2] = z;
                        t += dt;
                        q[3] = dt; q[4] = t - last_t;
                        last_t = t;
                        count++;
                    }
                }
            } else {
                while (t < far && count < n_step) {
                    if (s.probe(t, x, y, z, dt)) {
                        float* q = stage[threadIdx.x][count];
                        q[0] = x; q[1] = y; q[2] = z;
                        t += dt;
                        q[3] = dt; q[4] = t - last_t;
                        last_t = t;
                        count++;
                    }
                }
            }
        }
        // warp-aggregated allocation of `count` consecutive sample slots
        uint32_t incl = count;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += u;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t wbase = 0;
        if (lane == 0 && total) wbase = atomicAdd(&ctr->M, total);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        const uint32_t off = wbase + incl - count;
        if (n < n_alive) {
            B.slot[n] = make_int2((int)off, (int)count);
            float4 ri = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r_images && count) ri = *reinterpret_cast<const float4*>(r_images + 4 * (size_t)index);
            for (uint32_t j = 0; j < count; j++) {
                const float* q = stage[threadIdx.x][j];
                const size_t m = (size_t)off + j;
                B.s_xyz[3 * m] = q[0]; B.s_xyz[3 * m + 1] = q[1]; B.s_xyz[3 * m + 2] = q[2];
                B.s_dir[3 * m] = s.dx; B.s_dir[3 * m + 1] = s.dy; B.s_dir[3 * m + 2] = s.dz;
                *reinterpret_cast<float2*>(B.s_delta + 2 * m) = make_float2(q[3], q[4]);
                if (r_images) *reinterpret_cast<float4*>(B.s_rimg + 4 * m) = ri;
            }
        }
    }
}

// composite this iteration's samples into the per-ray accumulators, decide which rays stay alive, compact
// the alive list, and (last block) advance the iteration counters.
__global__ void __launch_bounds__(kMarchBlock) k_composite_compact(uint32_t N, float T_thresh, uint32_t max_steps, int geometry_only,
                                                                  int input_alpha, uint32_t n_step_floor, uint32_t n_step_cap, RenderBuffers B,
                                                                  RenderOutDev O) {
    Counters* ctr = B.ctr;
    const uint32_t n_alive = ctr->n_alive, n_step = ctr->n_step;
    const bool active = !(n_alive == 0 || ctr->step_total >= max_steps);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t log_base = ctr->total_samples_lo;       // samples of the earlier iterations (advanced by the last block, below)
    if (active) {
        const int32_t* __restrict__ alive = B.alive[ctr->iters & 1];
        int32_t* __restrict__ alive_next = B.alive[(ctr->iters & 1) ^ 1];
        for (uint32_t base = blockIdx.x * kMarchBlock; base < n_alive; base += gridDim.x * kMarchBlock) {
            const uint32_t n = base + threadIdx.x;
            bool keep = false;
            int index = -1;
            if (n < n_alive) {
                index = alive[n];
                const int2 sl = B.slot[n];
                const uint32_t cnt = (uint32_t)sl.y;
                float t = B.rays_t[index];
                float ws = O.weights_sum[index], d = O.depth[index];
                // every accumulator starts from its stored value and adds sample by sample, exactly like one
                // composite_rays call per image in the reference (cuda_ray.py:318-342)
                float acc[3], nrm[3] = {0, 0, 0}, cd[3] = {0, 0, 0}, cs[3] = {0, 0, 0}, rgh = 0;
                float* main_img = geometry_only ? O.normal_image : O.image;
                acc[0] = main_img[3 * index]; acc[1] = main_img[3 * index + 1]; acc[2] = main_img[3 * index + 2];
                const bool want_n = !geometry_only && O.normal_image;
                if (want_n) { const float* q = O.normal_image + 3 * index; nrm[0] = q[0]; nrm[1] = q[1]; nrm[2] = q[2]; }
                if (O.diffuse_image) { const float* q = O.diffuse_image + 3 * index; cd[0] = q[0]; cd[1] = q[1]; cd[2] = q[2]; }
                if (O.specular_image) { const float* q = O.specular_image + 3 * index; cs[0] = q[0]; cs[1] = q[1]; cs[2] = q[2]; }
                if (O.roughness_image) rgh = O.roughness_image[index];
                uint32_t step = 0, used = 0;
                const uint32_t seq0 = (O.log_ray && O.sample_count) ? (uint32_t)O.sample_count[index] : 0u;
                while (step < n_step) {
                    if (step >= cnt) break;                       // zero-delta slot in the reference layout
                    used++;
                    const size_t m = (size_t)sl.x + step;
                    const float2 dl = *reinterpret_cast<const float2*>(B.s_delta + 2 * m);
                    const float sg = B.s_sigma[m];
                    if (O.log_ray) {
                        const size_t gi = (size_t)log_base + m;
                        if (gi < O.log_cap) {
                            O.log_ray[gi] = index; O.log_seq[gi] = (int32_t)(seq0 + used - 1);
                            O.log_sigma[gi] = sg; *reinterpret_cast<float2*>(O.log_delta + 2 * gi) = dl;
                        }
                    }
                    const float alpha = input_alpha ? 0.0f + sg : 1.0f - __expf(-sg * dl.x);
                    const float T = 1 - ws;
                    const float w = alpha * T;
                    ws += w;
                    t = t + dl.y;
                    d += w * t;
                    const float* c = (geometry_only ? B.s_normal : B.s_rgb) + 3 * m;
                    acc[0] += w * c[0]; acc[1] += w * c[1]; acc[2] += w * c[2];
                    if (want_n) { const float* q = B.s_normal + 3 * m; nrm[0] += w * q[0]; nrm[1] += w * q[1]; nrm[2] += w * q[2]; }
                    if (O.diffuse_image) { const float* q = B.s_cd + 3 * m; cd[0] += w * q[0]; cd[1] += w * q[1]; cd[2] += w * q[2]; }
                    if (O.specular_image) { const float* q = B.s_cs + 3 * m; cs[0] += w * q[0]; cs[1] += w * q[1]; cs[2] += w * q[2]; }
                    if (O.roughness_image) rgh += w * B.s_rough[m];
                    if (T < T_thresh) break;
                    step++;
                }
                if (O.log_ray) {                                  // marched but not composited: not part of any replay
                    for (uint32_t s2 = used; s2 < cnt; s2++) {
                        const size_t gi = (size_t)log_base + (size_t)sl.x + s2;
                        if (gi < O.log_cap) O.log_ray[gi] = -1;
                    }
                }
                keep = !(step < n_step);
                if (keep) B.rays_t[index] = t;
                O.weights_sum[index] = ws; O.depth[index] = d;
                main_img[3 * index] = acc[0]; main_img[3 * index + 1] = acc[1]; main_img[3 * index + 2] = acc[2];
                if (want_n) { float* q = O.normal_image + 3 * index; q[0] = nrm[0]; q[1] = nrm[1]; q[2] = nrm[2]; }
                if (O.diffuse_image) { float* q = O.diffuse_image + 3 * index; q[0] = cd[0]; q[1] = cd[1]; q[2] = cd[2]; }
                if (O.specular_image) { float* q = O.specular_image + 3 * index; q[0] = cs[0]; q[1] = cs[1]; q[2] = cs[2]; }
                if (O.roughness_image) O.roughness_image[index] = rgh;
                if (O.sample_count) O.sample_count[index] += (int32_t)used;
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, keep);
            uint32_t wbase = 0;
            if (lane == 0 && mask) wbase = atomicAdd(&ctr->n_alive_next, (uint32_t)__popc(mask));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (keep) alive_next[wbase + __popc(mask & ((1u << lane) - 1))] = index;
        }
    }
    // last block to finish advances the iteration state
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&ctr->done_blocks, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        if (active) {
            const uint32_t next = atomicAdd(&ctr->n_alive_next, 0u);
            const uint32_t M = atomicAdd(&ctr->M, 0u);
            const uint64_t tot = (((uint64_t)ctr->total_samples_hi << 32) | ctr->total_samples_lo) + M;
            ctr->total_samples_lo = (uint32_t)tot; ctr->total_samples_hi = (uint32_t)(tot >> 32);
            ctr->step_total += n_step;
            ctr->iters += 1;
            ctr->n_alive = next;
            ctr->n_alive_next = 0;
            ctr->M = 0;
            uint32_t ns = next ? N / next : 1;
            ns = ns > n_step_cap ? n_step_cap : ns;
            ctr->n_step = ns < n_step_floor ? n_step_floor : ns;
        }
        ctr->done_blocks = 0;
    }
}

// sample log (iteration-major, one entry per marched sample) -> ray-contiguous arrays of the samples a later pass composites.
// One warp per log entry group: 8 lanes move the 128-byte record of one sample (float4 each), 4 samples per warp.
__global__ void __launch_bounds__(256) k_permute_log(const float4* __restrict__ rec, const float* __restrict__ sigma, const float2* __restrict__ delta,
                                                    const int32_t* __restrict__ ray, const int32_t* __restrict__ seq, uint64_t total,
                                                    const int32_t* __restrict__ ray_offset, float4* __restrict__ rec_out,
                                                    float* __restrict__ sigma_out, float2* __restrict__ delta_out, int32_t* __restrict__ idx_out,
                                                    uint32_t shift) {
    // shift = 3: 8 lanes per entry move its 128-byte record; shift = 0 (index mode, rec_out == NULL): one thread per entry
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t g = t >> shift;
    const uint32_t part = (uint32_t)t & ((1u << shift) - 1u);
    if (g >= total) return;
    const int32_t r = ray[g];
    if (r < 0) return;
    const int32_t off = ray_offset[r];
    if (off < 0) return;
    const uint64_t dst = (uint64_t)off + (uint32_t)seq[g];
    if (rec_out) rec_out[dst * 8 + part] = rec[g * 8 + part];
    if (part == 0) { sigma_out[dst] = sigma[g]; delta_out[dst] = delta[g]; if (idx_out) idx_out[dst] = (int32_t)g; }
}

// image += (1 - ws) * bg ; normal_image <- F.normalize(normal_image, eps=1e-10)   (cuda_ray.py:348-359)
__global__ void __launch_bounds__(256) k_render_finish(uint32_t N, float bg0, float bg1, float bg2, const float* __restrict__ bg_per_ray,
                                                      int geometry_only, RenderOutDev O) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (!geometry_only) {
        const float rem = 1 - O.weights_sum[n];
        const float b0 = bg_per_ray ? bg_per_ray[3 * n] : bg0, b1 = bg_per_ray ? bg_per_ray[3 * n + 1] : bg1,
                    b2 = bg_per_ray ? bg_per_ray[3 * n + 2] : bg2;
        O.image[3 * n] += rem * b0; O.image[3 * n + 1] += rem * b1; O.image[3 * n + 2] += rem * b2;
    }
    if (O.normal_image) {
        float* q = O.normal_image + 3 * n;
        const float den = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), 1e-10f);      // F.normalize: v / max(|v|, eps)
        q[0] = q[0] / den; q[1] = q[1] / den; q[2] = q[2] / den;
    }
}

static uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
    uint64_t nears, fars, rays_t, alive0, alive1, slot, s_xyz, s_dir, s_delta, s_rimg, s_sigma, s_rgb, s_normal, s_cd, s_cs, s_rough, ctr, occ_box, scratch, total;
};
static uint32_t clamp_floor(uint32_t f) { return f < 1 ? 1u : (f > (uint32_t)kRefNStepCap ? (uint32_t)kRefNStepCap : f); }
static uint32_t clamp_cap(uint32_t c) { return c == 0 ? (uint32_t)kRefNStepCap : (c > (uint32_t)kMaxNStep ? (uint32_t)kMaxNStep : c); }
static WsLayout ws_layout(uint32_t N, uint32_t n_step_floor = 1) {
    WsLayout L{};
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { const uint64_t o = off; off = align_up(off + bytes, 256); return o; };
    const uint64_t n = N, m = (uint64_t)N * clamp_floor(n_step_floor);     // m: samples of one iteration (n_alive * n_step <= N * floor)
    L.nears = take(4 * n); L.fars = take(4 * n); L.rays_t = take(4 * n);
    L.alive0 = take(4 * n); L.alive1 = take(4 * n); L.slot = take(8 * n);
    L.s_xyz = take(12 * m); L.s_dir = take(12 * m); L.s_delta = take(8 * m); L.s_rimg = take(16 * m);
    L.s_sigma = take(4 * m); L.s_rgb = take(12 * m); L.s_normal = take(12 * m); L.s_cd = take(12 * m); L.s_cs = take(12 * m);
    L.s_rough = take(4 * m);
    L.ctr = take(sizeof(Counters));
    L.occ_box = take(8 * sizeof(int));
    L.scratch = take(256 * m);             // tensor-core path: per-sample record + env features (2 x 32 floats)
    L.total = off;
    return L;
}

static uint32_t* g_host_flag = nullptr;     // pinned: [n_alive snapshot ring (8)] + stats[4]
static cudaEvent_t g_events[2];
static bool g_events_ok = false;

// instrumentation (bench.py): kernel launch counter and per-launch CUDA-event timing of the field kernel
uint64_t g_launches = 0;           // also bumped by the operator entry points the replay path uses (common.cuh)
static int g_timing = 0;
constexpr int kMaxTimed = 4096;
static cudaEvent_t g_tev[2 * kMaxTimed];
static int g_tev_created = 0, g_tev_used = 0;

// event pair for timing one field launch outside the fused loop (envidr_field_forward); nullptr when timing is off
cudaEvent_t* timing_acquire() {
    if (!g_timing || g_tev_used >= kMaxTimed) return nullptr;
    while (g_tev_created < 2 * (g_tev_used + 1)) cudaEventCreate(&g_tev[g_tev_created++]);
    return &g_tev[2 * g_tev_used];
}
void timing_commit() { g_tev_used++; }

}  // namespace envidr

using namespace envidr;

extern "C" {

uint64_t envidr_render_workspace_bytes(uint32_t N) { return ws_layout(N).total; }
uint64_t envidr_render_workspace_bytes_ex(uint32_t N, uint32_t n_step_floor) { return ws_layout(N, n_step_floor).total; }

int envidr_render_rays(const envidr_field* field, const uint8_t* bitfield, const float* rays_o, const float* rays_d,
                       const float* r_images, const float* noises, const float* bg_per_ray, uint32_t N,
                       const envidr_render_opts* opts, const envidr_render_out* out, void* workspace, uint64_t workspace_bytes,
                       envidr_stream_t stream) {
    ENVIDR_REQUIRE(field && bitfield && rays_o && rays_d && opts && out && workspace, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(out->weights_sum && out->depth && (opts->geometry_only ? out->normal_image != nullptr : out->image != nullptr),
                   ENVIDR_E_BADARG, "missing mandatory output buffer");
    ENVIDR_REQUIRE(opts->cascade >= 1 && opts->cascade <= 8 && opts->grid_size >= 1 && opts->grid_size <= 1024, ENVIDR_E_UNSUPPORTED,
                   "cascades must be 1..8, grid size <= 1024");
    if (N == 0) return 0;
    const uint32_t nsf = clamp_floor(opts->n_step_floor);
    const uint32_t nsc = clamp_cap(opts->n_step_cap) < nsf ? nsf : clamp_cap(opts->n_step_cap);
    const WsLayout L = ws_layout(N, nsf);
    ENVIDR_REQUIRE(workspace_bytes >= L.total, ENVIDR_E_WORKSPACE, "workspace too small (envidr_render_workspace_bytes)");
    ENVIDR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ENVIDR_E_BADARG, "workspace must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    if (!g_events_ok) {
        if (cudaHostAlloc(&g_host_flag, 64 * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) {
            set_error("render: pinned host allocation failed"); return (int)cudaErrorMemoryAllocation;
        }
        cudaEventCreateWithFlags(&g_events[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g_events[1], cudaEventDisableTiming);
        g_events_ok = true;
    }
    char* w = reinterpret_cast<char*>(workspace);
    RenderBuffers B{};
    B.nears = (float*)(w + L.nears); B.fars = (float*)(w + L.fars); B.rays_t = (float*)(w + L.rays_t);
    B.alive[0] = (int32_t*)(w + L.alive0); B.alive[1] = (int32_t*)(w + L.alive1); B.slot = (int2*)(w + L.slot);
    B.s_xyz = (float*)(w + L.s_xyz); B.s_dir = (float*)(w + L.s_dir); B.s_delta = (float*)(w + L.s_delta); B.s_rimg = (float*)(w + L.s_rimg);
    B.s_sigma = (float*)(w + L.s_sigma); B.s_rgb = (float*)(w + L.s_rgb); B.s_normal = (float*)(w + L.s_normal);
    B.s_cd = (float*)(w + L.s_cd); B.s_cs = (float*)(w + L.s_cs); B.s_rough = (float*)(w + L.s_rough);
    B.ctr = (Counters*)(w + L.ctr);
    B.occ_box = (int*)(w + L.occ_box);
    RenderOutDev O{out->image, out->depth, out->weights_sum, out->normal_image, out->diffuse_image, out->specular_image,
                   out->roughness_image, out->sample_count, nullptr, nullptr, nullptr, nullptr, 0u};
    RecCapture capture{nullptr, nullptr, 0u};
    if (out->log) {
        const envidr_sample_log* lg = out->log;
        ENVIDR_REQUIRE(opts->geometry_only && field->precision == 1, ENVIDR_E_BADARG, "sample log: geometry_only passes of the tensor-core field");
        ENVIDR_REQUIRE(lg->rec && lg->sigma && lg->delta && lg->ray && lg->seq && out->sample_count, ENVIDR_E_BADARG, "sample log: null buffer");
        O.log_sigma = lg->sigma; O.log_delta = lg->delta; O.log_ray = lg->ray; O.log_seq = lg->seq;
        O.log_cap = (uint32_t)(lg->capacity > 0xFFFFFFFFull ? 0xFFFFFFFFull : lg->capacity);
        capture = RecCapture{lg->rec, &B.ctr->total_samples_lo, O.log_cap};
    }
    if (opts->geometry_only) { O.image = out->normal_image; O.diffuse_image = nullptr; O.specular_image = nullptr; O.roughness_image = nullptr; }
    const float* a = opts->aabb;
    k_render_init<<<ceil_div(N, 256), 256, 0, st>>>(rays_o, rays_d, N, nsf, opts->min_near, a[0], a[1], a[2], a[3], a[4], a[5], B, O);
    // single-cascade, constant-step scenes take the specialised march (same samples, see Dda::probe_fast)
    const int fast = opts->cascade == 1 && opts->dt_gamma == 0.0f && opts->grid_size <= 256 && opts->grid_size % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(bitfield) & 3) == 0;
    if (fast) {
        const uint32_t H = opts->grid_size;
        k_occ_box<<<kSMs, 256, 0, st>>>(bitfield, H * H * H / 32, B.occ_box);
        g_launches += 1;
    }
    int rc = check_launch("render_init");
    if (rc) return rc;

    envidr_field fld = *field;                       // per-iteration sample count is bounded by N
    fld.scratch = w + L.scratch; fld.scratch_samples = (uint64_t)N * nsf;
    envidr_field_out fo{};
    fo.sigma = B.s_sigma; fo.normal = B.s_normal;
    if (!opts->geometry_only) {
        fo.rgb = B.s_rgb;
        if (out->diffuse_image) fo.c_diffuse = B.s_cd;
        if (out->specular_image) fo.c_specular = B.s_cs;
        if (out->roughness_image) fo.roughness = B.s_rough;
    }
    // The loop alternates between these small kernels and k_geom_tc (150 KB of dynamic shared memory per CTA).  Asking for the maximum
    // shared-memory carve-out for the small ones as well spares the SMs a shared-memory / L1 reconfiguration at every kernel boundary
    // (ENVIDR_LOOP_CARVEOUT=0 keeps the driver's default choice).
    static int carve = -1;
    if (carve < 0) {
        const char* e = getenv("ENVIDR_LOOP_CARVEOUT");
        carve = (e && e[0] == '0') ? 0 : 1;
        if (carve) {
            cudaFuncSetAttribute(k_march_compact, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_composite_compact, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_render_init, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_render_finish, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
    }
    const uint32_t march_grid = min(ceil_div(N, kMarchBlock), (uint32_t)kSMs * 8);
    const uint32_t max_iters = opts->max_steps;       // n_step >= 1 per iteration
    static const uint32_t batch = [] { const char* e = getenv("ENVIDR_LOOP_BATCH"); const int v = e ? atoi(e) : 0; return (uint32_t)(v >= 1 && v <= 64 ? v : 4); }();
    // iterations issued per host check: the host runs one batch ahead of the device, so a pass ends with up to 2 * batch - 1 no-op
    // iterations (3 launches each); measured 16.07 / 15.94 / 15.90 / 16.01 ms per frame for batch 2 / 3 / 4 / 8 (run 67)
    uint32_t it = 0, pending = 0;
    bool done = false;
    while (!done && it < max_iters) {
        for (uint32_t b = 0; b < batch && it < max_iters; b++, it++) {
            k_march_compact<<<march_grid, kMarchBlock, 0, st>>>(rays_o, rays_d, r_images, bitfield, opts->bound, opts->dt_gamma,
                                                               opts->max_steps, opts->cascade, opts->grid_size, noises, fast, B);
            const bool timed = g_timing && g_tev_used < kMaxTimed;
            int recorded = 0;
            if (timed) while (g_tev_created < 2 * (g_tev_used + 1)) cudaEventCreate(&g_tev[g_tev_created++]);
            rc = field_forward_launch(&fld, B.s_xyz, B.s_dir, r_images ? B.s_rimg : nullptr, &B.ctr->M, 0,
                                      opts->geometry_only ? (capture.rec ? 2 : 1) : 0, &fo, st, timed ? &g_tev[2 * g_tev_used] : nullptr,
                                      &recorded, capture.rec ? &capture : nullptr);
            if (rc) return rc;
            if (timed && recorded) g_tev_used++;
            g_launches += (fld.precision == 1 && !opts->geometry_only) ? 5 : 3;      // march, [geom, env, shade | field], composite
            k_composite_compact<<<march_grid, kMarchBlock, 0, st>>>(N, opts->T_thresh, opts->max_steps, opts->geometry_only,
                                                                   opts->input_alpha, nsf, nsc, B, O);
        }
        rc = check_launch("render_loop");
        if (rc) return rc;
        // snapshot n_alive / step_total after this batch; look at the snapshot of the PREVIOUS batch so the
        // host stays one batch ahead of the device and never stalls it
        const uint32_t slot = pending & 1;
        cudaMemcpyAsync(g_host_flag + 2 * slot, &B.ctr->n_alive, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(g_host_flag + 2 * slot + 1, &B.ctr->step_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        cudaEventRecord(g_events[slot], st);
        if (pending > 0) {
            const uint32_t prev = (pending - 1) & 1;
            cudaEventSynchronize(g_events[prev]);
            if (g_host_flag[2 * prev] == 0 || g_host_flag[2 * prev + 1] >= opts->max_steps) done = true;
        }
        pending++;
    }
    g_launches += 2;
    k_render_finish<<<ceil_div(N, 256), 256, 0, st>>>(N, opts->bg_color[0], opts->bg_color[1], opts->bg_color[2], bg_per_ray,
                                                      opts->geometry_only, O);
    cudaMemcpyAsync(g_host_flag + 8, &B.ctr->iters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(g_host_flag + 9, &B.ctr->total_samples_lo, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    cudaEventRecord(g_events[0], st);
    return check_launch("render_finish");
}

static int permute_log(const envidr_sample_log* log, uint64_t total, const int32_t* ray_offset, float* rec_out, float* sigma_out, float* delta_out,
                       int32_t* idx_out, envidr_stream_t stream);

int envidr_permute_sample_log(const envidr_sample_log* log, uint64_t total, const int32_t* ray_offset, float* rec_out,
                              float* sigma_out, float* delta_out, envidr_stream_t stream) {
    ENVIDR_REQUIRE(rec_out, ENVIDR_E_BADARG, "null pointer");
    return permute_log(log, total, ray_offset, rec_out, sigma_out, delta_out, nullptr, stream);
}

int envidr_permute_sample_log_index(const envidr_sample_log* log, uint64_t total, const int32_t* ray_offset, int32_t* idx_out,
                                    float* sigma_out, float* delta_out, envidr_stream_t stream) {
    ENVIDR_REQUIRE(idx_out && total <= 0x7FFFFFFFull, ENVIDR_E_BADARG, "null pointer / more than 2^31 log entries");
    return permute_log(log, total, ray_offset, nullptr, sigma_out, delta_out, idx_out, stream);
}

static int permute_log(const envidr_sample_log* log, uint64_t total, const int32_t* ray_offset, float* rec_out, float* sigma_out, float* delta_out,
                       int32_t* idx_out, envidr_stream_t stream) {
    ENVIDR_REQUIRE(log && ray_offset && sigma_out && delta_out, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(log->rec && log->sigma && log->delta && log->ray && log->seq, ENVIDR_E_BADARG, "sample log: null buffer");
    ENVIDR_REQUIRE(total <= log->capacity, ENVIDR_E_WORKSPACE, "sample log overflowed its capacity; the pass has to be marched again");
    if (total == 0) return 0;
    const uint32_t shift = rec_out ? 3u : 0u;
    const uint64_t threads = total << shift;
    k_permute_log<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(log->rec), log->sigma, reinterpret_cast<const float2*>(log->delta), log->ray, log->seq, total,
        ray_offset, reinterpret_cast<float4*>(rec_out), sigma_out, reinterpret_cast<float2*>(delta_out), idx_out, shift);
    g_launches += 1;
    return check_launch("permute_sample_log");
}

/* blocks until the last envidr_render_rays call of this process has finished; stats = {iterations, samples_lo, samples_hi, 0} */
int envidr_render_last_stats(uint32_t stats[4]) {
    ENVIDR_REQUIRE(stats, ENVIDR_E_BADARG, "null pointer");
    if (!g_events_ok) { stats[0] = stats[1] = stats[2] = stats[3] = 0; return 0; }
    cudaEventSynchronize(g_events[0]);
    stats[0] = g_host_flag[8]; stats[1] = g_host_flag[9]; stats[2] = g_host_flag[10]; stats[3] = 0;
    return 0;
}

/* Instrumentation for bench.py.  envidr_launch_count: kernels launched by envidr_render_rays so far in this process.
 * envidr_render_timing(1) makes every following field-kernel launch bracketed by CUDA events on its stream;
 * envidr_render_field_time synchronises, returns the summed duration (ms) and the number of timed launches, and resets. */
uint64_t envidr_launch_count(void) { return g_launches; }
int envidr_render_timing(int enable) { g_timing = enable; g_tev_used = 0; return 0; }
int envidr_render_field_time(float* total_ms, uint32_t* launches) {
    ENVIDR_REQUIRE(total_ms && launches, ENVIDR_E_BADARG, "null pointer");
    float sum = 0.f;
    for (int i = 0; i < g_tev_used; i++) {
        cudaEventSynchronize(g_tev[2 * i + 1]);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_tev[2 * i], g_tev[2 * i + 1]) == cudaSuccess) sum += ms;
    }
    (void)cudaGetLastError();
    *total_ms = sum; *launches = (uint32_t)g_tev_used;
    g_tev_used = 0;
    return 0;
}

}  // extern "C"
