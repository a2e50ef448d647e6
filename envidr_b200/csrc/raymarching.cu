// raymarching.cu -- occupancy-grid march, alpha compositing and support ops for sm_100a.
//
// Replaces the reference's `_raymarching` extension (raymarching/src/raymarching.cu, 11 entry points;
// raymarching/src/raymarching.h:7-18).  Integer results (Morton codes, bit fields, scatter indices,
// per-ray sample counts) and march sample positions are bit-identical to the reference; composited
// floats agree to rounding (the train compositor scans a ray with one warp instead of one thread).
#include <float.h>
#include "common.cuh"

namespace envidr {

constexpr int kRayBlock = 128;

// ------------------------------------------------------------------------------------------------
// support ops
// ------------------------------------------------------------------------------------------------

// reference: kernel_near_far_from_aabb (raymarching.cu:91-145) -- slab test, 1 thread / ray
__global__ void __launch_bounds__(kRayBlock) k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                       const float* __restrict__ aabb, uint32_t N, float min_near,
                                                       float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
    const float rdx = 1 / rays_d[3 * n], rdy = 1 / rays_d[3 * n + 1], rdz = 1 / rays_d[3 * n + 2];
    float lo = (aabb[0] - ox) * rdx, hi = (aabb[3] - ox) * rdx;
    if (lo > hi) { float s = lo; lo = hi; hi = s; }
    float lo2 = (aabb[1] - oy) * rdy, hi2 = (aabb[4] - oy) * rdy;
    if (lo2 > hi2) { float s = lo2; lo2 = hi2; hi2 = s; }
    bool miss = (lo > hi2) || (lo2 > hi);
    if (!miss) {
        if (lo2 > lo) lo = lo2;
        if (hi2 < hi) hi = hi2;
        lo2 = (aabb[2] - oz) * rdz; hi2 = (aabb[5] - oz) * rdz;
        if (lo2 > hi2) { float s = lo2; lo2 = hi2; hi2 = s; }
        miss = (lo > hi2) || (lo2 > hi);
        if (!miss) {
            if (lo2 > lo) lo = lo2;
            if (hi2 < hi) hi = hi2;
            if (lo < min_near) lo = min_near;
        }
    }
    nears[n] = miss ? FLT_MAX : lo;
    fars[n] = miss ? FLT_MAX : hi;
}

// reference: kernel_sph_from_ray (raymarching.cu:163-198)
__global__ void __launch_bounds__(kRayBlock) k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                           float radius, uint32_t N, float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
    const float dx = rays_d[3 * n], dy = rays_d[3 * n + 1], dz = rays_d[3 * n + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float Cq = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * Cq)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float RPI = 0.3183098861837907f;
    coords[2 * n] = 2 * atan2f(sqrtf(x * x + z * z), y) * RPI - 1;
    coords[2 * n + 1] = atan2f(z, x) * RPI;
}

__global__ void __launch_bounds__(256) k_morton(const int32_t* __restrict__ coords, uint32_t N, int32_t* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int32_t)morton3((uint32_t)coords[3 * n], (uint32_t)coords[3 * n + 1], (uint32_t)coords[3 * n + 2]);
}

__global__ void __launch_bounds__(256) k_morton_invert(const int32_t* __restrict__ indices, uint32_t N, int32_t* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int32_t ind = indices[n];
    coords[3 * n + 0] = (int32_t)compact3((uint32_t)(ind >> 0));
    coords[3 * n + 1] = (int32_t)compact3((uint32_t)(ind >> 1));
    coords[3 * n + 2] = (int32_t)compact3((uint32_t)(ind >> 2));
}

// reference: kernel_packbits (raymarching.cu:267-289).  One warp packs 32 consecutive bytes: each lane
// loads its 8 cells as two float4 (coalesced 1 KB per warp) and emits one byte.
template <bool kAligned>
__global__ void __launch_bounds__(256) k_packbits(const float* __restrict__ grid, uint32_t N, float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float4 a, b;
    if (kAligned) {
        a = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)n);
        b = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)n + 1);
    } else {
        const float* p = grid + 8 * (size_t)n;
        a = make_float4(p[0], p[1], p[2], p[3]);
        b = make_float4(p[4], p[5], p[6], p[7]);
    }
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;   bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;   bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;  bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;  bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// reference: kernel_get_scatter_idx (raymarching.cu:302-322).  One warp per ray: lanes stride the
// ray's contiguous sample range, so stores are coalesced.
// Rays dropped by march_rays_train (offset + count > M; the reference kernel has no such guard and writes out of bounds) are skipped.
__global__ void __launch_bounds__(256) k_scatter_idx(const int32_t* __restrict__ rays, uint32_t N, uint32_t M, int32_t* __restrict__ idx_map) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= N) return;
    const uint32_t id = rays[3 * w], off = rays[3 * w + 1], cnt = rays[3 * w + 2];
    if (cnt == 0 || (uint64_t)off + cnt > M) return;
    for (uint32_t s = lane; s < cnt; s += 32) idx_map[off + s] = (int32_t)id;
}

// ------------------------------------------------------------------------------------------------
// training march: count -> scan -> write   (reference: kernel_march_rays_train, raymarching.cu:340-509)
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float perturbed_start(const Dda& s, float near, float noise) {
    float t0 = near;
    t0 += s.step_size(t0) * noise;
    return t0;
}

__global__ void __launch_bounds__(kRayBlock) k_march_train_count(
        const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid, float bound,
        float dt_gamma, uint32_t max_steps, uint32_t early_stop_steps, uint32_t N, uint32_t C, uint32_t H,
        const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
        int32_t* __restrict__ rays) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    Dda s; s.init(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, grid);
    const float far = fars[n];
    float t = perturbed_start(s, nears[n], noises[n]);
    uint32_t count = 0;
    float x, y, z, dt;
    while (t < far && count < early_stop_steps) {
        if (s.probe(t, x, y, z, dt)) { count++; t += dt; }
    }
    rays[3 * n] = (int32_t)n;
    rays[3 * n + 2] = (int32_t)count;
}

// Exclusive scan of rays[:,2] into rays[:,1] (offset by counter[0]); advances counter.  Three small launches: per-block scan
// (1,024 rays per block, block totals to a scratch array), scan of the block totals by one block, add-back.  A single block
// walking all rays (the first version) cost 1.2 ms for the 640,000 rays of a frame; rays are few thousand in a training step.
__device__ __forceinline__ uint32_t block_inclusive_scan_1024(uint32_t v, uint32_t* warp_sums) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += u;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = warp_sums[lane];
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= (uint32_t)o) ws += u;
        }
        warp_sums[lane] = ws;  // inclusive over warps
    }
    __syncthreads();
    return inc + (wid ? warp_sums[wid - 1] : 0u);
}

__global__ void __launch_bounds__(1024) k_scan_blocks(int32_t* __restrict__ rays, uint32_t N, uint32_t* __restrict__ block_tot) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t n = blockIdx.x * 1024 + threadIdx.x;
    const uint32_t v = n < N ? (uint32_t)rays[3 * n + 2] : 0u;
    const uint32_t inc = block_inclusive_scan_1024(v, warp_sums);
    if (n < N) rays[3 * n + 1] = (int32_t)(inc - v);             // offset inside the block
    if (threadIdx.x == 1023) block_tot[blockIdx.x] = inc;
}

// one block: exclusive scan of the block totals in place (chunks of 1,024 with a running carry that starts at counter[0])
__global__ void __launch_bounds__(1024) k_scan_tops(uint32_t* __restrict__ block_tot, uint32_t nb, uint32_t N, int32_t* __restrict__ counter) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = (uint32_t)counter[0];
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t b = base + threadIdx.x;
        const uint32_t v = b < nb ? block_tot[b] : 0u;
        const uint32_t inc = block_inclusive_scan_1024(v, warp_sums);
        const uint32_t c = carry;
        if (b < nb) block_tot[b] = c + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counter[0] = (int32_t)carry;
        counter[1] += (int32_t)N;
    }
}

__global__ void __launch_bounds__(1024) k_scan_add(int32_t* __restrict__ rays, uint32_t N, const uint32_t* __restrict__ block_tot) {
    const uint32_t n = blockIdx.x * 1024 + threadIdx.x;
    if (n < N) rays[3 * n + 1] += (int32_t)block_tot[blockIdx.x];
}

// scratch for the block totals: the operator surface has no workspace argument to carry it -> stream_scratch (error.cu)
static int scan_ray_counts(int32_t* rays, uint32_t N, int32_t* counter, cudaStream_t st) {
    const uint32_t nb = ceil_div(N, 1024);
    uint32_t* g_scan_scratch = static_cast<uint32_t*>(stream_scratch(kScratchScan, (size_t)nb * sizeof(uint32_t), kScratchScanBytes, st));
    if (!g_scan_scratch) return (int)cudaErrorMemoryAllocation;
    k_scan_blocks<<<nb, 1024, 0, st>>>(rays, N, g_scan_scratch);
    k_scan_tops<<<1, 1024, 0, st>>>(g_scan_scratch, nb, N, counter);
    k_scan_add<<<nb, 1024, 0, st>>>(rays, N, g_scan_scratch);
    return 0;
}

__global__ void __launch_bounds__(kRayBlock) k_march_train_write(
        const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid, float bound,
        float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
        const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
        const int32_t* __restrict__ rays, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t offset = (uint32_t)rays[3 * n + 1], count = (uint32_t)rays[3 * n + 2];
    if (count == 0 || offset + count > M) return;
    Dda s; s.init(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, grid);
    const float far = fars[n], near = nears[n];
    float t = perturbed_start(s, near, noises ? noises[n] : 0.0f);
    float last_t = near;
    float* px = xyzs + 3 * (size_t)offset;
    float* pd = dirs + 3 * (size_t)offset;
    float* pl = deltas + 2 * (size_t)offset;
    uint32_t step = 0;
    float x, y, z, dt;
    while (t < far && step < count) {
        if (s.probe(t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = s.dx; pd[1] = s.dy; pd[2] = s.dz;
            t += dt;
            pl[0] = dt; pl[1] = t - last_t;
            last_t = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// training march in ONE launch with deterministic, ray-ordered offsets: count -> decoupled look-back scan across blocks ->
// write.  The reference also marches every ray twice inside one kernel (raymarching.cu:340-509) but takes its offsets from
// atomicAdd (run-to-run different sample order); the three-kernel count / scan / write formulation of this library paid the
// bit-field lookups of the second march from L2 (cold L1, new launch) and was 0.6x the reference at 640,000 rays.  Here the
// second march follows the first in the same thread (its occupancy bytes are L1-hot), and the block prefix comes from the
// chained-scan protocol: blocks take a ticket (so every predecessor is resident or done), publish (flag, value) words --
// 1 = block aggregate, 2 = inclusive prefix -- and warp 0 looks back 32 predecessors at a time.
// ------------------------------------------------------------------------------------------------
template <bool kFast>       // kFast: one cascade, dt_gamma == 0 -> Dda::probe_fast (same samples, a third of the instructions)
__global__ void __launch_bounds__(kRayBlock) k_march_train_fused(
        const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid, float bound,
        float dt_gamma, uint32_t max_steps, uint32_t early_stop_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
        const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
        int32_t* __restrict__ rays, int32_t* __restrict__ counter, unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket,
        const int* __restrict__ box, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas) {
    __shared__ uint32_t s_bid, s_prefix, s_base;
    __shared__ uint32_t warp_sums[kRayBlock / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_base = (uint32_t)counter[0];                 // read before any block can have finished (see the last block below)
        s_bid = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    const uint32_t bid = s_bid, nb = gridDim.x;
    const uint32_t n = bid * kRayBlock + tid;
    // ---- pass 1: count ----
    Dda s;
    float far = 0.0f, near = 0.0f, t_start = 0.0f;
    uint32_t count = 0;
    if (n < N) {
        s.init(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, grid);
        far = fars[n]; near = nears[n];
        t_start = perturbed_start(s, near, noises[n]);
        float t = t_start, x, y, z, dt;
        if (kFast) {
            // box of the occupied cells: a ray that misses it emits nothing (the reference steps through empty cells up to `far`);
            // one that has left it for good stops there -- same samples, see Dda::probe_fast
            if (box) { s.init_fast(box); if (s.misses_box()) t = far; }
            else s.init_fast_full();
        }
        while (t < far && count < early_stop_steps) {
            if (kFast ? s.probe_fast(t, x, y, z, dt, far) : s.probe(t, x, y, z, dt)) { count++; t += dt; }
        }
    }
    // ---- block scan of the counts ----
    uint32_t inc = count;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += u;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    uint32_t warp_excl = 0, total = 0;
    #pragma unroll
    for (uint32_t w = 0; w < kRayBlock / 32; w++) {
        if (w < wid) warp_excl += warp_sums[w];
        total += warp_sums[w];
    }
    // ---- chained scan across blocks (warp 0) ----
    if (wid == 0) {
        volatile unsigned long long* st = status;
        if (lane == 0 && bid > 0) { st[bid] = (1ull << 32) | total; __threadfence(); }
        uint32_t sum = 0;
        int j = (int)bid - 1;                          // nearest predecessor; index -1 = the virtual block holding counter[0]
        while (true) {
            const int idx = j - (int)lane;
            unsigned long long v;
            if (idx >= 0) {
                do { v = st[idx]; } while ((v >> 32) == 0ull);
            } else {
                v = (2ull << 32) | (idx == -1 ? (unsigned long long)s_base : 0ull);
            }
            const uint32_t incl = __ballot_sync(0xffffffffu, (v >> 32) == 2ull);
            const int first = incl ? __ffs((int)incl) - 1 : 32;
            uint32_t c = ((int)lane <= first) ? (uint32_t)(v & 0xffffffffull) : 0u;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            sum += c;
            if (incl) break;
            j -= 32;
        }
        if (lane == 0) {
            __threadfence();
            st[bid] = (2ull << 32) | (unsigned long long)(sum + total);
            s_prefix = sum;
            if (bid == nb - 1) {                       // every block has read counter[0] by now (it did so before publishing)
                counter[0] = (int32_t)(sum + total);
                counter[1] += (int32_t)N;
            }
        }
    }
    __syncthreads();
    if (n >= N) return;
    const uint32_t offset = s_prefix + warp_excl + inc - count;
    rays[3 * n] = (int32_t)n; rays[3 * n + 1] = (int32_t)offset; rays[3 * n + 2] = (int32_t)count;
    if (count == 0 || offset + count > M) return;
    // ---- pass 2: write (the same march; identical arithmetic to k_march_train_write) ----
    float t = t_start, last_t = near;
    float* px = xyzs + 3 * (size_t)offset;
    float* pd = dirs + 3 * (size_t)offset;
    float* pl = deltas + 2 * (size_t)offset;
    uint32_t step = 0;
    float x, y, z, dt;
    while (t < far && step < count) {
        if (kFast ? s.probe_fast(t, x, y, z, dt, far) : s.probe(t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = s.dx; pd[1] = s.dy; pd[2] = s.dz;
            t += dt;
            pl[0] = dt; pl[1] = t - last_t;
            last_t = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
}

__global__ void __launch_bounds__(32) k_occ_box_reset(int* __restrict__ box) {
    if (threadIdx.x < 6) box[threadIdx.x] = (threadIdx.x & 1) ? -1 : 0x7fffffff;
}

// scratch of the chained scan: one 8-byte status word per block + the ticket + the occupied-cell box -> stream_scratch (error.cu)

// ------------------------------------------------------------------------------------------------
// replay of a pass with known per-ray sample counts (see envidr_march_rays_replay in the header)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_replay_counts(const int32_t* __restrict__ counts, uint32_t N, int32_t* __restrict__ rays) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    rays[3 * n] = (int32_t)n;
    rays[3 * n + 2] = counts[n] > 0 ? counts[n] : 0;
}

// dst[off + s, :] = src[n, :] for every sample s of ray n (per-ray rows -> per-sample rows; one warp per ray, float4 rows)
__global__ void __launch_bounds__(256) k_scatter_rows4(const int32_t* __restrict__ rays, uint32_t N, uint32_t M, const float4* __restrict__ src,
                                                      float4* __restrict__ dst) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= N) return;
    const uint32_t id = rays[3 * w], off = rays[3 * w + 1], cnt = rays[3 * w + 2];
    if (off + cnt > M) return;
    const float4 v = src[id];
    for (uint32_t s = lane; s < cnt; s += 32) dst[off + s] = v;
}

// dst[p, :] = src[idx[p], :] for rows of `vec4` float4s (frame assembly after the all-gather of a sharded render: rows of 32 B per ray in
// rank order -> pixel order).  torch's index_select moves these 82 MB at 62 GB/s (1.33 ms per 1600x1600 frame, run r3_27); one float4 per
// thread moves them at HBM speed.
__global__ void __launch_bounds__(256) k_gather_rows(const float4* __restrict__ src, const int32_t* __restrict__ idx, uint64_t n_rows, uint32_t vec4,
                                                    float4* __restrict__ dst) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t row = t / vec4;
    if (row >= n_rows) return;
    const uint32_t part = (uint32_t)(t - row * vec4);
    dst[t] = __ldg(src + (size_t)idx[row] * vec4 + part);
}

// Inference compositor (reference kernel_composite_rays, raymarching.cu:957-1046) over ray-contiguous samples: the pre-sample
// transmittance T = 1 - sum w, w = alpha T, every image accumulated sample by sample in the same order.  One thread per ray.
__global__ void __launch_bounds__(128) k_composite_replay(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                         const float* __restrict__ normals, const float* __restrict__ cds,
                                                         const float* __restrict__ css, const float* __restrict__ roughs,
                                                         const float* __restrict__ deltas, const int32_t* __restrict__ rays,
                                                         const float* __restrict__ nears, uint32_t M,
                                                         uint32_t N, float T_thresh, uint32_t input_alpha, float* __restrict__ weights_sum,
                                                         float* __restrict__ depth, float* __restrict__ image, float* __restrict__ normal_image,
                                                         float* __restrict__ diffuse_image, float* __restrict__ specular_image,
                                                         float* __restrict__ roughness_image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * n], off = (uint32_t)rays[3 * n + 1];
    uint32_t cnt = (uint32_t)rays[3 * n + 2];
    if (off + cnt > M) cnt = 0;                                    // dropped ray (capacity), as in march_rays_train
    float ws = 0, d = 0, t = nears ? nears[index] : 0.0f, acc[3] = {0, 0, 0}, nrm[3] = {0, 0, 0}, cd[3] = {0, 0, 0}, cs[3] = {0, 0, 0}, rgh = 0;
    for (uint32_t s = 0; s < cnt; s++) {
        const size_t m = (size_t)off + s;
        const float2 dl = *reinterpret_cast<const float2*>(deltas + 2 * m);
        if (dl.x == 0) break;
        const float sg = sigmas[m];
        const float alpha = input_alpha ? 0.0f + sg : 1.0f - __expf(-sg * dl.x);
        const float T = 1 - ws;
        const float w = alpha * T;
        ws += w;
        t = t + dl.y;
        d += w * t;
        const float* c = rgbs + 3 * m;
        acc[0] += w * c[0]; acc[1] += w * c[1]; acc[2] += w * c[2];
        if (normal_image) { const float* q = normals + 3 * m; nrm[0] += w * q[0]; nrm[1] += w * q[1]; nrm[2] += w * q[2]; }
        if (diffuse_image) { const float* q = cds + 3 * m; cd[0] += w * q[0]; cd[1] += w * q[1]; cd[2] += w * q[2]; }
        if (specular_image) { const float* q = css + 3 * m; cs[0] += w * q[0]; cs[1] += w * q[1]; cs[2] += w * q[2]; }
        if (roughness_image) rgh += w * roughs[m];
        if (T < T_thresh) break;
    }
    weights_sum[index] = ws;
    if (depth) depth[index] = d;
    image[3 * index] = acc[0]; image[3 * index + 1] = acc[1]; image[3 * index + 2] = acc[2];
    if (normal_image) { float* q = normal_image + 3 * index; q[0] = nrm[0]; q[1] = nrm[1]; q[2] = nrm[2]; }
    if (diffuse_image) { float* q = diffuse_image + 3 * index; q[0] = cd[0]; q[1] = cd[1]; q[2] = cd[2]; }
    if (specular_image) { float* q = specular_image + 3 * index; q[0] = cs[0]; q[1] = cs[1]; q[2] = cs[2]; }
    if (roughness_image) roughness_image[index] = rgh;
}

// ------------------------------------------------------------------------------------------------
// training compositor: one warp per ray, 32 samples per pass, shuffle scan of the transmittance
// (reference: kernel_composite_rays_train_forward[_with_weight], raymarching.cu:529-701)
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float warp_sum(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive product scan
__device__ __forceinline__ float warp_scan_mul(float v, uint32_t lane) {
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v *= u;
    }
    return v;
}
// inclusive sum scan
__device__ __forceinline__ float warp_scan_add(float v, uint32_t lane) {
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += u;
    }
    return v;
}

template <bool kWriteWeights>
__global__ void __launch_bounds__(256) k_composite_train_fwd(
        const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
        const int32_t* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh, uint32_t accum_deltas, uint32_t input_alpha,
        float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image, float* __restrict__ weights) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= N) return;
    const uint32_t index = rays[3 * w], offset = rays[3 * w + 1], count = rays[3 * w + 2];
    if (count == 0 || offset + count > M) {
        if (lane == 0) { weights_sum[index] = 0; depth[index] = 0; }
        if (lane < 3) image[3 * index + lane] = 0;
        return;
    }
    float T_in = 1.0f, t_in = 0.0f;       // transmittance / accumulated depth parameter entering this pass
    float r = 0, g = 0, b = 0, ws = 0, d = 0;
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t s = base + lane;
        const bool live = s < count;
        const size_t i = (size_t)offset + s;
        float sg = 0, d0 = 0, d1 = 0, cr = 0, cg = 0, cb = 0;
        if (live) {
            sg = sigmas[i];
            const float2 dl = *reinterpret_cast<const float2*>(deltas + 2 * i);
            d0 = dl.x; d1 = dl.y;
            cr = rgbs[3 * i]; cg = rgbs[3 * i + 1]; cb = rgbs[3 * i + 2];
        }
        const float alpha = live ? (input_alpha ? 0.0f + sg : 1.0f - __expf(-sg * d0)) : 0.0f;
        const float T_incl = T_in * warp_scan_mul(1.0f - alpha, lane);       // T after this sample
        float T_before = __shfl_up_sync(0xffffffffu, T_incl, 1);
        if (lane == 0) T_before = T_in;
        const float t_incl = accum_deltas ? t_in + warp_scan_add(live ? d1 : 0.0f, lane) : d1;
        // the reference stops after the first sample whose updated T drops below T_thresh
        const uint32_t stop_mask = __ballot_sync(0xffffffffu, live && (T_incl < T_thresh));
        const uint32_t last = stop_mask ? (uint32_t)(__ffs(stop_mask) - 1) : 31u;
        const bool use = live && lane <= last;
        const float wgt = use ? alpha * T_before : 0.0f;
        if (kWriteWeights && use) weights[i] = wgt;
        r += wgt * cr; g += wgt * cg; b += wgt * cb;
        d += wgt * t_incl; ws += wgt;
        if (stop_mask) break;
        T_in = __shfl_sync(0xffffffffu, T_incl, 31);
        t_in = __shfl_sync(0xffffffffu, t_incl, 31);
    }
    r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); d = warp_sum(d); ws = warp_sum(ws);
    if (lane == 0) {
        weights_sum[index] = ws; depth[index] = d;
        image[3 * index] = r; image[3 * index + 1] = g; image[3 * index + 2] = b;
    }
}

// reference: kernel_composite_rays_train_backward (raymarching.cu:731-821).  Same warp-per-ray scan;
// prefix sums of (r,g,b,d) are exclusive-of-nothing (they include the current sample, as in the
// reference).  Reference quirk kept: depth / grad_depth are read at element 0 for every ray.
__global__ void __launch_bounds__(256) k_composite_train_bwd(
        const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image, const float* __restrict__ grad_depth,
        const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
        const int32_t* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ image,
        const float* __restrict__ depth, uint32_t M, uint32_t N, float T_thresh,
        float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, uint32_t accum_deltas, uint32_t input_alpha) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= N) return;
    const uint32_t index = rays[3 * w], offset = rays[3 * w + 1], count = rays[3 * w + 2];
    if (count == 0 || offset + count > M) return;
    const float gi0 = grad_image[3 * index], gi1 = grad_image[3 * index + 1], gi2 = grad_image[3 * index + 2];
    const float gws = grad_weights_sum[index], gd = grad_depth[0];
    const float r_final = image[3 * index], g_final = image[3 * index + 1], b_final = image[3 * index + 2];
    const float ws_final = weights_sum[index], d_final = depth[0];
    float T_in = 1.0f, t_in = 0.0f, r_in = 0, g_in = 0, b_in = 0, d_in = 0;
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t s = base + lane;
        const bool live = s < count;
        const size_t i = (size_t)offset + s;
        float sg = 0, d0 = 0, d1 = 0, cr = 0, cg = 0, cb = 0;
        if (live) {
            sg = sigmas[i];
            const float2 dl = *reinterpret_cast<const float2*>(deltas + 2 * i);
            d0 = dl.x; d1 = dl.y;
            cr = rgbs[3 * i]; cg = rgbs[3 * i + 1]; cb = rgbs[3 * i + 2];
        }
        const float alpha = live ? (input_alpha ? 0.0f + sg : 1.0f - __expf(-sg * d0)) : 0.0f;
        const float T_incl = T_in * warp_scan_mul(1.0f - alpha, lane);
        float T_before = __shfl_up_sync(0xffffffffu, T_incl, 1);
        if (lane == 0) T_before = T_in;
        const float t_incl = accum_deltas ? t_in + warp_scan_add(live ? d1 : 0.0f, lane) : d1;
        const uint32_t stop_mask = __ballot_sync(0xffffffffu, live && (T_incl < T_thresh));
        const uint32_t last = stop_mask ? (uint32_t)(__ffs(stop_mask) - 1) : 31u;
        const bool use = live && lane <= last;
        const float wgt = use ? alpha * T_before : 0.0f;
        const float r = r_in + warp_scan_add(wgt * cr, lane);
        const float g = g_in + warp_scan_add(wgt * cg, lane);
        const float b = b_in + warp_scan_add(wgt * cb, lane);
        const float d = d_in + warp_scan_add(wgt * t_incl, lane);
        if (use) {
            const float gscale = input_alpha ? (1.0f / (1.0f - alpha + 1e-4f)) : d0;
            grad_rgbs[3 * i] = gi0 * wgt; grad_rgbs[3 * i + 1] = gi1 * wgt; grad_rgbs[3 * i + 2] = gi2 * wgt;
            grad_sigmas[i] = gscale * (gi0 * (T_incl * cr - (r_final - r)) + gi1 * (T_incl * cg - (g_final - g)) +
                                       gi2 * (T_incl * cb - (b_final - b)) + gd * (T_incl * t_incl - (d_final - d)) +
                                       gws * (1 - ws_final));
        }
        if (stop_mask) break;
        T_in = __shfl_sync(0xffffffffu, T_incl, 31);
        t_in = __shfl_sync(0xffffffffu, t_incl, 31);
        r_in = __shfl_sync(0xffffffffu, r, 31); g_in = __shfl_sync(0xffffffffu, g, 31);
        b_in = __shfl_sync(0xffffffffu, b, 31); d_in = __shfl_sync(0xffffffffu, d, 31);
    }
}

// ------------------------------------------------------------------------------------------------
// inference march / composite (reference: kernel_march_rays :839-944, kernel_composite_rays :957-1046)
// ------------------------------------------------------------------------------------------------

// One thread marches one alive ray for <= n_step samples into shared memory; the block then streams
// its contiguous [128*n_step] slot range to global memory with coalesced stores (zeros past the ray end).
template <int kMaxStep>
__global__ void __launch_bounds__(kRayBlock) k_march_infer(
        uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive, const float* __restrict__ rays_t,
        const float* __restrict__ rays_o, const float* __restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps,
        uint32_t C, uint32_t H, const uint8_t* __restrict__ grid, const float* __restrict__ fars,
        float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, const float* __restrict__ noises) {
    extern __shared__ float smem[];
    float* s_xyz = smem;                                   // [128*n_step*3]
    float* s_dir = s_xyz + kRayBlock * n_step * 3;         // [128*n_step*3]
    float* s_del = s_dir + kRayBlock * n_step * 3;         // [128*n_step*2]
    const uint32_t n = blockIdx.x * kRayBlock + threadIdx.x;
    const uint32_t slot0 = threadIdx.x * n_step;
    uint32_t step = 0;
    if (n < n_alive) {
        const int index = rays_alive[n];
        Dda s; s.init(rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, dt_gamma, max_steps, C, H, grid);
        const float far = fars[index];
        float t = rays_t[index];
        float last_t = t;
        t += s.step_size(t) * noises[n];
        float x, y, z, dt;
        while (t < far && step < n_step) {
            if (s.probe(t, x, y, z, dt)) {
                const uint32_t k = slot0 + step;
                s_xyz[3 * k] = x; s_xyz[3 * k + 1] = y; s_xyz[3 * k + 2] = z;
                s_dir[3 * k] = s.dx; s_dir[3 * k + 1] = s.dy; s_dir[3 * k + 2] = s.dz;
                t += dt;
                s_del[2 * k] = dt; s_del[2 * k + 1] = t - last_t;
                last_t = t;
                step++;
            }
        }
    }
    for (uint32_t k = slot0 + step; k < slot0 + n_step; k++) {
        s_xyz[3 * k] = 0; s_xyz[3 * k + 1] = 0; s_xyz[3 * k + 2] = 0;
        s_dir[3 * k] = 0; s_dir[3 * k + 1] = 0; s_dir[3 * k + 2] = 0;
        s_del[2 * k] = 0; s_del[2 * k + 1] = 0;
    }
    __syncthreads();
    const size_t blk_slot = (size_t)blockIdx.x * kRayBlock * n_step;
    const uint32_t rays_here = min((uint32_t)kRayBlock, n_alive - blockIdx.x * kRayBlock);
    const uint32_t n3 = rays_here * n_step * 3, n2 = rays_here * n_step * 2;
    for (uint32_t k = threadIdx.x; k < n3; k += kRayBlock) {
        xyzs[blk_slot * 3 + k] = s_xyz[k];
        dirs[blk_slot * 3 + k] = s_dir[k];
    }
    for (uint32_t k = threadIdx.x; k < n2; k += kRayBlock) deltas[blk_slot * 2 + k] = s_del[k];
}

__global__ void __launch_bounds__(kRayBlock) k_composite_infer(
        uint32_t n_alive, uint32_t n_step, float T_thresh, uint32_t accum_deltas, uint32_t input_alpha,
        int32_t* __restrict__ rays_alive, float* __restrict__ rays_t, const float* __restrict__ sigmas,
        const float* __restrict__ rgbs, const float* __restrict__ deltas,
        float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float* sg = sigmas + (size_t)n * n_step;
    const float* cl = rgbs + (size_t)n * n_step * 3;
    const float* dl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    float ws = weights_sum[index], d = depth[index];
    float r = image[3 * index], g = image[3 * index + 1], b = image[3 * index + 2];
    uint32_t step = 0;
    while (step < n_step) {
        const float d0 = dl[0];
        if (d0 == 0) break;
        const float alpha = input_alpha ? 0.0f + sg[0] : 1.0f - __expf(-sg[0] * d0);
        const float T = 1 - ws;
        const float wgt = alpha * T;
        ws += wgt;
        t = accum_deltas ? t + dl[1] : dl[1];
        d += wgt * t;
        r += wgt * cl[0]; g += wgt * cl[1]; b += wgt * cl[2];
        if (T < T_thresh) break;
        sg++; cl += 3; dl += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = ws; depth[index] = d;
    image[3 * index] = r; image[3 * index + 1] = g; image[3 * index + 2] = b;
}

}  // namespace envidr

using namespace envidr;

extern "C" {

int envidr_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                              float* nears, float* fars, envidr_stream_t stream) {
    ENVIDR_REQUIRE(rays_o && rays_d && aabb && nears && fars, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    k_near_far<<<ceil_div(N, kRayBlock), kRayBlock, 0, as_stream(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    return check_launch("near_far_from_aabb");
}

int envidr_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, envidr_stream_t stream) {
    ENVIDR_REQUIRE(rays_o && rays_d && coords, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    k_sph_from_ray<<<ceil_div(N, kRayBlock), kRayBlock, 0, as_stream(stream)>>>(rays_o, rays_d, radius, N, coords);
    return check_launch("sph_from_ray");
}

int envidr_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, envidr_stream_t stream) {
    ENVIDR_REQUIRE(coords && indices, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    k_morton<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(coords, N, indices);
    return check_launch("morton3D");
}

int envidr_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, envidr_stream_t stream) {
    ENVIDR_REQUIRE(coords && indices, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    k_morton_invert<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(indices, N, coords);
    return check_launch("morton3D_invert");
}

int envidr_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, envidr_stream_t stream) {
    ENVIDR_REQUIRE(grid && bitfield, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(grid) & 15) == 0)
        k_packbits<true><<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(grid, N, density_thresh, bitfield);
    else
        k_packbits<false><<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(grid, N, density_thresh, bitfield);
    return check_launch("packbits");
}

int envidr_get_scatter_idx(const int32_t* rays, uint32_t N, uint32_t M, int32_t* idx_map, envidr_stream_t stream) {
    ENVIDR_REQUIRE(rays && idx_map, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    k_scatter_idx<<<ceil_div(N, 8), 256, 0, as_stream(stream)>>>(rays, N, M, idx_map);
    return check_launch("get_scatter_idx");
}

int envidr_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                            uint32_t max_steps, uint32_t early_stop_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                            const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays,
                            int32_t* counter, const float* noises, envidr_stream_t stream) {
    ENVIDR_REQUIRE(rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas && rays && counter && noises,
                   ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(C >= 1 && C <= 8 && H >= 1 && H <= 1024, ENVIDR_E_UNSUPPORTED, "cascades must be 1..8, grid size <= 1024");
    if (N == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const uint32_t nb = ceil_div(N, kRayBlock);
    // (allocated once per device and stream by the warm-up calls that precede a CUDA-graph capture; never moved afterwards)
    unsigned long long* g_chain = static_cast<unsigned long long*>(
        stream_scratch(kScratchChain, (size_t)(nb + 8) * sizeof(unsigned long long), kScratchChainBytes, st));
    if (!g_chain) return (int)cudaErrorMemoryAllocation;
    cudaMemsetAsync(g_chain, 0, (size_t)(nb + 1) * sizeof(unsigned long long), st);
    uint32_t* ticket = reinterpret_cast<uint32_t*>(g_chain + nb);
    int* box = reinterpret_cast<int*>(g_chain + nb + 1);
    const bool fast = C == 1 && dt_gamma == 0.0f && H <= 256;
    const bool with_box = fast && H % 4 == 0 && (reinterpret_cast<uintptr_t>(grid) & 3) == 0;
    if (with_box) {
        k_occ_box_reset<<<1, 32, 0, st>>>(box);
        k_occ_box<<<kSMs, 256, 0, st>>>(grid, H * H * H / 32, box);
        g_launches += 2;
    }
    if (fast)
        k_march_train_fused<true><<<nb, kRayBlock, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, early_stop_steps, N, C, H, M, nears,
                                                            fars, noises, rays, counter, g_chain, ticket, with_box ? box : nullptr, xyzs, dirs, deltas);
    else
        k_march_train_fused<false><<<nb, kRayBlock, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, early_stop_steps, N, C, H, M, nears,
                                                             fars, noises, rays, counter, g_chain, ticket, nullptr, xyzs, dirs, deltas);
    g_launches += 1;
    return check_launch("march_rays_train");
}

int envidr_march_rays_replay(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                             const float* fars, const int32_t* counts, float* xyzs, float* dirs, float* deltas,
                             int32_t* rays, int32_t* counter, envidr_stream_t stream) {
    ENVIDR_REQUIRE(rays_o && rays_d && grid && nears && fars && counts && xyzs && dirs && deltas && rays && counter,
                   ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(C >= 1 && C <= 8 && H >= 1 && H <= 1024, ENVIDR_E_UNSUPPORTED, "cascades must be 1..8, grid size <= 1024");
    if (N == 0) return 0;
    cudaStream_t st = as_stream(stream);
    k_replay_counts<<<ceil_div(N, 256), 256, 0, st>>>(counts, N, rays);
    { const int rc = scan_ray_counts(rays, N, counter, st); if (rc) return rc; }
    k_march_train_write<<<ceil_div(N, kRayBlock), kRayBlock, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M,
                                                                      nears, fars, nullptr, rays, xyzs, dirs, deltas);
    g_launches += 3;
    return check_launch("march_rays_replay");
}

int envidr_scatter_ray_rows4(const int32_t* rays, uint32_t N, uint32_t M, const float* src, float* dst, envidr_stream_t stream) {
    ENVIDR_REQUIRE(rays && src && dst, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    k_scatter_rows4<<<ceil_div(N, 8), 256, 0, as_stream(stream)>>>(rays, N, M, reinterpret_cast<const float4*>(src),
                                                                   reinterpret_cast<float4*>(dst));
    g_launches += 1;
    return check_launch("scatter_ray_rows4");
}

int envidr_gather_rows(const float* src, const int32_t* idx, uint64_t n_rows, uint32_t row_floats, float* dst, envidr_stream_t stream) {
    ENVIDR_REQUIRE(src && idx && dst, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(row_floats > 0 && row_floats % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                   ENVIDR_E_UNSUPPORTED, "rows must be whole float4s, 16-byte aligned");
    if (n_rows == 0) return 0;
    const uint32_t vec4 = row_floats / 4;
    const uint64_t threads = n_rows * vec4;
    k_gather_rows<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(src), idx, n_rows, vec4,
                                                                                 reinterpret_cast<float4*>(dst));
    g_launches += 1;
    return check_launch("gather_rows");
}

int envidr_composite_rays_replay(const float* sigmas, const float* rgbs, const float* normals, const float* c_diffuse,
                                 const float* c_specular, const float* roughness, const float* deltas, const int32_t* rays,
                                 const float* nears, uint32_t M, uint32_t N, float T_thresh, uint32_t input_alpha, float* weights_sum,
                                 float* depth, float* image, float* normal_image, float* diffuse_image,
                                 float* specular_image, float* roughness_image, envidr_stream_t stream) {
    ENVIDR_REQUIRE(sigmas && rgbs && deltas && rays && weights_sum && image, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE((!normal_image || normals) && (!diffuse_image || c_diffuse) && (!specular_image || c_specular) &&
                   (!roughness_image || roughness), ENVIDR_E_BADARG, "an output image needs its per-sample input");
    if (N == 0) return 0;
    k_composite_replay<<<ceil_div(N, 128), 128, 0, as_stream(stream)>>>(sigmas, rgbs, normals, c_diffuse, c_specular, roughness, deltas,
                                                                        rays, nears, M, N, T_thresh, input_alpha, weights_sum, depth, image,
                                                                        normal_image, diffuse_image, specular_image, roughness_image);
    g_launches += 1;
    return check_launch("composite_rays_replay");
}

int envidr_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                        uint32_t M, uint32_t N, float T_thresh, uint32_t accum_deltas, uint32_t input_alpha,
                                        float* weights_sum, float* depth, float* image, float* weights, envidr_stream_t stream) {
    ENVIDR_REQUIRE(sigmas && rgbs && deltas && rays && weights_sum && depth && image, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    const uint32_t blocks = ceil_div(N, 8);
    if (weights)
        k_composite_train_fwd<true><<<blocks, 256, 0, as_stream(stream)>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, accum_deltas,
                                                                          input_alpha, weights_sum, depth, image, weights);
    else
        k_composite_train_fwd<false><<<blocks, 256, 0, as_stream(stream)>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, accum_deltas,
                                                                           input_alpha, weights_sum, depth, image, nullptr);
    return check_launch("composite_rays_train_forward");
}

int envidr_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* grad_depth,
                                         const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                         const float* weights_sum, const float* image, const float* depth, uint32_t M, uint32_t N,
                                         float T_thresh, float* grad_sigmas, float* grad_rgbs, uint32_t accum_deltas,
                                         uint32_t input_alpha, envidr_stream_t stream) {
    ENVIDR_REQUIRE(grad_weights_sum && grad_image && grad_depth && sigmas && rgbs && deltas && rays && weights_sum && image &&
                   depth && grad_sigmas && grad_rgbs, ENVIDR_E_BADARG, "null pointer");
    if (N == 0) return 0;
    k_composite_train_bwd<<<ceil_div(N, 8), 256, 0, as_stream(stream)>>>(grad_weights_sum, grad_image, grad_depth, sigmas, rgbs, deltas,
                                                                        rays, weights_sum, image, depth, M, N, T_thresh, grad_sigmas,
                                                                        grad_rgbs, accum_deltas, input_alpha);
    return check_launch("composite_rays_train_backward");
}

int envidr_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                      const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                      const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                      const float* noises, envidr_stream_t stream) {
    (void)nears;
    ENVIDR_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas && noises,
                   ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(C >= 1 && C <= 8 && H >= 1 && H <= 1024, ENVIDR_E_UNSUPPORTED, "cascades must be 1..8, grid size <= 1024");
    ENVIDR_REQUIRE(n_step >= 1 && n_step <= 32, ENVIDR_E_UNSUPPORTED, "n_step must be 1..32");
    if (n_alive == 0) return 0;
    const size_t smem = (size_t)kRayBlock * n_step * 8 * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_march_infer<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRayBlock * 32 * 8 * (int)sizeof(float));
        attr_set = true;
    }
    k_march_infer<32><<<ceil_div(n_alive, kRayBlock), kRayBlock, smem, as_stream(stream)>>>(
        n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, fars, xyzs, dirs, deltas, noises);
    return check_launch("march_rays");
}

int envidr_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, uint32_t accum_deltas, uint32_t input_alpha,
                          int32_t* rays_alive, float* rays_t, const float* sigmas, const float* rgbs, const float* deltas,
                          float* weights_sum, float* depth, float* image, envidr_stream_t stream) {
    ENVIDR_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image, ENVIDR_E_BADARG, "null pointer");
    if (n_alive == 0) return 0;
    k_composite_infer<<<ceil_div(n_alive, kRayBlock), kRayBlock, 0, as_stream(stream)>>>(
        n_alive, n_step, T_thresh, accum_deltas, input_alpha, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image);
    return check_launch("composite_rays");
}

}  // extern "C"
