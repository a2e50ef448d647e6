// common.cuh -- shared helpers for libenvidr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/envidr_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libenvidr_b200 is written for sm_100a (B200) only"
#endif

namespace envidr {
extern uint64_t g_launches;                 // kernels launched through the library (bench.py gpu_launches), render.cu
cudaEvent_t* timing_acquire();              // render.cu: CUDA-event timing of the dominant field kernel
void timing_commit();


void set_error(const char* fmt, ...);
// error.cu: fixed-size scratch slot per (device, stream, kind) out of one never-freed allocation per device and kind; nullptr
// (+ last error) when bytes > fixed_bytes, the one-time allocation fails, or more than 8 streams ask
void* stream_scratch(int kind, size_t bytes, size_t fixed_bytes, cudaStream_t st);
constexpr int kScratchScan = 0, kScratchChain = 1;
constexpr size_t kScratchScanBytes = 256u << 10, kScratchChainBytes = 2u << 20;   // N <= 64 M / 32 M rays

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define ENVIDR_REQUIRE(cond, code, msg)              \
    do {                                             \
        if (!(cond)) {                               \
            ::envidr::set_error("%s: %s", __func__, msg); \
            return code;                             \
        }                                            \
    } while (0)

constexpr int kSMs = 148;  // B200

inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline cudaStream_t as_stream(envidr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// 10-bit-per-axis Morton code (reference semantics: raymarching/src/raymarching.cu:56-81)
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// ---------------------------------------------------------------------------------------------
// Occupancy-grid DDA, one ray per thread.  Restates the arithmetic of the reference march
// (raymarching/src/raymarching.cu:363-428, 869-942) operation by operation so that sample
// positions, counts and occupancy decisions are bit-identical under nvcc's default FMA contraction.
// ---------------------------------------------------------------------------------------------
struct Dda {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH, Hf;
    float H3f;
    uint32_t C, H;
    const uint8_t* __restrict__ grid;

    __device__ __forceinline__ void init(const float* __restrict__ o, const float* __restrict__ d, float bound_,
                                         float dt_gamma_, uint32_t max_steps, uint32_t C_, uint32_t H_,
                                         const uint8_t* __restrict__ grid_) {
        ox = o[0]; oy = o[1]; oz = o[2];
        dx = d[0]; dy = d[1]; dz = d[2];
        rdx = 1 / dx; rdy = 1 / dy; rdz = 1 / dz;
        bound = bound_; dt_gamma = dt_gamma_;
        C = C_; H = H_; grid = grid_;
        Hf = (float)H_;
        rH = 1 / (float)H_;
        H3f = (float)(H_ * H_ * H_);
        dt_min = 2 * 1.7320508075688772f / max_steps;
        dt_max = 2 * 1.7320508075688772f * (1 << (C_ - 1)) / H_;
    }

    __device__ __forceinline__ float step_size(float t) const { return clampf(t * dt_gamma, dt_min, dt_max); }

    __device__ __forceinline__ int level_of(float x, float y, float z, float dt) const {
        const float maxc = (float)C;
        int e_pos, e_dt;
        frexpf(fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))), &e_pos);
        const int lv_pos = (int)fminf(maxc - 1, fmaxf(0.0f, (float)e_pos));
        frexpf(dt * Hf * 0.5f, &e_dt);
        const int lv_dt = (int)fminf(maxc - 1, fmaxf(0.0f, (float)e_dt));
        return max(lv_pos, lv_dt);
    }

    // Probe the grid at parameter t.  Occupied: returns true with (x,y,z,dt) of the sample, t unchanged.
    // Empty: returns false and advances t past the voxel in dt-sized steps.
    __device__ __forceinline__ bool probe(float& t, float& x, float& y, float& z, float& dt) const {
        x = clampf(ox + t * dx, -bound, bound);
        y = clampf(oy + t * dy, -bound, bound);
        z = clampf(oz + t * dz, -bound, bound);
        dt = step_size(t);
        const int level = level_of(x, y, z, dt);
        const float mip_bound = fminf(scalbnf(1.0f, level), bound);
        const float mip_rbound = 1 / mip_bound;
        const int nx = (int)clampf(0.5f * (x * mip_rbound + 1) * Hf, 0.0f, (float)(H - 1));
        const int ny = (int)clampf(0.5f * (y * mip_rbound + 1) * Hf, 0.0f, (float)(H - 1));
        const int nz = (int)clampf(0.5f * (z * mip_rbound + 1) * Hf, 0.0f, (float)(H - 1));
        const uint32_t index = (uint32_t)(level * H3f + (float)morton3(nx, ny, nz));
        const bool occ = grid[index >> 3] & (1u << (index & 7u));
        if (occ) return true;
        const float tx = (((nx + 0.5f + 0.5f * copysignf(1.0f, dx)) * rH * 2 - 1) * mip_bound - x) * rdx;
        const float ty = (((ny + 0.5f + 0.5f * copysignf(1.0f, dy)) * rH * 2 - 1) * mip_bound - y) * rdy;
        const float tz = (((nz + 0.5f + 0.5f * copysignf(1.0f, dz)) * rH * 2 - 1) * mip_bound - z) * rdz;
        const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        do {
            t += step_size(t);
        } while (t < tt);
        return false;
    }

    // ---- single-cascade, constant-step variant (cascade == 1, dt_gamma == 0, H <= 256: every shipped config) ----------
    // With one cascade the mip level is always 0 and with dt_gamma == 0 the step is the constant min(dt_max, dt_min), so
    // probe() collapses to the arithmetic below: same operations on the same values, hence the same samples bit for bit.
    // box = {x0, x1, y0, y1, z0, z1}: bounding box of the occupied cells.  A ray whose current cell lies beyond the box
    // on an axis along which it moves away (cell indices are monotone in t) can never meet an occupied cell again: the
    // reference would step through empty cells up to `far` and emit nothing, so the march stops right there (t = far).
    float c_mipb, c_rmipb, c_dt, hx, hy, hz;
    int bx0, bx1, by0, by1, bz0, bz1;

    __device__ __forceinline__ void init_fast(const int* __restrict__ box) {
        c_mipb = fminf(scalbnf(1.0f, 0), bound);
        c_rmipb = 1 / c_mipb;
        c_dt = clampf(0.0f, dt_min, dt_max);
        hx = 0.5f + 0.5f * copysignf(1.0f, dx);
        hy = 0.5f + 0.5f * copysignf(1.0f, dy);
        hz = 0.5f + 0.5f * copysignf(1.0f, dz);
        bx0 = box[0]; bx1 = box[1]; by0 = box[2]; by1 = box[3]; bz0 = box[4]; bz1 = box[5];
    }

    // the same specialisation without an occupied-cell box (no early termination): the box is the whole grid
    __device__ __forceinline__ void init_fast_full() {
        c_mipb = fminf(scalbnf(1.0f, 0), bound);
        c_rmipb = 1 / c_mipb;
        c_dt = clampf(0.0f, dt_min, dt_max);
        hx = 0.5f + 0.5f * copysignf(1.0f, dx);
        hy = 0.5f + 0.5f * copysignf(1.0f, dy);
        hz = 0.5f + 0.5f * copysignf(1.0f, dz);
        bx0 = by0 = bz0 = 0; bx1 = by1 = bz1 = (int)H - 1;
    }

    // conservative slab test of the ray against the occupied box grown by one cell (sides on the grid border are open)
    __device__ __forceinline__ bool misses_box() const {
        if (bx0 > bx1) return true;                  // no occupied cell at all
        const float cell = 2.0f * c_mipb * rH;
        float t0 = -3.402823466e+38f, t1 = 3.402823466e+38f;
        const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
        const int lo[3] = {bx0, by0, bz0}, hi[3] = {bx1, by1, bz1};
        #pragma unroll
        for (int a = 0; a < 3; a++) {
            const float pl = lo[a] <= 0 ? -3.402823466e+38f : (lo[a] - 1) * cell - c_mipb;
            const float ph = hi[a] >= (int)H - 1 ? 3.402823466e+38f : (hi[a] + 2) * cell - c_mipb;
            if (d[a] == 0.0f) {
                if (o[a] < pl || o[a] > ph) return true;
            } else {
                const float r = 1 / d[a];
                float ta = (pl - o[a]) * r, tb = (ph - o[a]) * r;
                if (ta > tb) { const float sw = ta; ta = tb; tb = sw; }
                t0 = fmaxf(t0, ta); t1 = fminf(t1, tb);
            }
        }
        return t0 > t1 + 1e-3f;
    }

    __device__ __forceinline__ bool probe_fast(float& t, float& x, float& y, float& z, float& dt, float far) const {
        x = clampf(ox + t * dx, -bound, bound);
        y = clampf(oy + t * dy, -bound, bound);
        z = clampf(oz + t * dz, -bound, bound);
        dt = c_dt;
        const int nx = (int)clampf(0.5f * (x * c_rmipb + 1) * Hf, 0.0f, (float)(H - 1));
        const int ny = (int)clampf(0.5f * (y * c_rmipb + 1) * Hf, 0.0f, (float)(H - 1));
        const int nz = (int)clampf(0.5f * (z * c_rmipb + 1) * Hf, 0.0f, (float)(H - 1));
        const uint32_t index = morton3(nx, ny, nz);
        const bool occ = grid[index >> 3] & (1u << (index & 7u));
        if (occ) return true;
        if ((nx > bx1 && dx >= 0.0f) || (nx < bx0 && dx <= 0.0f) || (ny > by1 && dy >= 0.0f) || (ny < by0 && dy <= 0.0f) ||
            (nz > bz1 && dz >= 0.0f) || (nz < bz0 && dz <= 0.0f)) {
            t = far;
            return false;
        }
        const float tx = (((nx + hx) * rH * 2 - 1) * c_mipb - x) * rdx;
        const float ty = (((ny + hy) * rH * 2 - 1) * c_mipb - y) * rdy;
        const float tz = (((nz + hz) * rH * 2 - 1) * c_mipb - z) * rdz;
        const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        do {
            t += c_dt;
        } while (t < tt);
        return false;
    }
};

// bounding box of the occupied cells of cascade 0 (render.cu); box = {x0, x1, y0, y1, z0, z1} must hold {INT_MAX, -1, ...} on entry
__global__ void k_occ_box(const uint8_t* __restrict__ grid, uint32_t n_words, int* __restrict__ box);

}  // namespace envidr
