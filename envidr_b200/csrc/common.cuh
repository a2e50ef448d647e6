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

void set_error(const char* fmt, ...);

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
};

}  // namespace envidr
