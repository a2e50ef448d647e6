// gridenc.cuh -- cell geometry / indexing shared by the grid-encoder kernels (gridenc.cu) and the fused
// field kernel (field.cu).  See gridenc.cu for the reference citations.
#pragma once
#include "common.cuh"

namespace envidr {

template <int C> struct Row;
template <> struct Row<1> { float v[1]; };
template <> struct Row<2> { float v[2]; };
template <> struct Row<4> { float v[4]; };
template <> struct Row<8> { float v[8]; };

template <int C>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[C]) {
    if constexpr (C == 1) {
        v[0] = __ldg(p);
    } else if constexpr (C == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
        #pragma unroll
        for (int c = 0; c < C; c += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p + c));
            v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
        }
    }
}

template <int C>
__device__ __forceinline__ void atomic_add_row(float* p, const float (&v)[C]) {
    if constexpr (C == 1) {
        atomicAdd(p, v[0]);
    } else if constexpr (C == 2) {
        atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    } else {
        #pragma unroll
        for (int c = 0; c < C; c += 4)
            atomicAdd(reinterpret_cast<float4*>(p + c), make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
    }
}

struct EncMode {
    uint32_t smooth;         // 1: hashencoder semantics, 0: gridencoder semantics
    uint32_t gridtype;       // gridencoder only: 0 hash, 1 tiled
    uint32_t align_corners;  // gridencoder only
};

// Cell index (reference: get_grid_index, hashencoder.cu:54-72 / gridencoder.cu:54-72)
template <int D>
__device__ __forceinline__ uint32_t cell_index(const EncMode m, uint32_t hashmap_size, uint32_t resolution, const uint32_t (&pg)[D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    const uint32_t dim_stride = m.smooth ? resolution : (m.align_corners ? resolution : resolution + 1);
    uint32_t stride = 1, index = 0;
    #pragma unroll
    for (int d = 0; d < D; d++) {
        if (stride <= hashmap_size) {
            index += pg[d] * stride;
            stride *= dim_stride;
        }
    }
    if (stride > hashmap_size && (m.smooth || m.gridtype == 0)) {
        uint32_t h = 0;
        #pragma unroll
        for (int d = 0; d < D; d++) h ^= pg[d] * primes[d];
        index = h;
    }
    return index % hashmap_size;
}

// Per-level cell setup.  Returns false for inputs outside [0,1]^D.
template <int D>
struct Cell {
    float w[D];     // interpolation weight toward the +1 corner
    float dw[D];    // d w / d pos
    uint32_t pg[D];
    float scale;
    uint32_t resolution, hashmap_size;

    __device__ __forceinline__ bool setup(const EncMode m, const float* __restrict__ x, const int* __restrict__ offsets,
                                          uint32_t level, float S, uint32_t H) {
        bool inside = true;
        float xin[D];
        #pragma unroll
        for (int d = 0; d < D; d++) {
            xin[d] = x[d];
            if (xin[d] < 0 || xin[d] > 1) inside = false;
        }
        hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        scale = exp2f(level * S) * H - 1.0f;
        resolution = (uint32_t)ceilf(scale) + 1;
        #pragma unroll
        for (int d = 0; d < D; d++) {
            float p = m.smooth ? xin[d] * scale : xin[d] * scale + (m.align_corners ? 0.0f : 0.5f);
            pg[d] = (uint32_t)floorf(p);
            p -= (float)pg[d];
            if (m.smooth) {
                dw[d] = 6 * p * (1.0f - p);
                w[d] = p * p * (3.0f - 2.0f * p);
            } else {
                dw[d] = 1.0f;
                w[d] = p;
            }
        }
        return inside;
    }
};


}  // namespace envidr
