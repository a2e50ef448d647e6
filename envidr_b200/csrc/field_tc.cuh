// field_tc.cuh -- host-visible description of the tensor-core env_net kernel (field_tc.cu).
#pragma once
#include "common.cuh"

namespace envidr {

constexpr int kTcMaxLayers = ENVIDR_MAX_LAYERS;
constexpr int kTcRecFloats = 32;

struct TcLayer {
    uint32_t K, Kp, N, Np;        // Kp multiple of 16; Np = N (hidden, multiple of 32) or N rounded up to 16 (last layer)
    uint32_t img_off;             // byte offset of the layer image in the blob; stage s at img_off + s * Np * 64
    uint32_t bias_off;            // float offset in the bias region
};
struct TcEnv {
    const uint8_t* blob;          // weight images
    const float* bias;            // [sum Np] floats
    uint32_t n_layers, P, E;
    float kappa_diffuse, light_scale;
    TcLayer L[kTcMaxLayers];
};


// layout / packing / launch (field_tc.cu)
bool tc_layout(const envidr_field* f, uint64_t base_bytes, TcEnv* out, uint64_t* total_bytes);
int tc_pack(const envidr_field* f, const TcEnv& t, void* packed, cudaStream_t st);
int env_tc_launch(const TcEnv& t, uint32_t ide_degree, const float* rec, float* feat, const uint32_t* M_dev, uint32_t M_host, cudaStream_t st);

}  // namespace envidr
