// field_tc.cuh -- host-visible description of the tensor-core env_net kernel (field_tc.cu).
#pragma once
#include "common.cuh"

namespace envidr {

constexpr int kTcMaxLayers = ENVIDR_MAX_LAYERS;
constexpr int kTcRecFloats = 32;

struct TcLayer {
    uint32_t K, Kp, N, Np;        // Kp multiple of 16; Np = N (hidden, multiple of 32) or N rounded up to 16 (last layer)
    uint32_t img_off;             // byte offset of the layer image in the blob; stage s at img_off + s * Np * 64
    uint32_t img2_off;            // CTA-pair image: rank r at img2_off + r * (Kp/16) * (Np/2) * 64, K step s at + s * (Np/2) * 64
    uint32_t bias_off;            // float offset in the bias region
    uint32_t img8_off;            // fp8-correction image (f8 != 0): [phase A: Kp/32 chunks x (hi8 | lo8) x Np x 32 B][phase B: Kp/16 steps x hi16 x Np x 32 B]
    uint32_t f8;                  // this layer runs as hi16*hi16 (kind::f16) + e4m3 corrections (kind::f8f6f4), see field_tc.cu
    float dscale;                 // accumulator -> value scale of the layer (2^-3 for f8 layers, whose main weights are stored x 2^3)
};
struct TcEnv {
    const uint8_t* blob;          // weight images
    const float* bias;            // [sum Np] floats
    uint32_t n_layers, P, E;
    uint32_t ide_nb0;             // bands evaluated for the normal-direction (constant kappa) encoding, see tc_layout
    uint32_t debug;               // timing experiments (ENVIDR_ENV_TC_DEBUG bit mask; results are wrong): 1 no weight copies, 2 no IDE arithmetic, 4 hi*hi product only
    uint32_t reorder;             // issue the NEXT tile's layer 0 ahead of this tile's last layer (field_tc.cu, default 1)
    float kappa_diffuse, light_scale;
    int has_rot; float rot[9];    // records hold unrotated directions: d <- d @ rot before the encoding (envidr_field.rec_unrotated)
    TcLayer L[kTcMaxLayers];
};


// ---- geometry phase on tensor cores (geom_tc.cu) ---------------------------------------------------------
struct TcImg { uint32_t Kp, N, Np, off; };     // operand image of one layer inside the resident region (byte offset)
struct TcGeom {
    const uint8_t* blob;                       // start of the resident region in the packed blob
    uint64_t blob_off;                         // its byte offset inside field->packed
    uint32_t res_bytes, res_bytes_al, float_off, n_layers;
    TcImg F[4], R[4];                          // forward layers 0..n-1; reverse (transposed) images of layers 0..n-2
    const float* table; const int* offsets;
    uint32_t L, H; float S, bound; int enabled_levels; uint32_t geo_dim;
    float beta, density_scale, rough_bias, rough_act_scale, rough_scale;
    int has_rot; float rot[9];
    uint32_t stage_bytes;                      // shared-memory staging area for the coarsest level table (0 = off), set at launch
};
bool geom_tc_layout(const envidr_field* f, uint64_t base_bytes, TcGeom* out, uint64_t* total_bytes);
int geom_tc_pack(const envidr_field* f, const TcGeom& g, void* packed, cudaStream_t st);
int geom_tc_launch(const TcGeom& g, const float* xyzs, const float* dirs, const uint32_t* M_dev, uint32_t M_host, int mode, float* rec,
                   const envidr_field_out* out, cudaStream_t st, const uint32_t* rec_base_dev = nullptr, uint32_t rec_cap = 0xFFFFFFFFu);
// capture of the geometry records of a geometry-only pass (envidr_sample_log): external record buffer + running base index
struct RecCapture { float* rec; const uint32_t* base_dev; uint32_t cap; };

// ---- shading heads on tensor cores (shade_tc.cu) -----------------------------------------------------------
struct TcShade {
    const uint8_t* blob;
    uint64_t blob_off;
    uint32_t res_bytes, res_bytes_al, float_off;
    uint32_t net_layers[3];                    // 0 diffuse, 1 color, 2 renv
    TcImg img[3][4];
    uint32_t geo_dim, env_dim;
    float rough_scale, indir_rough_thresh, intensity_scale;
    int learn_blend;
};
bool shade_tc_layout(const envidr_field* f, uint64_t base_bytes, TcShade* out, uint64_t* total_bytes);
int shade_tc_pack(const envidr_field* f, const TcShade& s, void* packed, cudaStream_t st);
int shade_tc_launch(const TcShade& s, const float* rec, const float* feat, const float* r_images, const uint32_t* M_dev, uint32_t M_host,
                    const envidr_field_out* out, cudaStream_t st, const int32_t* ridx = nullptr);

// layout / packing / launch (field_tc.cu)
bool tc_layout(const envidr_field* f, uint64_t base_bytes, TcEnv* out, uint64_t* total_bytes);
int tc_pack(const envidr_field* f, const TcEnv& t, void* packed, cudaStream_t st);
// training forward (csrc/env_train_tc.cu): what the SAVE instantiation of k_env_tc leaves in HBM next to the features
struct TcSave {
    float* act[3];                // post-ReLU activations of the hidden layers, [2M, N_l] fp32, row = branch * M + sample (branch 0: normal direction)
    uint32_t* mask[3];            // their ReLU bit masks, [2M, N_l / 32]
    uint32_t M;
};
// ridx (optional): sample m reads record rec[ridx[m]] instead of rec[m] (records left in the geometry pass's log, render.py)
int env_tc_launch(const TcEnv& t, uint32_t ide_degree, const float* rec, float* feat, const uint32_t* M_dev, uint32_t M_host, cudaStream_t st,
                  const TcSave* save = nullptr, const int32_t* ridx = nullptr);
int env_tc_mode();               // 0: three fp16 products per K step; 1 (default): fp16 main product + two e4m3 correction products

}  // namespace envidr
