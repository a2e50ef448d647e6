/*
 * envidr_b200.h -- C ABI of libenvidr_b200.so, the B200 (sm_100a) implementation of the
 * ENVIDR volumetric-render hot path.
 *
 * Every entry point replaces one function the reference exports from its pybind11 / cpp_extension
 * modules (the drop-in boundary); the reference declaration it replaces is cited as
 * <reference path>:<line>.  Conventions:
 *   - plain device pointers + sizes; no torch types; all tensors fp32 / int32 / uint8, contiguous,
 *     row-major, resident on the current CUDA device.  Outputs are caller-allocated (as in the
 *     reference, where Python allocates every output tensor).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, what the
 *     reference launches on).  Calls are asynchronous; nothing synchronises the device.
 *   - return value: 0 on success; a positive cudaError_t if a launch failed; a negative
 *     ENVIDR_E_* code for an argument the implementation rejects.  envidr_last_error() returns a
 *     static description of the most recent failure of the calling thread.
 *   - The library has no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef ENVIDR_B200_H_
#define ENVIDR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define ENVIDR_E_UNSUPPORTED (-1) /* unsupported D / C / degree / dims   (reference: std::runtime_error) */
#define ENVIDR_E_BADARG      (-2) /* null pointer / inconsistent sizes    (reference: TORCH_CHECK)        */
#define ENVIDR_E_WORKSPACE   (-3) /* workspace too small                                                 */

typedef void* envidr_stream_t;

const char* envidr_last_error(void);
/* ABI version: major*100 + minor */
int envidr_version(void);
/* sizeof {envidr_mlp_layer, envidr_field, envidr_field_out, envidr_render_opts, envidr_render_out}: lets a
 * foreign-language binding verify its struct mirrors. */
int envidr_abi_sizes(uint32_t out[5]);
/* sizeof {envidr_density_opts, envidr_adam_tensor, envidr_loss_in, envidr_loss_opts, envidr_sample_log} */
int envidr_abi_sizes_aux(uint32_t out[5]);

/* ------------------------------------------------------------------------------------------------
 * raymarching  (reference: raymarching/src/raymarching.h:7-18, bindings.cpp:5-20)
 * ---------------------------------------------------------------------------------------------- */

/* raymarching.h:7  near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars) */
int envidr_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                              float min_near, float* nears, float* fars, envidr_stream_t stream);
/* raymarching.h:8  sph_from_ray(rays_o, rays_d, radius, N, coords[N,2]) */
int envidr_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                        envidr_stream_t stream);
/* raymarching.h:9-10  morton3D(coords[N,3] i32, N, indices[N]) / morton3D_invert */
int envidr_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, envidr_stream_t stream);
int envidr_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, envidr_stream_t stream);
/* raymarching.h:11  packbits(grid f32[8N], N, density_thresh, bitfield u8[N]) */
int envidr_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, envidr_stream_t stream);
/* raymarching.h:12  get_scatter_idx(rays i32[N,3], N, idx_map i32[M]).  M = idx_map's length (the pybind signature reads it off
 * the tensor): rays that march_rays_train dropped (offset + count > M) are skipped -- the reference kernel writes out of bounds there. */
int envidr_get_scatter_idx(const int32_t* rays, uint32_t N, uint32_t M, int32_t* idx_map, envidr_stream_t stream);

/* raymarching.h:14  march_rays_train(...).  rays[N,3] = (ray id, offset, count); counter[2] is
 * advanced by (total samples, N).  Unlike the reference (two atomicAdds, scheduling-dependent
 * offsets) slot r holds ray r and offsets are the exclusive prefix sum in ray order, so the
 * result is deterministic; per-ray counts and sample sequences are bit-identical. */
int envidr_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                            float dt_gamma, uint32_t max_steps, uint32_t early_stop_steps, uint32_t N,
                            uint32_t C, uint32_t H, uint32_t M, const float* nears, const float* fars,
                            float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                            const float* noises, envidr_stream_t stream);
/* raymarching.h:15  composite_rays_train_forward(...); weights may be NULL (reference: numel()==0) */
int envidr_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                        const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                        uint32_t accum_deltas, uint32_t input_alpha, float* weights_sum,
                                        float* depth, float* image, float* weights, envidr_stream_t stream);
/* raymarching.h:16  composite_rays_train_backward(...) */
int envidr_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                         const float* grad_depth, const float* sigmas, const float* rgbs,
                                         const float* deltas, const int32_t* rays, const float* weights_sum,
                                         const float* image, const float* depth, uint32_t M, uint32_t N,
                                         float T_thresh, float* grad_sigmas, float* grad_rgbs,
                                         uint32_t accum_deltas, uint32_t input_alpha, envidr_stream_t stream);
/* raymarching.h:17  march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma,
 * max_steps, C, H, grid, nears, fars, xyzs, dirs, deltas, noises).  Every slot of
 * xyzs/dirs/deltas[n_alive*n_step] is written (zeros past the end of a ray). */
int envidr_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                      const float* rays_o, const float* rays_d, float bound, float dt_gamma,
                      uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                      const float* fars, float* xyzs, float* dirs, float* deltas, const float* noises,
                      envidr_stream_t stream);
/* raymarching.h:18  composite_rays(...): in place on rays_alive, rays_t, weights_sum, depth, image */
int envidr_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, uint32_t accum_deltas,
                          uint32_t input_alpha, int32_t* rays_alive, float* rays_t, const float* sigmas,
                          const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                          float* image, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * hashencoder  (reference: hashencoder/src/hashencoder.h:13-15)  -- smoothstep, twice differentiable
 * ---------------------------------------------------------------------------------------------- */
/* outputs [L,B,C]; dy_dx [B, L*D*C] (ignored unless calc_grad_inputs).  D in {2,3}, C in {1,2,4,8}. */
int envidr_hash_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets,
                               float* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                               uint32_t H, int calc_grad_inputs, float* dy_dx, envidr_stream_t stream);
/* grad_embeddings (pre-zeroed by the caller, as in hashgrid.py:80) is accumulated into;
 * grad_inputs [B,D] is written when calc_grad_inputs. */
int envidr_hash_encode_backward(const float* grad, const float* inputs, const float* embeddings,
                                const int32_t* offsets, float* grad_embeddings, uint32_t B, uint32_t D,
                                uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                                const float* dy_dx, float* grad_inputs, envidr_stream_t stream);
/* grad_grad [L,B,C] is written; grad2_embeddings (pre-zeroed) is accumulated into. */
int envidr_hash_encode_second_backward(const float* grad, const float* inputs, const float* embeddings,
                                       const int32_t* offsets, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                       float S, uint32_t H, int calc_grad_inputs, const float* dy_dx,
                                       const float* grad_grad_inputs, float* grad_grad,
                                       float* grad2_embeddings, envidr_stream_t stream);

/* Test hook: the per-level fp32 `scale = exp2f(level*S)*H - 1` exactly as the kernels evaluate it (exp2f is the
 * ex2.approx instruction on the device, up to 2 ulp from libm).  Lets a CPU checker use the device's level geometry. */
int envidr_debug_level_scales(float S, uint32_t H, uint32_t L, float* scales /* device [L] */, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * gridencoder  (reference: gridencoder/src/gridencoder.h:12-13)  -- linear interp, hash / tiled
 * ---------------------------------------------------------------------------------------------- */
/* dy_dx may be NULL (reference: at::optional).  D in 1..5, C in {1,2,4,8}.  gridtype 0 = hash, 1 = tiled. */
int envidr_grid_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets,
                               float* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                               uint32_t H, float* dy_dx, uint32_t gridtype, int align_corners,
                               envidr_stream_t stream);
int envidr_grid_encode_backward(const float* grad, const float* inputs, const float* embeddings,
                                const int32_t* offsets, float* grad_embeddings, uint32_t B, uint32_t D,
                                uint32_t C, uint32_t L, float S, uint32_t H, const float* dy_dx,
                                float* grad_inputs, uint32_t gridtype, int align_corners,
                                envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * freqencoder / shencoder  (reference: freqencoder/src/freqencoder.h:7-10, shencoder/src/shencoder.h:9-10)
 * ---------------------------------------------------------------------------------------------- */
int envidr_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                               float* outputs, envidr_stream_t stream);
int envidr_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg,
                                uint32_t C, float* grad_inputs, envidr_stream_t stream);
/* degree in 1..8; dy_dx [B, 3*degree^2] may be NULL */
int envidr_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree,
                             float* dy_dx, envidr_stream_t stream);
/* accumulates (+=) into grad_inputs [B,D] like the reference */
int envidr_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree,
                              const float* dy_dx, float* grad_inputs, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * IDE  (reference: ide_encoder/ide_encoder.py:98-130, pure PyTorch there; no native symbol exists)
 * ---------------------------------------------------------------------------------------------- */
/* dirs [B,3]; kappa_inv: per-sample array [B] when kappa_inv_arr != NULL, else the scalar;
 * out [B, 2*P], P = 2^deg_view - 1 + deg_view ([Re | Im] halves), deg_view in 1..5; out *= scale. */
int envidr_ide_encode_forward(const float* dirs, const float* kappa_inv_arr, float kappa_inv_scalar,
                              uint32_t B, uint32_t deg_view, float scale, float* out, envidr_stream_t stream);
/* Backward of envidr_ide_encode_forward (the reference differentiates ide_encoder.py:98-130 with autograd): grad [B, 2*P] ->
 * grad_dirs [B,3] and, when grad_kappa != NULL, grad_kappa [B] (d/d kappa_inv; meaningful with a per-sample kappa array). */
int envidr_ide_encode_backward(const float* dirs, const float* kappa_inv_arr, float kappa_inv_scalar,
                               uint32_t B, uint32_t deg_view, float scale, const float* grad,
                               float* grad_dirs, float* grad_kappa, envidr_stream_t stream);
/* Host-only: the coefficient tables the kernel uses (ide_encoder.py:84-96): mat [(l_max+1), P] row-major,
 * sigma [P], ml [2, P] (row 0 = m, row 1 = l).  l_max = 2^(deg_view-1). */
int envidr_ide_tables(uint32_t deg_view, float* mat, float* sigma, int32_t* ml);

/* ------------------------------------------------------------------------------------------------
 * Fused per-sample field + fused inference render loop.
 * The reference has no operator boundary here (nn.Linear stacks driven from Python:
 * nerf/network.py:381-698, nerf/renderer.py:147-198, nerf/render_func/cuda_ray.py:238-359); the seams
 * these replace are NeRFNetwork.forward_sigma/forward_color and nerf.render_func.run_cuda.
 * ---------------------------------------------------------------------------------------------- */

typedef struct envidr_mlp_layer {
    const float* weight; /* [out_dim, in_dim] row-major (torch nn.Linear layout) */
    const float* bias;   /* [out_dim] or NULL */
    uint32_t in_dim, out_dim;
} envidr_mlp_layer;

#define ENVIDR_MAX_LAYERS 8

typedef struct envidr_field {
    /* hash grid (hashencoder.HashEncoder) */
    const float* embeddings;     /* [T, 2] */
    const int32_t* offsets;      /* [L+1] (device) */
    uint32_t num_levels;         /* L (<= 16) */
    uint32_t level_dim;          /* C (2) */
    uint32_t base_resolution;    /* H */
    float log2_per_level_scale;  /* S */
    float bound;
    int32_t enabled_levels;      /* <=0: all (network.py:390-393) */
    /* MLP stacks */
    uint32_t n_sdf, n_env, n_diffuse, n_color, n_renv;
    envidr_mlp_layer sdf[ENVIDR_MAX_LAYERS], env[ENVIDR_MAX_LAYERS], diffuse[ENVIDR_MAX_LAYERS],
                     color[ENVIDR_MAX_LAYERS], renv[ENVIDR_MAX_LAYERS];
    uint32_t geo_feat_dim;       /* 12 */
    uint32_t ide_degree;         /* deg_view: 4 or 5 */
    /* scalars */
    float beta;                  /* already clamped to [beta_min, beta_max] */
    float density_scale;
    float roughness_bias, roughness_act_scale, roughness_scale;
    float diffuse_kappa_inv, light_intensity_scale, intensity_scale;
    float indir_roughness_thresh;
    int32_t learn_indir_blend;
    int32_t has_env_rot; float env_rot[9]; /* row-major 3x3 R; w_r <- w_r @ R (renderer.py:160-161) */
    /* device buffer of repacked (K-major, padded) weights written by envidr_field_pack; must be
     * refreshed whenever a weight tensor changes.  Size: envidr_field_pack_bytes(). */
    const void* packed; uint64_t packed_bytes;
    /* arithmetic of the env_net passes: 0 = fp32 FFMA (bit-faithful path); 1 = tcgen05 tensor cores with every operand
     * split into two fp16 values and three MMAs per K step (fp32 accumulate; ~1e-6 relative to the fp32 path). */
    int32_t precision;
    int32_t rec_unrotated;  /* envidr_field_forward_records only: the records hold UNROTATED n_env / w_r (a geometry pass captured without
                             * env_rot) and env_rot is applied inside the env_net kernel -- one geometry pass serves every light
                             * rotation of a relight sweep (BASELINE config 5, utils.py:1297-1303) */
    /* precision = 1 only: device scratch of 256 bytes per sample for at least scratch_samples >= M samples */
    void* scratch; uint64_t scratch_samples;
} envidr_field;

/* Bytes needed for field->packed (0 if the field description is rejected; see envidr_last_error). */
uint64_t envidr_field_pack_bytes(const envidr_field* field);
/* Repack the torch-layout weights referenced by `field` into `packed` (device memory, async on stream). */
int envidr_field_pack(const envidr_field* field, void* packed, uint64_t packed_bytes, envidr_stream_t stream);

/* Per-sample outputs of the field; any pointer may be NULL.  r_images [M,4] may be NULL. */
typedef struct envidr_field_out {
    float* sigma;      /* [M]   density * density_scale            */
    float* rgb;        /* [M,3] (c_diffuse + c_specular) * intensity */
    float* normal;     /* [M,3] unit normal                        */
    float* sdf;        /* [M]                                      */
    float* c_diffuse;  /* [M,3] */
    float* c_specular; /* [M,3] */
    float* roughness;  /* [M]   */
    float* grad_x;     /* [M,3] un-normalised d sdf / d xyz        */
} envidr_field_out;

/* mode: 0 = full shading, 1 = geometry only (sigma, normal, sdf, roughness) */
int envidr_field_forward(const envidr_field* field, const float* xyzs, const float* dirs,
                         const float* r_images, uint32_t M, int mode, const envidr_field_out* out,
                         envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused inference render loop: replaces the `else` branch of run_cuda
 * (nerf/render_func/cuda_ray.py:238-359) -- near/far, march, field, composite, alive-list compaction and
 * the final background mix, all device-driven (no host sync per iteration).
 * ---------------------------------------------------------------------------------------------- */
typedef struct envidr_render_opts {
    float bound;            /* model.bound                                  */
    float dt_gamma;         /* opt.dt_gamma                                 */
    float T_thresh;         /* opt.T_thresh                                 */
    float min_near;         /* model.min_near                               */
    uint32_t max_steps;     /* opt.max_steps                                */
    uint32_t cascade;       /* model.cascade                                */
    uint32_t grid_size;     /* model.grid_size (128)                        */
    float aabb[6];          /* model.aabb_infer                             */
    float bg_color[3];      /* constant background (ignored when bg_per_ray is given) */
    int32_t geometry_only;  /* composite normals instead of colours (pass 1 of indir_ref) */
    int32_t input_alpha;    /* NeuS-style alpha input (use_neus_sdf)        */
    uint32_t n_step_floor;  /* 0 / 1: the reference schedule n_step = clamp(N / n_alive, 1, 8) (cuda_ray.py:287); f in 2..8: n_step >= f.
                             * Only the batching changes: every ray still composites the same samples (positions to the ulp-level
                             * re-synchronisation of rays_t), but a pass over few rays no longer starts with N / n_alive = 1, i.e. with
                             * launches of at most N samples.  The workspace must be sized with envidr_render_workspace_bytes_ex. */
    uint32_t n_step_cap;    /* 0 / 8: the reference's cap of n_step (cuda_ray.py:287); up to 16: a pass whose alive list has shrunk below N / 8
                             * marches up to that many samples per ray and iteration -- fewer, larger iterations in the long tail of a pass
                             * (again only the batching changes; at most cap - 1 samples per ray are marched past its termination) */
} envidr_render_opts;

/* Optional capture of a geometry-only pass (tensor-core field): one entry per MARCHED sample, in iteration-major order.
 * The main pass of the indirect-reflection scheme shades the same samples, so it can reuse the geometry records instead of
 * marching and evaluating the hash grid + sdf_net again: envidr_permute_sample_log orders them ray by ray,
 * envidr_field_forward_records runs env_net + the shading heads on them, envidr_composite_rays_replay composites. */
typedef struct envidr_sample_log {
    float* rec;             /* [capacity, 32] per-sample geometry record of the tensor-core path (geo feature, normal, n.w_o,
                             *                roughness, blend weight, rotated normal, reflected direction)                   */
    float* sigma;           /* [capacity]                                                                                     */
    float* delta;           /* [capacity, 2]                                                                                  */
    int32_t* ray;           /* [capacity] ray index of the sample; -1 = marched past the ray's termination (not composited)  */
    int32_t* seq;           /* [capacity] position of the sample among the composited samples of its ray                      */
    uint64_t capacity;      /* samples; entries past it are dropped (compare envidr_render_last_stats with it)                */
} envidr_sample_log;

typedef struct envidr_render_out {
    float* image;           /* [N,3]  (unused when geometry_only)           */
    float* depth;           /* [N]                                          */
    float* weights_sum;     /* [N]                                          */
    float* normal_image;    /* [N,3]  optional; unit-normalised as in cuda_ray.py:356-359 */
    float* diffuse_image;   /* [N,3]  optional ('diffuse' in visual_items)  */
    float* specular_image;  /* [N,3]  optional ('specular' in visual_items) */
    float* roughness_image; /* [N]    optional (composited roughness)       */
    int32_t* sample_count;  /* [N]    optional: samples composited per ray (consumed by envidr_march_rays_replay) */
    const envidr_sample_log* log;   /* optional (geometry_only passes of the tensor-core field): see envidr_sample_log */
} envidr_render_out;

uint64_t envidr_render_workspace_bytes(uint32_t N);
uint64_t envidr_render_workspace_bytes_ex(uint32_t N, uint32_t n_step_floor);
/* r_images [N,4] (reflected radiance + visibility per ray), noises [N] (perturb), bg_per_ray [N,3] may be NULL.
 * workspace: 256-byte aligned device memory of at least envidr_render_workspace_bytes(N). */
int envidr_render_rays(const envidr_field* field, const uint8_t* bitfield, const float* rays_o, const float* rays_d,
                       const float* r_images, const float* noises, const float* bg_per_ray, uint32_t N,
                       const envidr_render_opts* opts, const envidr_render_out* out, void* workspace,
                       uint64_t workspace_bytes, envidr_stream_t stream);
/* Replay of a pass whose per-ray sample counts are already known (the main pass of the indirect-reflection scheme,
 * renderer.py:439-513, visits the rays of the geometry pass again): instead of re-discovering ray termination iteration by
 * iteration, march counts[n] samples of every ray in one launch (same DDA as envidr_march_rays / _train; ray-contiguous
 * output like march_rays_train: rays[n] = (n, offset, counts[n]), counter[0] += sum), evaluate the field on the whole batch,
 * and composite with the inference compositor's arithmetic (raymarching.cu:996-1039: T = 1 - sum w before the sample).
 * Samples past capacity M are dropped ray-wise as in march_rays_train. */
int envidr_march_rays_replay(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                             const float* fars, const int32_t* counts, float* xyzs, float* dirs, float* deltas,
                             int32_t* rays, int32_t* counter, envidr_stream_t stream);
/* Ray-contiguous copy of a sample log: entry g (g < total) with ray r = log->ray[g] >= 0 and ray_offset[r] >= 0 goes to
 * slot ray_offset[r] + log->seq[g] of rec_out [*,32] / sigma_out / delta_out [*,2].  ray_offset [N]: exclusive scan of the
 * sample counts over the rays of the next pass, -1 for rays that are not part of it. */
int envidr_permute_sample_log(const envidr_sample_log* log, uint64_t total, const int32_t* ray_offset, float* rec_out,
                              float* sigma_out, float* delta_out, envidr_stream_t stream);
/* Index form: instead of copying the 128-byte records, idx_out [M] receives the log position of every ray-ordered sample (sigma / delta are
 * still gathered: the compositor walks them sequentially); use with envidr_field_forward_records_indexed. */
int envidr_permute_sample_log_index(const envidr_sample_log* log, uint64_t total, const int32_t* ray_offset, int32_t* idx_out,
                                    float* sigma_out, float* delta_out, envidr_stream_t stream);
/* env_net + shading heads on precomputed geometry records (tensor-core field only): rec [M,32], r_images [M,4] or NULL;
 * writes out->rgb (and c_diffuse / c_specular when set).  field->scratch needs 128 B per sample. */
int envidr_field_forward_records(const envidr_field* field, const float* rec, const float* r_images, uint32_t M,
                                 const envidr_field_out* out, envidr_stream_t stream);
/* The same with the records left where the geometry pass logged them: sample m uses rec[rec_index[m]] (rec_index from
 * envidr_permute_sample_log_index; saves copying 128 B per sample into ray order).  rec_index == NULL: as above. */
int envidr_field_forward_records_indexed(const envidr_field* field, const float* rec, const int32_t* rec_index, const float* r_images, uint32_t M,
                                         const envidr_field_out* out, envidr_stream_t stream);
/* dst[offset_n + s, 0:4] = src[ray_n, 0:4] for every sample of every ray in `rays` [N,3] (per-ray r_images -> per-sample rows). */
int envidr_scatter_ray_rows4(const int32_t* rays, uint32_t N, uint32_t M, const float* src /* [N,4] */, float* dst /* [M,4] */,
                             envidr_stream_t stream);
/* dst[p, 0:row_floats] = src[idx[p], 0:row_floats], p < n_rows (row_floats a multiple of 4, 16-byte aligned buffers): the de-interleave of a
 * sharded frame after its all-gather (envidr_b200/dist.py; no reference counterpart: the reference has no reachable multi-GPU path). */
int envidr_gather_rows(const float* src, const int32_t* idx, uint64_t n_rows, uint32_t row_floats, float* dst, envidr_stream_t stream);
/* images may be NULL except weights_sum / image; sigmas [M], rgbs / normals / c_diffuse / c_specular [M,3], roughness [M];
 * nears [N] (start of the depth accumulation, may be NULL = 0). */
int envidr_composite_rays_replay(const float* sigmas, const float* rgbs, const float* normals, const float* c_diffuse,
                                 const float* c_specular, const float* roughness, const float* deltas, const int32_t* rays,
                                 const float* nears, uint32_t M, uint32_t N, float T_thresh, uint32_t input_alpha, float* weights_sum,
                                 float* depth, float* image, float* normal_image, float* diffuse_image,
                                 float* specular_image, float* roughness_image, envidr_stream_t stream);

/* Waits for the most recent envidr_render_rays of this process; stats = {iterations, samples_lo, samples_hi, 0}. */
int envidr_render_last_stats(uint32_t stats[4]);

/* Test hook: D[128 x N] (fp32, row-major) = A[128 x K] * B[N x K]^T with both operands rounded to fp16, computed by one
 * tcgen05.mma chain (pins the UMMA descriptor / operand layout conventions of the tensor-core path on hardware). */
int envidr_tc_probe(const float* A, const float* B, float* D, uint32_t N, uint32_t K, uint32_t variant, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layer on tensor cores for the TRAINING branch (the reference: nn.Linear -> cuBLAS fp32, nerf/network.py:527-698).
 * Y[M,N] = act(X[M,K] W^T + bias), fp32 in / out, operands split into fp16 hi + lo (three tcgen05.mma per K step, fp32
 * accumulate): fp32-level accuracy.  N, K <= 256.  The weight image is built once per weight value by envidr_linear_tc_pack
 * (W [N,K] row-major, the torch layout); the data gradient dY W uses the same kernel on the image of W^T.
 * ---------------------------------------------------------------------------------------------- */
uint64_t envidr_linear_tc_image_bytes(uint32_t N, uint32_t K);
int envidr_linear_tc_pack(const float* W, uint32_t N, uint32_t K, void* img, envidr_stream_t stream);
int envidr_linear_tc(const float* X, uint32_t M, uint32_t K, const void* img, const float* bias /* [N] or NULL */, uint32_t N,
                     int relu, float* Y, envidr_stream_t stream);

/* Weight gradient dW [N,K] = dY^T X of the layer above (dY [M,N], X [M,K] fp32 row-major), same split precision; both operands
 * are fed MN-major, i.e. as they lie in memory.  partial: [envidr_wgrad_tc_partials(M), N, K] floats, one partial sum per CTA
 * (the caller adds them up); scales (device, 3 floats): power-of-two pre-scales of dY and X and the inverse of their product
 * (a loss gradient sits around 1e-7, below fp16's normal range).  variant: 0 (1 swaps the descriptor's LBO / SBO; test hook). */
uint32_t envidr_wgrad_tc_partials(uint32_t M);
/* The per-tensor scales envidr_wgrad_tc expects, in one launch: scales8 = device float[8], ZERO on first use (entries 4..6 are
 * scratch that the kernel leaves zeroed again); on return scales8[0..2] = {s_a, s_b, 1 / (s_a s_b)}, s = 2^(13 - floor(log2 max|.|)).
 * colsum != NULL: a is [na / N, N] and its column sums (the bias gradient when a = dY) are ADDED into colsum [N] in the same pass
 * (N a power of two in 4..256, 16-byte aligned operands; otherwise ENVIDR_E_UNSUPPORTED). */
int envidr_pow2_scales(const float* a, uint64_t na, const float* b, uint64_t nb, float* scales8, uint32_t N, float* colsum,
                       envidr_stream_t stream);
/* variant bit 1 (value 2): `partial` is dW [N, K] itself, zeroed by the caller; every CTA adds its partial sum with vector atomics
 * (no [grid, N, K] round trip through HBM; the summation order, hence the last bit, is not deterministic). */
int envidr_wgrad_tc(const float* dY, const float* X, uint32_t M, uint32_t N, uint32_t K, const float* scales, float* partial,
                    int variant, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * env_net of the TRAINING branch as one forward and one backward kernel (csrc/env_train_tc.cu).
 * Reference: the two env_net evaluations of get_color_mlp_extra_params / forward_color (nerf/network.py:527-541, 589-607:
 * IDE(n, diffuse_kappa_inv) and IDE(w_r, roughness) -> nn.Linear + ReLU stack -> F.normalize) and their autograd backward; the
 * reference has no operator boundary here (nn.Linear driven from Python), envidr_b200/env_train.py wraps these entries in a
 * torch.autograd.Function with the reference's argument meaning.
 *   forward : rec [M,32] sample records (floats 20 roughness, 22..24 normal, 25..27 reflected direction; the record of the inference path)
 *             -> feat [M,32]: unit-norm env feature of the normal direction at 0.., of the reflected direction at 16.. (slots 12 / 28 hold
 *             1 / |raw feature|), plus, per hidden layer l, act_l [2M, N_l] (post-ReLU, fp32) and mask_l [2M, N_l / 32] (ReLU bit masks);
 *             rows of the [2M] batch: [0, M) normal direction, [M, 2M) reflected direction.
 *   backward: gfeat [M,32] (gradient w.r.t. feat; slots >= env_feat ignored) -> gy [2M,16] (w.r.t. the raw feature), gact_l [2M, N_l]
 *             (w.r.t. the PRE-activation of hidden layer l: dW_l = gact_l^T act_{l-1}, db_l = column sums), gx0 [2M, input_cols]
 *             (w.r.t. the IDE features in the kernel's interleaved column order: column 2i = Re_i, 2i + 1 = Im_i, zero padded).
 * The weight images (forward + transposed) live in `blob` and are rebuilt by envidr_env_mlp_pack whenever the weights change.
 * ---------------------------------------------------------------------------------------------- */
typedef struct envidr_env_mlp {
    uint32_t n_layers;                           /* 2..4 */
    uint32_t dims[ENVIDR_MAX_LAYERS + 1];        /* dims[0] = 2 P (IDE features), dims[i + 1] = outputs of layer i; hidden widths multiples of 32 in 64..256, last <= 12 */
    const float* weight[ENVIDR_MAX_LAYERS];      /* [dims[i + 1], dims[i]] row-major (torch layout), device */
    const float* bias[ENVIDR_MAX_LAYERS];        /* [dims[i + 1]] or NULL */
    uint32_t ide_degree;
    float diffuse_kappa_inv, light_intensity_scale;
} envidr_env_mlp;
uint64_t envidr_env_mlp_blob_bytes(const envidr_env_mlp* d);      /* 0: shape outside the fused kernels */
uint64_t envidr_env_mlp_input_cols(const envidr_env_mlp* d);      /* columns of gx0 (2 P rounded up to 16) */
int envidr_env_mlp_pack(const envidr_env_mlp* d, void* blob, uint64_t blob_bytes, envidr_stream_t stream);
int envidr_env_mlp_forward(const envidr_env_mlp* d, const void* blob, const float* rec, uint32_t M, float* feat, float* act0, float* act1,
                           float* act2, uint32_t* mask0, uint32_t* mask1, uint32_t* mask2, envidr_stream_t stream);
int envidr_env_mlp_backward(const envidr_env_mlp* d, const void* blob, const float* gfeat, const float* feat, const uint32_t* mask0,
                            const uint32_t* mask1, const uint32_t* mask2, uint32_t M, float* gact0, float* gact1, float* gact2, float* gy,
                            float* gx0, envidr_stream_t stream);

/* The same pair of kernels for the small ReLU heads of the training branch (diffuse_net, color_net / specular_net, renv_net:
 * nerf/network.py:335-366, 555-607, 612-659; nn.Linear stacks under autograd in the reference): Y = W_n(... relu(W_1 X + b_1) ...) + b_n as
 * ONE forward kernel (masks, and optionally the hidden activations for the weight gradients, saved once) and ONE backward kernel for the
 * whole data-gradient chain.  The descriptor is envidr_env_mlp with the IDE fields ignored: 2..4 layers, dims[0] <= 64, hidden widths
 * multiples of 32 (<= 256), dims[n] <= 16.  X [rows, x_ld], Y [rows, 16] (zero padded), mask_l [rows, N_l / 32], act_l [rows, N_l] or NULL;
 * backward: gY [rows, gy_ld] -> gz_l [rows, N_l] (gradient w.r.t. the pre-activation of hidden layer l, NULL = not needed) and
 * gX [rows, dims[0] rounded up to 16]. */
uint64_t envidr_mlp_blob_bytes(const envidr_env_mlp* d);
int envidr_mlp_pack(const envidr_env_mlp* d, void* blob, uint64_t blob_bytes, envidr_stream_t stream);
int envidr_mlp_forward(const envidr_env_mlp* d, const void* blob, const float* X, uint32_t x_ld, uint32_t rows, float* Y, uint32_t* mask0,
                       uint32_t* mask1, uint32_t* mask2, float* act0, float* act1, float* act2, envidr_stream_t stream);
int envidr_mlp_backward(const envidr_env_mlp* d, const void* blob, const float* gY, uint32_t gy_ld, const uint32_t* mask0, const uint32_t* mask1,
                        const uint32_t* mask2, uint32_t rows, float* gz0, float* gz1, float* gz2, float* gX, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * NeuS-style opacity: NeuSDensity.forward (nerf/network.py:46-102), the density of the use_neus_sdf configs,
 * consumed by the compositors with input_alpha = 1.  variance: device scalar (the module's parameter); dists: [M] or NULL
 * (then dist_scalar, the module's base_dist); gradients [M,3] or NULL (the reference's gradient-free branch).
 * Backward: gradients of sum(grad_alpha * alpha) w.r.t. sdf [M], gradients [M,3] and variance (device scalar); any may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int envidr_neus_alpha_forward(const float* sdf, const float* dirs, const float* gradients, const float* dists, float dist_scalar,
                              const float* variance, float cos_anneal_ratio, uint32_t M, float* alpha, envidr_stream_t stream);
uint64_t envidr_neus_workspace_bytes(void);
int envidr_neus_alpha_backward(const float* grad_alpha, const float* sdf, const float* dirs, const float* gradients, const float* dists,
                               float dist_scalar, const float* variance, float cos_anneal_ratio, uint32_t M, float* grad_sdf,
                               float* grad_gradients, float* grad_variance, void* workspace, uint64_t workspace_bytes, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * NeuS-style geometry network (BASELINE config 4; nerf/network.py:154-222 construction, :415-421 forward with skip_layers,
 * nerf/renderer.py:182-198 normals): the kernels between the dense layers, which run on envidr_linear_tc.  The reference has no
 * operator boundary here (nn.Linear + Softplus(beta=100) + torch.cat + autograd.grad driven from Python);
 * envidr_b200/neus_field.py drives the chain.
 * ---------------------------------------------------------------------------------------------- */
/* The whole geometry network in ONE kernel per batch (csrc/neus_geom_tc.cu): frequency encoding -> the dense layers with Softplus(beta)
 * and an optional skip concat -> head [M,16] (row of the last layer: sdf, geo_feat, roughness, blend, zero padded) and the reverse pass
 * through the same layers -> grad_x [M,3] = d sdf / d xyz.  img / imgT: envidr_linear_tc_pack images of W [out,in] and of W^T [in,out]
 * (weight norm folded by the caller); hidden widths must pad (to 16) to multiples of 32, <= 256; multires <= 7; one skip layer at most.
 * scratch: envidr_neus_geometry_scratch_bytes(n_layers) (the softplus derivatives of one tile per SM, reused by every call). */
#define ENVIDR_NEUS_MAX_LAYERS 8
typedef struct envidr_neus_layer {
    const void* img;        /* image of W  [out_dim, in_dim]                         */
    const void* imgT;       /* image of W^T [in_dim, out_dim] (reverse pass)          */
    const float* bias;      /* [out_dim]                                              */
    uint32_t in_dim, out_dim;
} envidr_neus_layer;
typedef struct envidr_neus_net {
    envidr_neus_layer layers[ENVIDR_NEUS_MAX_LAYERS];
    const float* head_row;  /* W_last[0, :] (device, in_dim of the last layer floats) */
    uint32_t n_layers;
    int32_t skip_layer;     /* opt.skip_layers[0], or -1                              */
    uint32_t multires;      /* FreqEncoder degree                                     */
    float beta;             /* Softplus beta (100)                                    */
} envidr_neus_net;
uint64_t envidr_neus_geometry_scratch_bytes(uint32_t n_layers);
int envidr_neus_geometry(const envidr_neus_net* net, const float* xyzs, uint32_t M, float* head, float* grad_x, void* scratch,
                         uint64_t scratch_bytes, envidr_stream_t stream);
/* h = Softplus(beta, threshold 20)(z) and, when dh != NULL, dh = sigmoid(beta z) (its derivative); n elements; h may alias z. */
int envidr_softplus_forward(const float* z, uint64_t n, float beta, float* h, float* dh, envidr_stream_t stream);
/* out [M,N] = a [M,N] * b [M,N], or a_row [N] (broadcast over rows, a == NULL) * b.  The reverse pass g <- g . softplus'. */
int envidr_mul_rows(const float* a, const float* a_row, const float* b, uint64_t M, uint32_t N, float* out, envidr_stream_t stream);
/* out [M, Nh+Nx] = cat(h [M,Nh], x [M,Nx]) * scale (network.py:417-418, scale = 1/sqrt(2)), and its reverse: gh = g[:, :Nh] * scale,
 * gx (+)= g[:, Nh:] * scale. */
int envidr_skip_concat_forward(const float* h, const float* x, uint64_t M, uint32_t Nh, uint32_t Nx, float scale, float* out,
                               envidr_stream_t stream);
int envidr_skip_concat_backward(const float* g, uint64_t M, uint32_t Nh, uint32_t Nx, float scale, float* gh, float* gx, int accumulate,
                                envidr_stream_t stream);
/* Head of the geometry network -> inputs of the shading kernels: h [M, ld] = (sdf, geo_feat[G], roughness, [blend]) rows of the last
 * layer (network.py:424-448), grad_x [M,3] = d sdf / d x.  Outputs (any may be NULL): sdf [M], normal [M,3] (F.normalize, eps 1e-10),
 * roughness [M], rec [M,32] = the geometry record consumed by envidr_field_forward_records (reflected / normal directions of
 * renderer.py:147-180, rotated by rot9 = HOST float[9] rot_theta[:3,:3] when given). */
int envidr_neus_records(const float* h, uint32_t ld, const float* grad_x, const float* dirs, uint32_t M, uint32_t geo_feat_dim,
                        float roughness_bias, float roughness_act_scale, float roughness_scale, int has_blend, const float* rot9,
                        float* sdf, float* normal, float* roughness, float* rec, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Occupancy-grid maintenance: the caller either side of the march.
 * Replaces NeRFRenderer.update_extra_state (nerf/renderer.py:264-352) and NeRFRenderer.mark_untrained_grid
 * (nerf/renderer.py:200-262); the reference has no operator boundary here (torch ops + self.density() + morton3D + packbits
 * driven from Python, with a .item() sync for mean_density).
 * ---------------------------------------------------------------------------------------------- */
typedef struct envidr_density_opts {
    float bound;            /* model.bound                                                     */
    uint32_t cascade;       /* model.cascade                                                   */
    uint32_t grid_size;     /* model.grid_size (128)                                           */
    float decay;            /* update_extra_state(decay=0.95)                                  */
    float density_thresh;   /* model.density_thresh; the bit field uses min(mean_density, it)  */
} envidr_density_opts;

uint64_t envidr_density_workspace_bytes(uint32_t cascade, uint32_t grid_size, uint32_t n);
/* One update_extra_state: query the density (hash grid + sdf_net + Laplace density, x density_scale) at one jittered point per
 * visited cell, density_grid = max(density_grid * decay, tmp) where both are >= 0, mean_density = mean(clamp(density_grid, 0)),
 * bitfield = packbits(density_grid, min(mean_density, density_thresh)).  No host synchronisation.
 *   density_grid [cascade * H^3] in/out (Morton order within a cascade, as the reference), 16-byte aligned
 *   coords  NULL: full update (renderer.py:279-306), n must be H^3; else int32 [cascade, n, 3] cells to visit
 *           (partial update, renderer.py:310-336: the caller draws them as the reference does)
 *   noise   [cascade, n, 3] U[0,1) jitter, indexed like the reference's torch.rand_like(cas_xyzs): for the full update row
 *           (x * H + y) * H + z (meshgrid order), for the partial update row j of coords; NULL = cell centres (no jitter)
 *   bitfield [cascade * H^3 / 8] out;  stats: device float[2] out = {mean_density, threshold used}
 *   workspace: 256-byte aligned, envidr_density_workspace_bytes(cascade, grid_size, n) */
int envidr_density_grid_update(const envidr_field* field, float* density_grid, const int32_t* coords, const float* noise, uint32_t n,
                               const envidr_density_opts* opts, uint8_t* bitfield, float* stats, void* workspace,
                               uint64_t workspace_bytes, envidr_stream_t stream);
/* mark_untrained_grid: poses [B,4,4] camera-to-world (row-major, device), B <= 1024 per call.  count == NULL:
 * density_grid[cells seen by none of the B cameras] = -1.  count != NULL (int32 [cascade * H^3], zeroed by the caller): the
 * per-cell camera counts are accumulated into it instead (several calls for > 1024 poses), then envidr_mark_untrained_apply. */
int envidr_mark_untrained_grid(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, float bound, uint32_t cascade,
                               uint32_t grid_size, float* density_grid, int32_t* count, envidr_stream_t stream);
int envidr_mark_untrained_apply(const int32_t* count, uint32_t n, float* density_grid, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step for the trainable state of the path.
 * Replaces optimizer.zero_grad() + torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15).step() of the reference's training loop
 * (main_nerf.py:150, nerf/utils.py:1079-1087; torch/optim/adam.py, no weight decay / amsgrad / maximize) by one launch over
 * all parameter tensors.  step_size = lr / (1 - beta1^step) and bias_correction2_sqrt = sqrt(1 - beta2^step) are computed by
 * the caller (torch computes them as Python doubles on the host), per tensor because lr is per parameter group.
 * ---------------------------------------------------------------------------------------------- */
#define ENVIDR_ADAM_MAX_TENSORS 64 /* per launch; longer lists are split */
typedef struct envidr_adam_tensor {
    float* param;                 /* [n] updated in place        */
    float* grad;                  /* [n] (cleared when zero_grad) */
    float* exp_avg;               /* [n] first moment            */
    float* exp_avg_sq;            /* [n] second moment           */
    uint64_t n;
    float step_size;              /* lr / bias_correction1       */
    float bias_correction2_sqrt;
} envidr_adam_tensor;
/* tensors: HOST array of n_tensors descriptors (device pointers inside).  div_mode: 0 = sqrt(v) / bc2 as an IEEE division
 * (torch's foreach implementation, the CUDA default), 1 = sqrt(v) * fp32(1 / bc2) (torch's single-tensor implementation).
 * beta1 / beta2 / eps are doubles because torch derives the fp32 kernel constants from Python doubles: fp32(1 - beta1), not 1.0f - fp32(beta1). */
int envidr_adam_step(const envidr_adam_tensor* tensors, uint32_t n_tensors, double beta1, double beta2, double eps, int zero_grad,
                     int div_mode, envidr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Ray generation and loss epilogue: the steps either side of the training branch.
 * ---------------------------------------------------------------------------------------------- */
/* get_rays (nerf/utils.py:110-209).  poses [B,4,4] camera-to-world (device, row-major); intrinsics4 = HOST {fx, fy, cx, cy};
 * inds: device int64 [N] flat pixel indices h * W + w shared by all B poses (the reference's `inds.expand([B, N])`), or NULL for
 * all H * W pixels (N must be H * W).  rays_o, rays_d [B, N, 3] out.  The random choice of `inds` stays with the caller
 * (torch.randint / multinomial, utils.py:132-186), as does the gather of the ground-truth pixels. */
int envidr_get_rays(const float* poses, uint32_t B, const float* intrinsics4, uint32_t H, uint32_t W, const int64_t* inds, uint32_t N,
                    float* rays_o, float* rays_d, envidr_stream_t stream);

/* Loss terms of Trainer.train_step (nerf/utils.py:661-662 colour, :712-717 mask BCE, :735-747 back-sdf, :762-776 Cauchy,
 * :793-798 eikonal) over the outputs of run_cuda's training branch, including the auxiliary block that prepares them
 * (nerf/render_func/cuda_ray.py:173-211).  A term whose weight is 0 is skipped, as in the reference; backsdf_w > 0 switches the
 * auxiliary block on (point mask; the Cauchy term then averages over the masked samples, cuda_ray.py:209). */
typedef struct envidr_loss_in {
    const float* image;          /* [N,3] outputs['image']                                   */
    const float* weights_sum;    /* [N]   outputs['weights_sum'] (mask term)                 */
    const float* gt_rgb;         /* [N,3]                                                    */
    const float* gt_mask;        /* [N] alpha_mask, or NULL (no mask term)                   */
    const float* sdfs;           /* [M]                                                      */
    const float* sdf_gradients;  /* [M,3] or NULL (no eikonal term)                          */
    const float* weights;        /* [M] compositing weights (back-sdf; carry no gradient, raymarching.py:291) */
    const float* deltas;         /* [M,2]                                                    */
    const int32_t* rays;         /* [n_rays,3] (ray, offset, count) of march_rays_train      */
    const float* beta;           /* device scalar: sdf_density.get_beta()                    */
} envidr_loss_in;
typedef struct envidr_loss_opts {
    uint32_t N, M, n_rays;
    int32_t color_l1;            /* 1: L1Loss (color_l1_loss), 0: MSELoss (main_nerf.py:84-86) */
    int32_t backsdf_mean;        /* backsdf_mode != 'sum': divide by 1 + sum of the masked weights */
    float color_w, mask_w, cauchy_w, eikonal_w, backsdf_w, backsdf_thresh;
} envidr_loss_opts;
uint64_t envidr_train_loss_workspace_bytes(uint32_t M);
/* terms: device float[8] = {total, colour, mask, Cauchy, eikonal, back-sdf, #samples in the point mask, back-sdf denominator}. */
int envidr_train_loss_forward(const envidr_loss_in* in, const envidr_loss_opts* opts, float* terms, void* workspace, uint64_t workspace_bytes,
                              envidr_stream_t stream);
/* Gradients of grad_loss * total (grad_loss: device scalar or NULL = 1).  Needs `terms` and the SAME workspace as the forward
 * call (it holds the point-mask flags).  Any output pointer may be NULL.  d_image [N,3], d_weights_sum [N], d_sdfs [M],
 * d_sdf_gradients [M,3] are overwritten. */
int envidr_train_loss_backward(const envidr_loss_in* in, const envidr_loss_opts* opts, const float* terms, const float* grad_loss,
                               float* d_image, float* d_weights_sum, float* d_sdfs, float* d_sdf_gradients, void* workspace,
                               uint64_t workspace_bytes, envidr_stream_t stream);

/* Instrumentation (bench.py): launches issued by envidr_render_rays in this process; CUDA-event timing of the
 * field kernel (the dominant kernel) on its launch stream. */
uint64_t envidr_launch_count(void);
int envidr_render_timing(int enable);
int envidr_render_field_time(float* total_ms, uint32_t* launches);
/* Profiling hook: when `buf` (device, `capacity` uint64 slots) is non-NULL, CTA 0 of every following tensor-core env_net launch
 * records a clock64() timeline of its roles into it (layout in csrc/field_tc.cu, kProf*); NULL switches it off. */
int envidr_debug_env_tc_timeline(uint64_t* buf, uint32_t capacity);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ENVIDR_B200_H_ */
