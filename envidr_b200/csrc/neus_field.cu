// neus_field.cu -- the glue between the dense layers of the NeuS-style geometry network (BASELINE config 4).
//
// Reference: NeRFNetwork with use_neus_sdf + encoding_pos=frequency + geometric_init (nerf/network.py:154-222 construction,
// :415-421 forward with skip_layers, nerf/renderer.py:182-198 normal = autograd d sdf / d x): FreqEncoder(39) -> 8 x 256 weight-normed
// layers, Softplus(beta = 100) between them, h = cat([h, x]) / sqrt(2) in front of the skip layer.  The reference runs this as
// cuBLAS GEMMs + elementwise torch kernels and obtains the normal from an autograd graph.  Here the dense layers (forward, and the
// reverse pass g <- (g . softplus') W for the analytic normal) run on envidr_linear_tc (tcgen05, fp16 hi/lo split), and the kernels
// below do everything between them in one pass each; envidr_b200/neus_field.py drives the chain.
#include <math.h>
#include "common.cuh"

namespace envidr {
namespace {

constexpr int kNfBlock = 256;

// torch.nn.Softplus(beta, threshold = 20): x if beta x > 20 else log1p(exp(beta x)) / beta;  derivative = sigmoid(beta x)
__global__ void __launch_bounds__(kNfBlock) k_softplus_fwd(const float* __restrict__ z, uint64_t n, float beta, float* __restrict__ h,
                                                          float* __restrict__ s) {
    for (uint64_t i = (uint64_t)blockIdx.x * kNfBlock + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kNfBlock) {
        const float bz = beta * z[i];
        const float e = expf(-fabsf(bz));                                  // stable both ways
        const float sp = (bz > 20.0f) ? z[i] : (fmaxf(bz, 0.0f) + log1pf(e)) / beta;
        const float sg = (bz >= 0.0f) ? 1.0f / (1.0f + e) : e / (1.0f + e);
        h[i] = sp;
        if (s) s[i] = sg;
    }
}

// out[m, j] = a[m, j] * b[m, j]   or, with a_row (one row broadcast over m), a_row[j] * b[m, j]
__global__ void __launch_bounds__(kNfBlock) k_mul(const float* __restrict__ a, const float* __restrict__ a_row, const float* __restrict__ b,
                                                 uint64_t n, uint32_t N, float* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * kNfBlock + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kNfBlock)
        out[i] = (a_row ? a_row[i % N] : a[i]) * b[i];
}

// out[m] = cat(h[m, 0:Nh], x[m, 0:Nx]) * scale
__global__ void __launch_bounds__(kNfBlock) k_skip_cat(const float* __restrict__ h, const float* __restrict__ x, uint64_t M, uint32_t Nh, uint32_t Nx,
                                                      float scale, float* __restrict__ out) {
    const uint32_t N = Nh + Nx;
    const uint64_t n = M * N;
    for (uint64_t i = (uint64_t)blockIdx.x * kNfBlock + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kNfBlock) {
        const uint64_t m = i / N;
        const uint32_t j = (uint32_t)(i - m * N);
        out[i] = (j < Nh ? h[m * Nh + j] : x[m * Nx + (j - Nh)]) * scale;
    }
}

// reverse of k_skip_cat: gh[m] = g[m, 0:Nh] * scale ; gx[m] (+)= g[m, Nh:] * scale
__global__ void __launch_bounds__(kNfBlock) k_skip_split(const float* __restrict__ g, uint64_t M, uint32_t Nh, uint32_t Nx, float scale,
                                                        float* __restrict__ gh, float* __restrict__ gx, int accumulate) {
    const uint32_t N = Nh + Nx;
    const uint64_t n = M * N;
    for (uint64_t i = (uint64_t)blockIdx.x * kNfBlock + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kNfBlock) {
        const uint64_t m = i / N;
        const uint32_t j = (uint32_t)(i - m * N);
        const float v = g[i] * scale;
        if (j < Nh) gh[m * Nh + j] = v;
        else if (accumulate) gx[m * Nx + (j - Nh)] += v;
        else gx[m * Nx + (j - Nh)] = v;
    }
}

__device__ __forceinline__ float nf_softplus(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// Head of the geometry network -> what the shading kernels consume.  h [M, ld] = last layer output (sdf, geo_feat[G], roughness,
// blend: network.py:424-448), grad_x [M,3] = d sdf / d x.  Writes the unit normal (F.normalize eps 1e-10, renderer.py:192), the
// roughness and the 32-float geometry record of the tensor-core path (csrc/field.cu: geo 0-15, n 16-18, n.w_o 19, roughness 20,
// blend 21, n_env 22-24, w_r 25-27; renderer.py:147-180 get_color_mlp_extra_params with the optional env rotation).
__global__ void __launch_bounds__(kNfBlock) k_neus_records(const float* __restrict__ h, uint32_t ld, const float* __restrict__ grad_x,
                                                          const float* __restrict__ dirs, uint32_t M, uint32_t G, float rough_bias,
                                                          float rough_act_scale, float rough_scale, int has_blend, int has_rot, float r0, float r1,
                                                          float r2, float r3, float r4, float r5, float r6, float r7, float r8,
                                                          float* __restrict__ sdf, float* __restrict__ normal, float* __restrict__ roughness,
                                                          float* __restrict__ rec) {
    const uint32_t m = blockIdx.x * kNfBlock + threadIdx.x;
    if (m >= M) return;
    const float* q = h + (size_t)m * ld;
    const float gx = grad_x[3 * (size_t)m], gy = grad_x[3 * (size_t)m + 1], gz = grad_x[3 * (size_t)m + 2];
    const float gn = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-10f);
    const float nx = gx / gn, ny = gy / gn, nz = gz / gn;
    if (sdf) sdf[m] = q[0];
    if (normal) { normal[3 * (size_t)m] = nx; normal[3 * (size_t)m + 1] = ny; normal[3 * (size_t)m + 2] = nz; }
    const float rough = rough_act_scale * nf_softplus(q[1 + G] + rough_bias) * rough_scale;
    if (roughness) roughness[m] = rough;
    if (!rec) return;
    float ss = 0.f;
    for (uint32_t i = 0; i < G; i++) ss += q[1 + i] * q[1 + i];
    const float ginv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    const float blend = has_blend ? 1.0f / (1.0f + expf(-q[2 + G])) : 0.0f;
    const float wox = -dirs[3 * (size_t)m], woy = -dirs[3 * (size_t)m + 1], woz = -dirs[3 * (size_t)m + 2];
    const float ndot = nx * wox + ny * woy + nz * woz;
    float wrx = 2 * ndot * nx - wox, wry = 2 * ndot * ny - woy, wrz = 2 * ndot * nz - woz;
    float nex = nx, ney = ny, nez = nz;
    if (has_rot) {                                                         // v @ R (row vector times rot_theta[:3,:3])
        const float a0 = wrx * r0 + wry * r3 + wrz * r6, a1 = wrx * r1 + wry * r4 + wrz * r7, a2 = wrx * r2 + wry * r5 + wrz * r8;
        wrx = a0; wry = a1; wrz = a2;
        const float b0 = nx * r0 + ny * r3 + nz * r6, b1 = nx * r1 + ny * r4 + nz * r7, b2 = nx * r2 + ny * r5 + nz * r8;
        nex = b0; ney = b1; nez = b2;
    }
    float* o = rec + (size_t)m * 32;
    for (uint32_t i = 0; i < 16; i++) o[i] = (i < G) ? q[1 + i] * ginv : 0.0f;
    o[16] = nx; o[17] = ny; o[18] = nz; o[19] = ndot;
    o[20] = rough; o[21] = blend; o[22] = nex; o[23] = ney;
    o[24] = nez; o[25] = wrx; o[26] = wry; o[27] = wrz;
    o[28] = 0.f; o[29] = 0.f; o[30] = 0.f; o[31] = 0.f;
}

inline uint32_t nf_grid(uint64_t n) { const uint64_t b = (n + kNfBlock - 1) / kNfBlock; return (uint32_t)(b < (uint64_t)kSMs * 16 ? (b ? b : 1) : (uint64_t)kSMs * 16); }

}  // namespace
}  // namespace envidr

using namespace envidr;

extern "C" {

int envidr_softplus_forward(const float* z, uint64_t n, float beta, float* h, float* dh, envidr_stream_t stream) {
    if (n == 0) return 0;
    ENVIDR_REQUIRE(z && h && beta > 0, ENVIDR_E_BADARG, "null pointer / beta <= 0");
    k_softplus_fwd<<<nf_grid(n), kNfBlock, 0, as_stream(stream)>>>(z, n, beta, h, dh);
    g_launches += 1;
    return check_launch("softplus_forward");
}

int envidr_mul_rows(const float* a, const float* a_row, const float* b, uint64_t M, uint32_t N, float* out, envidr_stream_t stream) {
    if (M == 0 || N == 0) return 0;
    ENVIDR_REQUIRE((a || a_row) && b && out, ENVIDR_E_BADARG, "null pointer");
    k_mul<<<nf_grid(M * N), kNfBlock, 0, as_stream(stream)>>>(a, a_row, b, M * N, N, out);
    g_launches += 1;
    return check_launch("mul_rows");
}

int envidr_skip_concat_forward(const float* h, const float* x, uint64_t M, uint32_t Nh, uint32_t Nx, float scale, float* out,
                               envidr_stream_t stream) {
    if (M == 0) return 0;
    ENVIDR_REQUIRE(h && x && out, ENVIDR_E_BADARG, "null pointer");
    k_skip_cat<<<nf_grid(M * (Nh + Nx)), kNfBlock, 0, as_stream(stream)>>>(h, x, M, Nh, Nx, scale, out);
    g_launches += 1;
    return check_launch("skip_concat_forward");
}

int envidr_skip_concat_backward(const float* g, uint64_t M, uint32_t Nh, uint32_t Nx, float scale, float* gh, float* gx, int accumulate,
                                envidr_stream_t stream) {
    if (M == 0) return 0;
    ENVIDR_REQUIRE(g && gh && gx, ENVIDR_E_BADARG, "null pointer");
    k_skip_split<<<nf_grid(M * (Nh + Nx)), kNfBlock, 0, as_stream(stream)>>>(g, M, Nh, Nx, scale, gh, gx, accumulate);
    g_launches += 1;
    return check_launch("skip_concat_backward");
}

int envidr_neus_records(const float* h, uint32_t ld, const float* grad_x, const float* dirs, uint32_t M, uint32_t geo_feat_dim,
                        float roughness_bias, float roughness_act_scale, float roughness_scale, int has_blend, const float* rot9 /* host or NULL */,
                        float* sdf, float* normal, float* roughness, float* rec, envidr_stream_t stream) {
    if (M == 0) return 0;
    ENVIDR_REQUIRE(h && grad_x && dirs, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(geo_feat_dim <= 15 && ld >= 2 + geo_feat_dim + (has_blend ? 1u : 0u), ENVIDR_E_BADARG, "head layout: 1 + geo_feat_dim (+ roughness, blend) <= ld");
    const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const float* r = rot9 ? rot9 : I;
    k_neus_records<<<ceil_div(M, kNfBlock), kNfBlock, 0, as_stream(stream)>>>(h, ld, grad_x, dirs, M, geo_feat_dim, roughness_bias, roughness_act_scale,
                                                                            roughness_scale, has_blend, rot9 ? 1 : 0, r[0], r[1], r[2], r[3], r[4], r[5],
                                                                            r[6], r[7], r[8], sdf, normal, roughness, rec);
    g_launches += 1;
    return check_launch("neus_records");
}

}  // extern "C"
