// neus.cu -- NeuSDensity (reference nerf/network.py:46-102): per-sample opacity of the NeuS-style SDF configs (use_neus_sdf),
// consumed by the compositors with input_alpha = 1 (cuda_ray.py:121,154,308,318).  SURVEY.md 8 a-6.
//
//   inv_s  = clip(exp(10 * variance), 1e-6, 1e6)
//   cos    = dirs . gradients;  iter_cos = -(relu(-0.5 cos + 0.5) (1 - r) + relu(-cos) r)         (r = cos_anneal_ratio)
//   next   = sdf + iter_cos * dist / 2,  prev = sdf - iter_cos * dist / 2      (no gradients: next = sdf - dist / 2, prev = sdf + dist / 2)
//   alpha  = clip((sigmoid(prev inv_s) - sigmoid(next inv_s) + 1e-5) / (sigmoid(prev inv_s) + 1e-5), 0, 1)
//
// The reference evaluates this as ~25 elementwise torch kernels forward (and as many backward); here one kernel each way.
// Backward: d alpha / d {sdf, gradients, variance}; the variance gradient is a block-sum (fp64) combined in block order by the
// last block (deterministic).  Streaming, HBM-bound: 32 B in + 4 B out per sample forward.
#include "common.cuh"

namespace envidr {
namespace {

constexpr int kNBlock = 256;

struct NeusTerms { float prev, next, pc, nc, raw, dic_dcos; };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ NeusTerms neus_eval(float sdf, const float* __restrict__ d, const float* __restrict__ g, float dist, float inv_s, float r) {
    NeusTerms t;
    t.dic_dcos = 0.0f;
    if (g) {
        const float c = d[0] * g[0] + d[1] * g[1] + d[2] * g[2];
        const float a = -c * 0.5f + 0.5f, b = -c;
        const float ic = -(fmaxf(a, 0.0f) * (1.0f - r) + fmaxf(b, 0.0f) * r);
        // d iter_cos / d cos = 0.5 (1 - r) [a > 0] + r [b > 0]        (relu'(0) = 0 as in torch)
        t.dic_dcos = (a > 0.0f ? 0.5f * (1.0f - r) : 0.0f) + (b > 0.0f ? r : 0.0f);
        t.next = sdf + ic * dist * 0.5f;
        t.prev = sdf - ic * dist * 0.5f;
    } else {
        t.next = sdf - dist * 0.5f;
        t.prev = sdf + dist * 0.5f;
    }
    t.pc = sigmoidf_(t.prev * inv_s);
    t.nc = sigmoidf_(t.next * inv_s);
    t.raw = (t.pc - t.nc + 1e-5f) / (t.pc + 1e-5f);
    return t;
}

__device__ __forceinline__ float load_inv_s(const float* variance, bool* inside) {
    const float e = expf(*variance * 10.0f);
    *inside = e >= 1e-6f && e <= 1e6f;                                   // clip passes the gradient on [min, max]
    return fminf(fmaxf(e, 1e-6f), 1e6f);
}

__global__ void __launch_bounds__(kNBlock) k_neus_fwd(const float* __restrict__ sdf, const float* __restrict__ dirs, const float* __restrict__ grads,
                                                      const float* __restrict__ dists, float dist_scalar, const float* __restrict__ variance,
                                                      float r, uint32_t M, float* __restrict__ alpha) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    bool inside;
    const float inv_s = load_inv_s(variance, &inside);
    const NeusTerms t = neus_eval(sdf[i], dirs + 3 * (size_t)i, grads ? grads + 3 * (size_t)i : nullptr, dists ? dists[i] : dist_scalar, inv_s, r);
    alpha[i] = fminf(fmaxf(t.raw, 0.0f), 1.0f);
}

__global__ void __launch_bounds__(kNBlock) k_neus_bwd(const float* __restrict__ g_alpha, const float* __restrict__ sdf, const float* __restrict__ dirs,
                                                      const float* __restrict__ grads, const float* __restrict__ dists, float dist_scalar,
                                                      const float* __restrict__ variance, float r, uint32_t M, float* __restrict__ g_sdf,
                                                      float* __restrict__ g_grads, double* __restrict__ partial, uint32_t* __restrict__ ticket,
                                                      float* __restrict__ g_variance) {
    __shared__ double red[kNBlock / 32];
    __shared__ double fin[kNBlock];
    __shared__ bool last;
    bool inside;
    const float inv_s = load_inv_s(variance, &inside);
    double acc = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float* d = dirs + 3 * (size_t)i;
        const float dist = dists ? dists[i] : dist_scalar;
        const NeusTerms t = neus_eval(sdf[i], d, grads ? grads + 3 * (size_t)i : nullptr, dist, inv_s, r);
        float gs = 0.0f, gic = 0.0f;
        if (t.raw >= 0.0f && t.raw <= 1.0f) {
            const float go = g_alpha[i];
            const float den = t.pc + 1e-5f;
            const float d_pc = go * t.nc / (den * den);                               // d raw / d pc = nc / (pc + e)^2  (the e terms cancel)
            const float d_nc = -go / den;
            const float s_p = t.pc * (1.0f - t.pc), s_n = t.nc * (1.0f - t.nc);       // sigmoid'
            const float d_prev = d_pc * s_p * inv_s, d_next = d_nc * s_n * inv_s;
            gs = d_prev + d_next;
            gic = (d_next - d_prev) * dist * 0.5f;                                    // next = sdf + ic d / 2, prev = sdf - ic d / 2
            acc += (double)(d_pc * s_p * t.prev + d_nc * s_n * t.next);               // d / d inv_s
        }
        if (g_sdf) g_sdf[i] = gs;
        if (g_grads) {
            const float k = grads ? gic * t.dic_dcos : 0.0f;                          // cos = dirs . gradients
            g_grads[3 * (size_t)i] = k * d[0]; g_grads[3 * (size_t)i + 1] = k * d[1]; g_grads[3 * (size_t)i + 2] = k * d[2];
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < kNBlock / 32; w++) b += red[w];
        partial[blockIdx.x] = b;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double t = 0.0;
        for (uint32_t b = threadIdx.x; b < gridDim.x; b += kNBlock) t += __ldcg(partial + b);
        fin[threadIdx.x] = t;
        __syncthreads();
        for (int o = kNBlock / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) fin[threadIdx.x] += fin[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (g_variance) *g_variance = inside ? (float)(fin[0] * 10.0 * (double)inv_s) : 0.0f;   // d inv_s / d variance = 10 inv_s
            *ticket = 0;
        }
    }
}

constexpr uint32_t kNeusBlocks = kSMs * 4;

}  // namespace
}  // namespace envidr

using namespace envidr;

extern "C" {

int envidr_neus_alpha_forward(const float* sdf, const float* dirs, const float* gradients, const float* dists, float dist_scalar,
                              const float* variance, float cos_anneal_ratio, uint32_t M, float* alpha, envidr_stream_t stream) {
    if (M == 0) return 0;
    ENVIDR_REQUIRE(sdf && dirs && variance && alpha, ENVIDR_E_BADARG, "null pointer");
    k_neus_fwd<<<ceil_div(M, kNBlock), kNBlock, 0, as_stream(stream)>>>(sdf, dirs, gradients, dists, dist_scalar, variance, cos_anneal_ratio, M, alpha);
    g_launches += 1;
    return check_launch("neus_alpha_forward");
}

uint64_t envidr_neus_workspace_bytes(void) { return (uint64_t)kNeusBlocks * 8 + 256; }

int envidr_neus_alpha_backward(const float* grad_alpha, const float* sdf, const float* dirs, const float* gradients, const float* dists,
                               float dist_scalar, const float* variance, float cos_anneal_ratio, uint32_t M, float* grad_sdf,
                               float* grad_gradients, float* grad_variance, void* workspace, uint64_t workspace_bytes, envidr_stream_t stream) {
    ENVIDR_REQUIRE(workspace && workspace_bytes >= envidr_neus_workspace_bytes() && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                   ENVIDR_E_WORKSPACE, "workspace: envidr_neus_workspace_bytes(), 8-byte aligned");
    cudaStream_t st = as_stream(stream);
    if (M == 0) {
        if (grad_variance) cudaMemsetAsync(grad_variance, 0, 4, st);
        return 0;
    }
    ENVIDR_REQUIRE(grad_alpha && sdf && dirs && variance, ENVIDR_E_BADARG, "null pointer");
    double* partial = reinterpret_cast<double*>(workspace);
    uint32_t* ticket = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + (uint64_t)kNeusBlocks * 8);
    cudaMemsetAsync(ticket, 0, 4, st);
    const uint32_t blocks = min(kNeusBlocks, ceil_div(M, kNBlock));
    k_neus_bwd<<<blocks, kNBlock, 0, st>>>(grad_alpha, sdf, dirs, gradients, dists, dist_scalar, variance, cos_anneal_ratio, M, grad_sdf,
                                           grad_gradients, partial, ticket, grad_variance);
    g_launches += 1;
    return check_launch("neus_alpha_backward");
}

}  // extern "C"
