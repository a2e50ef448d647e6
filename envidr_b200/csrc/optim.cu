// optim.cu -- fused multi-tensor Adam step for the trainable state of the path.
//
// The reference trains with torch.optim.Adam(model.get_params(...), betas=(0.9, 0.99), eps=1e-15) (main_nerf.py:150,
// nerf/network.py:772-819) and calls optimizer.zero_grad() + optimizer.step() every iteration (nerf/utils.py:1079-1087).  The
// parameters are the hash table (12.2 M floats at toaster dims) plus ~25 small MLP tensors.  torch's default (foreach) Adam walks
// the parameter set once per elementary op (lerp, mul, addcmul, sqrt, div, add, addcdiv: seven passes of 2-3 streams each);
// here ONE launch reads p, g, m, v and writes p, m, v (and clears g) -- 28 (32) algorithmic bytes per parameter, the HBM floor.
//
// Arithmetic per element = torch/optim/adam.py::_single_tensor_adam / _multi_tensor_adam, non-capturable, no weight decay,
// no amsgrad, maximize off:
//     exp_avg.lerp_(grad, 1 - beta1)                                  -> m + w * (g - m)              (ATen Lerp.h, w < 0.5)
//     exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)    -> v * beta2 + (1 - beta2) * (g * g)
//     denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
//     param.addcdiv_(exp_avg, denom, value=-step_size)                -> p + (-step_size) * (m / denom)
// step_size = lr / (1 - beta1^step) and bias_correction2_sqrt = sqrt(1 - beta2^step) are Python doubles on the host in torch;
// the host side of this library computes them the same way and passes them rounded to fp32 (what the torch kernels receive).
#include "common.cuh"

namespace envidr {
namespace {

constexpr int kABlock = 256;
constexpr int kAVec = 4;
constexpr uint32_t kAChunk = kABlock * kAVec * 4;      // elements per block: 4 float4 per thread

struct AdamTable {
    envidr_adam_tensor t[ENVIDR_ADAM_MAX_TENSORS];
    uint32_t chunk_end[ENVIDR_ADAM_MAX_TENSORS];         // inclusive prefix sum of chunks
    uint32_t n_tensors;
};

__device__ __forceinline__ void adam1(float& p, float& g, float& m, float& v, float w1, float beta2, float w2, float eps, float rbc2,
                                      float bc2, float neg_step, int div_mode) {
    m = fmaf(w1, __fsub_rn(g, m), m);
    v = fmaf(w2, __fmul_rn(g, g), __fmul_rn(v, beta2));
    const float s = __fsqrt_rn(v);
    const float denom = __fadd_rn(div_mode ? __fmul_rn(s, rbc2) : __fdiv_rn(s, bc2), eps);
    p = fmaf(neg_step, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(kABlock) k_adam(const __grid_constant__ AdamTable T, float w1, float beta2, float w2, float eps, int zero_grad,
                                                  int div_mode) {
    // block -> tensor: binary search in the chunk prefix sums
    uint32_t lo = 0, hi = T.n_tensors - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (blockIdx.x < T.chunk_end[mid]) hi = mid; else lo = mid + 1;
    }
    const envidr_adam_tensor& d = T.t[lo];
    const uint32_t chunk = blockIdx.x - (lo ? T.chunk_end[lo - 1] : 0);
    const uint64_t base = (uint64_t)chunk * kAChunk;
    const float neg_step = -d.step_size, bc2 = d.bias_correction2_sqrt, rbc2 = __frcp_rn(d.bias_correction2_sqrt);
    float* __restrict__ P = d.param; float* __restrict__ G = d.grad; float* __restrict__ M = d.exp_avg; float* __restrict__ V = d.exp_avg_sq;
    const bool vec = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(M) |
                       reinterpret_cast<uintptr_t>(V)) & 15) == 0;
    if (vec && base + kAChunk <= d.n) {
        // whole chunk: 16 independent 16-byte loads per thread in flight before the first use
        float4 p[4], g[4], m[4], v[4];
        #pragma unroll
        for (int it = 0; it < 4; it++) {
            const uint64_t e = base + ((uint64_t)it * kABlock + threadIdx.x) * kAVec;
            p[it] = *reinterpret_cast<const float4*>(P + e); g[it] = *reinterpret_cast<const float4*>(G + e);
            m[it] = *reinterpret_cast<const float4*>(M + e); v[it] = *reinterpret_cast<const float4*>(V + e);
        }
        #pragma unroll
        for (int it = 0; it < 4; it++) {
            const uint64_t e = base + ((uint64_t)it * kABlock + threadIdx.x) * kAVec;
            adam1(p[it].x, g[it].x, m[it].x, v[it].x, w1, beta2, w2, eps, rbc2, bc2, neg_step, div_mode);
            adam1(p[it].y, g[it].y, m[it].y, v[it].y, w1, beta2, w2, eps, rbc2, bc2, neg_step, div_mode);
            adam1(p[it].z, g[it].z, m[it].z, v[it].z, w1, beta2, w2, eps, rbc2, bc2, neg_step, div_mode);
            adam1(p[it].w, g[it].w, m[it].w, v[it].w, w1, beta2, w2, eps, rbc2, bc2, neg_step, div_mode);
            *reinterpret_cast<float4*>(P + e) = p[it]; *reinterpret_cast<float4*>(M + e) = m[it]; *reinterpret_cast<float4*>(V + e) = v[it];
            if (zero_grad) *reinterpret_cast<float4*>(G + e) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    for (int it = 0; it < 4; it++) {                     // ragged tail / unaligned tensor
        const uint64_t e = base + ((uint64_t)it * kABlock + threadIdx.x) * kAVec;
        for (uint64_t k = e; k < e + kAVec && k < d.n; k++) {
            float p = P[k], g = G[k], m = M[k], v = V[k];
            adam1(p, g, m, v, w1, beta2, w2, eps, rbc2, bc2, neg_step, div_mode);
            P[k] = p; M[k] = m; V[k] = v;
            if (zero_grad) G[k] = 0.f;
        }
    }
}

}  // namespace
}  // namespace envidr

using namespace envidr;

extern "C" int envidr_adam_step(const envidr_adam_tensor* tensors, uint32_t n_tensors, double beta1, double beta2, double eps, int zero_grad,
                                int div_mode, envidr_stream_t stream) {
    ENVIDR_REQUIRE(tensors || n_tensors == 0, ENVIDR_E_BADARG, "null pointer");
    cudaStream_t st = as_stream(stream);
    uint32_t i = 0;
    while (i < n_tensors) {
        AdamTable T;
        T.n_tensors = 0;
        uint32_t chunks = 0;
        for (; i < n_tensors && T.n_tensors < ENVIDR_ADAM_MAX_TENSORS; i++) {
            const envidr_adam_tensor& d = tensors[i];
            if (d.n == 0) continue;
            ENVIDR_REQUIRE(d.param && d.grad && d.exp_avg && d.exp_avg_sq, ENVIDR_E_BADARG, "null tensor pointer");
            ENVIDR_REQUIRE(d.bias_correction2_sqrt > 0.0f, ENVIDR_E_BADARG, "bias_correction2_sqrt must be positive (step >= 1)");
            const uint64_t c = (d.n + kAChunk - 1) / kAChunk;
            ENVIDR_REQUIRE(chunks + c < 0x7fffffffu, ENVIDR_E_BADARG, "too many elements in one call");
            chunks += (uint32_t)c;
            T.t[T.n_tensors] = d;
            T.chunk_end[T.n_tensors] = chunks;
            T.n_tensors++;
        }
        if (chunks == 0) continue;
        // torch hands its kernels the Python doubles 1 - beta1, beta2, 1 - beta2, eps rounded to fp32 (not 1.0f - (float)beta1)
        k_adam<<<chunks, kABlock, 0, st>>>(T, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, zero_grad, div_mode);
        g_launches += 1;
    }
    return check_launch("adam_step");
}
