// gridenc.cu -- multi-resolution hash / tiled grid encoders for sm_100a.
//
// Replaces the reference's `_hashencoder` (smoothstep interpolation, first + second order backward;
// hashencoder/src/hashencoder.cu:103-595) and `_gridencoder` (linear interpolation, hash | tiled,
// align_corners; gridencoder/src/gridencoder.cu:75-342) extensions with one templated kernel family.
//
// Layouts are the reference's: inputs [B,D] in [0,1], table [T,C], offsets [L+1], outputs [L,B,C],
// dy_dx [B, L*D*C].  Level geometry is computed in-kernel with the reference's fp32 expression
// (scale = exp2f(l*S)*H - 1, res = ceil(scale)+1) so cell indices are bit-identical; the hash uses
// uint32 wrap-around arithmetic.
//
// B200 notes: the forward is a pure gather (2^D random C*4-byte reads per sample and level); the grid is
// (sample tiles) x (levels) so one CTA touches one level's table (dense levels stay L1/L2 resident, the whole
// 48.8 MB table of the default config fits the 126 MB L2).  Table rows are fetched with one vector load
// per corner (float2 for C=2, float4 for C>=4).  The backward scatters with vector red.global.add
// (float2/float4 atomics, sm_90+), halving/quartering the atomic count of the scalar reference kernel.
#include "gridenc.cuh"

namespace envidr {

// forward: one thread per (sample, level)
template <int D, int C>
__global__ void __launch_bounds__(256) k_encode_fwd(const EncMode m, const float* __restrict__ inputs, const float* __restrict__ table,
                                                   const int* __restrict__ offsets, float* __restrict__ outputs, uint32_t B,
                                                   uint32_t L, float S, uint32_t H, float* __restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const float* grid = table + (size_t)(uint32_t)offsets[level] * C;
    float* out = outputs + ((size_t)level * B + b) * C;
    float* jac = dy_dx ? dy_dx + (size_t)b * D * L * C + (size_t)level * D * C : nullptr;
    Cell<D> cell;
    if (!cell.setup(m, inputs + (size_t)b * D, offsets, level, S, H)) {
        #pragma unroll
        for (int c = 0; c < C; c++) out[c] = 0;
        if (jac) {
            #pragma unroll
            for (int i = 0; i < D * C; i++) jac[i] = 0;
        }
        return;
    }
    // gather all 2^D corner rows first (independent loads in flight), then blend
    float rows[1 << D][C];
    #pragma unroll
    for (uint32_t corner = 0; corner < (1u << D); corner++) {
        uint32_t pl[D];
        #pragma unroll
        for (int d = 0; d < D; d++) pl[d] = cell.pg[d] + ((corner >> d) & 1u);
        load_row<C>(grid + (size_t)cell_index<D>(m, cell.hashmap_size, cell.resolution, pl) * C, rows[corner]);
    }
    float acc[C];
    #pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0;
    #pragma unroll
    for (uint32_t corner = 0; corner < (1u << D); corner++) {
        float wt = 1;
        #pragma unroll
        for (int d = 0; d < D; d++) wt *= ((corner >> d) & 1u) ? cell.w[d] : 1 - cell.w[d];
        #pragma unroll
        for (int c = 0; c < C; c++) acc[c] += wt * rows[corner][c];
    }
    #pragma unroll
    for (int c = 0; c < C; c++) out[c] = acc[c];
    if (!jac) return;
    float jrow[D * C];          // the (sample, level) block of dy_dx: written with the widest stores its alignment allows (below)
    #pragma unroll
    for (int gd = 0; gd < D; gd++) {
        float g[C];
        #pragma unroll
        for (int c = 0; c < C; c++) g[c] = 0;
        #pragma unroll
        for (uint32_t sub = 0; sub < (1u << (D - 1)); sub++) {
            float wt = cell.scale;
            uint32_t corner = 0;
            #pragma unroll
            for (int nd = 0; nd < D - 1; nd++) {
                const int d = nd >= gd ? nd + 1 : nd;
                if ((sub >> nd) & 1u) { wt *= cell.w[d]; corner |= 1u << d; }
                else                  { wt *= 1 - cell.w[d]; }
            }
            #pragma unroll
            for (int c = 0; c < C; c++) {
                const float diff = rows[corner | (1u << gd)][c] - rows[corner][c];
                if (m.smooth) g[c] += wt * diff * cell.dw[gd];
                else          g[c] += wt * diff;
            }
        }
        #pragma unroll
        for (int c = 0; c < C; c++) jrow[gd * C + c] = g[c];
    }
    // dy_dx rows of neighbouring threads lie D*L*C floats apart, so every scalar store is its own 32-byte sector write at L2: the
    // six 4-byte stores of the D = 3, C = 2 case cost 4x the kernel's gather time.  A 24-byte block is 8-byte aligned and either
    // its first or its last 16 bytes are 16-byte aligned: one 16-byte and one 8-byte store.
    if constexpr (D * C == 6) {
        if ((reinterpret_cast<uintptr_t>(jac) & 15) == 0) {
            *reinterpret_cast<float4*>(jac) = make_float4(jrow[0], jrow[1], jrow[2], jrow[3]);
            *reinterpret_cast<float2*>(jac + 4) = make_float2(jrow[4], jrow[5]);
        } else if ((reinterpret_cast<uintptr_t>(jac) & 7) == 0) {
            *reinterpret_cast<float2*>(jac) = make_float2(jrow[0], jrow[1]);
            *reinterpret_cast<float4*>(jac + 2) = make_float4(jrow[2], jrow[3], jrow[4], jrow[5]);
        } else {
            #pragma unroll
            for (int i = 0; i < 6; i++) jac[i] = jrow[i];
        }
    } else {
        #pragma unroll
        for (int i = 0; i < D * C; i++) jac[i] = jrow[i];
    }
}

// backward into the table: one thread per (sample, level), vector atomics per corner
template <int D, int C>
__global__ void __launch_bounds__(256) k_encode_bwd_table(const EncMode m, const float* __restrict__ grad, const float* __restrict__ inputs,
                                                         const int* __restrict__ offsets, float* __restrict__ grad_table, uint32_t B,
                                                         uint32_t L, float S, uint32_t H) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    Cell<D> cell;
    if (!cell.setup(m, inputs + (size_t)b * D, offsets, level, S, H)) return;
    float* gt = grad_table + (size_t)(uint32_t)offsets[level] * C;
    float g[C];
    load_row<C>(grad + ((size_t)level * B + b) * C, g);
    #pragma unroll
    for (uint32_t corner = 0; corner < (1u << D); corner++) {
        float wt = 1;
        uint32_t pl[D];
        #pragma unroll
        for (int d = 0; d < D; d++) {
            const uint32_t bit = (corner >> d) & 1u;
            wt *= bit ? cell.w[d] : 1 - cell.w[d];
            pl[d] = cell.pg[d] + bit;
        }
        float v[C];
        #pragma unroll
        for (int c = 0; c < C; c++) v[c] = wt * g[c];
        atomic_add_row<C>(gt + (size_t)cell_index<D>(m, cell.hashmap_size, cell.resolution, pl) * C, v);
    }
}

// grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]   (hashencoder.cu:346-372)
template <int D, int C>
__global__ void __launch_bounds__(256) k_encode_bwd_input(const float* __restrict__ grad, const float* __restrict__ dy_dx,
                                                         float* __restrict__ grad_inputs, uint32_t B, uint32_t L) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float acc[D];
    #pragma unroll
    for (int d = 0; d < D; d++) acc[d] = 0;
    const float* jac = dy_dx + (size_t)b * L * D * C;
    for (uint32_t l = 0; l < L; l++) {
        float g[C];
        load_row<C>(grad + ((size_t)l * B + b) * C, g);
        #pragma unroll
        for (int d = 0; d < D; d++) {
            #pragma unroll
            for (int c = 0; c < C; c++) acc[d] += g[c] * jac[l * D * C + d * C + c];
        }
    }
    #pragma unroll
    for (int d = 0; d < D; d++) grad_inputs[(size_t)b * D + d] = acc[d];
}

// second-order backward of the hash encoder (hashencoder.cu:375-595)
template <int D, int C>
__global__ void __launch_bounds__(256) k_hash_second_bwd(const EncMode m, const float* __restrict__ grad, const float* __restrict__ inputs,
                                                        const int* __restrict__ offsets, const float* __restrict__ grad_grad_inputs,
                                                        const float* __restrict__ dy_dx, float* __restrict__ grad_grad,
                                                        float* __restrict__ grad2_table, uint32_t B, uint32_t L, float S, uint32_t H) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    float ggx[D];
    #pragma unroll
    for (int d = 0; d < D; d++) ggx[d] = grad_grad_inputs[(size_t)b * D + d];
    {   // d(grad_inputs)/d(grad): written for every sample (also out-of-range ones, where dy_dx is 0)
        const float* jac = dy_dx + (size_t)b * L * D * C + (size_t)level * D * C;
        float* og = grad_grad + ((size_t)level * B + b) * C;
        #pragma unroll
        for (int c = 0; c < C; c++) {
            float r = 0;
            #pragma unroll
            for (int d = 0; d < D; d++) r += ggx[d] * jac[d * C + c];
            og[c] = r;
        }
    }
    Cell<D> cell;
    if (!cell.setup(m, inputs + (size_t)b * D, offsets, level, S, H)) return;
    float g[C];
    load_row<C>(grad + ((size_t)level * B + b) * C, g);
    float cache[1 << D][C];
    #pragma unroll
    for (uint32_t k = 0; k < (1u << D); k++) {
        #pragma unroll
        for (int c = 0; c < C; c++) cache[k][c] = 0;
    }
    #pragma unroll
    for (int gd = 0; gd < D; gd++) {
        #pragma unroll
        for (uint32_t sub = 0; sub < (1u << (D - 1)); sub++) {
            float wt = cell.scale;
            uint32_t corner = 0;
            #pragma unroll
            for (int nd = 0; nd < D - 1; nd++) {
                const int d = nd >= gd ? nd + 1 : nd;
                if ((sub >> nd) & 1u) { wt *= cell.w[d]; corner |= 1u << d; }
                else                  { wt *= 1 - cell.w[d]; }
            }
            #pragma unroll
            for (int c = 0; c < C; c++) {
                const float v = wt * g[c] * ggx[gd] * cell.dw[gd];
                cache[corner | (1u << gd)][c] += v;
                cache[corner][c] -= v;
            }
        }
    }
    float* gt = grad2_table + (size_t)(uint32_t)offsets[level] * C;
    #pragma unroll
    for (uint32_t corner = 0; corner < (1u << D); corner++) {
        uint32_t pl[D];
        #pragma unroll
        for (int d = 0; d < D; d++) pl[d] = cell.pg[d] + ((corner >> d) & 1u);
        atomic_add_row<C>(gt + (size_t)cell_index<D>(m, cell.hashmap_size, cell.resolution, pl) * C, cache[corner]);
    }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
template <int D, int C>
static int launch_fwd(const EncMode m, const float* inputs, const float* table, const int* offsets, float* outputs, uint32_t B,
                      uint32_t L, float S, uint32_t H, float* dy_dx, cudaStream_t st) {
    const dim3 grid(ceil_div(B, 256), L);
    k_encode_fwd<D, C><<<grid, 256, 0, st>>>(m, inputs, table, offsets, outputs, B, L, S, H, dy_dx);
    return check_launch("encode_forward");
}
template <int D, int C>
static int launch_bwd(const EncMode m, const float* grad, const float* inputs, const int* offsets, float* grad_table, uint32_t B,
                      uint32_t L, float S, uint32_t H, const float* dy_dx, float* grad_inputs, cudaStream_t st) {
    const dim3 grid(ceil_div(B, 256), L);
    k_encode_bwd_table<D, C><<<grid, 256, 0, st>>>(m, grad, inputs, offsets, grad_table, B, L, S, H);
    if (dy_dx && grad_inputs) k_encode_bwd_input<D, C><<<ceil_div(B, 256), 256, 0, st>>>(grad, dy_dx, grad_inputs, B, L);
    return check_launch("encode_backward");
}
template <int D, int C>
static int launch_second(const EncMode m, const float* grad, const float* inputs, const int* offsets, const float* ggx,
                         const float* dy_dx, float* grad_grad, float* grad2_table, uint32_t B, uint32_t L, float S, uint32_t H,
                         cudaStream_t st) {
    const dim3 grid(ceil_div(B, 256), L);
    k_hash_second_bwd<D, C><<<grid, 256, 0, st>>>(m, grad, inputs, offsets, ggx, dy_dx, grad_grad, grad2_table, B, L, S, H);
    return check_launch("hash_encode_second_backward");
}

#define DISPATCH_C(D_, CALL)                                                       \
    switch (C) {                                                                   \
        case 1: return CALL(D_, 1);                                                \
        case 2: return CALL(D_, 2);                                                \
        case 4: return CALL(D_, 4);                                                \
        case 8: return CALL(D_, 8);                                                \
        default: set_error("GridEncoding: C must be 1, 2, 4, or 8."); return ENVIDR_E_UNSUPPORTED; \
    }

// per-level `scale` exactly as the encoder kernels compute it (exp2f is ex2.approx on the device)
__global__ void k_level_scales(float S, uint32_t H, uint32_t L, float* __restrict__ out) {
    const uint32_t level = threadIdx.x;
    if (level < L) out[level] = exp2f(level * S) * H - 1.0f;
}

}  // namespace envidr

using namespace envidr;

extern "C" {

int envidr_debug_level_scales(float S, uint32_t H, uint32_t L, float* scales, envidr_stream_t stream) {
    ENVIDR_REQUIRE(scales && L <= 64, ENVIDR_E_BADARG, "bad arguments");
    k_level_scales<<<1, 64, 0, as_stream(stream)>>>(S, H, L, scales);
    return check_launch("debug_level_scales");
}

int envidr_hash_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs, uint32_t B,
                               uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, float* dy_dx,
                               envidr_stream_t stream) {
    ENVIDR_REQUIRE(inputs && embeddings && offsets && outputs, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(!calc_grad_inputs || dy_dx, ENVIDR_E_BADARG, "dy_dx required when calc_grad_inputs");
    if (B == 0 || L == 0) return 0;
    const EncMode m{1, 0, 0};
    cudaStream_t st = as_stream(stream);
    float* jac = calc_grad_inputs ? dy_dx : nullptr;
#define CALL(D_, C_) launch_fwd<D_, C_>(m, inputs, embeddings, offsets, outputs, B, L, S, H, jac, st)
    switch (D) {
        case 2: DISPATCH_C(2, CALL)
        case 3: DISPATCH_C(3, CALL)
        default: set_error("HashEncoding: D must be 2 or 3."); return ENVIDR_E_UNSUPPORTED;
    }
#undef CALL
}

int envidr_hash_encode_backward(const float* grad, const float* inputs, const float* embeddings, const int32_t* offsets,
                                float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                int calc_grad_inputs, const float* dy_dx, float* grad_inputs, envidr_stream_t stream) {
    (void)embeddings;
    ENVIDR_REQUIRE(grad && inputs && offsets && grad_embeddings, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(!calc_grad_inputs || (dy_dx && grad_inputs), ENVIDR_E_BADARG, "dy_dx / grad_inputs required when calc_grad_inputs");
    if (B == 0 || L == 0) return 0;
    const EncMode m{1, 0, 0};
    cudaStream_t st = as_stream(stream);
    const float* jac = calc_grad_inputs ? dy_dx : nullptr;
#define CALL(D_, C_) launch_bwd<D_, C_>(m, grad, inputs, offsets, grad_embeddings, B, L, S, H, jac, grad_inputs, st)
    switch (D) {
        case 2: DISPATCH_C(2, CALL)
        case 3: DISPATCH_C(3, CALL)
        default: set_error("HashEncoding: D must be 2 or 3."); return ENVIDR_E_UNSUPPORTED;
    }
#undef CALL
}

int envidr_hash_encode_second_backward(const float* grad, const float* inputs, const float* embeddings, const int32_t* offsets,
                                       uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                                       const float* dy_dx, const float* grad_grad_inputs, float* grad_grad, float* grad2_embeddings,
                                       envidr_stream_t stream) {
    (void)embeddings; (void)calc_grad_inputs;
    ENVIDR_REQUIRE(grad && inputs && offsets && dy_dx && grad_grad_inputs && grad_grad && grad2_embeddings, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(C != 1, ENVIDR_E_UNSUPPORTED, "second backward: C must be 2, 4, or 8 (as in the reference)");
    if (B == 0 || L == 0) return 0;
    const EncMode m{1, 0, 0};
    cudaStream_t st = as_stream(stream);
#define CALL(D_, C_) launch_second<D_, C_>(m, grad, inputs, offsets, grad_grad_inputs, dy_dx, grad_grad, grad2_embeddings, B, L, S, H, st)
    switch (D) {
        case 2: DISPATCH_C(2, CALL)
        case 3: DISPATCH_C(3, CALL)
        default: set_error("HashEncoding: D must be 2 or 3."); return ENVIDR_E_UNSUPPORTED;
    }
#undef CALL
}

int envidr_grid_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs, uint32_t B,
                               uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, float* dy_dx, uint32_t gridtype,
                               int align_corners, envidr_stream_t stream) {
    ENVIDR_REQUIRE(inputs && embeddings && offsets && outputs, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(gridtype <= 1, ENVIDR_E_UNSUPPORTED, "gridtype must be 0 (hash) or 1 (tiled)");
    if (B == 0 || L == 0) return 0;
    const EncMode m{0, gridtype, align_corners ? 1u : 0u};
    cudaStream_t st = as_stream(stream);
#define CALL(D_, C_) launch_fwd<D_, C_>(m, inputs, embeddings, offsets, outputs, B, L, S, H, dy_dx, st)
    switch (D) {
        case 1: DISPATCH_C(1, CALL)
        case 2: DISPATCH_C(2, CALL)
        case 3: DISPATCH_C(3, CALL)
        case 4: DISPATCH_C(4, CALL)
        case 5: DISPATCH_C(5, CALL)
        default: set_error("GridEncoding: D must be 1, 2, 3, 4, or 5."); return ENVIDR_E_UNSUPPORTED;
    }
#undef CALL
}

int envidr_grid_encode_backward(const float* grad, const float* inputs, const float* embeddings, const int32_t* offsets,
                                float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                const float* dy_dx, float* grad_inputs, uint32_t gridtype, int align_corners, envidr_stream_t stream) {
    (void)embeddings;
    ENVIDR_REQUIRE(grad && inputs && offsets && grad_embeddings, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(gridtype <= 1, ENVIDR_E_UNSUPPORTED, "gridtype must be 0 (hash) or 1 (tiled)");
    if (B == 0 || L == 0) return 0;
    const EncMode m{0, gridtype, align_corners ? 1u : 0u};
    cudaStream_t st = as_stream(stream);
#define CALL(D_, C_) launch_bwd<D_, C_>(m, grad, inputs, offsets, grad_embeddings, B, L, S, H, dy_dx, grad_inputs, st)
    switch (D) {
        case 1: DISPATCH_C(1, CALL)
        case 2: DISPATCH_C(2, CALL)
        case 3: DISPATCH_C(3, CALL)
        case 4: DISPATCH_C(4, CALL)
        case 5: DISPATCH_C(5, CALL)
        default: set_error("GridEncoding: D must be 1, 2, 3, 4, or 5."); return ENVIDR_E_UNSUPPORTED;
    }
#undef CALL
}

}  // extern "C"
