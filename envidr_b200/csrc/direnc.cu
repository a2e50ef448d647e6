// direnc.cu -- directional / positional encoders for sm_100a: frequency, spherical harmonics, IDE.
//
// Replaces the reference's `_freqencoder` (freqencoder/src/freqencoder.cu:30-94) and `_shencoder`
// (shencoder/src/shencoder.cu:27-383) extensions and gives the pure-PyTorch IntegratedDirEncoder
// (ide_encoder/ide_encoder.py:57-130) a kernel.
#include <math.h>
#include <string.h>
#include "common.cuh"
#include "ide_tables.cuh"

namespace envidr {

// ------------------------------------------------------------------------------------------------
// frequency encoding: out[b] = [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] in D-wide blocks.
// One thread per output element (coalesced stores); cos is sin(. + pi/2) with the fast intrinsic,
// exactly as the reference computes it.
// ------------------------------------------------------------------------------------------------
// kIdx = uint32_t whenever B * C fits (always, in practice): the element -> (sample, column) split is then a 32-bit division
// instead of a 64-bit one, which was most of the kernel's instructions.
template <typename kIdx>
__global__ void __launch_bounds__(256) k_freq_fwd(const float* __restrict__ inputs, uint32_t B, uint32_t D, uint32_t C,
                                                 float* __restrict__ outputs) {
    const kIdx t = (kIdx)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (kIdx)B * C) return;
    const uint32_t b = (uint32_t)(t / C), c = (uint32_t)(t - (kIdx)b * C);
    const float* x = inputs + (size_t)b * D;
    float v;
    if (c < D) {
        v = x[c];
    } else {
        const uint32_t col = c / D - 1, d = c % D, freq = col / 2;
        const float phase_shift = (col % 2) * (3.141592653589793f / 2);
        // scalbnf(x, f) of the reference == x * 2^f exactly (a power-of-two product rounds nowhere short of overflow); the exponent-field
        // constant saves scalbnf's special-case code, which was half of the kernel's instructions
        v = __sinf(x[d] * __uint_as_float((127u + freq) << 23) + phase_shift);
    }
    outputs[t] = v;
}

__global__ void __launch_bounds__(256) k_freq_bwd(const float* __restrict__ grad, const float* __restrict__ outputs, uint32_t B,
                                                 uint32_t D, uint32_t deg, uint32_t C, float* __restrict__ grad_inputs) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * D) return;
    const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (size_t)b * D);
    const float* g = grad + (size_t)b * C;
    const float* o = outputs + (size_t)b * C;
    float result = g[d];
    g += D; o += D;
    for (uint32_t f = 0; f < deg; f++) {
        result += scalbnf(1.0f, f) * (g[d] * o[D + d] - g[D + d] * o[d]);
        g += 2 * D; o += 2 * D;
    }
    grad_inputs[t] = result;
}

// ------------------------------------------------------------------------------------------------
// real spherical harmonics, degree <= 8, evaluated by recurrence instead of the reference's table of
// 64 hard-coded polynomials (same polynomials in x,y,z; also for non-unit inputs):
//   Y[l*l+l+m] = N_l^m Q_l^m(z) Re (x+iy)^m,  Y[l*l+l-m] = N_l^m Q_l^m(z) Im (x+iy)^m,
//   Q_m^m = (-1)^m (2m-1)!!,  Q_{m+1}^m = (2m+1) z Q_m^m,  (l-m) Q_l^m = (2l-1) z Q_{l-1}^m - (l+m-1) Q_{l-2}^m.
// ------------------------------------------------------------------------------------------------
__constant__ float c_sh_norm[8 * 8];   // N_l^m (incl. sqrt(2) for m > 0), [l][m]

template <bool kGrad>
__global__ void __launch_bounds__(128) k_sh_fwd(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B, uint32_t D,
                                               uint32_t deg, float* __restrict__ dy_dx) {
    // outputs of a thread (its sample's row of deg^2 values) are staged in shared memory (row stride deg^2 + 1: conflict-free)
    // and written by the whole block as one contiguous, coalesced run; the per-thread row stores of the first version were one
    // 32-byte sector write per value.
    __shared__ float s_out[128 * 65];
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t C2 = deg * deg;
    const uint32_t bc = min(b, B - 1);                                   // threads past the end compute a duplicate, store nothing
    const float x = inputs[(size_t)bc * D], y = inputs[(size_t)bc * D + 1], z = inputs[(size_t)bc * D + 2];
    float* out = s_out + threadIdx.x * (C2 + 1);
    float* gx = kGrad ? dy_dx + (size_t)bc * 3 * C2 : nullptr;
    float* gy = kGrad ? gx + C2 : nullptr;
    float* gz = kGrad ? gy + C2 : nullptr;
    float A = 1.0f, Bi = 0.0f;           // Re / Im (x+iy)^m
    float Ap = 0.0f, Bp = 0.0f;          // previous power (m-1)
    float qmm = 1.0f;
    for (uint32_t m = 0; m < deg; m++) {
        if (m > 0) {
            Ap = A; Bp = Bi;
            const float An = x * A - y * Bi;
            Bi = x * Bi + y * A;
            A = An;
            qmm *= -(2.0f * m - 1.0f);
        }
        // Q_l^m and Q_l^{m+1} (for d/dz) by upward recurrence in l
        float q_prev2 = 0.0f, q_prev = 0.0f, q = 0.0f;            // order m
        float r_prev2 = 0.0f, r_prev = 0.0f, r = 0.0f;            // order m+1
        const float rmm = qmm * -(2.0f * m + 1.0f);
        for (uint32_t l = m; l < deg; l++) {
            if (l == m) q = qmm;
            else if (l == m + 1) q = (2.0f * m + 1.0f) * z * qmm;
            else q = ((2.0f * l - 1.0f) * z * q_prev - (float)(l + m - 1) * q_prev2) / (float)(l - m);
            if (kGrad) {
                if (l < m + 1) r = 0.0f;
                else if (l == m + 1) r = rmm;
                else if (l == m + 2) r = (2.0f * m + 3.0f) * z * rmm;
                else r = ((2.0f * l - 1.0f) * z * r_prev - (float)(l + m) * r_prev2) / (float)(l - m - 1);
            }
            const float Nlm = c_sh_norm[l * 8 + m];
            const float nq = Nlm * q;
            const uint32_t ip = l * l + l + m, in = l * l + l - m;
            out[ip] = nq * A;
            if (m) out[in] = nq * Bi;
            if (kGrad && b < B) {
                const float ndq = -Nlm * r;                          // d/dz Q_l^m = -Q_l^{m+1}
                gx[ip] = m ? nq * m * Ap : 0.0f;
                gy[ip] = m ? -nq * m * Bp : 0.0f;
                gz[ip] = ndq * A;
                if (m) {
                    gx[in] = nq * m * Bp;
                    gy[in] = nq * m * Ap;
                    gz[in] = ndq * Bi;
                }
            }
            q_prev2 = q_prev; q_prev = q;
            r_prev2 = r_prev; r_prev = r;
        }
    }
    __syncthreads();
    const uint32_t b0 = blockIdx.x * blockDim.x;
    const uint32_t rows = min((uint32_t)blockDim.x, B - b0);
    float* dst = outputs + (size_t)b0 * C2;
    for (uint32_t i = threadIdx.x; i < rows * C2; i += blockDim.x) dst[i] = s_out[(i / C2) * (C2 + 1) + i % C2];
}

// grad_inputs[b,d] += sum_k grad[b,k] * dy_dx[b,d,k]   (shencoder.cu:358-383)
__global__ void __launch_bounds__(256) k_sh_bwd(const float* __restrict__ grad, const float* __restrict__ dy_dx, uint32_t B, uint32_t D,
                                               uint32_t C2, float* __restrict__ grad_inputs) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * D) return;
    const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (size_t)b * D);
    const float* g = grad + (size_t)b * C2;
    const float* j = dy_dx + (size_t)b * D * C2 + (size_t)d * C2;
    float r = 0;
    for (uint32_t k = 0; k < C2; k++) r += g[k] * j[k];
    grad_inputs[t] += r;
}

static void upload_sh_norm() {
    static bool done = false;
    if (done) return;
    float h[64];
    memset(h, 0, sizeof(h));
    for (int l = 0; l < 8; l++)
        for (int m = 0; m <= l; m++) {
            double fac = 1.0;
            for (int k = l - m + 1; k <= l + m; k++) fac /= (double)k;
            h[l * 8 + m] = (float)(sqrt((2.0 * l + 1.0) / (4.0 * M_PI) * fac) * (m ? sqrt(2.0) : 1.0));
        }
    cudaMemcpyToSymbol(c_sh_norm, h, sizeof(h));
    done = true;
}

// ------------------------------------------------------------------------------------------------
// IDE.  One thread per direction; results staged in shared memory so the [B, 2P] rows leave coalesced.
// ------------------------------------------------------------------------------------------------
__constant__ IdeTables c_ide;      // tables of the degree last uploaded (see ide_tables.cuh)
static int g_ide_deg_uploaded = 0;

int ensure_ide_tables(uint32_t deg_view) {
    if ((int)deg_view == g_ide_deg_uploaded) return 0;
    IdeTables t;
    if (!ide_build_tables((int)deg_view, &t)) return ENVIDR_E_UNSUPPORTED;
    cudaError_t e = cudaMemcpyToSymbol(c_ide, &t, sizeof(t));
    if (e != cudaSuccess) { set_error("ide tables: %s", cudaGetErrorString(e)); return (int)e; }
    g_ide_deg_uploaded = (int)deg_view;
    return 0;
}

__global__ void __launch_bounds__(128) k_ide_fwd(const float* __restrict__ dirs, const float* __restrict__ kappa_arr, float kappa_scalar,
                                                uint32_t B, float scale, float* __restrict__ out) {
    extern __shared__ float s_out[];                 // [128][2P]
    const uint32_t P = c_ide.P;
    const uint32_t b = blockIdx.x * 128 + threadIdx.x;
    if (b < B) {
        const float kap = kappa_arr ? kappa_arr[b] : kappa_scalar;
        ide_eval(c_ide, dirs[3 * (size_t)b], dirs[3 * (size_t)b + 1], dirs[3 * (size_t)b + 2], kap, scale,
                 s_out + (size_t)threadIdx.x * 2 * P, 1, s_out + (size_t)threadIdx.x * 2 * P + P, 1);
    }
    __syncthreads();
    const uint32_t rows = min(128u, B - blockIdx.x * 128);
    float* dst = out + (size_t)blockIdx.x * 128 * 2 * P;
    for (uint32_t k = threadIdx.x; k < rows * 2 * P; k += 128) dst[k] = s_out[k];
}

// Backward of the encoding w.r.t. the direction and kappa_inv (the reference gets it from autograd over ~20 torch ops,
// ide_encoder.py:98-130).  With P_m = (x+iy)^m, q = K_l^m Q_l^m(z), a_l = exp(-sigma_l kappa) * scale:
//   out_re = Re(P_m) q a,  out_im = Im(P_m) q a
//   d/dx: m Re(P_{m-1}) q a | m Im(P_{m-1}) q a      d/dy: -m Im(P_{m-1}) q a | m Re(P_{m-1}) q a
//   d/dz: q' from the differentiated recurrence q'_l = ra (q_{l-1} + z q'_{l-1}) - rb q'_{l-2}, q'_m = 0
//   d/dkappa: -sigma_l * out
// One thread per direction; the gradient row [2P] is read through shared memory (coalesced).
__global__ void __launch_bounds__(128) k_ide_bwd(const float* __restrict__ dirs, const float* __restrict__ kappa_arr, float kappa_scalar,
                                                uint32_t B, float scale, const float* __restrict__ grad, float* __restrict__ grad_dirs,
                                                float* __restrict__ grad_kappa) {
    extern __shared__ float s_g[];                   // [128][2P + 1] (padded: conflict-free row access)
    const IdeTables& T = c_ide;
    const uint32_t P = T.P, W = 2 * P + 1;
    const uint32_t rows = min(128u, B - blockIdx.x * 128);
    const float* src = grad + (size_t)blockIdx.x * 128 * 2 * P;
    for (uint32_t k = threadIdx.x; k < rows * 2 * P; k += 128) s_g[(k / (2 * P)) * W + k % (2 * P)] = src[k];
    __syncthreads();
    const uint32_t b = blockIdx.x * 128 + threadIdx.x;
    if (b >= B) return;
    const float* g = s_g + threadIdx.x * W;
    float x = dirs[3 * (size_t)b], y = dirs[3 * (size_t)b + 1];
    const float z = dirs[3 * (size_t)b + 2];
    const float kap = kappa_arr ? kappa_arr[b] : kappa_scalar;
    if (x == 0.0f && y == 0.0f) y += 1.0f;
    float att[5];
    #pragma unroll
    for (int d = 0; d < 5; d++) att[d] = d < (int)T.deg ? expf(-T.band_sigma[d] * kap) * scale : 0.f;
    const int l_max = (int)T.l_max;
    float gx = 0.f, gy = 0.f, gz = 0.f, gk = 0.f;
    float re = 1.0f, im = 0.0f, pre = 0.0f, pim = 0.0f;          // P_m and P_{m-1}
    for (int m = 0; m <= l_max; m++) {
        if (m > 0) { pre = re; pim = im; const float nr = re * x - im * y; im = re * y + im * x; re = nr; }
        float q2 = 0.0f, q1 = T.qmm[m], d2 = 0.0f, d1 = 0.0f;      // q_{l-2}, q_{l-1} and their z-derivatives
        for (int l = m; l <= l_max; l++) {
            float q = q1, dq = d1;
            if (l > m) {
                const float ra = T.ra[l][m], rb = T.rb[l][m];
                q = ra * z * q1 - rb * q2;
                dq = ra * (q1 + z * d1) - rb * d2;
                q2 = q1; q1 = q; d2 = d1; d1 = dq;
            }
            if (l > 0 && (l & (l - 1)) == 0) {
                const int band = 31 - __clz(l);
                const int i = T.band_base[band] + m;
                const float a = att[band], gr = g[i], gi = g[P + i];
                const float qa = q * a;
                gx += (float)m * (pre * gr + pim * gi) * qa;
                gy += (float)m * (pre * gi - pim * gr) * qa;
                gz += (re * gr + im * gi) * dq * a;
                gk -= T.band_sigma[band] * (re * gr + im * gi) * qa;
            }
        }
    }
    grad_dirs[3 * (size_t)b] = gx; grad_dirs[3 * (size_t)b + 1] = gy; grad_dirs[3 * (size_t)b + 2] = gz;
    if (grad_kappa) grad_kappa[b] = gk;
}

}  // namespace envidr

using namespace envidr;

extern "C" {

int envidr_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs, envidr_stream_t stream) {
    ENVIDR_REQUIRE(inputs && outputs, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(C == D + 2 * D * deg, ENVIDR_E_BADARG, "C must equal D + 2*D*deg");
    if (B == 0) return 0;
    const size_t total = (size_t)B * C;
    if (total + 256 < 0xffffffffull)
        k_freq_fwd<uint32_t><<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(inputs, B, D, C, outputs);
    else
        k_freq_fwd<size_t><<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(inputs, B, D, C, outputs);
    return check_launch("freq_encode_forward");
}

int envidr_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                float* grad_inputs, envidr_stream_t stream) {
    ENVIDR_REQUIRE(grad && outputs && grad_inputs, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(C == D + 2 * D * deg, ENVIDR_E_BADARG, "C must equal D + 2*D*deg");
    if (B == 0) return 0;
    const size_t total = (size_t)B * D;
    k_freq_bwd<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(grad, outputs, B, D, deg, C, grad_inputs);
    return check_launch("freq_encode_backward");
}

int envidr_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree, float* dy_dx,
                             envidr_stream_t stream) {
    ENVIDR_REQUIRE(inputs && outputs, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(D == 3, ENVIDR_E_UNSUPPORTED, "SH encoder: D must be 3");
    ENVIDR_REQUIRE(degree >= 1 && degree <= 8, ENVIDR_E_UNSUPPORTED, "SH encoder: degree must be 1..8");
    if (B == 0) return 0;
    upload_sh_norm();
    if (dy_dx) k_sh_fwd<true><<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(inputs, outputs, B, D, degree, dy_dx);
    else       k_sh_fwd<false><<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(inputs, outputs, B, D, degree, nullptr);
    return check_launch("sh_encode_forward");
}

int envidr_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree, const float* dy_dx,
                              float* grad_inputs, envidr_stream_t stream) {
    (void)inputs;
    ENVIDR_REQUIRE(grad && dy_dx && grad_inputs, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(D == 3, ENVIDR_E_UNSUPPORTED, "SH encoder: D must be 3");
    if (B == 0) return 0;
    const size_t total = (size_t)B * D;
    k_sh_bwd<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(grad, dy_dx, B, D, degree * degree, grad_inputs);
    return check_launch("sh_encode_backward");
}

int envidr_ide_encode_forward(const float* dirs, const float* kappa_inv_arr, float kappa_inv_scalar, uint32_t B, uint32_t deg_view,
                              float scale, float* out, envidr_stream_t stream) {
    ENVIDR_REQUIRE(dirs && out, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(deg_view >= 1 && deg_view <= 5, ENVIDR_E_UNSUPPORTED, "Only deg_view of at most 5 is numerically stable.");
    if (B == 0) return 0;
    int rc = ensure_ide_tables(deg_view);
    if (rc) return rc;
    const uint32_t P = (1u << deg_view) - 1 + deg_view;
    const size_t smem = (size_t)128 * 2 * P * sizeof(float);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_ide_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 2 * 36 * 4); attr = true; }
    k_ide_fwd<<<ceil_div(B, 128), 128, smem, as_stream(stream)>>>(dirs, kappa_inv_arr, kappa_inv_scalar, B, scale, out);
    return check_launch("ide_encode_forward");
}

int envidr_ide_encode_backward(const float* dirs, const float* kappa_inv_arr, float kappa_inv_scalar, uint32_t B, uint32_t deg_view,
                               float scale, const float* grad, float* grad_dirs, float* grad_kappa, envidr_stream_t stream) {
    ENVIDR_REQUIRE(dirs && grad && grad_dirs, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(deg_view >= 1 && deg_view <= 5, ENVIDR_E_UNSUPPORTED, "Only deg_view of at most 5 is numerically stable.");
    if (B == 0) return 0;
    int rc = ensure_ide_tables(deg_view);
    if (rc) return rc;
    const uint32_t P = (1u << deg_view) - 1 + deg_view;
    const size_t smem = (size_t)128 * (2 * P + 1) * sizeof(float);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_ide_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * (2 * 36 + 1) * 4); attr = true; }
    k_ide_bwd<<<ceil_div(B, 128), 128, smem, as_stream(stream)>>>(dirs, kappa_inv_arr, kappa_inv_scalar, B, scale, grad, grad_dirs, grad_kappa);
    return check_launch("ide_encode_backward");
}

// host-only helper (no GPU needed): the IDE coefficient tables, for tests against ide_encoder.py:84-96
int envidr_ide_tables(uint32_t deg_view, float* mat /*[(l_max+1)*P]*/, float* sigma /*[P]*/, int32_t* ml /*[2*P]*/) {
    IdeTables t;
    if (!ide_build_tables((int)deg_view, &t)) { set_error("Only deg_view of at most 5 is numerically stable."); return ENVIDR_E_UNSUPPORTED; }
    for (uint32_t k = 0; k <= t.l_max; k++)
        for (uint32_t i = 0; i < t.P; i++) mat[k * t.P + i] = t.mat[k][i];
    for (uint32_t i = 0; i < t.P; i++) { sigma[i] = t.sigma[i]; ml[i] = t.m[i]; ml[t.P + i] = t.l[i]; }
    return 0;
}

}  // extern "C"
