// ide_tables.cuh -- Integrated Directional Encoding (Ref-NeRF eq. 6-8) tables + per-direction evaluation.
//
// Restates IntegratedDirEncoder (reference: ide_encoder/ide_encoder.py:5-130).  The coefficient matrix
// `mat`, the (m,l) list and the attenuation exponents sigma_l = l(l+1)/2 are built on the host in double
// precision with the reference's formulas and rounded to fp32, exactly like the reference's numpy code.
//
// Evaluation: the reference evaluates the z-polynomial of every (l, m) pair in the power basis
// (vmz @ mat, fp32).  For the l = 16 band (deg_view 5) that sum cancels catastrophically (|coeff| ~ 1e5):
// the reference's own fp32 result is up to ~5e-4 away from the exact value, and two fp32 evaluations of
// the same formula (torch CPU / torch CUDA / this GPU) disagree by as much -- measured on B200 in round 1.
// The kernel therefore evaluates the SAME polynomials (mat[:, i] are the power-basis coefficients of
// K_l^m * Q_l^m(z), Q the phase-carrying associated Legendre polynomial without its sin^m factor) with the
// fully normalised three-term recurrence in l, which is stable in fp32 (~1e-6 relative).  The result is the
// exact encoding; its distance to the reference is the reference's own cancellation error.
#pragma once
#include <math.h>
#include <stdint.h>

namespace envidr {

constexpr int kIdeMaxP = 36;    // deg_view 5: 2^5 - 1 + 5
constexpr int kIdeMaxL = 16;

struct IdeTables {
    uint32_t P, l_max, m_max, deg;
    float mat[kIdeMaxL + 1][kIdeMaxP];   // mat[k][i]: coefficient of z^k for pair i (0 for k > l - m)
    float sigma[kIdeMaxP];
    int32_t m[kIdeMaxP], l[kIdeMaxP];
    // normalised recurrence  q_l^m = ra[l][m] * z * q_{l-1}^m - rb[l][m] * q_{l-2}^m,  q_m^m = qmm[m]
    float qmm[kIdeMaxL + 1];
    float ra[kIdeMaxL + 1][kIdeMaxL + 1], rb[kIdeMaxL + 1][kIdeMaxL + 1];
    float band_sigma[5];
    int32_t band_base[5];
};

inline double ide_factorial(int n) { double r = 1; for (int i = 2; i <= n; i++) r *= i; return r; }

// generalized binomial coefficient prod_{j<k}(a - j) / k!   (ide_encoder.py:5-8)
inline double ide_gen_binom(double a, int k) {
    double p = 1;
    for (int j = 0; j < k; j++) p *= (a - j);
    return p / ide_factorial(k);
}

// ide_encoder.py:10-42
inline double ide_sph_harm_coeff(int l, int m, int k) {
    const double legendre = ((m & 1) ? -1.0 : 1.0) * pow(2.0, l) * ide_factorial(l) / ide_factorial(k) / ide_factorial(l - k - m) *
                            ide_gen_binom(0.5 * (l + k + m - 1.0), l);
    return sqrt((2.0 * l + 1.0) * ide_factorial(l - m) / (4.0 * M_PI * ide_factorial(l + m))) * legendre;
}

inline bool ide_build_tables(int deg_view, IdeTables* t) {
    if (deg_view < 1 || deg_view > 5) return false;
    *t = IdeTables{};
    t->deg = (uint32_t)deg_view;
    t->l_max = 1u << (deg_view - 1);
    uint32_t i = 0;
    for (int d = 0; d < deg_view; d++) {
        const int l = 1 << d;
        for (int m = 0; m <= l; m++, i++) {
            t->m[i] = m; t->l[i] = l;
            t->sigma[i] = (float)(0.5 * l * (l + 1));
            for (int k = 0; k <= l - m; k++) t->mat[k][i] = (float)ide_sph_harm_coeff(l, m, k);
        }
    }
    t->P = i;
    t->m_max = t->l_max;
    for (int d = 0, base = 0; d < deg_view; d++) {
        const int l = 1 << d;
        t->band_base[d] = base;
        t->band_sigma[d] = (float)(0.5 * l * (l + 1));
        base += l + 1;
    }
    for (int m = 0; m <= (int)t->l_max; m++) {
        double dfact = 1.0;                                   // (2m-1)!!
        for (int k = 2 * m - 1; k > 1; k -= 2) dfact *= k;
        t->qmm[m] = (float)(sqrt((2.0 * m + 1.0) / (4.0 * M_PI) / ide_factorial(2 * m)) * ((m & 1) ? -1.0 : 1.0) * dfact);
        for (int l = m + 1; l <= (int)t->l_max; l++) {
            t->ra[l][m] = (float)sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
            t->rb[l][m] = (l - m - 1 > 0)
                ? (float)sqrt((2.0 * l + 1.0) * (l + m - 1.0) * (l - m - 1.0) / ((2.0 * l - 3.0) * (l - m) * (double)(l + m))) : 0.0f;
        }
    }
    return true;
}

#ifdef __CUDACC__
// emit(i, re_i, im_i), i < P, with [Re | Im]_i = (x+iy)^m * K_l^m Q_l^m(z) * exp(-sigma_l * kappa_inv) * scale.
// Only the orders m = m_start, m_start + m_step, ... are visited (lets several threads share one direction).
template <class Emit>
__device__ __forceinline__ void ide_eval_emit(const IdeTables& T, float x, float y, float z, float kappa_inv, float scale, Emit&& emit,
                                              int m_start = 0, int m_step = 1) {
    if (x == 0.0f && y == 0.0f) y += 1.0f;         // "avoid 0 + 0j exponentiation" (ide_encoder.py:113-115)
    float att0 = 0.f, att1 = 0.f, att2 = 0.f, att3 = 0.f, att4 = 0.f;
    const int deg = (int)T.deg;
    if (deg > 0) att0 = expf(-T.band_sigma[0] * kappa_inv) * scale;
    if (deg > 1) att1 = expf(-T.band_sigma[1] * kappa_inv) * scale;
    if (deg > 2) att2 = expf(-T.band_sigma[2] * kappa_inv) * scale;
    if (deg > 3) att3 = expf(-T.band_sigma[3] * kappa_inv) * scale;
    if (deg > 4) att4 = expf(-T.band_sigma[4] * kappa_inv) * scale;
    auto att_of = [&](int b) { return b == 0 ? att0 : (b == 1 ? att1 : (b == 2 ? att2 : (b == 3 ? att3 : att4))); };
    const int l_max = (int)T.l_max;
    // (x+iy)^m_start and the per-iteration multiplier (x+iy)^m_step
    float re = 1.0f, im = 0.0f, sr = 1.0f, si = 0.0f;
    for (int k = 0; k < m_start; k++) { const float nr = re * x - im * y; im = re * y + im * x; re = nr; }
    for (int k = 0; k < m_step; k++) { const float nr = sr * x - si * y; si = sr * y + si * x; sr = nr; }
    for (int m = m_start; m <= l_max; m += m_step) {
        if (m > m_start) {
            const float nr = re * sr - im * si;
            im = re * si + im * sr;
            re = nr;
        }
        float p2 = 0.0f, p1 = T.qmm[m];
        if (m > 0 && (m & (m - 1)) == 0) {            // l == m is itself a band (m = 1, 2, 4, 8, 16)
            const int b = 31 - __clz(m);
            const float a = att_of(b);
            emit(T.band_base[b] + m, re * p1 * a, im * p1 * a);
        }
        for (int l = m + 1; l <= l_max; l++) {
            const float q = T.ra[l][m] * z * p1 - T.rb[l][m] * p2;
            p2 = p1; p1 = q;
            if ((l & (l - 1)) == 0) {
                const int b = 31 - __clz(l);
                const float a = att_of(b);
                emit(T.band_base[b] + m, re * q * a, im * q * a);
            }
        }
    }
}
// Compile-time-unrolled variant for the tensor-core kernel: DEG fixes every loop bound, so after unrolling all table indices are
// immediates (ra / rb / qmm become constant-bank operands of the FMULs, no index arithmetic, no band tests at run time) and
// the feature index handed to `emit` is a constant.  Same recurrence, same operation order as ide_eval_emit.
__host__ __device__ constexpr bool ide_is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
__host__ __device__ constexpr int ide_ilog2(int v) { return v <= 1 ? 0 : 1 + ide_ilog2(v >> 1); }
__host__ __device__ constexpr int ide_band_base(int b) { return (1 << b) - 1 + b; }

// NB <= DEG: only the first NB bands (l = 1 .. 2^(NB-1)) are evaluated; the caller guarantees the others are negligible
// (exp(-sigma_l kappa) below its threshold) and keeps their outputs at zero.
template <int DEG, int MSTART, int MSTEP, int NB = DEG, class Emit>
__device__ __forceinline__ void ide_eval_emit_static(const IdeTables& T, float x, float y, float z, float kappa_inv, float scale, Emit&& emit) {
    static_assert(NB >= 1 && NB <= DEG, "band count");
    constexpr int LMAX = 1 << (NB - 1);
    if (x == 0.0f && y == 0.0f) y += 1.0f;
    float att[NB];
    #pragma unroll
    for (int b = 0; b < NB; b++) att[b] = expf(-T.band_sigma[b] * kappa_inv) * scale;
    float re = 1.0f, im = 0.0f, sr = 1.0f, si = 0.0f;
    #pragma unroll
    for (int k = 0; k < MSTART; k++) { const float nr = re * x - im * y; im = re * y + im * x; re = nr; }
    #pragma unroll
    for (int k = 0; k < MSTEP; k++) { const float nr = sr * x - si * y; si = sr * y + si * x; sr = nr; }
    #pragma unroll
    for (int m = MSTART; m <= LMAX; m += MSTEP) {
        if (m > MSTART) {
            const float nr = re * sr - im * si;
            im = re * si + im * sr;
            re = nr;
        }
        float p2 = 0.0f, p1 = T.qmm[m];
        if (ide_is_pow2(m)) {
            const float a = att[ide_ilog2(m)];
            emit(ide_band_base(ide_ilog2(m)) + m, re * p1 * a, im * p1 * a);
        }
        #pragma unroll
        for (int l = m + 1; l <= LMAX; l++) {
            const float q = T.ra[l][m] * z * p1 - T.rb[l][m] * p2;
            p2 = p1; p1 = q;
            if (ide_is_pow2(l)) {
                const float a = att[ide_ilog2(l)];
                emit(ide_band_base(ide_ilog2(l)) + m, re * q * a, im * q * a);
            }
        }
    }
}
// out_re[i*sr], out_im[i*si], i < P
__device__ __forceinline__ void ide_eval(const IdeTables& T, float x, float y, float z, float kappa_inv, float scale,
                                         float* out_re, int sr, float* out_im, int si) {
    ide_eval_emit(T, x, y, z, kappa_inv, scale, [&](int i, float re, float im) { out_re[i * sr] = re; out_im[i * si] = im; });
}
#endif

}  // namespace envidr
