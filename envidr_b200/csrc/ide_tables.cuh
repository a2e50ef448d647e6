// ide_tables.cuh -- Integrated Directional Encoding (Ref-NeRF eq. 6-8) tables + per-direction evaluation.
//
// Restates IntegratedDirEncoder (reference: ide_encoder/ide_encoder.py:5-130).  The coefficient matrix
// `mat`, the (m,l) list and the attenuation exponents sigma_l = l(l+1)/2 are built on the host in double
// precision with the reference's formulas and rounded to fp32, exactly like the reference's numpy code.
//
// Evaluation mirrors the reference's fp32 arithmetic on purpose: z^k through powf, the z-polynomial as a
// sequential fp32 dot product over the power basis.  For the l = 16 band (deg_view 5) that polynomial
// cancels catastrophically (|coeff| ~ 1e5), so the reference's own result is ~5e-4 away from the exact
// value; matching its arithmetic keeps us within ~1e-4 of what the reference produces.
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
    return true;
}

#ifdef __CUDACC__
// out_re[i*sr], out_im[i*si], i < P:  [Re | Im] of  (x+iy)^m * (sum_k z^k mat[k][i]) * exp(-sigma_i * kappa_inv) * scale
__device__ __forceinline__ void ide_eval(const IdeTables& T, float x, float y, float z, float kappa_inv, float scale,
                                         float* out_re, int sr, float* out_im, int si) {
    if (x == 0.0f && y == 0.0f) y += 1.0f;         // "avoid 0 + 0j exponentiation" (ide_encoder.py:113-115)
    float zp[kIdeMaxL + 1], re[kIdeMaxL + 1], im[kIdeMaxL + 1];
    zp[0] = 1.0f; re[0] = 1.0f; im[0] = 0.0f;
    #pragma unroll
    for (int k = 1; k <= kIdeMaxL; k++) {
        if (k <= (int)T.l_max) {
            zp[k] = powf(z, (float)k);
            re[k] = re[k - 1] * x - im[k - 1] * y;
            im[k] = re[k - 1] * y + im[k - 1] * x;
        } else {
            zp[k] = 0.0f; re[k] = 0.0f; im[k] = 0.0f;
        }
    }
    // pairs are grouped by band l = 1, 2, 4, ...; within a band m = 0..l
    uint32_t i = 0;
    #pragma unroll
    for (int band = 0; band < 5; band++) {
        if (band >= (int)T.deg) break;
        const int l = 1 << band;
        const float att = expf(-T.sigma[i] * kappa_inv) * 1.0f;
        #pragma unroll
        for (int m = 0; m <= l; m++, i++) {
            float acc = 0.0f;
            #pragma unroll
            for (int k = 0; k <= l - m; k++) acc = fmaf(zp[k], T.mat[k][i], acc);
            out_re[i * sr] = re[m] * acc * att * scale;
            out_im[i * si] = im[m] * acc * att * scale;
        }
    }
}
#endif

}  // namespace envidr
