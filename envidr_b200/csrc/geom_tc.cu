// geom_tc.cu -- geometry phase of the per-sample field on tensor cores (sm_100a):
//   hash-grid gather (+ d/dx)  ->  sdf_net forward  ->  reverse pass for d sdf / d enc  ->  normal, Laplace density,
//   unit-norm geo feature, roughness, blend, reflected direction ... (reference: hashencoder.cu:103-254,
//   nerf/network.py:381-522, nerf/renderer.py:20-39,147-198).
//
// Same arithmetic contract as field_tc.cu: every dense operand is split into two fp16 values (hi + lo) and each
// 16-wide K step issues three tcgen05.mma (hi*hi + lo*hi + hi*lo) with fp32 accumulation in tensor memory, which
// keeps the SDF to ~2e-7 (the density amplifies SDF error by 1/(2 beta^2)).
//
// One persistent CTA per SM; all sdf_net weight images (forward and transposed, ~52 KB) stay resident in shared
// memory.  Per 128-sample tile: 8 worker warps gather the grid (thread = (sample, level parity)), write the encoding
// as the first A operand, then alternate with the MMA-issuing warp through n forward and n-1 reverse stages
// (tcgen05.ld -> bias / ReLU / ReLU-mask -> next A operand); the ReLU masks live in registers.
#include <math.h>
#include "gridenc.cuh"
#include "tc_common.cuh"
#include "field_tc.cuh"

namespace envidr {

constexpr int kGThreads = 384;                     // 4 control warps + 8 worker warps
constexpr uint32_t kGOperand = 32768;              // one A-operand buffer: 128 rows x 64 K x 2 B x (hi, lo)
constexpr uint32_t kGOperandHalf = 16384;
constexpr int kGJacLd = 100;                       // floats per row of the jacobian tile (96 + pad)

struct GeomOutDev { float *sigma, *normal, *sdf, *roughness, *grad_x; };

__device__ __forceinline__ float g_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float g_softplus(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

__device__ __forceinline__ void g_split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kGThreads, 1)
k_geom_tc(const TcGeom G, const float* __restrict__ xyzs, const float* __restrict__ dirs, const uint32_t* __restrict__ M_dev,
          uint32_t M_host, int mode, float* __restrict__ rec, const GeomOutDev O) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem;                                           // resident weight images + float region
    uint8_t* s_op = smem + G.res_bytes_al;                         // two operand buffers
    float* s_jac = reinterpret_cast<float*>(s_op + 2 * kGOperand); // [128][100]
    float* s_side = s_jac + 128 * kGJacLd;                         // [16][128] last-layer outputs
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_side + 16 * 128);
    uint64_t* w_full = bars;          // resident weights landed
    uint64_t* acc_ready = bars + 1;   // issuer -> workers
    uint64_t* a_ready = bars + 2;     // workers -> issuer (256 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t M = M_dev ? *M_dev : M_host;
    const uint32_t n_tiles = (M + 127) / 128;
    if (blockIdx.x >= n_tiles) return;             // nothing to do for this CTA (tail iterations of the render loop)
    const int n = (int)G.n_layers;                 // forward layers; n - 1 reverse stages
    const int n_stages = 2 * n - 1;
    const float* s_f = reinterpret_cast<const float*>(s_w + G.float_off);     // biases [n][64], then w_row0[64]

    if (tid == 0) {
        tc::mbar_init(w_full, 1);
        tc::mbar_init(acc_ready, 1);
        tc::mbar_init(a_ready, 256);
        tc::mbar_fence_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0 && lane == 0) {                  // one-time: pull the resident region in with a few bulk copies
        tc::mbar_arrive_expect_tx(w_full, G.res_bytes);
        for (uint32_t o = 0; o < G.res_bytes; o += 16384) {
            const uint32_t b = min(16384u, G.res_bytes - o);
            tc::bulk_g2s(s_w + o, G.blob + o, b, w_full);
        }
    }
    tc::mbar_wait(w_full, 0);

    if (warp == 0) {
        // ===================== MMA issuer =====================
        uint32_t a_par = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int st = 0; st < n_stages; st++) {
                const TcImg& I = (st < n) ? G.F[st] : G.R[n_stages - 1 - st];      // reverse stages use layers n-2 .. 0
                tc::mbar_wait(a_ready, a_par); a_par ^= 1;
                tc::tc_fence_after();
                if (lane == 0) {
                    const uint32_t idesc = tc::make_idesc_f16(128, I.Np);
                    const uint32_t a_hi0 = tc::smem_u32(s_op + (st & 1) * kGOperand), a_lo0 = a_hi0 + kGOperandHalf;
                    const uint32_t b0 = tc::smem_u32(s_w + I.off);
                    for (uint32_t s = 0; s < I.Kp / 16; s++) {
                        const uint32_t b_hi = b0 + s * I.Np * 64, b_lo = b_hi + I.Np * 32;
                        const uint64_t da_hi = tc::make_smem_desc(a_hi0 + s * 4096, 2048, 128), da_lo = tc::make_smem_desc(a_lo0 + s * 4096, 2048, 128);
                        const uint64_t db_hi = tc::make_smem_desc(b_hi, I.Np * 16, 128), db_lo = tc::make_smem_desc(b_lo, I.Np * 16, 128);
                        tc::mma_f16_ss(tmem, da_hi, db_hi, idesc, s > 0);
                        tc::mma_f16_ss(tmem, da_lo, db_hi, idesc, 1);
                        tc::mma_f16_ss(tmem, da_hi, db_lo, idesc, 1);
                    }
                    tc::mma_commit(acc_ready);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== workers =====================
        const uint32_t wt = tid - 128;                     // 0..255
        const uint32_t quarter = warp & 3, g = (warp - 4) >> 2;
        const uint32_t row = quarter * 32 + lane;          // accumulator row of this thread (2 threads per row: g = 0, 1)
        const uint32_t lane_addr = (quarter * 32u) << 16;
        const uint32_t Hd = G.F[0].N;                      // hidden width (32 or 64)
        const uint32_t hchunks = Hd / 32;
        uint32_t acc_par = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t m0 = tile * 128;
            // ---- gather: thread = (sample, level parity); encoding -> operand buffer 0 (fp16 hi/lo), jacobian -> smem fp32
            {
                const uint32_t s = wt & 127, par = wt >> 7;
                const uint32_t m = m0 + s;
                const bool valid = m < M;
                float x01[3] = {0.f, 0.f, 0.f};
                if (valid) {
                    #pragma unroll
                    for (int d = 0; d < 3; d++) x01[d] = (xyzs[3 * (size_t)m + d] + G.bound) / (2 * G.bound);
                }
                const EncMode em{1, 0, 0};
                for (uint32_t l = par; l < G.L; l += 2) {
                    float e0 = 0.f, e1 = 0.f, j[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    Cell<3> cell;
                    const bool lvl_on = !(G.enabled_levels > 0 && (int)l >= G.enabled_levels);
                    if (valid && lvl_on && cell.setup(em, x01, G.offsets, l, G.S, G.H)) {
                        const float* grid = G.table + (size_t)(uint32_t)G.offsets[l] * 2;
                        float rows[8][2];
                        #pragma unroll
                        for (uint32_t corner = 0; corner < 8; corner++) {
                            uint32_t pl[3];
                            #pragma unroll
                            for (int d = 0; d < 3; d++) pl[d] = cell.pg[d] + ((corner >> d) & 1u);
                            load_row<2>(grid + (size_t)cell_index<3>(em, cell.hashmap_size, cell.resolution, pl) * 2, rows[corner]);
                        }
                        #pragma unroll
                        for (uint32_t corner = 0; corner < 8; corner++) {
                            float wt_ = 1;
                            #pragma unroll
                            for (int d = 0; d < 3; d++) wt_ *= ((corner >> d) & 1u) ? cell.w[d] : 1 - cell.w[d];
                            e0 += wt_ * rows[corner][0];
                            e1 += wt_ * rows[corner][1];
                        }
                        #pragma unroll
                        for (int gd = 0; gd < 3; gd++) {
                            #pragma unroll
                            for (uint32_t sub = 0; sub < 4; sub++) {
                                float wt_ = cell.scale;
                                uint32_t corner = 0;
                                #pragma unroll
                                for (int nd = 0; nd < 2; nd++) {
                                    const int d = nd >= gd ? nd + 1 : nd;
                                    if ((sub >> nd) & 1u) { wt_ *= cell.w[d]; corner |= 1u << d; }
                                    else                  { wt_ *= 1 - cell.w[d]; }
                                }
                                j[gd * 2 + 0] += wt_ * (rows[corner | (1u << gd)][0] - rows[corner][0]) * cell.dw[gd];
                                j[gd * 2 + 1] += wt_ * (rows[corner | (1u << gd)][1] - rows[corner][1]) * cell.dw[gd];
                            }
                        }
                    }
                    uint32_t hi, lo;
                    g_split2(e0, e1, hi, lo);
                    const uint32_t off = tc::op_off(128, s, 2 * l);
                    *reinterpret_cast<uint32_t*>(s_op + off) = hi;
                    *reinterpret_cast<uint32_t*>(s_op + kGOperandHalf + off) = lo;
                    float* jq = s_jac + s * kGJacLd + 6 * l;
                    #pragma unroll
                    for (int q = 0; q < 6; q++) jq[q] = j[q];
                }
                // zero the encoder columns between 2L and the padded K of the first layer
                if (par == 0) for (uint32_t k = 2 * G.L; k < G.F[0].Kp; k += 2) {
                    const uint32_t off = tc::op_off(128, s, k);
                    *reinterpret_cast<uint32_t*>(s_op + off) = 0u;
                    *reinterpret_cast<uint32_t*>(s_op + kGOperandHalf + off) = 0u;
                }
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(a_ready);

            uint32_t mask[2] = {0u, 0u};                   // ReLU masks of hidden layers 0 and 1 for this thread's 32 columns
            for (int st = 0; st < n_stages; st++) {
                tc::mbar_wait(acc_ready, acc_par); acc_par ^= 1;
                tc::tc_fence_after();
                uint8_t* dst = s_op + ((st + 1) & 1) * kGOperand;          // operand buffer of the next stage
                if (st < n - 1) {
                    // hidden forward layer st: h = relu(D + b), keep the mask, write the next operand
                    if (g < hchunks) {
                        uint32_t r[32];
                        tc::tmem_ld32(tmem + lane_addr + g * 32, r);
                        tc::tmem_ld_wait();
                        const float* bias = s_f + st * 64 + g * 32;
                        uint32_t mk = 0;
                        #pragma unroll
                        for (int jj = 0; jj < 4; jj++) {
                            uint32_t ph[4], pl[4];
                            #pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const int c = jj * 8 + e * 2;
                                const float v0 = __uint_as_float(r[c]) + bias[c], v1 = __uint_as_float(r[c + 1]) + bias[c + 1];
                                mk |= (v0 > 0.f ? 1u : 0u) << c;
                                mk |= (v1 > 0.f ? 1u : 0u) << (c + 1);
                                g_split2(fmaxf(v0, 0.f), fmaxf(v1, 0.f), ph[e], pl[e]);
                            }
                            const uint32_t off = tc::op_off(128, row, g * 32 + jj * 8);
                            *reinterpret_cast<uint4*>(dst + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(dst + kGOperandHalf + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                        mask[st] = mk;
                    }
                } else if (st == n - 1) {
                    // last forward layer: keep its outputs, start the reverse pass: g = W_last[0,:] * relu'(h_{n-2})
                    if (g == 0) {
                        uint32_t r[16];
                        tc::tmem_ld16(tmem + lane_addr, r);
                        tc::tmem_ld_wait();
                        const float* bias = s_f + st * 64;
                        #pragma unroll
                        for (int i = 0; i < 16; i++) s_side[i * 128 + row] = __uint_as_float(r[i]) + bias[i];
                    }
                    if (g < hchunks) {
                        const float* w0 = s_f + n * 64 + g * 32;
                        const uint32_t mk = mask[n - 2];
                        #pragma unroll
                        for (int jj = 0; jj < 4; jj++) {
                            uint32_t ph[4], pl[4];
                            #pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const int c = jj * 8 + e * 2;
                                g_split2(((mk >> c) & 1u) ? w0[c] : 0.f, ((mk >> (c + 1)) & 1u) ? w0[c + 1] : 0.f, ph[e], pl[e]);
                            }
                            const uint32_t off = tc::op_off(128, row, g * 32 + jj * 8);
                            *reinterpret_cast<uint4*>(dst + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(dst + kGOperandHalf + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                    }
                } else if (st < n_stages - 1) {
                    // reverse stage through layer i = n_stages - 1 - st (>= 1): g_{i-1} = D * relu'(h_{i-1})
                    const int i = n_stages - 1 - st;
                    if (g < hchunks) {
                        uint32_t r[32];
                        tc::tmem_ld32(tmem + lane_addr + g * 32, r);
                        tc::tmem_ld_wait();
                        const uint32_t mk = mask[i - 1];
                        #pragma unroll
                        for (int jj = 0; jj < 4; jj++) {
                            uint32_t ph[4], pl[4];
                            #pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const int c = jj * 8 + e * 2;
                                g_split2(((mk >> c) & 1u) ? __uint_as_float(r[c]) : 0.f, ((mk >> (c + 1)) & 1u) ? __uint_as_float(r[c + 1]) : 0.f,
                                         ph[e], pl[e]);
                            }
                            const uint32_t off = tc::op_off(128, row, g * 32 + jj * 8);
                            *reinterpret_cast<uint4*>(dst + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(dst + kGOperandHalf + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                    }
                } else {
                    // last reverse stage: D = d sdf / d enc; contract with the jacobian, finish the sample
                    if (g == 0) {
                        uint32_t r[32];
                        tc::tmem_ld32(tmem + lane_addr, r);
                        tc::tmem_ld_wait();
                        const uint32_t m = m0 + row;
                        const float* jq = s_jac + row * kGJacLd;
                        float gx = 0.f, gy = 0.f, gz = 0.f;
                        for (uint32_t l = 0; l < G.L; l++) {
                            const float g0 = __uint_as_float(r[2 * l]), g1 = __uint_as_float(r[2 * l + 1]);
                            gx += g0 * jq[6 * l + 0] + g1 * jq[6 * l + 1];
                            gy += g0 * jq[6 * l + 2] + g1 * jq[6 * l + 3];
                            gz += g0 * jq[6 * l + 4] + g1 * jq[6 * l + 5];
                        }
                        const float inv2b = 1.0f / (2 * G.bound);
                        gx *= inv2b; gy *= inv2b; gz *= inv2b;
                        const float gn = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-10f);
                        const float nx = gx / gn, ny = gy / gn, nz = gz / gn;
                        const int Gd = (int)G.geo_dim;
                        const float sdf = s_side[0 * 128 + row];
                        const float sg = (sdf > 0.f) ? 1.f : ((sdf < 0.f) ? -1.f : 0.f);
                        const float sigma = (1.0f / G.beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(sdf) / G.beta)) * G.density_scale;
                        if (m < M) {
                            if (O.sigma) O.sigma[m] = sigma;
                            if (O.sdf) O.sdf[m] = sdf;
                            if (O.normal) { O.normal[3 * (size_t)m] = nx; O.normal[3 * (size_t)m + 1] = ny; O.normal[3 * (size_t)m + 2] = nz; }
                            if (O.grad_x) { O.grad_x[3 * (size_t)m] = gx; O.grad_x[3 * (size_t)m + 1] = gy; O.grad_x[3 * (size_t)m + 2] = gz; }
                            const float rough = G.rough_act_scale * g_softplus(s_side[(1 + Gd) * 128 + row] + G.rough_bias) * G.rough_scale;
                            if (O.roughness) O.roughness[m] = rough;
                            if (rec && mode != 1) {
                                float ss = 0.f;
                                for (int i = 0; i < Gd; i++) { const float v = s_side[(1 + i) * 128 + row]; ss += v * v; }
                                const float ginv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                                const float blend = g_sigmoid(s_side[(2 + Gd) * 128 + row]);
                                const float dx = dirs[3 * (size_t)m], dy = dirs[3 * (size_t)m + 1], dz = dirs[3 * (size_t)m + 2];
                                const float wox = -dx, woy = -dy, woz = -dz;
                                const float ndot = nx * wox + ny * woy + nz * woz;
                                float wrx = 2 * ndot * nx - wox, wry = 2 * ndot * ny - woy, wrz = 2 * ndot * nz - woz;
                                float nex = nx, ney = ny, nez = nz;
                                if (G.has_rot) {
                                    const float* R = G.rot;
                                    const float a0 = wrx * R[0] + wry * R[3] + wrz * R[6], a1 = wrx * R[1] + wry * R[4] + wrz * R[7],
                                                a2 = wrx * R[2] + wry * R[5] + wrz * R[8];
                                    wrx = a0; wry = a1; wrz = a2;
                                    const float b0 = nx * R[0] + ny * R[3] + nz * R[6], b1 = nx * R[1] + ny * R[4] + nz * R[7],
                                                b2 = nx * R[2] + ny * R[5] + nz * R[8];
                                    nex = b0; ney = b1; nez = b2;
                                }
                                float* q = rec + (size_t)m * kTcRecFloats;
                                for (int i = 0; i < Gd; i++) q[i] = s_side[(1 + i) * 128 + row] * ginv;
                                q[16] = nx; q[17] = ny; q[18] = nz; q[19] = ndot; q[20] = rough; q[21] = blend;
                                q[22] = nex; q[23] = ney; q[24] = nez; q[25] = wrx; q[26] = wry; q[27] = wrz;
                            }
                        }
                    }
                }
                if (st < n_stages - 1) {
                    tc::tc_fence_before();
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive(a_ready);
                } else {
                    tc::tc_fence_before();
                    __syncwarp();
                }
            }
            // all 8 worker warps must be done with s_jac / s_side / TMEM before the next tile's gather overwrites them
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 64);
}

// image of B[n][k] (N x K, K-major) for one layer; transpose = 1 reads W as [K][N] (reverse pass: B = W^T)
__global__ void k_pack_tc2(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np,
                           int transpose) {
    const uint32_t total = Kp * Np;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t nn = i / Kp, k = i - nn * Kp;
        float v = 0.0f;
        if (nn < N && k < K) v = transpose ? W[(size_t)k * N + nn] : W[(size_t)nn * K + k];
        __half h, lo;
        tc::split_f16(v, h, lo);
        const uint32_t s = k >> 4, kk = k & 15;
        const size_t base = (size_t)s * Np * 64 + (kk >> 3) * (Np * 16) + nn * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + base) = h;
        *reinterpret_cast<__half*>(img + base + (size_t)Np * 32) = lo;
    }
}
__global__ void k_pack_floats(const float* __restrict__ src, float* __restrict__ dst, uint32_t n, uint32_t n_pad) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) dst[i] = (src && i < n) ? src[i] : 0.0f;
}

static uint32_t rup2(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

bool geom_tc_layout(const envidr_field* f, uint64_t base_bytes, TcGeom* out, uint64_t* total_bytes) {
    TcGeom& g = *out;
    g = TcGeom{};
    const uint32_t n = f->n_sdf;
    if (n < 2 || n > 3 || f->level_dim != 2) return false;
    const uint32_t in0 = f->num_levels * 2, Hd = f->sdf[0].out_dim;
    if (!(in0 == 16 || in0 == 32) || !(Hd == 32 || Hd == 64)) return false;
    if (f->sdf[n - 1].out_dim > 16 || f->sdf[n - 1].out_dim < 3 + f->geo_feat_dim) return false;
    for (uint32_t i = 0; i + 1 < n; i++) if (f->sdf[i].out_dim != Hd) return false;
    uint32_t off = 0;
    auto img = [&](TcImg& I, uint32_t K, uint32_t N, uint32_t Np) {
        I.Kp = rup2(K, 16); I.N = N; I.Np = Np; I.off = off;
        off += (I.Kp / 16) * Np * 64;
    };
    for (uint32_t i = 0; i < n; i++) img(g.F[i], f->sdf[i].in_dim, f->sdf[i].out_dim, i == n - 1 ? 16 : Hd);
    for (uint32_t i = 0; i + 1 < n; i++) img(g.R[i], f->sdf[i].out_dim, f->sdf[i].in_dim, f->sdf[i].in_dim);    // B = W_i^T: K = out_i, N = in_i
    g.float_off = off;
    off += (n + 1) * 64 * 4;                   // biases [n][64] + w_row0[64]
    g.res_bytes = off;
    g.res_bytes_al = rup2(off, 1024);
    g.n_layers = n;
    const uint64_t start = rup2((uint32_t)base_bytes, 1024);
    if (f->packed) g.blob = reinterpret_cast<const uint8_t*>(f->packed) + start;
    g.blob_off = start;
    g.table = f->embeddings; g.offsets = f->offsets; g.L = f->num_levels; g.H = f->base_resolution; g.S = f->log2_per_level_scale;
    g.bound = f->bound; g.enabled_levels = f->enabled_levels; g.geo_dim = f->geo_feat_dim;
    g.beta = f->beta; g.density_scale = f->density_scale; g.rough_bias = f->roughness_bias; g.rough_act_scale = f->roughness_act_scale;
    g.rough_scale = f->roughness_scale; g.has_rot = f->has_env_rot;
    for (int i = 0; i < 9; i++) g.rot[i] = f->env_rot[i];
    *total_bytes = start + g.res_bytes_al;
    return true;
}

int geom_tc_pack(const envidr_field* f, const TcGeom& g, void* packed, cudaStream_t st) {
    uint8_t* blob = reinterpret_cast<uint8_t*>(packed) + g.blob_off;
    const uint32_t n = g.n_layers;
    for (uint32_t i = 0; i < n; i++)
        k_pack_tc2<<<32, 256, 0, st>>>(f->sdf[i].weight, blob + g.F[i].off, f->sdf[i].in_dim, f->sdf[i].out_dim, g.F[i].Kp, g.F[i].Np, 0);
    for (uint32_t i = 0; i + 1 < n; i++)
        k_pack_tc2<<<32, 256, 0, st>>>(f->sdf[i].weight, blob + g.R[i].off, f->sdf[i].out_dim, f->sdf[i].in_dim, g.R[i].Kp, g.R[i].Np, 1);
    float* fl = reinterpret_cast<float*>(blob + g.float_off);
    for (uint32_t i = 0; i < n; i++) k_pack_floats<<<1, 64, 0, st>>>(f->sdf[i].bias, fl + i * 64, f->sdf[i].out_dim, 64);
    k_pack_floats<<<1, 64, 0, st>>>(f->sdf[n - 1].weight, fl + n * 64, f->sdf[n - 1].in_dim, 64);        // row 0 of the last layer
    return check_launch("geom_tc_pack");
}

int geom_tc_launch(const TcGeom& g, const float* xyzs, const float* dirs, const uint32_t* M_dev, uint32_t M_host, int mode, float* rec,
                   const envidr_field_out* out, cudaStream_t st) {
    const size_t smem = (size_t)g.res_bytes_al + 2 * kGOperand + 128 * kGJacLd * 4 + 16 * 128 * 4 + 64;
    static size_t attr_set = 0;
    if (attr_set < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_geom_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("geom_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = smem;
    }
    uint32_t grid = kSMs;
    if (!M_dev) grid = min((uint32_t)kSMs, (M_host + 127) / 128);
    if (grid == 0) return 0;
    GeomOutDev O{out->sigma, out->normal, out->sdf, out->roughness, out->grad_x};
    k_geom_tc<<<grid, kGThreads, smem, st>>>(g, xyzs, dirs, M_dev, M_host, mode, rec, O);
    return check_launch("geom_tc");
}

}  // namespace envidr
