// neus_geom_tc.cu -- the geometry network of the NeuS-style field (BASELINE config 4) as ONE kernel per batch of samples:
// frequency encoding -> 8 x 256 weight-normed layers with Softplus(beta = 100) and a skip connection -> head (sdf, features), and the
// REVERSE pass through the same layers for the analytic normal  g <- (g . softplus'(z)) W  down to d sdf / d x -- forward + backward of
// the stack back to back per 128-sample tile, activations never leaving the SM except the softplus derivatives (see below).
//
// Reference: nerf/network.py:154-222 (construction), :415-421 (forward with skip_layers), nerf/renderer.py:182-198 (normal by
// autograd.grad), freqencoder/src/freqencoder.cu:30-94.  The reference runs 8 cuBLAS GEMMs + elementwise kernels forward and an
// autograd graph of the same size backward per render iteration; round 2's first version (envidr_b200/neus_field.py, composed from
// envidr_linear_tc + csrc/neus_field.cu) still round-tripped every activation through HBM (7 % of the tensor roofline).
//
// Structure = k_env_tc's (csrc/field_tc.cu): persistent CTA per SM, warp-specialised, mbarrier-only hand-offs
//   warp 0      producer: weight images of the 15 GEMMs of a tile (8 forward, 7 reverse = images of W^T) from L2 through a
//               4 x 16 KB ring with 1-D bulk copies (k_linear_pack layout: per 16-wide K step [hi | lo], fp16)
//   warp 1      issuer: tcgen05.mma kind::f16, M = 128, three MMAs per K step (hi*hi + lo*hi + hi*lo, fp32 accumulate), two 256-column
//               accumulators ping-pong in TMEM so that the epilogue of GEMM i overlaps the MMAs of GEMM i + 1 chunk by chunk
//   warps 4-19  epilogue: tcgen05.ld -> bias + Softplus and its derivative (forward) or x saved derivative (reverse) -> fp16 hi/lo
//               re-split -> next A operand in shared memory, 32 columns at a time; skip concat / split, head, final transpose-Jacobian
//               product of the frequency encoding (d sdf / d x)
//   warps 20-23 frequency encoding of the next tile's layer-0 operand
// The softplus derivatives s_l = sigmoid(beta z_l) are needed again in the reverse pass: 7 x 128 x 256 fp32 = 896 KB per tile do not
// fit next to the operands, so each CTA spills them to its own slice of a scratch buffer (L2-resident; every thread reads back
// exactly what it wrote).  Everything else (activations, operands, gradients) stays in shared / tensor memory.
#include <math.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace envidr {

constexpr int kNgEpiWarps = 16, kNgEpiGroups = kNgEpiWarps / 4;
constexpr int kNgThreads = (4 + kNgEpiWarps + 4) * 32;       // 4 control warps, 16 epilogue warps, 4 encode warps
constexpr int kNgStages = 4;
constexpr uint32_t kNgStageBytes = 16384;
constexpr uint32_t kNgARegion = 65536;             // 128 rows x 256 K x 2 B
constexpr uint32_t kNgERegion = 12288;             // 128 rows x 48 K x 2 B (layer-0 operand); both halves are re-used as g_skip [48][128] fp32
constexpr int kNgMaxFwd = 8, kNgMaxGemm = 2 * kNgMaxFwd;
constexpr uint32_t kNgMaxIn = 48;

struct NgLayer {
    const uint8_t* img;
    const float* bias;         // forward layers
    uint32_t Kp, Np, N;        // padded GEMM shape; N = valid output columns
    uint32_t kind;             // 0 forward hidden, 1 forward head, 2 reverse hidden, 3 reverse into the encoding
    uint32_t out_chunks;       // 32-column chunks of the next GEMM's A operand this epilogue publishes
    uint32_t s_slot;           // kind 0: slot the derivative is saved to; kinds 1, 2: slot it is read from
    uint32_t s_cols;           // valid columns of that slot
    uint32_t append_enc;       // kind 0: the operand is cat([h, enc]) (the layer in front of the skip layer)
    uint32_t split_skip;       // kind 2: columns >= Nh are the gradient of the concatenated encoding
    uint32_t Nh;
    float scale;               // 1/sqrt(2) at the skip concat / split, else 1
};
struct NgParams {
    NgLayer L[kNgMaxGemm];
    const float* head_row;     // W_last[0, :]  (d sdf / d h of the last hidden layer)
    uint32_t n_gemm, n_fwd, in_dim, multires, has_skip, head_cols, n_slots;
    float beta;
};

__device__ __forceinline__ float exp2f_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float log2f_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ float ng_enc(float x, float y, float z, uint32_t e) {
    if (e < 3) return e == 0 ? x : (e == 1 ? y : z);
    const uint32_t col = e / 3 - 1, d = e % 3, f = col >> 1;
    const float v = d == 0 ? x : (d == 1 ? y : z);
    return __sinf(v * __uint_as_float((127u + f) << 23) + (float)(col & 1u) * (3.141592653589793f / 2));   // csrc/direnc.cu k_freq_fwd
}

__global__ void __launch_bounds__(kNgThreads, 1)
k_neus_geom_tc(const NgParams P, const float* __restrict__ xyzs, uint32_t M, float* __restrict__ head, float* __restrict__ grad_x,
               float* __restrict__ S) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA_hi = smem;
    uint8_t* sA_lo = smem + kNgARegion;
    uint8_t* sE_hi = smem + 2 * kNgARegion;
    uint8_t* sE_lo = sE_hi + kNgERegion;
    float* sGskip = reinterpret_cast<float*>(sE_hi);                        // [in_dim][128], alive from the skip split to the end of the tile
    uint8_t* ring = sE_lo + kNgERegion;
    float* s_bias = reinterpret_cast<float*>(ring + kNgStages * kNgStageBytes);      // [kNgMaxFwd][256]
    float* s_wrow = s_bias + kNgMaxFwd * 256;                               // [256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_wrow + 256);
    uint64_t* full = bars;                          // [4]
    uint64_t* empty = bars + kNgStages;             // [4]
    uint64_t* acc_ready = bars + 2 * kNgStages;     // [2]
    uint64_t* enc_full = acc_ready + 2;             // encode warps -> issuer (128 arrivals)
    uint64_t* tile_done = acc_ready + 3;            // epilogue warps -> encode warps (all epilogue threads)
    uint64_t* a_rdy = acc_ready + 4;                // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_rdy + 8);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t n_tiles = (M + 127) / 128;
    if (blockIdx.x >= n_tiles) return;
    const int nL = (int)P.n_gemm;

    if (tid == 0) {
        for (int i = 0; i < kNgStages; i++) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&acc_ready[0], 1); tc::mbar_init(&acc_ready[1], 1);
        tc::mbar_init(enc_full, 128); tc::mbar_init(tile_done, kNgEpiWarps * 32);
        for (int i = 0; i < 8; i++) tc::mbar_init(&a_rdy[i], 128);
        tc::mbar_fence_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
    for (uint32_t i = tid; i < P.n_fwd * 256; i += kNgThreads) {
        const uint32_t l = i >> 8, c = i & 255;
        s_bias[i] = (P.L[l].bias && c < P.L[l].N) ? __ldg(P.L[l].bias + c) : 0.0f;
    }
    for (uint32_t i = tid; i < 256; i += kNgThreads) s_wrow[i] = (i < P.head_cols) ? __ldg(P.head_row + i) : 0.0f;
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== producer =====================
        uint32_t stage = 0, phase = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int i = 0; i < nL; i++) {
                const uint32_t ksteps = P.L[i].Kp / 16, kbytes = P.L[i].Np * 64;
                const uint32_t kper = max(1u, kNgStageBytes / kbytes);
                const uint8_t* src = P.L[i].img;
                for (uint32_t s = 0; s < ksteps; s += kper) {
                    const uint32_t bytes = min(kper, ksteps - s) * kbytes;
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    if (lane == 0) {
                        tc::mbar_arrive_expect_tx(&full[stage], bytes);
                        tc::bulk_g2s(ring + stage * kNgStageBytes, src + (size_t)s * kbytes, bytes, &full[stage]);
                    }
                    __syncwarp();
                    if (++stage == kNgStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t stage = 0, phase = 0, enc_par = 0, chunk_par = 0, gl = 0;
        const uint32_t ring0 = tc::smem_u32(ring);
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int i = 0; i < nL; i++) {
                const uint32_t ksteps = P.L[i].Kp / 16, Np = P.L[i].Np;
                const uint32_t idesc = tc::make_idesc_f16(128, Np);
                const uint32_t d_tmem = tmem + (gl & 1u) * 256u;
                const uint32_t buf = gl & 1u;
                gl++;
                uint64_t da_hi, da_lo;
                if (i == 0) {
                    tc::mbar_wait(enc_full, enc_par); enc_par ^= 1;
                    da_hi = tc::make_smem_desc(tc::smem_u32(sE_hi), 2048, 128);
                    da_lo = tc::make_smem_desc(tc::smem_u32(sE_lo), 2048, 128);
                } else {
                    da_hi = tc::make_smem_desc(tc::smem_u32(sA_hi), 2048, 128);
                    da_lo = tc::make_smem_desc(tc::smem_u32(sA_lo), 2048, 128);
                }
                const uint64_t db0 = tc::make_smem_desc(ring0, Np * 16, 128);
                const uint32_t lo_off = Np * 32, kbytes = Np * 64;
                const uint32_t kper = max(1u, kNgStageBytes / kbytes);
                for (uint32_t s0 = 0; s0 < ksteps; s0 += kper) {
                    tc::mbar_wait(&full[stage], phase);
                    const uint32_t kend = min(ksteps, s0 + kper);
                    uint64_t db_hi = tc::desc_advance(db0, stage * kNgStageBytes);
                    for (uint32_t s = s0; s < kend; s++) {
                        if (i > 0 && (s & 1u) == 0) {
                            const uint32_t c = s >> 1;
                            tc::mbar_wait(&a_rdy[c], (chunk_par >> c) & 1u);
                            chunk_par ^= 1u << c;
                        }
                        tc::tc_fence_after();
                        __syncwarp();
                        const uint64_t db_lo = tc::desc_advance(db_hi, lo_off);
                        tc::mma_f16_ss_w(d_tmem, da_hi, db_hi, idesc, s > 0);
                        tc::mma_f16_ss_w(d_tmem, da_lo, db_hi, idesc, 1);
                        tc::mma_f16_ss_w(d_tmem, da_hi, db_lo, idesc, 1);
                        da_hi = tc::desc_advance(da_hi, 4096); da_lo = tc::desc_advance(da_lo, 4096);
                        db_hi = tc::desc_advance(db_hi, kbytes);
                    }
                    tc::mma_commit_w(&empty[stage]);
                    if (++stage == kNgStages) { stage = 0; phase ^= 1; }
                }
                tc::mma_commit_w(&acc_ready[buf]);
            }
        }
    } else if (warp >= 4 && warp < 4 + kNgEpiWarps) {
        // ===================== epilogue warps =====================
        // 16 warps = 4 groups x 4 TMEM lane quarters; group g owns the 32-column chunks cb = g (mod 4) and works through each in two
        // 16-column halves (register budget of a 768-thread CTA).  Measured with 8 warps and 32-column steps (ncu, run 18): the kernel
        // was bound by these warps' instruction issue (0.94 IPC per SM, the issuer waiting 77 % of its time for A-operand chunks).
        const uint32_t quarter = warp & 3, g = (warp - 4) >> 2;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        float* S_cta = S + (size_t)blockIdx.x * P.n_slots * 128 * 256;
        const uint32_t in_dim = P.in_dim;
        uint32_t acc_par = 0, gl = 0;
        const float inv_beta = 1.0f / P.beta;
        auto store16 = [&](uint32_t c0, const float (&v)[16]) {           // 16 columns of the next A operand, starting at column c0
            #pragma unroll
            for (int q = 0; q < 2; q++) {
                float w[8];
                #pragma unroll
                for (int j = 0; j < 8; j++) w[j] = v[8 * q + j];
                tc::store_chunk8(sA_hi, sA_lo, row, c0 + q * 8, w);
            }
        };
        auto publish = [&](uint32_t cb) {
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&a_rdy[cb]);
        };
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t m = tile * 128 + row;
            const bool valid = m < M;
            float px = 0.f, py = 0.f, pz = 0.f;
            if (valid) { px = xyzs[3 * (size_t)m]; py = xyzs[3 * (size_t)m + 1]; pz = xyzs[3 * (size_t)m + 2]; }
            for (int i = 0; i < nL; i++) {
                const NgLayer& L = P.L[i];
                const uint32_t buf = gl & 1u;
                gl++;
                tc::mbar_wait(&acc_ready[buf], (acc_par >> buf) & 1u); acc_par ^= 1u << buf;
                tc::tc_fence_after();
                const uint32_t acc = tmem + lane_addr + buf * 256u;
                const uint32_t nin = L.Np / 32;
                if (L.kind == 0) {
                    // ---- forward hidden layer: z = D + b, h = softplus(z), s = sigmoid(beta z) saved for the reverse pass
                    const float* bias = s_bias + i * 256;
                    float* Srow = S_cta + ((size_t)L.s_slot * 128 + row) * 256;
                    const uint32_t nmax = max(nin, L.out_chunks);
                    const float scale = L.scale;
                    const float k_t = P.beta * 1.4426950408889634f, k_out = 0.6931471805599453f * inv_beta * scale;
                    for (uint32_t cb = g; cb < nmax; cb += kNgEpiGroups) {
                        #pragma unroll 1
                        for (uint32_t hf = 0; hf < 2; hf++) {
                            const uint32_t c0 = cb * 32 + hf * 16;
                            float v[16];
                            if (cb < nin) {
                                uint32_t r[16];
                                tc::tmem_ld16(acc + c0, r);
                                tc::tmem_ld_wait();
                                float sv[16];
                                // Softplus(beta, threshold 20) and its derivative from ONE exponential: e = exp(-|beta z|) in (0, 1], so
                                // log(1 + e) needs no log1p (absolute error 6e-8, i.e. 6e-10 after / beta) and the fast intrinsics (MUFU
                                // ex2 / lg2 / rcp, relative 2^-21) keep h and s within 1e-6 of the libm formulation
                                // In base 2: t = beta z log2(e) (one FMA on the accumulator, bias pre-scaled), e = 2^-|t|,
                                // softplus = (max(t, 0) + log2(1 + e)) ln2 / beta.  No threshold branch: for beta z > 20, 1 + e rounds to 1 and
                                // the expression returns beta z / beta, torch's threshold value to 1 ulp.
                                #pragma unroll
                                for (int j = 0; j < 16; j++) {
                                    const float t = fmaf(__uint_as_float(r[j]), k_t, bias[c0 + j] * k_t);
                                    const float e = exp2f_fast(-fabsf(t));
                                    const float ope = 1.0f + e;
                                    const float rc = __frcp_rn(ope);
                                    sv[j] = (t >= 0.0f) ? rc : e * rc;
                                    v[j] = (fmaxf(t, 0.0f) + log2f_fast(ope)) * k_out;
                                }
                                if (c0 + 16 > L.N) {                     // padded columns of the operand are zero (or the encoding, below)
                                    #pragma unroll
                                    for (int j = 0; j < 16; j++) if (c0 + j >= L.N) v[j] = 0.0f;
                                }
                                float4* s4 = reinterpret_cast<float4*>(Srow + c0);
                                #pragma unroll
                                for (int q = 0; q < 4; q++) s4[q] = make_float4(sv[4 * q], sv[4 * q + 1], sv[4 * q + 2], sv[4 * q + 3]);
                            } else {
                                #pragma unroll
                                for (int j = 0; j < 16; j++) v[j] = 0.0f;
                            }
                            if (L.append_enc && c0 + 16 > L.N) {         // h = cat([h, x_enc]) / sqrt(2)   (network.py:417-418)
                                #pragma unroll
                                for (int j = 0; j < 16; j++) {
                                    const uint32_t col = c0 + j;
                                    if (col >= L.N && col < L.N + in_dim) v[j] = ng_enc(px, py, pz, col - L.N) * scale;
                                }
                            }
                            if (cb < L.out_chunks) store16(c0, v);
                        }
                        if (cb < L.out_chunks) publish(cb);
                    }
                } else if (L.kind == 1) {
                    // ---- forward head: (sdf, features) out; first reverse operand g = W_last[0, :] . s of the last hidden layer
                    if (g == 0) {
                        uint32_t r[16];
                        tc::tmem_ld16(acc, r);
                        tc::tmem_ld_wait();
                        if (valid) {
                            const float* bias = s_bias + i * 256;
                            float4* dst = reinterpret_cast<float4*>(head + (size_t)m * 16);
                            #pragma unroll
                            for (int q = 0; q < 4; q++)
                                dst[q] = make_float4(__uint_as_float(r[4 * q]) + bias[4 * q], __uint_as_float(r[4 * q + 1]) + bias[4 * q + 1],
                                                     __uint_as_float(r[4 * q + 2]) + bias[4 * q + 2], __uint_as_float(r[4 * q + 3]) + bias[4 * q + 3]);
                        }
                    }
                    const float* Srow = S_cta + ((size_t)L.s_slot * 128 + row) * 256;
                    for (uint32_t cb = g; cb < L.out_chunks; cb += kNgEpiGroups) {
                        #pragma unroll 1
                        for (uint32_t hf = 0; hf < 2; hf++) {
                            const uint32_t c0 = cb * 32 + hf * 16;
                            float v[16];
                            const float4* s4 = reinterpret_cast<const float4*>(Srow + c0);
                            #pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const float4 sq = s4[q];
                                const uint32_t c = c0 + 4 * q;
                                v[4 * q] = sq.x * s_wrow[c]; v[4 * q + 1] = sq.y * s_wrow[c + 1];
                                v[4 * q + 2] = sq.z * s_wrow[c + 2]; v[4 * q + 3] = sq.w * s_wrow[c + 3];
                            }
                            if (c0 + 16 > L.s_cols) {
                                #pragma unroll
                                for (int j = 0; j < 16; j++) if (c0 + j >= L.s_cols) v[j] = 0.0f;
                            }
                            store16(c0, v);
                        }
                        publish(cb);
                    }
                } else if (L.kind == 2) {
                    // ---- reverse hidden: D = g_in = g_z W_l; next g_z = g_in . s_{l-1}; at the skip layer the trailing columns are d / d enc
                    const float* Srow = S_cta + ((size_t)L.s_slot * 128 + row) * 256;
                    const uint32_t nmax = max(nin, L.out_chunks);
                    const float scale = L.scale;
                    const uint32_t lim = L.split_skip ? L.Nh : L.s_cols;   // columns < lim continue down the stack
                    for (uint32_t cb = g; cb < nmax; cb += kNgEpiGroups) {
                        #pragma unroll 1
                        for (uint32_t hf = 0; hf < 2; hf++) {
                            const uint32_t c0 = cb * 32 + hf * 16;
                            float v[16], sv[16];
                            uint32_t r[16];
                            if (c0 < lim) {                              // the saved derivatives first: their L2 latency overlaps the tcgen05.ld
                                const float4* s4 = reinterpret_cast<const float4*>(Srow + c0);
                                #pragma unroll
                                for (int q = 0; q < 4; q++) { const float4 sq = s4[q]; sv[4 * q] = sq.x; sv[4 * q + 1] = sq.y; sv[4 * q + 2] = sq.z; sv[4 * q + 3] = sq.w; }
                            } else {
                                #pragma unroll
                                for (int j = 0; j < 16; j++) sv[j] = 0.0f;
                            }
                            if (cb < nin) { tc::tmem_ld16(acc + c0, r); tc::tmem_ld_wait(); }
                            else {
                                #pragma unroll
                                for (int j = 0; j < 16; j++) r[j] = 0u;
                            }
                            #pragma unroll
                            for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]) * scale * sv[j];
                            if (c0 + 16 > lim) {
                                #pragma unroll
                                for (int j = 0; j < 16; j++) {
                                    const uint32_t col = c0 + j;
                                    if (col >= lim) {
                                        if (L.split_skip && col < L.Nh + in_dim) sGskip[(col - L.Nh) * 128 + row] = __uint_as_float(r[j]) * scale;
                                        v[j] = 0.0f;
                                    }
                                }
                            }
                            if (cb < L.out_chunks) store16(c0, v);
                        }
                        if (cb < L.out_chunks) publish(cb);
                    }
                } else if (g == 0) {
                    // ---- reverse into the encoding: g_enc = D (+ skip part); d sdf / d x = J_freq^T g_enc   (freqencoder.cu:82-90)
                    float ge[kNgMaxIn];
                    #pragma unroll
                    for (int t3 = 0; t3 < 3; t3++) {
                        uint32_t r[16];
                        if ((uint32_t)t3 * 16 < L.Np) { tc::tmem_ld16(acc + t3 * 16, r); tc::tmem_ld_wait(); }
                        #pragma unroll
                        for (int j = 0; j < 16; j++) ge[t3 * 16 + j] = ((uint32_t)t3 * 16 < L.Np) ? __uint_as_float(r[j]) : 0.0f;
                    }
                    if (P.has_skip) {
                        #pragma unroll
                        for (int e = 0; e < (int)kNgMaxIn; e++) if ((uint32_t)e < in_dim) ge[e] += sGskip[e * 128 + row];
                    }
                    float gx = ge[0], gy = ge[1], gz = ge[2];
                    #pragma unroll
                    for (int f = 0; f < 7; f++) {
                        if ((uint32_t)f < P.multires) {
                            const float sc = __uint_as_float((127u + f) << 23);
                            const int b = 3 + 6 * f;
                            const float sx = __sinf(px * sc), cx = __sinf(px * sc + 3.141592653589793f / 2);
                            const float sy = __sinf(py * sc), cy = __sinf(py * sc + 3.141592653589793f / 2);
                            const float sz = __sinf(pz * sc), cz = __sinf(pz * sc + 3.141592653589793f / 2);
                            gx += sc * (ge[b] * cx - ge[b + 3] * sx);
                            gy += sc * (ge[b + 1] * cy - ge[b + 4] * sy);
                            gz += sc * (ge[b + 2] * cz - ge[b + 5] * sz);
                        }
                    }
                    if (valid) { grad_x[3 * (size_t)m] = gx; grad_x[3 * (size_t)m + 1] = gy; grad_x[3 * (size_t)m + 2] = gz; }
                    tc::tc_fence_before();
                }
            }
            tc::mbar_arrive(tile_done);
        }
    } else if (warp >= 4 + kNgEpiWarps) {
        // ===================== frequency encoding of the next tile's layer-0 operand =====================
        const uint32_t row = tid - (4 + kNgEpiWarps) * 32;
        const uint32_t Kp0 = P.L[0].Kp;
        uint32_t done_par = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const uint32_t m = tile * 128 + row;
            float px = 0.f, py = 0.f, pz = 0.f;
            if (m < M) { px = xyzs[3 * (size_t)m]; py = xyzs[3 * (size_t)m + 1]; pz = xyzs[3 * (size_t)m + 2]; }
            if (it > 0) { tc::mbar_wait(tile_done, done_par); done_par ^= 1; }     // the operand buffer doubles as g_skip of the tile before
            for (uint32_t k0 = 0; k0 < Kp0; k0 += 8) {
                float v[8];
                #pragma unroll
                for (int j = 0; j < 8; j++) v[j] = (k0 + j < P.in_dim) ? ng_enc(px, py, pz, k0 + j) : 0.0f;
                tc::store_chunk8(sE_hi, sE_lo, row, k0, v);
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(enc_full);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem, 512);
}

constexpr size_t kNgSmem = 2 * kNgARegion + 2 * kNgERegion + kNgStages * kNgStageBytes + (kNgMaxFwd * 256 + 256) * sizeof(float) + 256;

}  // namespace envidr

using namespace envidr;

extern "C" {

uint64_t envidr_neus_geometry_scratch_bytes(uint32_t n_layers) {
    if (n_layers < 2 || n_layers > (uint32_t)kNgMaxFwd) return 0;
    return (uint64_t)kSMs * (n_layers - 1) * 128 * 256 * sizeof(float);
}

int envidr_neus_geometry(const envidr_neus_net* net, const float* xyzs, uint32_t M, float* head, float* grad_x, void* scratch,
                         uint64_t scratch_bytes, envidr_stream_t stream) {
    ENVIDR_REQUIRE(net, ENVIDR_E_BADARG, "null pointer");
    const uint32_t nl = net->n_layers;
    ENVIDR_REQUIRE(nl >= 2 && nl <= (uint32_t)kNgMaxFwd, ENVIDR_E_UNSUPPORTED, "neus_geometry: 2..8 layers");
    const uint32_t in_dim = 3 + 6 * net->multires;
    ENVIDR_REQUIRE(net->multires >= 1 && net->multires <= 7 && in_dim <= kNgMaxIn, ENVIDR_E_UNSUPPORTED, "neus_geometry: frequency degree 1..7");
    ENVIDR_REQUIRE(net->layers[0].in_dim == in_dim, ENVIDR_E_BADARG, "neus_geometry: layer 0 must take the frequency encoding");
    const int ls = net->skip_layer;
    ENVIDR_REQUIRE(ls < 0 || (ls >= 1 && ls < (int)nl - 1), ENVIDR_E_UNSUPPORTED, "neus_geometry: skip layer must be a hidden layer > 0");
    ENVIDR_REQUIRE(net->layers[nl - 1].out_dim >= 1 && net->layers[nl - 1].out_dim <= 16, ENVIDR_E_UNSUPPORTED, "neus_geometry: head of at most 16 columns");
    NgParams P{};
    auto rup = [](uint32_t v, uint32_t m) { return (v + m - 1) / m * m; };
    uint32_t n = 0;
    for (uint32_t l = 0; l < nl; l++) {                                      // forward GEMMs
        const envidr_neus_layer& s = net->layers[l];
        ENVIDR_REQUIRE(s.img && s.bias && (l == 0 || s.imgT), ENVIDR_E_BADARG, "neus_geometry: null layer image / bias");
        const uint32_t in_expected = (l == 0) ? in_dim : net->layers[l - 1].out_dim + ((int)l == ls ? in_dim : 0);
        ENVIDR_REQUIRE(s.in_dim == in_expected, ENVIDR_E_BADARG, "neus_geometry: layer dims do not chain");
        ENVIDR_REQUIRE(s.in_dim <= 256 && s.out_dim <= 256, ENVIDR_E_UNSUPPORTED, "neus_geometry: widths <= 256");
        NgLayer& L = P.L[n++];
        L.img = reinterpret_cast<const uint8_t*>(s.img); L.bias = s.bias;
        L.Kp = rup(s.in_dim, 16); L.Np = rup(s.out_dim, 16); L.N = s.out_dim;
        L.scale = 1.0f;
        if (l + 1 < nl) {
            ENVIDR_REQUIRE(L.Np % 32 == 0 && (l == 0 || L.Kp % 32 == 0), ENVIDR_E_UNSUPPORTED, "neus_geometry: hidden widths must pad to multiples of 32");
            L.kind = 0; L.s_slot = l; L.s_cols = s.out_dim;
            L.append_enc = ((int)l + 1 == ls) ? 1u : 0u;
            if (L.append_enc) L.scale = 0.70710678118654752440f;
            L.out_chunks = rup(net->layers[l + 1].in_dim, 16) / 32;
            ENVIDR_REQUIRE(rup(net->layers[l + 1].in_dim, 16) % 32 == 0, ENVIDR_E_UNSUPPORTED, "neus_geometry: hidden widths must pad to multiples of 32");
        } else {
            L.kind = 1; L.s_slot = l - 1; L.s_cols = net->layers[l - 1].out_dim;
            L.out_chunks = rup(net->layers[l - 1].out_dim, 16) / 32;         // K of the first reverse GEMM
        }
    }
    for (int l = (int)nl - 2; l >= 0; l--) {                                 // reverse GEMMs: g_in = g_z W_l, image of W_l^T [in, out]
        const envidr_neus_layer& s = net->layers[l];
        NgLayer& L = P.L[n++];
        L.img = reinterpret_cast<const uint8_t*>(l == 0 ? s.imgT : s.imgT);
        ENVIDR_REQUIRE(s.imgT, ENVIDR_E_BADARG, "neus_geometry: null transposed image");
        L.Kp = rup(s.out_dim, 16); L.Np = rup(s.in_dim, 16); L.N = s.in_dim;
        L.scale = 1.0f;
        if (l > 0) {
            L.kind = 2; L.s_slot = l - 1; L.s_cols = net->layers[l - 1].out_dim;
            if (l == ls) { L.split_skip = 1; L.Nh = net->layers[l - 1].out_dim; L.scale = 0.70710678118654752440f; }
            L.out_chunks = rup(net->layers[l - 1].out_dim, 16) / 32;
            ENVIDR_REQUIRE(L.Np % 32 == 0, ENVIDR_E_UNSUPPORTED, "neus_geometry: hidden widths must pad to multiples of 32");
        } else {
            L.kind = 3; L.out_chunks = 0;
            ENVIDR_REQUIRE(L.Np <= 48, ENVIDR_E_UNSUPPORTED, "neus_geometry: encoding width");
        }
    }
    P.n_gemm = n; P.n_fwd = nl; P.in_dim = in_dim; P.multires = net->multires; P.has_skip = ls >= 0 ? 1u : 0u;
    P.head_row = net->head_row; P.head_cols = net->layers[nl - 1].in_dim; P.n_slots = nl - 1; P.beta = net->beta;
    ENVIDR_REQUIRE(P.head_row && P.beta > 0, ENVIDR_E_BADARG, "neus_geometry: head_row / beta");
    if (M == 0) return 0;
    ENVIDR_REQUIRE(xyzs && head && grad_x && scratch, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(scratch_bytes >= envidr_neus_geometry_scratch_bytes(nl) && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0, ENVIDR_E_WORKSPACE,
                   "scratch: envidr_neus_geometry_scratch_bytes(n_layers), 16-byte aligned");
    ENVIDR_REQUIRE((reinterpret_cast<uintptr_t>(head) & 15) == 0, ENVIDR_E_BADARG, "head must be 16-byte aligned");
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_neus_geom_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNgSmem);
        if (e != cudaSuccess) { set_error("neus_geom_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    const uint32_t n_tiles = (M + 127) / 128;
    const uint32_t grid = n_tiles < (uint32_t)kSMs ? n_tiles : (uint32_t)kSMs;
    k_neus_geom_tc<<<grid, kNgThreads, kNgSmem, as_stream(stream)>>>(P, xyzs, M, head, grad_x, reinterpret_cast<float*>(scratch));
    g_launches += 1;
    return check_launch("neus_geometry");
}

}  // extern "C"
