// shade_tc.cu -- shading heads of the per-sample field on tensor cores (sm_100a):
//   c_d = sigmoid(diffuse_net([geo | f_n])),  c_s = sigmoid(color_net([geo | n | f_r | n.w_o])),
//   inter-reflection branch: f_e = unitNorm(renv_net([r_rgb * vis | rho])), c_e = sigmoid(color_net([geo | n | f_e | n.w_o])),
//   c_s <- w c_s + (1 - w) c_e on the masked samples, rgb = (c_d + c_s) * intensity   (reference: nerf/network.py:524-698).
//
// Inputs are the per-sample record written by the geometry kernel and the unit-normalised env features written by
// k_env_tc.  All weight images (~75 KB) stay resident in shared memory; per 128-sample tile the worker warps assemble the
// input operand of each net, then alternate with the MMA-issuing warp layer by layer (same fp16 hi/lo split, 3 MMAs per K
// step, fp32 accumulation in TMEM as the other tensor-core kernels).
#include <math.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "field_tc.cuh"

namespace envidr {

constexpr int kSThreads = 384;                     // 4 control warps + 8 worker warps
constexpr uint32_t kSOperand = 32768;              // one A-operand buffer: 128 rows x 64 K x 2 B x (hi, lo)
constexpr uint32_t kSOperandHalf = 16384;

struct ShadeOutDev { float *rgb, *c_diffuse, *c_specular; };

__device__ __forceinline__ float s_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kSThreads, 1)
k_shade_tc(const TcShade S, const float* __restrict__ rec, const float* __restrict__ feat, const float* __restrict__ r_images,
           const uint32_t* __restrict__ M_dev, uint32_t M_host, const ShadeOutDev O) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_op = smem + S.res_bytes_al;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_op + 2 * kSOperand);
    uint64_t* w_full = bars;
    uint64_t* acc_ready = bars + 1;
    uint64_t* a_ready = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t M = M_dev ? *M_dev : M_host;
    const uint32_t n_tiles = (M + 127) / 128;
    if (blockIdx.x >= n_tiles) return;             // nothing to do for this CTA (tail iterations of the render loop)
    const bool do_renv = (r_images != nullptr) && S.net_layers[2] > 0;
    const int n_nets = do_renv ? 4 : 2;            // diffuse, color, [renv, color again]
    const float* s_f = reinterpret_cast<const float*>(s_w + S.float_off);     // biases: [net 0..2][layer][64]

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
    if (warp == 0 && lane == 0) {
        tc::mbar_arrive_expect_tx(w_full, S.res_bytes);
        for (uint32_t o = 0; o < S.res_bytes; o += 16384) tc::bulk_g2s(s_w + o, S.blob + o, min(16384u, S.res_bytes - o), w_full);
    }
    tc::mbar_wait(w_full, 0);

    if (warp == 0) {
        // ===================== MMA issuer =====================
        uint32_t a_par = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            uint32_t st = 0;
            for (int net = 0; net < n_nets; net++) {
                const int w = (net == 3) ? 1 : net;                    // weight set: 0 diffuse, 1 color, 2 renv
                for (uint32_t l = 0; l < S.net_layers[w]; l++, st++) {
                    const TcImg& I = S.img[w][l];
                    tc::mbar_wait(a_ready, a_par); a_par ^= 1;
                    tc::tc_fence_after();
                    if (lane == 0) {
                        const uint32_t idesc = tc::make_idesc_f16(128, I.Np);
                        const uint32_t a_hi0 = tc::smem_u32(s_op + (st & 1) * kSOperand), a_lo0 = a_hi0 + kSOperandHalf;
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
        }
    } else if (warp >= 4) {
        // ===================== workers =====================
        const uint32_t quarter = warp & 3, g = (warp - 4) >> 2;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        const int Gd = (int)S.geo_dim, Ed = (int)S.env_dim;
        uint32_t acc_par = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t m = tile * 128 + row;
            const bool valid = m < M;
            const float* q = rec + (size_t)min(m, M - 1) * kTcRecFloats;
            const float* ft = feat + (size_t)min(m, M - 1) * kTcRecFloats;
            float outv[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};      // raw heads: diffuse, color, color(renv)
            float fe[16];
            float rr = 0.f, vis = 0.f;
            uint32_t st = 0;
            for (int net = 0; net < n_nets; net++) {
                const int w = (net == 3) ? 1 : net;
                const uint32_t nl = S.net_layers[w];
                // ---- assemble this net's input row (K <= 32) into the operand buffer of its first stage -----------------
                if (g == 0) {
                    float in[32];
                    #pragma unroll
                    for (int i = 0; i < 32; i++) in[i] = 0.f;
                    if (net == 0) {
                        for (int i = 0; i < Gd; i++) in[i] = q[i];
                        for (int i = 0; i < Ed; i++) in[Gd + i] = ft[i];
                    } else if (net == 1 || net == 3) {
                        for (int i = 0; i < Gd; i++) in[i] = q[i];
                        in[Gd] = q[16]; in[Gd + 1] = q[17]; in[Gd + 2] = q[18];
                        for (int i = 0; i < Ed; i++) in[Gd + 3 + i] = (net == 1) ? ft[16 + i] : fe[i];
                        in[Gd + 3 + Ed] = q[19];
                    } else {
                        float4 ri = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid) ri = *reinterpret_cast<const float4*>(r_images + 4 * (size_t)m);
                        vis = ri.w;
                        rr = sqrtf(q[20] / S.rough_scale / 0.75f);
                        in[0] = ri.x * vis; in[1] = ri.y * vis; in[2] = ri.z * vis; in[3] = rr;
                    }
                    uint8_t* dst = s_op + (st & 1) * kSOperand;
                    const uint32_t Kp = S.img[w][0].Kp;
                    #pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if ((uint32_t)c * 8 < Kp) {
                            float v8[8];
                            #pragma unroll
                            for (int e = 0; e < 8; e++) v8[e] = in[c * 8 + e];
                            tc::store_chunk8(dst, dst + kSOperandHalf, row, c * 8, v8);
                        }
                    }
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(a_ready);
                // ---- layers ------------------------------------------------------------------------------------------------
                for (uint32_t l = 0; l < nl; l++, st++) {
                    tc::mbar_wait(acc_ready, acc_par); acc_par ^= 1;
                    tc::tc_fence_after();
                    const float* bias = s_f + (w * 4 + l) * 64;
                    if (l + 1 < nl) {
                        const uint32_t chunks = S.img[w][l].N / 32;
                        uint8_t* dst = s_op + ((st + 1) & 1) * kSOperand;
                        if (g < chunks) tc::hidden_epilogue32(tmem + lane_addr + g * 32, bias + g * 32, dst, dst + kSOperandHalf, row, g * 32);
                        tc::tc_fence_before();
                        tc::fence_proxy_async_smem();
                        tc::mbar_arrive(a_ready);
                    } else {
                        if (g == 0) {
                            uint32_t r[16];
                            tc::tmem_ld16(tmem + lane_addr, r);
                            tc::tmem_ld_wait();
                            if (net == 2) {
                                float ss = 0.f;
                                #pragma unroll
                                for (int i = 0; i < 16; i++) { fe[i] = (i < Ed) ? __uint_as_float(r[i]) + bias[i] : 0.f; ss += fe[i] * fe[i]; }
                                const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                                #pragma unroll
                                for (int i = 0; i < 16; i++) fe[i] *= inv;
                            } else {
                                const int slot = net == 0 ? 0 : (net == 1 ? 1 : 2);
                                #pragma unroll
                                for (int i = 0; i < 3; i++) outv[slot][i] = __uint_as_float(r[i]) + bias[i];
                            }
                        }
                        tc::tc_fence_before();
                        // every worker warp must be past its TMEM read before the next net's first MMA may overwrite D:
                        // that MMA waits for a_ready, which all 256 workers arrive on after this point (program order)
                    }
                }
            }
            if (g == 0 && valid) {
                float cd[3], cs[3];
                #pragma unroll
                for (int i = 0; i < 3; i++) { cd[i] = s_sigmoid(outv[0][i]); cs[i] = s_sigmoid(outv[1][i]); }
                if (do_renv && q[20] < S.indir_rough_thresh && vis > 0.9f) {
                    const float bw = S.learn_blend ? 0.98f * q[21] : 0.95f * s_sigmoid(80.0f * (rr - 0.18f));
                    #pragma unroll
                    for (int i = 0; i < 3; i++) cs[i] = cs[i] * bw + s_sigmoid(outv[2][i]) * (1 - bw);
                }
                #pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (O.rgb) O.rgb[3 * (size_t)m + i] = (cd[i] + cs[i]) * S.intensity_scale;
                    if (O.c_diffuse) O.c_diffuse[3 * (size_t)m + i] = cd[i];
                    if (O.c_specular) O.c_specular[3 * (size_t)m + i] = cs[i];
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 64);
}

__global__ void k_pack_tc3(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np);
__global__ void k_pack_floats3(const float* __restrict__ src, float* __restrict__ dst, uint32_t n, uint32_t n_pad);

__global__ void k_pack_tc3(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np) {
    const uint32_t total = Kp * Np;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t nn = i / Kp, k = i - nn * Kp;
        const float v = (nn < N && k < K) ? W[(size_t)nn * K + k] : 0.0f;
        __half h, lo;
        tc::split_f16(v, h, lo);
        const uint32_t s = k >> 4, kk = k & 15;
        const size_t base = (size_t)s * Np * 64 + (kk >> 3) * (Np * 16) + nn * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + base) = h;
        *reinterpret_cast<__half*>(img + base + (size_t)Np * 32) = lo;
    }
}
__global__ void k_pack_floats3(const float* __restrict__ src, float* __restrict__ dst, uint32_t n, uint32_t n_pad) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) dst[i] = (src && i < n) ? src[i] : 0.0f;
}

static uint32_t rup3(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

bool shade_tc_layout(const envidr_field* f, uint64_t base_bytes, TcShade* out, uint64_t* total_bytes) {
    TcShade& s = *out;
    s = TcShade{};
    const envidr_mlp_layer* nets[3] = {f->diffuse, f->color, f->renv};
    const uint32_t nl[3] = {f->n_diffuse, f->n_color, f->n_renv};
    const uint32_t G = f->geo_feat_dim, E = f->env[f->n_env - 1].out_dim;
    if (G > 13 || E > 16 || G + 4 + E > 32) return false;
    uint32_t off = 0;
    for (int w = 0; w < 3; w++) {
        if (nl[w] > 4 || (w < 2 && nl[w] < 1)) return false;
        for (uint32_t l = 0; l < nl[w]; l++) {
            const bool last = (l + 1 == nl[w]);
            const uint32_t K = nets[w][l].in_dim, N = nets[w][l].out_dim;
            if (l == 0 && K > 32) return false;
            if (!last && !(N == 32 || N == 64)) return false;
            if (last && N > 16) return false;
            if (l > 0 && K != nets[w][l - 1].out_dim) return false;
            TcImg& I = s.img[w][l];
            I.Kp = rup3(K, 16); I.N = N; I.Np = last ? 16 : N; I.off = off;
            off += (I.Kp / 16) * I.Np * 64;
        }
        s.net_layers[w] = nl[w];
    }
    if (nl[0] && nets[0][0].in_dim != G + E) return false;
    if (nets[1][0].in_dim != G + 3 + E + 1) return false;
    if (nl[2] && (nets[2][0].in_dim != 4 || nets[2][nl[2] - 1].out_dim != E)) return false;
    s.float_off = off;
    off += 3 * 4 * 64 * 4;
    s.res_bytes = off;
    s.res_bytes_al = rup3(off, 1024);
    const uint64_t start = rup3((uint32_t)base_bytes, 1024);
    if (f->packed) s.blob = reinterpret_cast<const uint8_t*>(f->packed) + start;
    s.blob_off = start;
    s.geo_dim = G; s.env_dim = E;
    s.rough_scale = f->roughness_scale; s.indir_rough_thresh = f->indir_roughness_thresh; s.learn_blend = f->learn_indir_blend;
    s.intensity_scale = f->intensity_scale;
    *total_bytes = start + s.res_bytes_al;
    return true;
}

int shade_tc_pack(const envidr_field* f, const TcShade& s, void* packed, cudaStream_t st) {
    uint8_t* blob = reinterpret_cast<uint8_t*>(packed) + s.blob_off;
    const envidr_mlp_layer* nets[3] = {f->diffuse, f->color, f->renv};
    float* fl = reinterpret_cast<float*>(blob + s.float_off);
    for (int w = 0; w < 3; w++)
        for (uint32_t l = 0; l < s.net_layers[w]; l++) {
            const TcImg& I = s.img[w][l];
            k_pack_tc3<<<32, 256, 0, st>>>(nets[w][l].weight, blob + I.off, nets[w][l].in_dim, nets[w][l].out_dim, I.Kp, I.Np);
            k_pack_floats3<<<1, 64, 0, st>>>(nets[w][l].bias, fl + (w * 4 + l) * 64, nets[w][l].out_dim, 64);
        }
    return check_launch("shade_tc_pack");
}

int shade_tc_launch(const TcShade& s, const float* rec, const float* feat, const float* r_images, const uint32_t* M_dev, uint32_t M_host,
                    const envidr_field_out* out, cudaStream_t st) {
    const size_t smem = (size_t)s.res_bytes_al + 2 * kSOperand + 64;
    static size_t attr_set = 0;
    if (attr_set < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_shade_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("shade_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = smem;
    }
    uint32_t grid = kSMs;
    if (!M_dev) grid = min((uint32_t)kSMs, (M_host + 127) / 128);
    if (grid == 0) return 0;
    ShadeOutDev O{out->rgb, out->c_diffuse, out->c_specular};
    k_shade_tc<<<grid, kSThreads, smem, st>>>(s, rec, feat, r_images, M_dev, M_host, O);
    return check_launch("shade_tc");
}

}  // namespace envidr
