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

constexpr int kSThreads = 576;                     // warps 0-15 chain (4 groups x 4), 16 MMA issuer, 17 TMEM / weights
constexpr int kSGroups = 4;                        // tiles in flight per CTA
constexpr uint32_t kSOperand = 32768;              // one A-operand buffer per group: 128 rows x 64 K x 2 B x (hi, lo), rewritten in place
constexpr uint32_t kSOperandHalf = 16384;

struct ShadeOutDev { float *rgb, *c_diffuse, *c_specular; };

__device__ __forceinline__ float s_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// Operand column layout of the first layers (the weight images are packed with the same permutation, see k_pack_tc3):
//   diffuse_net : [ geo 0..11 | 0 0 0 0 | f_n 0..15 ]
//   color_net   : [ geo 0..11 | n.x n.y n.z n.w_o | f_r (or f_e) 0..15 ]
//   renv_net    : [ r*vis (3) rho | 0 ... ]   (K = 16)
__global__ void __launch_bounds__(kSThreads, 1)
k_shade_tc(const TcShade S, const float* __restrict__ rec, const float* __restrict__ feat, const float* __restrict__ r_images,
           const uint32_t* __restrict__ M_dev, uint32_t M_host, const ShadeOutDev O, const int32_t* __restrict__ ridx) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_op = smem + S.res_bytes_al;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_op + kSGroups * kSOperand);
    uint64_t* w_full = bars;
    uint64_t* acc_ready = bars + 1;                // [4] issuer -> chain group
    uint64_t* a_ready = bars + 1 + kSGroups;       // [4] chain group -> issuer (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kSGroups);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: role branches stay uniform (UR datapath)
    const uint32_t M = M_dev ? *M_dev : M_host;
    const uint32_t n_tiles = (M + 127) / 128;
    if (blockIdx.x >= n_tiles) return;             // nothing to do for this CTA (tail iterations of the render loop)
    const uint32_t T = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;       // tiles of this CTA: blockIdx.x + j * gridDim.x
    const bool do_renv = (r_images != nullptr) && S.net_layers[2] > 0;
    const int n_nets = do_renv ? 4 : 2;            // diffuse, color, [renv, color again]
    const float* s_f = reinterpret_cast<const float*>(s_w + S.float_off);     // biases: [net 0..2][layer][64]

    if (tid == 0) {
        tc::mbar_init(w_full, 1);
        for (int i = 0; i < kSGroups; i++) { tc::mbar_init(acc_ready + i, 1); tc::mbar_init(a_ready + i, 128); }
        tc::mbar_fence_init();
    }
    if (warp == 17) tc::tmem_alloc(tmem_slot, 64 * kSGroups);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 17 && lane == 0) {
        tc::mbar_arrive_expect_tx(w_full, S.res_bytes);
        for (uint32_t o = 0; o < S.res_bytes; o += 16384) tc::bulk_g2s(s_w + o, S.blob + o, min(16384u, S.res_bytes - o), w_full);
    }

    if (warp == 16) {
        // ===================== MMA issuer: up to 4 tiles in flight, stages interleaved round-robin =====================
        tc::mbar_wait(w_full, 0);
        uint32_t a_par = 0;                        // bit u = parity of the next a_ready[u] phase
        for (uint32_t p = 0; p < T; p += kSGroups) {
            const uint32_t ntp = min((uint32_t)kSGroups, T - p);
            for (int net = 0; net < n_nets; net++) {
                const int w = (net == 3) ? 1 : net;                    // weight set: 0 diffuse, 1 color, 2 renv
                for (uint32_t l = 0; l < S.net_layers[w]; l++) {
                    const TcImg& I = S.img[w][l];
                    for (uint32_t u = 0; u < ntp; u++) {
                        tc::mbar_wait(a_ready + u, (a_par >> u) & 1u); a_par ^= 1u << u;
                        tc::tc_fence_after();
                        __syncwarp();
                        {
                            const uint32_t idesc = tc::make_idesc_f16(128, I.Np);
                            const uint32_t a_hi0 = tc::smem_u32(s_op + u * kSOperand), a_lo0 = a_hi0 + kSOperandHalf;
                            const uint32_t d = tmem + u * 64;
                            uint64_t da_hi = tc::make_smem_desc(a_hi0, 2048, 128), da_lo = tc::make_smem_desc(a_lo0, 2048, 128);
                            uint64_t db_hi = tc::make_smem_desc(tc::smem_u32(s_w + I.off), I.Np * 16, 128);
                            for (uint32_t s = 0; s < I.Kp / 16; s++) {
                                const uint64_t db_lo = tc::desc_advance(db_hi, I.Np * 32);
                                tc::mma_f16_ss_w(d, da_hi, db_hi, idesc, s > 0);
                                tc::mma_f16_ss_w(d, da_lo, db_hi, idesc, 1);
                                tc::mma_f16_ss_w(d, da_hi, db_lo, idesc, 1);
                                da_hi = tc::desc_advance(da_hi, 4096); da_lo = tc::desc_advance(da_lo, 4096);
                                db_hi = tc::desc_advance(db_hi, I.Np * 64);
                            }
                            tc::mma_commit_w(acc_ready + u);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp < 16) {
        // ===================== chain: group u = warp / 4 owns every 4th tile; thread = accumulator row =====================
        const uint32_t u = warp >> 2, quarter = warp & 3;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t acc_t = tmem + ((quarter * 32u) << 16) + u * 64;
        uint8_t* op = s_op + u * kSOperand;
        const int Gd = (int)S.geo_dim, Ed = (int)S.env_dim;
        tc::mbar_wait(w_full, 0);
        uint32_t acc_par = 0;
        for (uint32_t j = u; j < T; j += kSGroups) {
            const uint32_t tile = blockIdx.x + j * gridDim.x;
            const uint32_t m = tile * 128 + row;
            const bool valid = m < M;
            const uint32_t mc = min(m, M - 1);
            const float4* q4 = reinterpret_cast<const float4*>(rec + (size_t)(ridx ? (uint32_t)ridx[mc] : mc) * kTcRecFloats);
            const float4* f4 = reinterpret_cast<const float4*>(feat + (size_t)min(m, M - 1) * kTcRecFloats);
            float geo[16];                         // geo 0..11, then n.xyz, n.w_o
            {
                const float4 a = __ldg(q4), b = __ldg(q4 + 1), c = __ldg(q4 + 2), d = __ldg(q4 + 4);
                const float t[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
                #pragma unroll
                for (int i = 0; i < 12; i++) geo[i] = (i < Gd) ? t[i] : 0.f;
                geo[12] = d.x; geo[13] = d.y; geo[14] = d.z; geo[15] = d.w;
            }
            const float4 q5 = __ldg(q4 + 5);       // rough, blend, ...
            const float rough = q5.x, blend = q5.y;
            float outv[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};      // raw heads: diffuse, color, color(renv)
            float fe[16];
            #pragma unroll
            for (int i = 0; i < 16; i++) fe[i] = 0.f;
            float rr = 0.f, vis = 0.f;
            for (int net = 0; net < n_nets; net++) {
                const int w = (net == 3) ? 1 : net;
                const uint32_t nl = S.net_layers[w];
                // ---- assemble this net's input row into the operand buffer (the previous MMA reading it has completed) ---------
                if (net == 2) {
                    float4 ri = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) ri = __ldg(reinterpret_cast<const float4*>(r_images) + m);
                    vis = ri.w;
                    rr = sqrtf(rough / S.rough_scale / 0.75f);
                    const float v0[8] = {ri.x * vis, ri.y * vis, ri.z * vis, rr, 0.f, 0.f, 0.f, 0.f};
                    const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    tc::store_chunk8(op, op + kSOperandHalf, row, 0, v0);
                    tc::store_chunk8(op, op + kSOperandHalf, row, 8, z8);
                } else {
                    float v0[8], v1[8];
                    #pragma unroll
                    for (int i = 0; i < 8; i++) { v0[i] = geo[i]; v1[i] = (net == 0 && i >= 4) ? 0.f : geo[8 + i]; }
                    tc::store_chunk8(op, op + kSOperandHalf, row, 0, v0);
                    tc::store_chunk8(op, op + kSOperandHalf, row, 8, v1);
                    float f[16];
                    if (net == 3) {
                        #pragma unroll
                        for (int i = 0; i < 16; i++) f[i] = fe[i];
                    } else {
                        const float4* src = f4 + (net == 1 ? 4 : 0);           // f_r at float 16, f_n at float 0
                        #pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const float4 t = __ldg(src + i);
                            f[4 * i] = t.x; f[4 * i + 1] = t.y; f[4 * i + 2] = t.z; f[4 * i + 3] = t.w;
                        }
                        #pragma unroll
                        for (int i = 0; i < 16; i++) f[i] = (i < Ed) ? f[i] : 0.f;
                    }
                    float v2[8], v3[8];
                    #pragma unroll
                    for (int i = 0; i < 8; i++) { v2[i] = f[i]; v3[i] = f[8 + i]; }
                    tc::store_chunk8(op, op + kSOperandHalf, row, 16, v2);
                    tc::store_chunk8(op, op + kSOperandHalf, row, 24, v3);
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(a_ready + u);
                // ---- layers ------------------------------------------------------------------------------------------------
                for (uint32_t l = 0; l < nl; l++) {
                    tc::mbar_wait(acc_ready + u, acc_par); acc_par ^= 1;
                    tc::tc_fence_after();
                    const float* bias = s_f + (w * 4 + l) * 64;
                    if (l + 1 < nl) {
                        const uint32_t chunks = S.img[w][l].N / 32;
                        #pragma unroll
                        for (uint32_t ch = 0; ch < 2; ch++)
                            if (ch < chunks) tc::hidden_epilogue32(acc_t + ch * 32, bias + ch * 32, op, op + kSOperandHalf, row, ch * 32);
                        tc::tc_fence_before();
                        tc::fence_proxy_async_smem();
                        tc::mbar_arrive(a_ready + u);
                    } else {
                        uint32_t r[16];
                        tc::tmem_ld16(acc_t, r);
                        tc::tmem_ld_wait();
                        if (net == 2) {
                            float ss = 0.f;
                            #pragma unroll
                            for (int i = 0; i < 16; i++) { fe[i] = (i < Ed) ? __uint_as_float(r[i]) + bias[i] : 0.f; ss += fe[i] * fe[i]; }
                            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                            #pragma unroll
                            for (int i = 0; i < 16; i++) fe[i] *= inv;
                        } else {
                            #pragma unroll
                            for (int i = 0; i < 3; i++) {
                                const float v = __uint_as_float(r[i]) + bias[i];
                                if (net == 0) outv[0][i] = v; else if (net == 1) outv[1][i] = v; else outv[2][i] = v;
                            }
                        }
                        tc::tc_fence_before();     // the next MMA into this accumulator waits for this thread's next a_ready arrival
                    }
                }
            }
            if (valid) {
                float cd[3], cs[3];
                #pragma unroll
                for (int i = 0; i < 3; i++) { cd[i] = s_sigmoid(outv[0][i]); cs[i] = s_sigmoid(outv[1][i]); }
                if (do_renv && rough < S.indir_rough_thresh && vis > 0.9f) {
                    const float bw = S.learn_blend ? 0.98f * blend : 0.95f * s_sigmoid(80.0f * (rr - 0.18f));
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
    if (warp == 17) tc::tmem_dealloc(tmem, 64 * kSGroups);
}

__global__ void k_pack_tc3(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np, int kmap,
                           uint32_t Gd, uint32_t Ed);
__global__ void k_pack_floats3(const float* __restrict__ src, float* __restrict__ dst, uint32_t n, uint32_t n_pad);

// kmap: 0 identity; 1 diffuse_net layer 0 ([geo | f] -> geo at 0.., f at 16..); 2 color_net layer 0
// ([geo | n | f | n.w_o] -> geo at 0.., n at 12..14, n.w_o at 15, f at 16..).  k below is the operand column; src the weight column.
__global__ void k_pack_tc3(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np, int kmap,
                           uint32_t Gd, uint32_t Ed) {
    const uint32_t total = Kp * Np;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t nn = i / Kp, k = i - nn * Kp;
        int src = (int)k;
        if (kmap == 1) {
            src = k < 12 ? (k < Gd ? (int)k : -1) : (k < 16 ? -1 : (k - 16 < Ed ? (int)(Gd + k - 16) : -1));
        } else if (kmap == 2) {
            src = k < 12 ? (k < Gd ? (int)k : -1) : (k < 15 ? (int)(Gd + k - 12) : (k == 15 ? (int)(Gd + 3 + Ed) : (k - 16 < Ed ? (int)(Gd + 3 + k - 16) : -1)));
        }
        const float v = (nn < N && src >= 0 && (uint32_t)src < K) ? W[(size_t)nn * K + src] : 0.0f;
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
    if (G > 12 || E > 16) return false;
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
            I.Kp = (l == 0 && w < 2) ? 32 : rup3(K, 16); I.N = N; I.Np = last ? 16 : N; I.off = off;
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
            k_pack_tc3<<<32, 256, 0, st>>>(nets[w][l].weight, blob + I.off, nets[w][l].in_dim, nets[w][l].out_dim, I.Kp, I.Np,
                                       (l == 0 && w < 2) ? w + 1 : 0, s.geo_dim, s.env_dim);
            k_pack_floats3<<<1, 64, 0, st>>>(nets[w][l].bias, fl + (w * 4 + l) * 64, nets[w][l].out_dim, 64);
        }
    return check_launch("shade_tc_pack");
}

int shade_tc_launch(const TcShade& s, const float* rec, const float* feat, const float* r_images, const uint32_t* M_dev, uint32_t M_host,
                    const envidr_field_out* out, cudaStream_t st, const int32_t* ridx) {
    const size_t smem = (size_t)s.res_bytes_al + kSGroups * kSOperand + 128;
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
    k_shade_tc<<<grid, kSThreads, smem, st>>>(s, rec, feat, r_images, M_dev, M_host, O, ridx);
    return check_launch("shade_tc");
}

}  // namespace envidr
