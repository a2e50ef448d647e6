// env_train_tc.cu -- env_net of the TRAINING branch as ONE forward and ONE backward kernel.
//
// Reference: get_color_mlp_extra_params / forward_color evaluate env_net twice per sample (IDE(n, 0.64) and IDE(w_r, roughness),
// nerf/network.py:527-541, 589-607) as 8 cuBLAS fp32 GEMMs + ~100 elementwise kernels, and autograd replays as many in the backward.
// Round 1 / 2 ran every dense layer on tcgen05 but as one launch per layer and direction (k_linear_tc x 8, k_ide_fwd x 2, normalize,
// threshold_backward x 3, ...), activations round-tripping HBM between the launches.  Here:
//   forward  = k_env_tc<SAVE> (field_tc.cu): IDE -> all layers -> unit-norm feature in one kernel, activations stay on chip; what the
//              backward needs goes to HBM once: post-ReLU activations (fp32, the operand of the weight-gradient GEMMs), ReLU bit masks,
//              the inverse norm of the raw feature;
//   backward = k_env_bwd_tc (this file): backward of F.normalize -> data-gradient chain through all layers (tcgen05, transposed weight
//              images streamed through the same 3-stage ring, ReLU masks applied in the epilogue, operands kept in shared memory) ->
//              gradient w.r.t. the IDE features.  It writes the pre-activation gradients of every layer (fp32) for the weight-gradient
//              GEMMs (k_wgrad_tc, one launch per layer: they contract over the SAMPLE dimension and share nothing with this chain).
// Gradients of a loss sit around 1e-7, below fp16's normal range: every row is pre-scaled by a power of two that brings its largest
// |d y| into [1, 2) and un-scaled on the way out (exact); operands are split fp16 hi/lo as everywhere else (three MMAs per K step).
#include <math.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "field_tc.cuh"

namespace envidr {

constexpr int kBwThreads = 12 * 32;               // warp 0 producer, 1 issuer, 2 TMEM, 3 idle, 4-11 epilogue (2 groups x 4 lane quarters)
constexpr int kBwStages = 3;
constexpr uint32_t kBwStageBytes = 16384;
constexpr uint32_t kBwARegion = 65536;            // 128 rows x 256 K x 2 B
constexpr uint32_t kBwGRegion = 16384;            // 128 rows x 64 K x 2 B: operand of the first GEMM (d y: K = 16; generic forward input: K <= 64)

struct BwLayer { uint32_t Kp, Np, img_off; };      // B operand image: [Np rows = inputs of the forward layer][Kp = its outputs]
struct EnvBwd {
    const uint8_t* blob;
    uint32_t n_layers;                             // chain length = forward layers; chain layer j = transpose of forward layer n-1-j
    BwLayer L[4];
    uint32_t E;                                    // env feature width (<= 12)
    const float* gfeat;                            // [M, 32] gradient w.r.t. the unit-norm features (normal branch at 0.., reflected at 16..)
    const float* feat;                             // [M, 32] forward output (inverse norm of the raw feature in slots 12 / 28)
    const uint32_t* mask[3];                       // ReLU masks of forward hidden layer l, [2M, N_l / 32]
    float* gact[3];                                // out: gradient w.r.t. the PRE-activation of forward hidden layer l, [2M, N_l]
    float* gy;                                     // out: gradient w.r.t. the raw feature, [2M, 16] (zero padded)
    float* gx0;                                    // out: gradient w.r.t. the layer-0 input in ITS column order, [2M, L[n-1].Np]
    uint32_t M;
    // ---- generic chain (MODE 1: backward of a small ReLU MLP, MODE 2: its forward; envidr_mlp_* entries) ----
    const float* in;                               // MODE 1: d Y [rows, in_ld] (in_cols <= 16 used); MODE 2: X [rows, in_ld] (in_cols <= 64 used)
    uint32_t in_ld, in_cols, rows;
    const float* bias[4];                          // MODE 2: bias of forward layer l or NULL
    uint32_t nout[4];                              // MODE 2: true output width of forward layer l
    uint32_t* mask_out[3];                         // MODE 2: ReLU masks of hidden layer l, [rows, N_l / 32]
    float* act_out[3];                             // MODE 2: post-ReLU activations [rows, N_l] or NULL (only needed when the layer above trains)
    float* y;                                      // MODE 2: output [rows, 16], zero padded
};

// MODE 0: backward of env_net (input = gradient of the unit-norm features, rows = the [2M] batch of the forward kernel)
// MODE 1: backward of a generic small MLP (input = d Y rows);  MODE 2: forward of a generic small MLP (bias + ReLU epilogues, masks saved)
template <int MODE>
__global__ void __launch_bounds__(kBwThreads, 1)
k_chain_tc(const EnvBwd B) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA_hi = smem;
    uint8_t* sA_lo = smem + kBwARegion;
    uint8_t* sG_hi = smem + 2 * kBwARegion;
    uint8_t* sG_lo = sG_hi + kBwGRegion;
    uint8_t* ring = sG_lo + kBwGRegion;
    float* s_inv = reinterpret_cast<float*>(ring + kBwStages * kBwStageBytes);       // [128] 1 / row scale
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 128);
    uint64_t* full = bars;                        // [3]
    uint64_t* empty = bars + kBwStages;           // [3]
    uint64_t* acc_ready = bars + 2 * kBwStages;   // [2]
    uint64_t* g_full = acc_ready + 2;             // d y operand of this tile is in shared memory (128 arrivals)
    uint64_t* a_rdy = acc_ready + 3;              // [8] 32-column chunk c of the next A operand (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_rdy + 8);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t rows_total = MODE == 0 ? 2 * B.M : B.rows;
    const uint32_t n_tiles = (rows_total + 127) / 128;
    const int nl = (int)B.n_layers;
    if (blockIdx.x >= n_tiles) return;

    if (tid == 0) {
        for (int i = 0; i < kBwStages; i++) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&acc_ready[0], 1);
        tc::mbar_init(&acc_ready[1], 1);
        tc::mbar_init(g_full, 128);
        for (int i = 0; i < 8; i++) tc::mbar_init(&a_rdy[i], 128);
        tc::mbar_fence_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: transposed weight images, layer after layer, tile after tile =====================
        uint32_t stage = 0, phase = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int j = 0; j < nl; j++) {
                const uint32_t ksteps = B.L[j].Kp / 16, kbytes = B.L[j].Np * 64;
                const uint32_t kper = max(1u, kBwStageBytes / kbytes);
                const uint8_t* src = B.blob + B.L[j].img_off;
                for (uint32_t s = 0; s < ksteps; s += kper) {
                    const uint32_t bytes = min(kper, ksteps - s) * kbytes;
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    if (lane == 0) {
                        tc::mbar_arrive_expect_tx(&full[stage], bytes);
                        tc::bulk_g2s(ring + stage * kBwStageBytes, src + (size_t)s * kbytes, bytes, &full[stage]);
                    }
                    __syncwarp();
                    if (++stage == kBwStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t stage = 0, phase = 0, g_par = 0, chunk_par = 0, gl = 0;
        const uint32_t ring0 = tc::smem_u32(ring);
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int j = 0; j < nl; j++) {
                const uint32_t ksteps = B.L[j].Kp / 16, Np = B.L[j].Np;
                const uint32_t idesc = tc::make_idesc_f16(128, Np);
                const uint32_t d_tmem = tmem + (gl & 1u) * 256u;
                const uint32_t buf = gl & 1u;
                gl++;
                uint64_t da_hi, da_lo;
                if (j == 0) {
                    tc::mbar_wait(g_full, g_par); g_par ^= 1;
                    da_hi = tc::make_smem_desc(tc::smem_u32(sG_hi), 2048, 128);
                    da_lo = tc::make_smem_desc(tc::smem_u32(sG_lo), 2048, 128);
                } else {
                    da_hi = tc::make_smem_desc(tc::smem_u32(sA_hi), 2048, 128);
                    da_lo = tc::make_smem_desc(tc::smem_u32(sA_lo), 2048, 128);
                }
                const uint64_t db0 = tc::make_smem_desc(ring0, Np * 16, 128);
                const uint32_t lo_off = Np * 32, kbytes = Np * 64;
                const uint32_t kper = max(1u, kBwStageBytes / kbytes);
                for (uint32_t s0 = 0; s0 < ksteps; s0 += kper) {
                    tc::mbar_wait(&full[stage], phase);
                    const uint32_t kend = min(ksteps, s0 + kper);
                    uint64_t db_hi = tc::desc_advance(db0, stage * kBwStageBytes);
                    for (uint32_t s = s0; s < kend; s++) {
                        if (j > 0 && (s & 1u) == 0) {                        // K steps 2c, 2c+1 read chunk c of the A operand
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
                    if (++stage == kBwStages) { stage = 0; phase ^= 1; }
                }
                tc::mma_commit_w(&acc_ready[buf]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps (group 0 also prepares d y at the start of a tile) =====================
        const uint32_t quarter = warp & 3, g = (warp - 4) >> 2;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        uint32_t acc_par = 0, gl = 0;
        const int Ef = (int)B.E;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t R = tile * 128 + row;                  // row of the [2M] batch: branch * M + sample
            const bool valid = R < rows_total;
            const uint32_t branch = (valid && R >= B.M) ? 1u : 0u, m = valid ? R - branch * B.M : 0u;
            if (g == 0 && MODE == 2) {
                // ---- generic forward: the row of X as the first operand (K = L[0].Kp <= 64, zero padded), no scaling
                const uint32_t K0 = B.L[0].Kp;
                const float* xr = B.in + (size_t)R * B.in_ld;
                for (uint32_t k0 = 0; k0 < K0; k0 += 8) {
                    float c[8];
                    #pragma unroll
                    for (int i = 0; i < 8; i++) c[i] = (valid && k0 + i < B.in_cols) ? __ldg(xr + k0 + i) : 0.f;
                    tc::store_chunk8(sG_hi, sG_lo, row, k0, c);
                }
                s_inv[row] = 1.0f;
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(g_full);
            } else if (g == 0) {
                // ---- backward of F.normalize(y): d y = (g - f (f . g)) / |y|, per-row power-of-two scale, fp16 hi/lo operand (K = 16)
                float gyv[16];
                #pragma unroll
                for (int i = 0; i < 16; i++) gyv[i] = 0.f;
                float inv_scale = 1.0f;
                if (valid && MODE == 1) {
                    // generic backward: the row of d Y as it is
                    const float* gr = B.in + (size_t)R * B.in_ld;
                    float mx = 0.f;
                    #pragma unroll
                    for (int i = 0; i < 16; i++) {
                        gyv[i] = (uint32_t)i < B.in_cols ? __ldg(gr + i) : 0.f;
                        mx = fmaxf(mx, fabsf(gyv[i]));
                    }
                    if (mx > 0.f && mx < 3.0e38f) {
                        int ex;
                        frexpf(mx, &ex);
                        ex = max(-120, min(120, ex - 1));
                        const float sc = ldexpf(1.0f, -ex);
                        inv_scale = ldexpf(1.0f, ex);
                        #pragma unroll
                        for (int i = 0; i < 16; i++) gyv[i] *= sc;
                    }
                } else if (valid) {
                    const float4* gp = reinterpret_cast<const float4*>(B.gfeat + (size_t)m * kTcRecFloats + 16 * branch);
                    const float4* fp = reinterpret_cast<const float4*>(B.feat + (size_t)m * kTcRecFloats + 16 * branch);
                    float gv[16], fv[16];
                    #pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float4 a = __ldg(gp + i), b = __ldg(fp + i);
                        gv[4 * i] = a.x; gv[4 * i + 1] = a.y; gv[4 * i + 2] = a.z; gv[4 * i + 3] = a.w;
                        fv[4 * i] = b.x; fv[4 * i + 1] = b.y; fv[4 * i + 2] = b.z; fv[4 * i + 3] = b.w;
                    }
                    const float inv_norm = fv[12];
                    float dot = 0.f;
                    #pragma unroll
                    for (int i = 0; i < 16; i++) if (i < Ef) dot += fv[i] * gv[i];
                    float mx = 0.f;
                    #pragma unroll
                    for (int i = 0; i < 16; i++) {
                        gyv[i] = (i < Ef) ? (gv[i] - fv[i] * dot) * inv_norm : 0.f;
                        mx = fmaxf(mx, fabsf(gyv[i]));
                    }
                    float4* go = reinterpret_cast<float4*>(B.gy + (size_t)R * 16);
                    #pragma unroll
                    for (int i = 0; i < 4; i++) go[i] = make_float4(gyv[4 * i], gyv[4 * i + 1], gyv[4 * i + 2], gyv[4 * i + 3]);
                    if (mx > 0.f && mx < 3.0e38f) {
                        int ex;
                        frexpf(mx, &ex);                                   // mx = f * 2^ex, f in [0.5, 1)
                        ex = max(-120, min(120, ex - 1));
                        const float sc = ldexpf(1.0f, -ex);               // mx * sc in [1, 2)
                        inv_scale = ldexpf(1.0f, ex);
                        #pragma unroll
                        for (int i = 0; i < 16; i++) gyv[i] *= sc;
                    }
                }
                s_inv[row] = inv_scale;
                float c0[8], c1[8];
                #pragma unroll
                for (int i = 0; i < 8; i++) { c0[i] = gyv[i]; c1[i] = gyv[8 + i]; }
                tc::store_chunk8(sG_hi, sG_lo, row, 0, c0);
                tc::store_chunk8(sG_hi, sG_lo, row, 8, c1);
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(g_full);
            }
            // every epilogue thread needs its row's un-scale factor: written by the group-0 thread of the same row
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float inv_scale = s_inv[row];
            for (int j = 0; j < nl; j++) {
                const uint32_t buf = gl & 1u;
                gl++;
                // backward: forward hidden layer whose pre-activation gradient this chain layer produces; forward: the hidden layer itself
                const int fl = MODE == 2 ? (j < nl - 1 ? j : -1) : nl - 2 - j;
                const uint32_t Np = B.L[j].Np;
                uint32_t mk[4] = {0u, 0u, 0u, 0u};
                if (MODE != 2 && fl >= 0 && valid) {              // masks of this thread's chunks, fetched before the accumulator is ready
                    const uint32_t nch = Np / 32;
                    #pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t cb = g + 2 * q;
                        if (cb < nch) mk[q] = __ldg(B.mask[fl] + (size_t)R * nch + cb);
                    }
                }
                tc::mbar_wait(&acc_ready[buf], (acc_par >> buf) & 1u); acc_par ^= 1u << buf;
                tc::tc_fence_after();
                const uint32_t acc = tmem + lane_addr + buf * 256u;
                if (fl >= 0) {
                    const uint32_t nch = Np / 32;
                    #pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t cb = g + 2 * q;
                        if (cb >= nch) break;
                        uint32_t r[32];
                        tc::tmem_ld32(acc + cb * 32, r);
                        tc::tmem_ld_wait();
                        if (MODE == 2) {
                            // forward hidden layer: h = relu(D + b); mask (and, if a layer above trains, h itself) to HBM; next operand
                            const float* bl = B.bias[fl];
                            uint32_t bits2 = 0;
                            #pragma unroll
                            for (int jj = 0; jj < 4; jj++) {
                                float v[8];
                                #pragma unroll
                                for (int e = 0; e < 8; e++) {
                                    const uint32_t c = cb * 32 + 8 * jj + e;
                                    v[e] = fmaxf(__uint_as_float(r[8 * jj + e]) + ((bl && c < B.nout[fl]) ? __ldg(bl + c) : 0.f), 0.f);
                                    bits2 |= (v[e] > 0.f ? 1u : 0u) << (8 * jj + e);
                                }
                                if (valid && B.act_out[fl]) {
                                    float* adst = B.act_out[fl] + (size_t)R * Np + cb * 32 + 8 * jj;
                                    reinterpret_cast<float4*>(adst)[0] = make_float4(v[0], v[1], v[2], v[3]);
                                    reinterpret_cast<float4*>(adst)[1] = make_float4(v[4], v[5], v[6], v[7]);
                                }
                                tc::store_chunk8(sA_hi, sA_lo, row, cb * 32 + jj * 8, v);
                            }
                            if (valid) B.mask_out[fl][(size_t)R * nch + cb] = bits2;
                            tc::tc_fence_before();
                            tc::fence_proxy_async_smem();
                            tc::mbar_arrive(&a_rdy[cb]);
                            continue;
                        }
                        const uint32_t bits = mk[q];
                        float* gdst = B.gact[fl] ? B.gact[fl] + (size_t)R * Np + cb * 32 : nullptr;
                        #pragma unroll
                        for (int jj = 0; jj < 4; jj++) {
                            float v[8];
                            #pragma unroll
                            for (int e = 0; e < 8; e++) v[e] = ((bits >> (8 * jj + e)) & 1u) ? __uint_as_float(r[8 * jj + e]) : 0.f;
                            if (valid && gdst) {
                                reinterpret_cast<float4*>(gdst + 8 * jj)[0] = make_float4(v[0] * inv_scale, v[1] * inv_scale, v[2] * inv_scale, v[3] * inv_scale);
                                reinterpret_cast<float4*>(gdst + 8 * jj)[1] = make_float4(v[4] * inv_scale, v[5] * inv_scale, v[6] * inv_scale, v[7] * inv_scale);
                            }
                            tc::store_chunk8(sA_hi, sA_lo, row, cb * 32 + jj * 8, v);
                        }
                        tc::tc_fence_before();
                        tc::fence_proxy_async_smem();
                        tc::mbar_arrive(&a_rdy[cb]);
                    }
                } else if (MODE == 2) {
                    // last forward layer: y = D + b, 16 columns (zero padded)
                    if (g == 0) {
                        uint32_t r[16];
                        tc::tmem_ld16(acc, r);
                        tc::tmem_ld_wait();
                        if (valid) {
                            const float* bl = B.bias[nl - 1];
                            float o[16];
                            #pragma unroll
                            for (int i = 0; i < 16; i++) o[i] = (uint32_t)i < B.nout[nl - 1] ? __uint_as_float(r[i]) + (bl ? __ldg(bl + i) : 0.f) : 0.f;
                            float4* dst = reinterpret_cast<float4*>(B.y + (size_t)R * 16);
                            #pragma unroll
                            for (int i = 0; i < 4; i++) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                        }
                    }
                    tc::tc_fence_before();
                } else {
                    // last chain layer: gradient w.r.t. the layer-0 input, Np columns in 16-column units (no mask)
                    const uint32_t nu = Np / 16;
                    for (uint32_t u = g; u < nu; u += 2) {
                        uint32_t r[16];
                        tc::tmem_ld16(acc + u * 16, r);
                        tc::tmem_ld_wait();
                        if (valid) {
                            float4* dst = reinterpret_cast<float4*>(B.gx0 + (size_t)R * Np + u * 16);
                            #pragma unroll
                            for (int i = 0; i < 4; i++)
                                dst[i] = make_float4(__uint_as_float(r[4 * i]) * inv_scale, __uint_as_float(r[4 * i + 1]) * inv_scale,
                                                     __uint_as_float(r[4 * i + 2]) * inv_scale, __uint_as_float(r[4 * i + 3]) * inv_scale);
                        }
                    }
                    tc::tc_fence_before();
                }
            }
            // the next tile's d y overwrites sG / s_inv: every thread of this tile is past its reads (s_inv read above; sG was consumed by the
            // first GEMM, whose completion acc_ready signalled long ago)
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem, 512);
}

// transposed operand image of one forward layer W [N_out, K_in] (torch layout): B operand rows n = forward INPUT index (Np = K_in rounded
// up to 16), K = forward OUTPUT index (Kp = N_out rounded up to 16).  `interleave` = P > 0 (forward layer 0): row n < 2P is input column
// (n & 1) * P + (n >> 1), the K order the forward kernel's IDE warps emit ([Re_0, Im_0, Re_1, ...]), so that d x0 comes out in that order too.
// transpose = 0: the plain forward image, rows n = forward OUTPUT index (Np), K = forward INPUT index (Kp).
__global__ void k_pack_tcT(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K_in, uint32_t N_out, uint32_t Np, uint32_t Kp,
                           uint32_t interleave, int transpose = 1) {
    const uint32_t total = Np * Kp;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t n = i / Kp, k = i - n * Kp;
        const uint32_t ns = (interleave && n < 2 * interleave) ? (n & 1u) * interleave + (n >> 1) : n;
        const float v = transpose ? ((ns < K_in && k < N_out) ? W[(size_t)k * K_in + ns] : 0.0f)
                                  : ((n < N_out && k < K_in) ? W[(size_t)n * K_in + k] : 0.0f);
        __half h, lo;
        tc::split_f16(v, h, lo);
        const uint32_t s = k >> 4, kk = k & 15;
        const size_t base = (size_t)s * Np * 64 + (kk >> 3) * (Np * 16) + n * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + base) = h;
        *reinterpret_cast<__half*>(img + base + (size_t)Np * 32) = lo;
    }
}

static uint32_t rup_t(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// host description of the trainable env_net: a throw-away envidr_field with only the env stack filled, so that the inference kernel's
// own layout / pack code (tc_layout, tc_pack) builds the forward images
struct EnvTrainLayout {
    envidr_field f;
    TcEnv fwd;
    uint64_t fwd_bytes;
    BwLayer bw[4];
    uint64_t total_bytes;
};

static bool env_train_layout(const envidr_env_mlp* d, const void* blob, EnvTrainLayout* out) {
    EnvTrainLayout& L = *out;
    L = EnvTrainLayout{};
    if (d->n_layers < 2 || d->n_layers > 4) return false;
    L.f.n_env = d->n_layers;
    for (uint32_t i = 0; i < d->n_layers; i++) {
        L.f.env[i].weight = d->weight[i];
        L.f.env[i].bias = d->bias[i];
        L.f.env[i].in_dim = d->dims[i];
        L.f.env[i].out_dim = d->dims[i + 1];
    }
    L.f.ide_degree = d->ide_degree;
    L.f.diffuse_kappa_inv = d->diffuse_kappa_inv;
    L.f.light_intensity_scale = d->light_intensity_scale;
    L.f.packed = const_cast<void*>(blob);
    if (!tc_layout(&L.f, 0, &L.fwd, &L.fwd_bytes)) return false;
    if (L.fwd.E > 12) return false;
    uint64_t off = rup_t((uint32_t)L.fwd_bytes, 1024);
    const uint32_t n = d->n_layers;
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t fl = n - 1 - j;                      // forward layer transposed by chain layer j
        BwLayer& b = L.bw[j];
        b.Kp = rup_t(d->dims[fl + 1], 16);
        b.Np = (fl == 0) ? L.fwd.L[0].Kp : d->dims[fl];      // hidden widths are multiples of 32 (tc_layout)
        b.img_off = (uint32_t)off;
        off += (uint64_t)(b.Kp / 16) * b.Np * 64;
        if (b.Np > 256 || b.Kp > 256 || b.Np % 16 != 0) return false;
    }
    L.total_bytes = off;
    return true;
}

// generic small MLP (colour / diffuse / renv heads of the training branch): forward images + transposed images in one blob
struct MlpLayout { BwLayer fw[4], bw[4]; uint64_t total_bytes; uint32_t n; };
static bool mlp_layout(const envidr_env_mlp* d, MlpLayout* out) {
    MlpLayout& L = *out;
    L = MlpLayout{};
    const uint32_t n = d->n_layers;
    if (n < 2 || n > 4 || d->dims[0] < 1 || d->dims[0] > 64 || d->dims[n] < 1 || d->dims[n] > 16) return false;
    for (uint32_t i = 1; i < n; i++) if (d->dims[i] % 32 != 0 || d->dims[i] < 32 || d->dims[i] > 256) return false;
    uint64_t off = 0;
    for (uint32_t l = 0; l < n; l++) {                       // forward layer l: rows = outputs, K = inputs
        L.fw[l].Kp = rup_t(d->dims[l], 16);
        L.fw[l].Np = (l == n - 1) ? 16 : d->dims[l + 1];
        L.fw[l].img_off = (uint32_t)off;
        off += (uint64_t)(L.fw[l].Kp / 16) * L.fw[l].Np * 64;
    }
    for (uint32_t j = 0; j < n; j++) {                       // chain layer j of the backward = transpose of forward layer n-1-j
        const uint32_t fl = n - 1 - j;
        L.bw[j].Kp = rup_t(d->dims[fl + 1], 16);
        L.bw[j].Np = (fl == 0) ? rup_t(d->dims[0], 16) : d->dims[fl];
        L.bw[j].img_off = (uint32_t)off;
        off += (uint64_t)(L.bw[j].Kp / 16) * L.bw[j].Np * 64;
    }
    L.total_bytes = off;
    L.n = n;
    return true;
}

constexpr size_t kBwSmem = 2 * kBwARegion + 2 * kBwGRegion + kBwStages * kBwStageBytes + 128 * sizeof(float) + 256;

}  // namespace envidr

using namespace envidr;

extern "C" {

uint64_t envidr_env_mlp_blob_bytes(const envidr_env_mlp* d) {
    EnvTrainLayout L;
    if (!d || !env_train_layout(d, nullptr, &L)) return 0;
    return L.total_bytes;
}

int envidr_env_mlp_pack(const envidr_env_mlp* d, void* blob, uint64_t blob_bytes, envidr_stream_t stream) {
    ENVIDR_REQUIRE(d && blob, ENVIDR_E_BADARG, "null argument");
    EnvTrainLayout L;
    ENVIDR_REQUIRE(env_train_layout(d, blob, &L), ENVIDR_E_UNSUPPORTED,
                   "env_net outside the fused training kernels (2..4 layers, hidden widths multiples of 32 in 64..256, IDE input, env_feat <= 12)");
    ENVIDR_REQUIRE(blob_bytes >= L.total_bytes, ENVIDR_E_WORKSPACE, "blob too small (envidr_env_mlp_blob_bytes)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = tc_pack(&L.f, L.fwd, blob, st);
    if (rc) return rc;
    const uint32_t n = d->n_layers;
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t fl = n - 1 - j;
        k_pack_tcT<<<64, 256, 0, st>>>(d->weight[fl], reinterpret_cast<uint8_t*>(blob) + L.bw[j].img_off, d->dims[fl], d->dims[fl + 1], L.bw[j].Np,
                                       L.bw[j].Kp, fl == 0 ? L.fwd.P : 0u);
    }
    return check_launch("env_mlp_pack");
}

int envidr_env_mlp_forward(const envidr_env_mlp* d, const void* blob, const float* rec, uint32_t M, float* feat, float* act0, float* act1,
                           float* act2, uint32_t* mask0, uint32_t* mask1, uint32_t* mask2, envidr_stream_t stream) {
    ENVIDR_REQUIRE(d && blob && rec && feat, ENVIDR_E_BADARG, "null argument");
    EnvTrainLayout L;
    ENVIDR_REQUIRE(env_train_layout(d, blob, &L), ENVIDR_E_UNSUPPORTED, "env_net outside the fused training kernels");
    if (M == 0) return 0;
    TcSave sv{};
    float* acts[3] = {act0, act1, act2};
    uint32_t* masks[3] = {mask0, mask1, mask2};
    for (uint32_t i = 0; i + 1 < d->n_layers; i++) {
        ENVIDR_REQUIRE(acts[i] && masks[i], ENVIDR_E_BADARG, "activation / mask buffer missing");
        sv.act[i] = acts[i]; sv.mask[i] = masks[i];
    }
    sv.M = M;
    return env_tc_launch(L.fwd, d->ide_degree, rec, feat, nullptr, M, reinterpret_cast<cudaStream_t>(stream), &sv);
}

int envidr_env_mlp_backward(const envidr_env_mlp* d, const void* blob, const float* gfeat, const float* feat, const uint32_t* mask0,
                            const uint32_t* mask1, const uint32_t* mask2, uint32_t M, float* gact0, float* gact1, float* gact2, float* gy,
                            float* gx0, envidr_stream_t stream) {
    ENVIDR_REQUIRE(d && blob && gfeat && feat && gy && gx0, ENVIDR_E_BADARG, "null argument");
    EnvTrainLayout L;
    ENVIDR_REQUIRE(env_train_layout(d, blob, &L), ENVIDR_E_UNSUPPORTED, "env_net outside the fused training kernels");
    if (M == 0) return 0;
    EnvBwd B{};
    B.blob = reinterpret_cast<const uint8_t*>(blob);
    B.n_layers = d->n_layers;
    for (uint32_t j = 0; j < d->n_layers; j++) B.L[j] = L.bw[j];
    B.E = L.fwd.E;
    B.gfeat = gfeat; B.feat = feat;
    const uint32_t* masks[3] = {mask0, mask1, mask2};
    float* gacts[3] = {gact0, gact1, gact2};
    for (uint32_t i = 0; i + 1 < d->n_layers; i++) {
        ENVIDR_REQUIRE(masks[i] && gacts[i], ENVIDR_E_BADARG, "mask / gradient buffer missing");
        B.mask[i] = masks[i]; B.gact[i] = gacts[i];
    }
    B.gy = gy; B.gx0 = gx0; B.M = M;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwSmem);
        if (e != cudaSuccess) { set_error("env_bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    const uint32_t n_tiles = (2 * M + 127) / 128;
    k_chain_tc<0><<<min((uint32_t)kSMs, n_tiles), kBwThreads, kBwSmem, reinterpret_cast<cudaStream_t>(stream)>>>(B);
    return check_launch("env_mlp_backward");
}

uint64_t envidr_mlp_blob_bytes(const envidr_env_mlp* d) {
    MlpLayout L;
    return (d && mlp_layout(d, &L)) ? L.total_bytes : 0;
}

int envidr_mlp_pack(const envidr_env_mlp* d, void* blob, uint64_t blob_bytes, envidr_stream_t stream) {
    ENVIDR_REQUIRE(d && blob, ENVIDR_E_BADARG, "null argument");
    MlpLayout L;
    ENVIDR_REQUIRE(mlp_layout(d, &L), ENVIDR_E_UNSUPPORTED, "MLP outside the fused chain kernels (2..4 layers, inputs <= 64, hidden widths multiples of 32 <= 256, outputs <= 16)");
    ENVIDR_REQUIRE(blob_bytes >= L.total_bytes, ENVIDR_E_WORKSPACE, "blob too small (envidr_mlp_blob_bytes)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    uint8_t* b = reinterpret_cast<uint8_t*>(blob);
    for (uint32_t l = 0; l < L.n; l++)
        k_pack_tcT<<<16, 256, 0, st>>>(d->weight[l], b + L.fw[l].img_off, d->dims[l], d->dims[l + 1], L.fw[l].Np, L.fw[l].Kp, 0u, 0);
    for (uint32_t j = 0; j < L.n; j++) {
        const uint32_t fl = L.n - 1 - j;
        k_pack_tcT<<<16, 256, 0, st>>>(d->weight[fl], b + L.bw[j].img_off, d->dims[fl], d->dims[fl + 1], L.bw[j].Np, L.bw[j].Kp, 0u, 1);
    }
    return check_launch("mlp_pack");
}

static int chain_attr() {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwSmem);
        if (e != cudaSuccess) { set_error("mlp chain smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    return 0;
}

int envidr_mlp_forward(const envidr_env_mlp* d, const void* blob, const float* X, uint32_t x_ld, uint32_t rows, float* Y, uint32_t* mask0,
                       uint32_t* mask1, uint32_t* mask2, float* act0, float* act1, float* act2, envidr_stream_t stream) {
    ENVIDR_REQUIRE(d && blob && X && Y, ENVIDR_E_BADARG, "null argument");
    MlpLayout L;
    ENVIDR_REQUIRE(mlp_layout(d, &L), ENVIDR_E_UNSUPPORTED, "MLP outside the fused chain kernels");
    ENVIDR_REQUIRE(x_ld >= d->dims[0], ENVIDR_E_BADARG, "row stride of X smaller than the input width");
    if (rows == 0) return 0;
    EnvBwd B{};
    B.blob = reinterpret_cast<const uint8_t*>(blob);
    B.n_layers = L.n;
    for (uint32_t l = 0; l < L.n; l++) { B.L[l] = L.fw[l]; B.bias[l] = d->bias[l]; B.nout[l] = d->dims[l + 1]; }
    uint32_t* masks[3] = {mask0, mask1, mask2};
    float* acts[3] = {act0, act1, act2};
    for (uint32_t l = 0; l + 1 < L.n; l++) {
        ENVIDR_REQUIRE(masks[l], ENVIDR_E_BADARG, "mask buffer missing");
        B.mask_out[l] = masks[l]; B.act_out[l] = acts[l];
    }
    B.in = X; B.in_ld = x_ld; B.in_cols = d->dims[0]; B.rows = rows; B.y = Y;
    int rc = chain_attr();
    if (rc) return rc;
    const uint32_t n_tiles = (rows + 127) / 128;
    k_chain_tc<2><<<min((uint32_t)kSMs, n_tiles), kBwThreads, kBwSmem, reinterpret_cast<cudaStream_t>(stream)>>>(B);
    return check_launch("mlp_forward");
}

int envidr_mlp_backward(const envidr_env_mlp* d, const void* blob, const float* gY, uint32_t gy_ld, const uint32_t* mask0, const uint32_t* mask1,
                        const uint32_t* mask2, uint32_t rows, float* gz0, float* gz1, float* gz2, float* gX, envidr_stream_t stream) {
    ENVIDR_REQUIRE(d && blob && gY && gX, ENVIDR_E_BADARG, "null argument");
    MlpLayout L;
    ENVIDR_REQUIRE(mlp_layout(d, &L), ENVIDR_E_UNSUPPORTED, "MLP outside the fused chain kernels");
    ENVIDR_REQUIRE(gy_ld >= d->dims[L.n], ENVIDR_E_BADARG, "row stride of dY smaller than the output width");
    if (rows == 0) return 0;
    EnvBwd B{};
    B.blob = reinterpret_cast<const uint8_t*>(blob);
    B.n_layers = L.n;
    for (uint32_t j = 0; j < L.n; j++) B.L[j] = L.bw[j];
    const uint32_t* masks[3] = {mask0, mask1, mask2};
    float* gzs[3] = {gz0, gz1, gz2};
    for (uint32_t l = 0; l + 1 < L.n; l++) {
        ENVIDR_REQUIRE(masks[l], ENVIDR_E_BADARG, "mask buffer missing");
        B.mask[l] = masks[l]; B.gact[l] = gzs[l];
    }
    B.in = gY; B.in_ld = gy_ld; B.in_cols = d->dims[L.n]; B.rows = rows; B.gx0 = gX;
    int rc = chain_attr();
    if (rc) return rc;
    const uint32_t n_tiles = (rows + 127) / 128;
    k_chain_tc<1><<<min((uint32_t)kSMs, n_tiles), kBwThreads, kBwSmem, reinterpret_cast<cudaStream_t>(stream)>>>(B);
    return check_launch("mlp_backward");
}

uint64_t envidr_env_mlp_input_cols(const envidr_env_mlp* d) {
    EnvTrainLayout L;
    if (!d || !env_train_layout(d, nullptr, &L)) return 0;
    return L.fwd.L[0].Kp;
}

}  // extern "C"
