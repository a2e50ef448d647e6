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

constexpr int kGThreads = 576;                     // warps 0-7 chain (2 groups x 4), 8-15 gather, 16 MMA issuer, 17 TMEM / weights
constexpr uint32_t kGEnc = 16384, kGEncHalf = 8192;    // encoding operand: 128 rows x 32 K x 2 B, hi | lo
constexpr uint32_t kGHid = 32768, kGHidHalf = 16384;   // hidden operand:   128 rows x 64 K x 2 B, hi | lo
constexpr uint32_t kGJacBase = 128, kGJacCols = 96;    // TMEM columns: 2 accumulators x 64, then 4 jacobian slots x 96
// Shared-memory copy of the coarsest level table (north star: "shared-memory staging of level tables").  Level 0 of the shipped grid is
// 16^3 x 2 floats = 32 KB and fits next to the 150 KB of weight images and operands; level 1 (23^3 x 8 B = 97 KB) does not.  Every
// persistent CTA pulls it in once with a bulk copy; its 8 corner loads per sample then never leave the SM.  Measured (run r3_17, ncu over
// 12 mid-loop launches, staged vs not): 49.3 vs 48.1 us per launch, L1 hit rate 70.0 vs 71.3 %, lts throughput 7.2 vs 7.3 %, issue slots
// 22.3 vs 22.8 %, frame 15.90 vs 15.78 ms -- no gain: the 32 KB table already lives in L1 (the kernel leaves ~100 KB of it), and the
// staging area takes that capacity away from the finer levels.  Kept behind ENVIDR_GEOM_STAGE=1, default off.
constexpr uint32_t kGStageBytes = 32768;

struct GeomOutDev { float *sigma, *normal, *sdf, *roughness, *grad_x; };
struct GLevel { uint32_t off2, size, res, hashed, magic, on; float scale; uint32_t res2; };

__device__ __forceinline__ float g_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float g_softplus(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// raw % size with magic = floor(2^32 / size): the estimated quotient is exact or one short, so one conditional
// subtraction completes it (same value as the reference's `index % hashmap_size`, hashencoder.cu:71).
__device__ __forceinline__ uint32_t g_mod(uint32_t raw, uint32_t size, uint32_t magic) {
    uint32_t r = raw - __umulhi(raw, magic) * size;
    if (r >= size) r -= size;
    return r;
}

// write `vals` (32 columns col0..col0+31 of row `row`) into a 128-row hidden operand as fp16 hi / lo
__device__ __forceinline__ void g_store32(uint8_t* hid, uint32_t row, uint32_t col0, const float (&v)[32]) {
    #pragma unroll
    for (int jj = 0; jj < 4; jj++) {
        uint32_t ph[4], pl[4];
        #pragma unroll
        for (int e = 0; e < 4; e++) tc::split2(v[jj * 8 + 2 * e], v[jj * 8 + 2 * e + 1], ph[e], pl[e]);
        const uint32_t off = tc::op_off(128, row, col0 + jj * 8);
        *reinterpret_cast<uint4*>(hid + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(hid + kGHidHalf + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
}

__global__ void __launch_bounds__(kGThreads, 1)
k_geom_tc(const TcGeom G, const float* __restrict__ xyzs, const float* __restrict__ dirs, const uint32_t* __restrict__ M_dev,
          uint32_t M_host, int mode, float* __restrict__ rec, const GeomOutDev O, const uint32_t* __restrict__ rec_base_dev,
          uint32_t rec_cap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem;                                           // resident weight images + float region
    uint8_t* s_enc = smem + G.res_bytes_al;                        // two encoding operands (gather -> stage 0)
    uint8_t* s_hid = s_enc + 2 * kGEnc;                            // one hidden operand per chain group (rewritten in place)
    GLevel* s_lvl = reinterpret_cast<GLevel*>(s_hid + 2 * kGHid);  // [16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_lvl + 16);
    const float2* s_tab = reinterpret_cast<const float2*>(reinterpret_cast<uint8_t*>(s_lvl + 16) + 16 * 8);     // staged level-0 table (G.stage_bytes > 0)
    uint64_t* w_full = bars;              // resident weights landed
    uint64_t* enc_full = bars + 1;        // [2] gather -> issuer (256 arrivals)
    uint64_t* enc_free = bars + 3;        // [2] issuer (commit of the stage-0 MMA) -> gather
    uint64_t* acc_ready = bars + 5;       // [2] issuer -> chain group
    uint64_t* a_ready = bars + 7;         // [2] chain group -> issuer (128 arrivals)
    uint64_t* jac_free = bars + 9;        // [4] chain group -> gather (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: role branches stay uniform (UR datapath)
    const uint32_t M = M_dev ? *M_dev : M_host;
    const uint32_t n_tiles = (M + 127) / 128;
    if (blockIdx.x >= n_tiles) return;             // nothing to do for this CTA (tail iterations of the render loop)
    const uint32_t T = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;       // tiles of this CTA: blockIdx.x + j * gridDim.x
    const int n = (int)G.n_layers;                 // forward layers; n - 1 reverse stages
    const int n_stages = 2 * n - 1;
    const float* s_f = reinterpret_cast<const float*>(s_w + G.float_off);     // biases [n][64], then w_row0[64]

    if (tid == 0) {
        tc::mbar_init(w_full, 1);
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(enc_full + i, 256);
            tc::mbar_init(enc_free + i, 1);
            tc::mbar_init(acc_ready + i, 1);
            tc::mbar_init(a_ready + i, 128);
        }
        for (int i = 0; i < 4; i++) tc::mbar_init(jac_free + i, 128);
        tc::mbar_fence_init();
    }
    if (tid < G.L) {                               // per-level constants (reference: hashencoder.cu:118-124, 54-72)
        GLevel lv;
        const uint32_t l = tid;
        lv.off2 = (uint32_t)G.offsets[l] * 2;
        lv.size = (uint32_t)(G.offsets[l + 1] - G.offsets[l]);
        lv.scale = exp2f(l * G.S) * G.H - 1.0f;
        lv.res = (uint32_t)ceilf(lv.scale) + 1;
        uint32_t stride = 1;
        for (int d = 0; d < 3; d++) if (stride <= lv.size) stride *= lv.res;
        lv.hashed = stride > lv.size ? 1u : 0u;
        lv.magic = lv.size > 1 ? (uint32_t)(0x100000000ull / lv.size) : 0xFFFFFFFFu;
        lv.on = !(G.enabled_levels > 0 && (int)l >= G.enabled_levels) ? 1u : 0u;
        lv.res2 = lv.res * lv.res;
        s_lvl[l] = lv;
    }
    if (warp == 17) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // level 0 is staged when it is a dense level that fits the staging area (decided identically by every thread from the offsets)
    const uint32_t lvl0_bytes = (uint32_t)(G.offsets[1] - G.offsets[0]) * 8u;
    const bool stage0 = G.stage_bytes > 0 && lvl0_bytes <= G.stage_bytes && (lvl0_bytes & 15u) == 0 && G.offsets[0] == 0;
    if (warp == 17 && lane == 0) {                 // one-time: pull the resident region in with a few bulk copies
        tc::mbar_arrive_expect_tx(w_full, G.res_bytes + (stage0 ? lvl0_bytes : 0u));
        if (stage0)
            for (uint32_t o = 0; o < lvl0_bytes; o += 16384)
                tc::bulk_g2s(const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(s_tab)) + o, reinterpret_cast<const uint8_t*>(G.table) + o,
                             min(16384u, lvl0_bytes - o), w_full);
        for (uint32_t o = 0; o < G.res_bytes; o += 16384) {
            const uint32_t b = min(16384u, G.res_bytes - o);
            tc::bulk_g2s(s_w + o, G.blob + o, b, w_full);
        }
    }

    if (warp == 16) {
        // ===================== MMA issuer: two tiles in flight, stages interleaved =====================
        tc::mbar_wait(w_full, 0);
        uint32_t encf_par = 0, a_par = 0;          // bit u = parity of the next phase to wait for
        for (uint32_t p = 0; p < T; p += 2) {
            const uint32_t ntp = min(2u, T - p);
            for (int st = 0; st < n_stages; st++) {
                const TcImg& I = (st < n) ? G.F[st] : G.R[n_stages - 1 - st];      // reverse stages use layers n-2 .. 0
                for (uint32_t u = 0; u < ntp; u++) {
                    if (st == 0) {
                        tc::mbar_wait(enc_full + u, (encf_par >> u) & 1u); encf_par ^= 1u << u;
                        if (p >= 2) { tc::mbar_wait(a_ready + u, (a_par >> u) & 1u); a_par ^= 1u << u; }   // accumulator drained
                    } else {
                        tc::mbar_wait(a_ready + u, (a_par >> u) & 1u); a_par ^= 1u << u;
                    }
                    tc::tc_fence_after();
                    __syncwarp();
                    {
                        const uint32_t idesc = tc::make_idesc_f16(128, I.Np);
                        const uint32_t a_hi0 = tc::smem_u32(st == 0 ? s_enc + u * kGEnc : s_hid + u * kGHid);
                        const uint32_t a_lo0 = a_hi0 + (st == 0 ? kGEncHalf : kGHidHalf);
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
                        if (st == 0) tc::mma_commit_w(enc_free + u);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 8 && warp < 16) {
        // ===================== gather: thread = (sample, level parity) =====================
        const uint32_t quarter = warp & 3, par = (warp - 8) >> 2;
        const uint32_t s = quarter * 32 + lane;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        float xn[3] = {0.f, 0.f, 0.f};
        if (stage0) tc::mbar_wait(w_full, 0);     // the staged table arrives with the weight images
        {
            const uint32_t m = blockIdx.x * 128 + s;
            if (m < M) {
                #pragma unroll
                for (int d = 0; d < 3; d++) xn[d] = __ldg(xyzs + 3 * (size_t)m + d);
            }
        }
        for (uint32_t j = 0; j < T; j++) {
            const uint32_t tile = blockIdx.x + j * gridDim.x;
            const uint32_t m = tile * 128 + s;
            const bool valid = m < M;
            float x01[3];
            bool inside = valid;
            #pragma unroll
            for (int d = 0; d < 3; d++) {
                x01[d] = (xn[d] + G.bound) / (2 * G.bound);
                if (x01[d] < 0 || x01[d] > 1) inside = false;
            }
            if (j + 1 < T) {                       // prefetch the next tile's position
                const uint32_t mn = (tile + gridDim.x) * 128 + s;
                #pragma unroll
                for (int d = 0; d < 3; d++) xn[d] = mn < M ? __ldg(xyzs + 3 * (size_t)mn + d) : 0.f;
            }
            const uint32_t b = j & 1, q = j & 3;
            if (j >= 2) tc::mbar_wait(enc_free + b, ((j >> 1) - 1) & 1u);
            if (j >= 4) tc::mbar_wait(jac_free + q, ((j >> 2) - 1) & 1u);
            tc::tc_fence_after();
            uint8_t* enc = s_enc + b * kGEnc;
            const uint32_t jac_t = tmem + lane_addr + kGJacBase + q * kGJacCols + par * 48;
            for (uint32_t l0 = par; l0 < G.L; l0 += 4) {
                float rows[2][8][2], w[2][3], dw[2][3], sc[2];
                #pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint32_t l = l0 + 2 * u;
                    #pragma unroll
                    for (int c = 0; c < 8; c++) { rows[u][c][0] = 0.f; rows[u][c][1] = 0.f; }
                    #pragma unroll
                    for (int d = 0; d < 3; d++) { w[u][d] = 0.f; dw[u][d] = 0.f; }
                    sc[u] = 0.f;
                    if (l < G.L) {
                        const GLevel lv = s_lvl[l];
                        if (inside && lv.on) {
                            uint32_t pg[3];
                            #pragma unroll
                            for (int d = 0; d < 3; d++) {
                                float p = x01[d] * lv.scale;
                                pg[d] = (uint32_t)floorf(p);
                                p -= (float)pg[d];
                                dw[u][d] = 6 * p * (1.0f - p);
                                w[u][d] = p * p * (3.0f - 2.0f * p);
                            }
                            sc[u] = lv.scale;
                            const float2* grid = reinterpret_cast<const float2*>(G.table + lv.off2);
                            // corner c = (x bit 0, y bit 1, z bit 2); hashencoder.cu:54-72 with the y / z products shared
                            uint32_t raw[8];
                            if (lv.hashed) {
                                const uint32_t y0 = pg[1] * 2654435761u, y1 = y0 + 2654435761u;
                                const uint32_t z0 = pg[2] * 805459861u, z1 = z0 + 805459861u;
                                const uint32_t yz[4] = {y0 ^ z0, y1 ^ z0, y0 ^ z1, y1 ^ z1};
                                #pragma unroll
                                for (uint32_t c = 0; c < 8; c++) raw[c] = (pg[0] + (c & 1u)) ^ yz[c >> 1];
                            } else {
                                const uint32_t y0 = pg[1] * lv.res, z0 = pg[2] * lv.res2;
                                const uint32_t yz[4] = {y0 + z0, y0 + lv.res + z0, y0 + z0 + lv.res2, y0 + lv.res + z0 + lv.res2};
                                #pragma unroll
                                for (uint32_t c = 0; c < 8; c++) raw[c] = pg[0] + (c & 1u) + yz[c >> 1];
                            }
                            if (stage0 && l == 0) {            // warp-uniform: the level is a function of the warp's parity and loop counter
                                #pragma unroll
                                for (uint32_t c = 0; c < 8; c++) {
                                    const float2 t = s_tab[g_mod(raw[c], lv.size, lv.magic)];
                                    rows[u][c][0] = t.x; rows[u][c][1] = t.y;
                                }
                            } else {
                                #pragma unroll
                                for (uint32_t c = 0; c < 8; c++) {
                                    const float2 t = __ldg(grid + g_mod(raw[c], lv.size, lv.magic));
                                    rows[u][c][0] = t.x; rows[u][c][1] = t.y;
                                }
                            }
                        }
                    }
                }
                #pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint32_t l = l0 + 2 * u;
                    if (l < G.L) {
                        // tri-"linear" interpolation with smoothstep weights, factored x -> y -> z; the x / y / z differences
                        // that fall out are exactly the terms of d enc / d x (hashencoder.cu:210-253)
                        float ev[2], jv[6];
                        #pragma unroll
                        for (int ch = 0; ch < 2; ch++) {
                            float a[4], dx[4];
                            #pragma unroll
                            for (int j = 0; j < 4; j++) {
                                dx[j] = rows[u][2 * j + 1][ch] - rows[u][2 * j][ch];
                                a[j] = fmaf(w[u][0], dx[j], rows[u][2 * j][ch]);
                            }
                            float b[2], bx[2], by[2];
                            #pragma unroll
                            for (int k = 0; k < 2; k++) {
                                by[k] = a[2 * k + 1] - a[2 * k];
                                b[k] = fmaf(w[u][1], by[k], a[2 * k]);
                                bx[k] = fmaf(w[u][1], dx[2 * k + 1] - dx[2 * k], dx[2 * k]);
                            }
                            const float ez = b[1] - b[0];
                            ev[ch] = fmaf(w[u][2], ez, b[0]);
                            jv[0 + ch] = fmaf(w[u][2], bx[1] - bx[0], bx[0]) * (dw[u][0] * sc[u]);
                            jv[2 + ch] = fmaf(w[u][2], by[1] - by[0], by[0]) * (dw[u][1] * sc[u]);
                            jv[4 + ch] = ez * (dw[u][2] * sc[u]);
                        }
                        const float e0 = ev[0], e1 = ev[1];
                        uint32_t hi, lo;
                        tc::split2(e0, e1, hi, lo);
                        const uint32_t off = tc::op_off(128, s, 2 * l);
                        *reinterpret_cast<uint32_t*>(enc + off) = hi;
                        *reinterpret_cast<uint32_t*>(enc + kGEncHalf + off) = lo;
                        const uint32_t jt = jac_t + 6 * (l >> 1);
                        tmem_st4(jt, jv[0], jv[1], jv[2], jv[3]);
                        tmem_st2(jt + 4, jv[4], jv[5]);
                    }
                }
            }
            tmem_st_wait();
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(enc_full + b);
        }
    } else if (warp < 8) {
        // ===================== chain: group u = warp / 4 owns every other tile; thread = accumulator row =====================
        const uint32_t u = warp >> 2, quarter = warp & 3;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        const uint32_t acc_t = tmem + lane_addr + u * 64;
        uint8_t* hid = s_hid + u * kGHid;
        const uint32_t Hd = G.F[0].N;                      // hidden width (32 or 64)
        const uint32_t hchunks = Hd / 32;
        const int Gd = (int)G.geo_dim;
        const bool want_rec = rec && mode != 1;
        tc::mbar_wait(w_full, 0);
        uint32_t acc_par = 0;
        for (uint32_t j = u; j < T; j += 2) {
            const uint32_t tile = blockIdx.x + j * gridDim.x;
            const uint32_t m = tile * 128 + row;
            float dx = 0.f, dy = 0.f, dz = 0.f;
            if (want_rec && m < M) { dx = __ldg(dirs + 3 * (size_t)m); dy = __ldg(dirs + 3 * (size_t)m + 1); dz = __ldg(dirs + 3 * (size_t)m + 2); }
            uint32_t mk[2][2] = {{0u, 0u}, {0u, 0u}};      // ReLU masks [hidden layer][32-column chunk]
            float outv[16];
            #pragma unroll
            for (int i = 0; i < 16; i++) outv[i] = 0.f;
            for (int st = 0; st < n_stages; st++) {
                tc::mbar_wait(acc_ready + u, acc_par); acc_par ^= 1;
                tc::tc_fence_after();
                if (st < n - 1) {
                    // hidden forward layer st: h = relu(D + b), keep the mask, write the next operand (in place)
                    #pragma unroll
                    for (uint32_t ch = 0; ch < 2; ch++) if (ch < hchunks) {
                        const uint32_t v = tc::hidden_epilogue32(acc_t + ch * 32, s_f + st * 64 + ch * 32, hid, hid + kGHidHalf, row, ch * 32);
                        if (st == 0) mk[0][ch] = v; else mk[1][ch] = v;
                    }
                } else if (st == n - 1) {
                    // last forward layer: keep its outputs, start the reverse pass: g = W_last[0,:] * relu'(h_{n-2})
                    uint32_t r[16];
                    tc::tmem_ld16(acc_t, r);
                    tc::tmem_ld_wait();
                    const float* bias = s_f + st * 64;
                    #pragma unroll
                    for (int i = 0; i < 16; i++) outv[i] = __uint_as_float(r[i]) + bias[i];
                    #pragma unroll
                    for (uint32_t ch = 0; ch < 2; ch++) if (ch < hchunks) {
                        const float* w0 = s_f + n * 64 + ch * 32;
                        const uint32_t mask = (n - 2 == 0) ? mk[0][ch] : mk[1][ch];
                        float v[32];
                        #pragma unroll
                        for (int c = 0; c < 32; c++) v[c] = ((mask >> c) & 1u) ? w0[c] : 0.f;
                        g_store32(hid, row, ch * 32, v);
                    }
                } else if (st < n_stages - 1) {
                    // reverse stage through layer i = n_stages - 1 - st (>= 1): g_{i-1} = D * relu'(h_{i-1})
                    const int i = n_stages - 1 - st;
                    #pragma unroll
                    for (uint32_t ch = 0; ch < 2; ch++) if (ch < hchunks) {
                        uint32_t r[32];
                        tc::tmem_ld32(acc_t + ch * 32, r);
                        tc::tmem_ld_wait();
                        const uint32_t mask = (i - 1 == 0) ? mk[0][ch] : mk[1][ch];
                        float v[32];
                        #pragma unroll
                        for (int c = 0; c < 32; c++) v[c] = ((mask >> c) & 1u) ? __uint_as_float(r[c]) : 0.f;
                        g_store32(hid, row, ch * 32, v);
                    }
                } else {
                    // last reverse stage: D = d sdf / d enc; contract with the jacobian (TMEM slot j & 3), finish the sample
                    uint32_t g[32];
                    tc::tmem_ld32(acc_t, g);
                    float gx = 0.f, gy = 0.f, gz = 0.f;
                    const uint32_t jt = tmem + lane_addr + kGJacBase + (j & 3) * kGJacCols;
                    #pragma unroll
                    for (int cc = 0; cc < 3; cc++) {
                        uint32_t r[32];
                        tc::tmem_ld32(jt + cc * 32, r);
                        tc::tmem_ld_wait();
                        #pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const int col = cc * 32 + i, jp = col / 48, rem = col % 48, li = rem / 6, qq = rem % 6;
                            const int l = 2 * li + jp;
                            if (l < (int)G.L) {
                                const float t = __uint_as_float(g[2 * l + (qq & 1)]) * __uint_as_float(r[i]);
                                if ((qq >> 1) == 0) gx += t; else if ((qq >> 1) == 1) gy += t; else gz += t;
                            }
                        }
                    }
                    tc::tc_fence_before();
                    tc::mbar_arrive(a_ready + u);          // accumulator and jacobian slot are drained
                    tc::mbar_arrive(jac_free + (j & 3));
                    const float inv2b = 1.0f / (2 * G.bound);
                    gx *= inv2b; gy *= inv2b; gz *= inv2b;
                    const float gn = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-10f);
                    const float nx = gx / gn, ny = gy / gn, nz = gz / gn;
                    const float sdf = outv[0];
                    const float sg = (sdf > 0.f) ? 1.f : ((sdf < 0.f) ? -1.f : 0.f);
                    const float sigma = (1.0f / G.beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(sdf) / G.beta)) * G.density_scale;
                    if (m < M) {
                        if (O.sigma) O.sigma[m] = sigma;
                        if (O.sdf) O.sdf[m] = sdf;
                        if (O.normal) { O.normal[3 * (size_t)m] = nx; O.normal[3 * (size_t)m + 1] = ny; O.normal[3 * (size_t)m + 2] = nz; }
                        if (O.grad_x) { O.grad_x[3 * (size_t)m] = gx; O.grad_x[3 * (size_t)m + 1] = gy; O.grad_x[3 * (size_t)m + 2] = gz; }
                        float rough_raw = 0.f, blend_raw = 0.f, ss = 0.f;
                        #pragma unroll
                        for (int i = 1; i < 16; i++) {
                            if (i <= Gd) ss += outv[i] * outv[i];
                            if (i == 1 + Gd) rough_raw = outv[i];
                            if (i == 2 + Gd) blend_raw = outv[i];
                        }
                        const float rough = G.rough_act_scale * g_softplus(rough_raw + G.rough_bias) * G.rough_scale;
                        if (O.roughness) O.roughness[m] = rough;
                        if (want_rec) {
                            const float ginv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                            const float blend = g_sigmoid(blend_raw);
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
                            // record slot: batch-local index, or (capture of a geometry pass) offset by the running sample total
                            const size_t ri = (size_t)(rec_base_dev ? *rec_base_dev : 0u) + m;
                            if (ri < rec_cap) {
                                float4* q4 = reinterpret_cast<float4*>(rec + ri * kTcRecFloats);
                                float gq[16];
                                #pragma unroll
                                for (int i = 0; i < 15; i++) gq[i] = (i < Gd) ? outv[1 + i] * ginv : 0.f;
                                gq[15] = 0.f;
                                #pragma unroll
                                for (int i = 0; i < 4; i++) q4[i] = make_float4(gq[4 * i], gq[4 * i + 1], gq[4 * i + 2], gq[4 * i + 3]);
                                q4[4] = make_float4(nx, ny, nz, ndot);
                                q4[5] = make_float4(rough, blend, nex, ney);
                                q4[6] = make_float4(nez, wrx, wry, wrz);
                            }
                        }
                    }
                }
                if (st < n_stages - 1) {
                    tc::tc_fence_before();
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive(a_ready + u);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 17) tc::tmem_dealloc(tmem, 512);
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
                   const envidr_field_out* out, cudaStream_t st, const uint32_t* rec_base_dev, uint32_t rec_cap) {
    static int stage = -1;
    if (stage < 0) { const char* e = getenv("ENVIDR_GEOM_STAGE"); stage = (e && e[0] == '1') ? 1 : 0; }
    TcGeom gs = g;
    gs.stage_bytes = (stage && (reinterpret_cast<uintptr_t>(g.table) & 15u) == 0) ? kGStageBytes : 0u;
    const size_t smem = (size_t)g.res_bytes_al + 2 * kGEnc + 2 * kGHid + 16 * sizeof(GLevel) + 16 * 8 + gs.stage_bytes;
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
    k_geom_tc<<<grid, kGThreads, smem, st>>>(gs, xyzs, dirs, M_dev, M_host, mode, rec, O, rec_base_dev, rec_cap);
    return check_launch("geom_tc");
}

}  // namespace envidr
