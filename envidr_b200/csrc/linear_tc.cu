// linear_tc.cu -- Y[M x N] = act(X[M x K] W^T + b) on the 5th-generation tensor cores, fp32 in / fp32 out, with the same
// split-precision contract as the inference kernels (every operand x = hi + lo in fp16, three tcgen05.mma per K step, fp32
// accumulation in tensor memory; see field_tc.cu).  Used by the TRAINING branch (envidr_b200/train.py) for the forward and the
// data-gradient GEMMs of the env / colour / diffuse / renv MLPs, which the reference runs as cuBLAS fp32 GEMMs
// (nn.Linear, nerf/network.py:527-698): dY W is the same kernel on the image of W^T.
//
// Persistent CTA per SM, 640 threads:
//   warp 0       producer: streams the packed weight image (k_pack_tc layout: per 16-wide K step [hi | lo], chunk-major) through a
//                3 x 16 KB ring with 1-D bulk copies, once per 128-row tile
//   warp 1       issuer (warp-uniform code, elect.sync inside the asm)
//   warp 2       TMEM allocator (2 accumulators of up to 256 columns, ping-pong between tiles)
//   warps 4-11   loaders: read the next 128 rows of X (fp32, row-major), split to fp16 hi / lo, write the A operand
//   warps 12-19  epilogue: tcgen05.ld -> + bias -> ReLU -> fp32 rows of Y, overlapping the next tile's MMAs
#include "common.cuh"
#include "tc_common.cuh"

namespace envidr {

constexpr int kLinThreads = 640;
constexpr int kLinStages = 3;
constexpr uint32_t kLinStageBytes = 16384;
constexpr uint32_t kLinARegion = 65536;              // 128 rows x 256 K x 2 B

static uint32_t lin_rup(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

__global__ void __launch_bounds__(kLinThreads, 1)
k_linear_tc(const float* __restrict__ X, uint32_t M, uint32_t K, uint32_t Kp, const uint8_t* __restrict__ img, const float* __restrict__ bias,
            uint32_t N, uint32_t Np, int relu, float* __restrict__ Y) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA_hi = smem;
    uint8_t* sA_lo = smem + kLinARegion;
    uint8_t* ring = smem + 2 * kLinARegion;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kLinStages * kLinStageBytes);
    uint64_t* full = bars;                          // [3] producer -> issuer
    uint64_t* empty = bars + kLinStages;            // [3] issuer -> producer
    uint64_t* a_full = bars + 2 * kLinStages;       // loaders -> issuer (256 arrivals)
    uint64_t* a_free = a_full + 1;                  // issuer -> loaders (MMAs of the tile retired)
    uint64_t* acc_ready = a_full + 2;               // [2] issuer -> epilogue
    uint64_t* acc_free = a_full + 4;                // [2] epilogue -> issuer (256 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 6);
    float* s_max = reinterpret_cast<float*>(a_full + 8);      // [2][128] partial row maxima of the two loader threads of a row
    float* s_inv = s_max + 256;                               // [4][128] 1 / row scale of the tiles in flight (power of two)

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t n_tiles = (M + 127) / 128;
    if (blockIdx.x >= n_tiles) return;
    const uint32_t ksteps = Kp / 16, kb = Np * 64;
    const uint32_t kper = max(1u, kLinStageBytes / kb);         // K steps per ring stage (narrow layers)

    if (tid == 0) {
        for (int i = 0; i < kLinStages; i++) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(a_full, 256); tc::mbar_init(a_free, 1);
        for (int i = 0; i < 2; i++) { tc::mbar_init(&acc_ready[i], 1); tc::mbar_init(&acc_free[i], 256); }
        tc::mbar_fence_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== producer =====================
        uint32_t stage = 0, phase = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (uint32_t s = 0; s < ksteps; s += kper) {
                const uint32_t bytes = min(kper, ksteps - s) * kb;
                tc::mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    tc::mbar_arrive_expect_tx(&full[stage], bytes);
                    tc::bulk_g2s(ring + stage * kLinStageBytes, img + (size_t)s * kb, bytes, &full[stage]);
                }
                __syncwarp();
                if (++stage == kLinStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== issuer =====================
        uint32_t stage = 0, phase = 0, af_par = 0, accf_par = 0, it = 0;
        const uint32_t idesc = tc::make_idesc_f16(128, Np);
        const uint64_t da_hi0 = tc::make_smem_desc(tc::smem_u32(sA_hi), 2048, 128), da_lo0 = tc::make_smem_desc(tc::smem_u32(sA_lo), 2048, 128);
        const uint64_t db0 = tc::make_smem_desc(tc::smem_u32(ring), Np * 16, 128);
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const uint32_t buf = it & 1u;
            tc::mbar_wait(a_full, af_par); af_par ^= 1;
            if (it >= 2) { tc::mbar_wait(&acc_free[buf], (accf_par >> buf) & 1u); accf_par ^= 1u << buf; }
            tc::tc_fence_after();
            const uint32_t d = tmem + buf * 256u;
            uint64_t da_hi = da_hi0, da_lo = da_lo0;
            for (uint32_t s0 = 0; s0 < ksteps; s0 += kper) {
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after();
                __syncwarp();
                uint64_t db = tc::desc_advance(db0, stage * kLinStageBytes);
                const uint32_t kend = min(ksteps, s0 + kper);
                for (uint32_t s = s0; s < kend; s++) {
                    const uint64_t db_lo = tc::desc_advance(db, kb / 2);
                    tc::mma_f16_ss_w(d, da_hi, db, idesc, s > 0);
                    tc::mma_f16_ss_w(d, da_lo, db, idesc, 1);
                    tc::mma_f16_ss_w(d, da_hi, db_lo, idesc, 1);
                    da_hi = tc::desc_advance(da_hi, 4096); da_lo = tc::desc_advance(da_lo, 4096);
                    db = tc::desc_advance(db, kb);
                }
                tc::mma_commit_w(&empty[stage]);
                if (++stage == kLinStages) { stage = 0; phase ^= 1; }
            }
            tc::mma_commit_w(a_free);                 // the A operand may be overwritten with the next tile
            tc::mma_commit_w(&acc_ready[buf]);
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================== loaders: X rows -> fp16 hi / lo A operand =====================
        const uint32_t t2 = tid - 4 * 32;
        const uint32_t row = t2 & 127, part = t2 >> 7;          // two threads per row: 8-wide K chunks c = part (mod 2)
        const uint32_t nchunks = Kp / 8;
        const bool vec = (K % 4) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
        uint32_t free_par = 1, it = 0;                           // first wait passes
        auto load8 = [&](const float* x, uint32_t m, uint32_t k0, float (&v)[8]) {
            if (m < M && vec && k0 + 8 <= K) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(x + k0)), b = __ldg(reinterpret_cast<const float4*>(x + k0 + 4));
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
                #pragma unroll
                for (int j = 0; j < 8; j++) v[j] = (m < M && k0 + j < K) ? __ldg(x + k0 + j) : 0.0f;
            }
        };
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const uint32_t m = tile * 128 + row;
            const float* x = X + (size_t)m * K;
            // Per-row power-of-two scale so that the row's largest magnitude lands in [2^13, 2^14): fp16 has 5 exponent bits, and the
            // rows of a gradient (dY of a loss averaged over 1e5 samples) sit around 1e-7, where hi would be subnormal and lo zero.
            // The scale is exact and undone on the accumulator row in the epilogue.
            float mx = 0.0f;
            for (uint32_t c = part; c < nchunks; c += 8) {           // 4 chunks (8 vector loads) in flight per thread
                float v[4][8];
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (c + 2 * u < nchunks) load8(x, m, (c + 2 * u) * 8, v[u]);
                    else { for (int j = 0; j < 8; j++) v[u][j] = 0.0f; }
                }
                #pragma unroll
                for (int u = 0; u < 4; u++)
                    #pragma unroll
                    for (int j = 0; j < 8; j++) mx = fmaxf(mx, fabsf(v[u][j]));
            }
            s_max[part * 128 + row] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");       // the 8 loader warps
            mx = fmaxf(s_max[row], s_max[128 + row]);
            float scale = 1.0f, inv = 1.0f;
            const int e = (int)((__float_as_uint(mx) >> 23) & 255u) - 127;         // floor(log2(mx)) for normal mx
            if (mx > 1e-30f && mx < 1e30f) { scale = __uint_as_float((uint32_t)(127 + 13 - e) << 23); inv = __uint_as_float((uint32_t)(127 - 13 + e) << 23); }
            tc::mbar_wait(a_free, free_par); free_par ^= 1;
            for (uint32_t c = part; c < nchunks; c += 8) {
                float v[4][8];
                #pragma unroll
                for (int u = 0; u < 4; u++) if (c + 2 * u < nchunks) load8(x, m, (c + 2 * u) * 8, v[u]);
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (c + 2 * u < nchunks) {
                        #pragma unroll
                        for (int j = 0; j < 8; j++) v[u][j] *= scale;
                        tc::store_chunk8(sA_hi, sA_lo, row, (c + 2 * u) * 8, v[u]);
                    }
                }
            }
            if (part == 0) s_inv[(it & 3u) * 128 + row] = inv;
            asm volatile("bar.sync 1, 256;" ::: "memory");       // s_max may be rewritten for the next tile
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(a_full);
        }
    } else if (warp >= 12) {
        // ===================== epilogue: accumulator -> + bias -> ReLU -> Y =====================
        const uint32_t quarter = warp & 3, g = (warp - 12) >> 2;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        const uint32_t nchunks = (N + 31) / 32;
        const bool vec = (N % 4) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0;
        uint32_t acc_par = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const uint32_t buf = it & 1u;
            const uint32_t m = tile * 128 + row;
            tc::mbar_wait(&acc_ready[buf], (acc_par >> buf) & 1u); acc_par ^= 1u << buf;
            tc::tc_fence_after();
            const float inv = s_inv[(it & 3u) * 128 + row];
            const uint32_t acc = tmem + lane_addr + buf * 256u;
            for (uint32_t cb = g; cb < nchunks; cb += 2) {
                uint32_t r[32];
                if (cb * 32 + 32 <= Np) {
                    tc::tmem_ld32(acc + cb * 32, r);
                } else {                                         // Np is a multiple of 16: a 16-column tail
                    uint32_t r16[16];
                    tc::tmem_ld16(acc + cb * 32, r16);
                    #pragma unroll
                    for (int j = 0; j < 16; j++) { r[j] = r16[j]; r[16 + j] = 0; }
                }
                tc::tmem_ld_wait();
                if (m < M) {
                    float* y = Y + (size_t)m * N + cb * 32;
                    #pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float o[4];
                        #pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const uint32_t col = cb * 32 + j + e;
                            float t = __uint_as_float(r[j + e]) * inv + ((bias && col < N) ? __ldg(bias + col) : 0.0f);
                            o[e] = relu ? fmaxf(t, 0.0f) : t;
                        }
                        const uint32_t col0 = cb * 32 + j;
                        if (vec && col0 + 4 <= N) *reinterpret_cast<float4*>(y + j) = make_float4(o[0], o[1], o[2], o[3]);
                        else {
                            #pragma unroll
                            for (int e = 0; e < 4; e++) if (col0 + e < N) y[j + e] = o[e];
                        }
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&acc_free[buf]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem, 512);
}

// image of W [N x K] (row-major): per 16-wide K step s: [hi: chunk 0 | chunk 1][lo: chunk 0 | chunk 1], chunk = [Np][8] halfs
__global__ void k_linear_pack(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np) {
    const uint32_t total = Kp * Np;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t n = i / Kp, k = i - n * Kp;
        const float v = (n < N && k < K) ? W[(size_t)n * K + k] : 0.0f;
        __half h, lo;
        tc::split_f16(v, h, lo);
        const uint32_t s = k >> 4, kk = k & 15;
        const size_t base = (size_t)s * Np * 64 + (kk >> 3) * (Np * 16) + n * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + base) = h;
        *reinterpret_cast<__half*>(img + base + (size_t)Np * 32) = lo;
    }
}

constexpr size_t kLinSmem = 2 * kLinARegion + kLinStages * kLinStageBytes + 256 + (256 + 512) * sizeof(float);

}  // namespace envidr

using namespace envidr;

extern "C" {

uint64_t envidr_linear_tc_image_bytes(uint32_t N, uint32_t K) {
    if (N == 0 || K == 0 || N > 256 || K > 256) return 0;
    return (uint64_t)lin_rup(K, 16) * lin_rup(N, 16) * 4;
}

int envidr_linear_tc_pack(const float* W, uint32_t N, uint32_t K, void* img, envidr_stream_t stream) {
    ENVIDR_REQUIRE(W && img, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(N >= 1 && N <= 256 && K >= 1 && K <= 256, ENVIDR_E_UNSUPPORTED, "linear_tc: 1 <= N, K <= 256");
    k_linear_pack<<<64, 256, 0, as_stream(stream)>>>(W, reinterpret_cast<uint8_t*>(img), K, N, lin_rup(K, 16), lin_rup(N, 16));
    g_launches += 1;
    return check_launch("linear_tc_pack");
}

int envidr_linear_tc(const float* X, uint32_t M, uint32_t K, const void* img, const float* bias, uint32_t N, int relu, float* Y,
                     envidr_stream_t stream) {
    ENVIDR_REQUIRE(X && img && Y, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(N >= 1 && N <= 256 && K >= 1 && K <= 256, ENVIDR_E_UNSUPPORTED, "linear_tc: 1 <= N, K <= 256");
    ENVIDR_REQUIRE((reinterpret_cast<uintptr_t>(img) & 15) == 0, ENVIDR_E_BADARG, "weight image must be 16-byte aligned");
    if (M == 0) return 0;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLinSmem);
        if (e != cudaSuccess) { set_error("linear_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    const uint32_t grid = min((uint32_t)kSMs, (M + 127) / 128);
    k_linear_tc<<<grid, kLinThreads, kLinSmem, as_stream(stream)>>>(X, M, K, lin_rup(K, 16), reinterpret_cast<const uint8_t*>(img), bias, N,
                                                                    lin_rup(N, 16), relu, Y);
    g_launches += 1;
    return check_launch("linear_tc");
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------------
// Weight gradient dW[N x K] = dY^T X  (dY [M x N], X [M x K], both fp32 row-major; the contraction runs over the M samples).
// Both operands are "MN-major" for the tensor core (the M / N index is the contiguous one in memory), so the rows of dY and X
// go to shared memory as they are -- 16-byte stores of 8 consecutive features of one sample -- in the canonical no-swizzle
// MN-major layout (cute: ((8,1,m),(8,k)):((1,8,SBO),(8,LBO)) elements): element (mn, s) of a K step at
//   (s / 8) * LBO + (mn / 8) * 128 + (s % 8) * 16 + (mn % 8) * 2 bytes,  LBO = rows * 16, SBO = 128.
// Each CTA owns a contiguous range of samples, accumulates its partial dW in tensor memory (two 128-row halves of N x up to 256
// columns = all 512 columns) and writes it to partial[cta]; the host sums the partials.  Same fp16 hi / lo split (3 MMAs per
// 16 samples and half); the operands are pre-scaled by per-tensor powers of two (scales[0..1], undone with scales[2]) because
// a loss gradient sits around 1e-7, below fp16's normal range.
// ------------------------------------------------------------------------------------------------------------------------
namespace envidr {

constexpr int kWgThreads = 384;                      // 4 control warps + 8 loader / epilogue warps
constexpr uint32_t kWgSamples = 32;                  // samples per stage (two K steps)
constexpr uint32_t kWgABytes = 2 * 2 * 128 * kWgSamples * 2;    // two n-halves x (hi, lo) x 128 rows x 32 samples x 2 B = 32 KB
constexpr uint32_t kWgBBytes = 2 * 256 * kWgSamples * 2;        // (hi, lo) x 256 rows x 32 samples x 2 B = 32 KB
constexpr uint32_t kWgStage = kWgABytes + kWgBBytes;

__host__ __device__ constexpr uint32_t make_idesc_f16_mn(uint32_t M, uint32_t N) {     // both operands MN-major
    return (1u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(kWgThreads, 1)
k_wgrad_tc(const float* __restrict__ G, const float* __restrict__ X, uint32_t M, uint32_t N, uint32_t K, uint32_t Kp,
           const float* __restrict__ scales, float* __restrict__ partial, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kWgStage);
    uint64_t* full = bars;                       // [2] loaders -> issuer (256 arrivals)
    uint64_t* empty = bars + 2;                  // [2] issuer -> loaders
    uint64_t* acc_ready = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t n_halves = (N + 127) / 128;
    // sample range of this CTA (multiples of the stage size)
    const uint32_t n_st_total = (M + kWgSamples - 1) / kWgSamples;
    const uint32_t per = (n_st_total + gridDim.x - 1) / gridDim.x;
    const uint32_t st0 = min(n_st_total, blockIdx.x * per), st1 = min(n_st_total, st0 + per);
    const bool accumulate = (variant & 2) != 0;  // partial is dW [N, K] itself (zeroed by the caller): CTAs add their sums atomically
    float* out = accumulate ? partial : partial + (size_t)blockIdx.x * N * K;
    if (st0 >= st1) {                            // no samples: the partial is zero
        if (!accumulate)
            for (uint32_t i = tid; i < N * K; i += kWgThreads) out[i] = 0.0f;
        return;
    }
    if (tid == 0) {
        for (int i = 0; i < 2; i++) { tc::mbar_init(&full[i], 256); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(acc_ready, 1);
        tc::mbar_fence_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lboA = 128 * 16, lboB = Kp * 16;                 // bytes between the two 8-sample groups of a K step

    if (warp == 1) {
        // ===================== issuer =====================
        const uint32_t idesc = make_idesc_f16_mn(128, Kp);
        uint32_t phase = 0;
        for (uint32_t st = st0, it = 0; st < st1; st++, it++) {
            const uint32_t b = it & 1u;
            tc::mbar_wait(&full[b], phase);
            tc::tc_fence_after();
            __syncwarp();
            const uint32_t sbase = tc::smem_u32(smem + b * kWgStage);
            for (uint32_t j = 0; j < kWgSamples / 16; j++) {
                // B operand (X): hi block then lo block, each Kp rows x 32 samples
                const uint32_t b_hi = sbase + kWgABytes + j * 2 * lboB, b_lo = b_hi + Kp * kWgSamples * 2;
                const uint64_t db_hi = (variant & 1) ? tc::make_smem_desc(b_hi, 128, lboB) : tc::make_smem_desc(b_hi, lboB, 128);
                const uint64_t db_lo = (variant & 1) ? tc::make_smem_desc(b_lo, 128, lboB) : tc::make_smem_desc(b_lo, lboB, 128);
                for (uint32_t h = 0; h < n_halves; h++) {
                    const uint32_t a_hi = sbase + h * (2 * 128 * kWgSamples * 2) + j * 2 * lboA, a_lo = a_hi + 128 * kWgSamples * 2;
                    const uint64_t da_hi = (variant & 1) ? tc::make_smem_desc(a_hi, 128, lboA) : tc::make_smem_desc(a_hi, lboA, 128);
                    const uint64_t da_lo = (variant & 1) ? tc::make_smem_desc(a_lo, 128, lboA) : tc::make_smem_desc(a_lo, lboA, 128);
                    const uint32_t d = tmem + h * 256u;
                    const uint32_t acc = (it > 0 || j > 0) ? 1u : 0u;
                    tc::mma_f16_ss_w(d, da_hi, db_hi, idesc, acc);
                    tc::mma_f16_ss_w(d, da_lo, db_hi, idesc, 1);
                    tc::mma_f16_ss_w(d, da_hi, db_lo, idesc, 1);
                }
            }
            tc::mma_commit_w(&empty[b]);
            if (b == 1) phase ^= 1;
        }
        tc::mma_commit_w(acc_ready);
    } else if (warp >= 4) {
        // ===================== loaders (lane = sample of the stage), then epilogue =====================
        const uint32_t w = warp - 4;                                 // 0..7
        const float sg = scales[0], sx = scales[1], inv = scales[2];
        const uint32_t ga = n_halves * 16, gb = Kp / 8;              // groups of 8 features: dY (padded to 128 per half), X
        const bool vecG = (N % 4) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0;
        const bool vecX = (K % 4) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
        uint32_t phase = 1;                                          // first waits pass
        for (uint32_t st = st0, it = 0; st < st1; st++, it++) {
            const uint32_t b = it & 1u;
            tc::mbar_wait(&empty[b], phase);
            uint8_t* sb = smem + b * kWgStage;
            const uint32_t m = st * kWgSamples + lane;
            const uint32_t soff = (lane & 7) * 16;                      // sample inside its group of 8
            for (uint32_t g0 = w; g0 < ga + gb; g0 += 32) {             // 4 groups (8 vector loads) in flight per thread
                float v[4][8];
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t g = g0 + 8 * u;
                    if (g < ga + gb) {
                        const bool isA = g < ga;
                        const uint32_t f0 = (isA ? g : g - ga) * 8, F = isA ? N : K;
                        const float* src = isA ? G + (size_t)m * N : X + (size_t)m * K;
                        if (m < M && (isA ? vecG : vecX) && f0 + 8 <= F) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(src + f0)), c = __ldg(reinterpret_cast<const float4*>(src + f0 + 4));
                            v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = c.x; v[u][5] = c.y; v[u][6] = c.z; v[u][7] = c.w;
                        } else {
                            #pragma unroll
                            for (int e = 0; e < 8; e++) v[u][e] = (m < M && f0 + e < F) ? __ldg(src + f0 + e) : 0.0f;
                        }
                    }
                }
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t g = g0 + 8 * u;
                    if (g < ga + gb) {
                        const bool isA = g < ga;
                        const uint32_t gg = isA ? g : g - ga;
                        const float sc = isA ? sg : sx;
                        uint32_t ph[4], pl[4];
                        #pragma unroll
                        for (int e = 0; e < 4; e++) tc::split2(v[u][2 * e] * sc, v[u][2 * e + 1] * sc, ph[e], pl[e]);
                        uint32_t hi_off, lo_off;
                        if (isA) {
                            const uint32_t h = gg >> 4, gh = gg & 15;           // n-half, group inside the half
                            hi_off = h * (2 * 128 * kWgSamples * 2) + (lane >> 3) * lboA + gh * 128 + soff;
                            lo_off = hi_off + 128 * kWgSamples * 2;
                        } else {
                            hi_off = kWgABytes + (lane >> 3) * lboB + gg * 128 + soff;
                            lo_off = hi_off + Kp * kWgSamples * 2;
                        }
                        *reinterpret_cast<uint4*>(sb + hi_off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                        *reinterpret_cast<uint4*>(sb + lo_off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    }
                }
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&full[b]);
            if (b == 1) phase ^= 1;
        }
        // epilogue: warp -> (TMEM lane quarter, n-half); lane -> row n of dW
        tc::mbar_wait(acc_ready, 0);
        tc::tc_fence_after();
        const uint32_t quarter = warp & 3, h = w >> 2;
        if (h < n_halves) {
            const uint32_t n = h * 128 + quarter * 32 + lane;
            const uint32_t acc = tmem + ((quarter * 32u) << 16) + h * 256u;
            for (uint32_t c0 = 0; c0 < Kp; c0 += 16) {
                uint32_t r[16];
                tc::tmem_ld16(acc + c0, r);
                tc::tmem_ld_wait();
                if (n < N) {
                    if (!accumulate) {
                        #pragma unroll
                        for (int e = 0; e < 16; e++) if (c0 + e < K) out[(size_t)n * K + c0 + e] = __uint_as_float(r[e]) * inv;
                    } else if ((K & 3) == 0 && c0 + 16 <= K) {       // 16-byte aligned rows: four vector reductions (red.global.add.v4.f32)
                        #pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            atomicAdd(reinterpret_cast<float4*>(out + (size_t)n * K + c0 + e),
                                      make_float4(__uint_as_float(r[e]) * inv, __uint_as_float(r[e + 1]) * inv, __uint_as_float(r[e + 2]) * inv,
                                                  __uint_as_float(r[e + 3]) * inv));
                    } else {
                        #pragma unroll
                        for (int e = 0; e < 16; e++) if (c0 + e < K) atomicAdd(out + (size_t)n * K + c0 + e, __uint_as_float(r[e]) * inv);
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem, 512);
}

constexpr size_t kWgSmem = 2 * kWgStage + 256;

// Per-tensor power-of-two scales of the two wgrad operands in one launch (the torch formulation is ~14 tiny kernels):
// scales[0] = s_a, [1] = s_b with s = 2^(13 - floor(log2(max|.|))) (largest magnitude into [2^13, 2^14)), [2] = 1 / (s_a s_b).
// scales[4..6] are scratch (bit patterns of the two maxima, ticket) and must be zero on entry; the last block restores that.
// The pass over `a` (= dY [M, N]) can also produce its column sums (the bias gradient), added into colsum [N] (zeroed by the
// caller): with 16-byte loads and a grid stride that is a multiple of N every thread keeps seeing the same four columns, so
// the sums stay in registers; one shared-memory and one global atomic per column and block at the end.
constexpr int kPsBlock = 256;
__global__ void __launch_bounds__(kPsBlock) k_pow2_scales(const float* __restrict__ a, uint64_t na, const float* __restrict__ b, uint64_t nb,
                                                          float* __restrict__ scales, uint32_t N, float* __restrict__ colsum, int vec) {
    __shared__ float s_col[256];
    uint32_t* scratch = reinterpret_cast<uint32_t*>(scales + 4);
    float ma = 0.0f, mb = 0.0f;
    const uint64_t T = (uint64_t)gridDim.x * blockDim.x, t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (colsum) {
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) s_col[i] = 0.0f;
        __syncthreads();
    }
    if (vec) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        const uint64_t na4 = na / 4, nb4 = nb / 4;
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        uint64_t i = t0;
        for (; i + 3 * T < na4; i += 4 * T) {                // four independent 16-byte loads in flight
            const float4 v0 = __ldg(a4 + i), v1 = __ldg(a4 + i + T), v2 = __ldg(a4 + i + 2 * T), v3 = __ldg(a4 + i + 3 * T);
            ma = fmaxf(ma, fmaxf(fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))),
                                 fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w)))));
            ma = fmaxf(ma, fmaxf(fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w))),
                                 fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w)))));
            cs.x += (v0.x + v1.x) + (v2.x + v3.x); cs.y += (v0.y + v1.y) + (v2.y + v3.y);
            cs.z += (v0.z + v1.z) + (v2.z + v3.z); cs.w += (v0.w + v1.w) + (v2.w + v3.w);
        }
        for (; i < na4; i += T) {
            const float4 v = __ldg(a4 + i);
            ma = fmaxf(ma, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
        }
        if (colsum) {                                        // (4 * T) % N == 0 and N % 4 == 0 (checked by the host): fixed columns
            const uint32_t c0 = (uint32_t)((4 * t0) % N);
            atomicAdd(&s_col[c0], cs.x); atomicAdd(&s_col[c0 + 1], cs.y); atomicAdd(&s_col[c0 + 2], cs.z); atomicAdd(&s_col[c0 + 3], cs.w);
        }
        if (t0 == 0)
            for (uint64_t k = na4 * 4; k < na; k++) ma = fmaxf(ma, fabsf(a[k]));      // (no column sums here: na % 4 == 0 when colsum is set)
        i = t0;
        for (; i + 3 * T < nb4; i += 4 * T) {
            const float4 v0 = __ldg(b4 + i), v1 = __ldg(b4 + i + T), v2 = __ldg(b4 + i + 2 * T), v3 = __ldg(b4 + i + 3 * T);
            mb = fmaxf(mb, fmaxf(fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))),
                                 fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w)))));
            mb = fmaxf(mb, fmaxf(fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w))),
                                 fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w)))));
        }
        for (; i < nb4; i += T) {
            const float4 v = __ldg(b4 + i);
            mb = fmaxf(mb, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        if (t0 == 0)
            for (uint64_t k = nb4 * 4; k < nb; k++) mb = fmaxf(mb, fabsf(b[k]));
    } else {
        for (uint64_t i = t0; i < na; i += T) ma = fmaxf(ma, fabsf(__ldg(a + i)));
        for (uint64_t i = t0; i < nb; i += T) mb = fmaxf(mb, fabsf(__ldg(b + i)));
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    }
    if ((threadIdx.x & 31) == 0) {                       // non-negative floats order like their bit patterns
        atomicMax(scratch + 0, __float_as_uint(ma));
        atomicMax(scratch + 1, __float_as_uint(mb));
    }
    __shared__ bool last;
    __syncthreads();
    if (colsum)
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(colsum + i, s_col[i]);
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(scratch + 2, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float s[2];
        for (int k = 0; k < 2; k++) {
            float m = __uint_as_float(atomicAdd(scratch + k, 0u));
            m = fminf(fmaxf(m, 1e-30f), 1e30f);          // clamp_min(1e-30); inf / nan gradients keep a finite scale
            s[k] = exp2f(13.0f - floorf(log2f(m)));
        }
        scales[0] = s[0]; scales[1] = s[1]; scales[2] = 1.0f / (s[0] * s[1]);
        scratch[0] = 0; scratch[1] = 0; scratch[2] = 0;
    }
}

}  // namespace envidr

extern "C" {

int envidr_pow2_scales(const float* a, uint64_t na, const float* b, uint64_t nb, float* scales8, uint32_t N, float* colsum,
                       envidr_stream_t stream) {
    ENVIDR_REQUIRE(a && b && scales8, ENVIDR_E_BADARG, "null pointer");
    const uint64_t n = na > nb ? na : nb;
    uint32_t blocks = (uint32_t)(n / 4096 < 1 ? 1 : (n / 4096 > 8u * envidr::kSMs ? 8u * envidr::kSMs : n / 4096));
    const int vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    if (colsum) {
        // fixed columns per thread need 16-byte loads, N | row length, and a grid stride (4 * blocks * 256 elements) divisible by N
        ENVIDR_REQUIRE(vec && N >= 4 && N <= 256 && (N & (N - 1)) == 0 && na % N == 0, ENVIDR_E_UNSUPPORTED,
                       "pow2_scales: column sums need 16-byte aligned operands and N a power of two in 4..256");
    }
    envidr::k_pow2_scales<<<blocks, envidr::kPsBlock, 0, envidr::as_stream(stream)>>>(a, na, b, nb, scales8, N, colsum, vec);
    envidr::g_launches += 1;
    return envidr::check_launch("pow2_scales");
}

/* partial: [grid, N, K] floats with grid = envidr_wgrad_tc_partials(M); scales (device): {s_dY, s_X, 1 / (s_dY * s_X)}, powers of two. */
uint32_t envidr_wgrad_tc_partials(uint32_t M) {
    const uint32_t stages = (M + envidr::kWgSamples - 1) / envidr::kWgSamples;
    return stages < (uint32_t)envidr::kSMs ? (stages ? stages : 1u) : (uint32_t)envidr::kSMs;
}

int envidr_wgrad_tc(const float* dY, const float* X, uint32_t M, uint32_t N, uint32_t K, const float* scales, float* partial,
                    int variant, envidr_stream_t stream) {
    ENVIDR_REQUIRE(dY && X && scales && partial, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(N >= 1 && N <= 256 && K >= 1 && K <= 256, ENVIDR_E_UNSUPPORTED, "wgrad_tc: 1 <= N, K <= 256");
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(envidr::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)envidr::kWgSmem);
        if (e != cudaSuccess) { envidr::set_error("wgrad_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    const uint32_t grid = envidr_wgrad_tc_partials(M);
    envidr::k_wgrad_tc<<<grid, envidr::kWgThreads, envidr::kWgSmem, envidr::as_stream(stream)>>>(dY, X, M, N, K, envidr::lin_rup(K, 16), scales,
                                                                                               partial, variant);
    envidr::g_launches += 1;
    return envidr::check_launch("wgrad_tc");
}

}  // extern "C"
