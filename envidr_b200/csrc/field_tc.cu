// field_tc.cu -- env_net (IDE -> hidden x (n-1) -> env_feat, evaluated for the normal and the reflected direction of
// every sample) on the 5th-generation tensor cores of sm_100a.
//
// This is >90 % of the FLOPs of the render path (610,304 of 651,008 FLOP/sample at toaster dims).
// The reference runs it as 8 cuBLAS fp32 GEMMs + 8 elementwise launches per render iteration (network.py:527-607).
//
// Design (one persistent CTA per SM, 896 threads, warp-specialised):
//   warp 0      producer: streams the pre-packed weight images of every layer from L2 into a 3-stage x 16 KB shared-memory
//               ring with 1-D bulk async copies (TMA engine, cp.async.bulk -> UBLKCP), as many K steps per stage as fit
//   warp 1      issuer (warp-uniform code, elect.sync inside the asm): tcgen05.mma.cta_group::1 (M = 128 rows, N = layer
//               width) with two fp32 accumulator tiles [128 x 256] ping-ponging in tensor memory (TMEM, 512 columns)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue (2 threads per accumulator row): after each layer read their 32-column slices of the
//               accumulator with tcgen05.ld, apply bias + ReLU, re-split and write the next layer's A operand to shared
//               memory, published per 32-column chunk (8 mbarriers) so that the next layer's MMAs start under the epilogue;
//               the last layer's epilogue unit-normalises the env feature
//   warps 12-27 IDE (16 warps): directional encoding of the NEXT tile straight into a separate layer-0 A-operand buffer
//               while the current tile occupies the tensor pipe
//   Synchronisation is mbarrier-only between roles (full/empty ring, accumulator-ready / IDE-buffer-free via
//   tcgen05.commit, operand-ready via arrive after fence.proxy.async).
//   Issue order: the NEXT tile's layer 0 is issued ahead of this tile's 16-wide last layer (it depends on nothing of this tile and
//   fills the tensor pipe while the epilogue drains the last hidden layer): 8.04 -> 7.61 ms of k_env_tc per frame (run r3_03).
//
// What bounds a tile (clock64 timelines + timing experiments of round 2, ENVIDR_ENV_TC_DEBUG, runs r3_10 / r3_11): a tile takes ~21.5 k
// cycles for 14.2 k cycles of MMA work.  Removing the weight copies altogether saves 2 %, removing the IDE arithmetic 4 %, and issuing ONE
// product instead of three (1/3 of the tensor work) only 10 %: the tile is a latency chain, layer after layer -- commit -> epilogue wakes ->
// tcgen05.ld (TMEM read-out of a 128 x 256 fp32 accumulator: 128 KB, ~64 B/clk) -> convert -> st.shared -> fence -> arrive -> issuer
// wakes -> MMAs -- with a chunk pair published every ~900 cycles whatever the number of epilogue warps (8 or 16, run r3_12), a ring
// handshake of ~300 cycles per K step even without data movement, and a ~7 k-cycle tail (epilogue of the last hidden layer, the 16-wide
// layer, its epilogue, first chunk of the next tile) in which only the next tile's layer 0 has independent MMA work.  Shared memory (A
// operand 128 KB in place) and tensor memory (two 256-column accumulators) are full, so a second tile cannot be interleaved.
// Variants kept behind environment switches, all parity-green, none faster: e4m3 corrections (ENVIDR_ENV_TC_MODE=1), CTA pair
// with cta_group::2 MMAs (ENVIDR_ENV_TC_CTAS=2; round 2: the peer's weight halves arrive by tensor-map TMA that signals the leader's
// barrier, operand-ready by named barrier + one remote arrive, no relay warps: 9.39 -> 8.68 ms, still behind), weight multicast
// across a cluster of 2 / 4 CTAs (ENVIDR_ENV_TC_MULTICAST).
//
// Precision: the reference computes these layers in fp32 and the parity bar is 1e-4 on RGB, which a single fp16/bf16/
// tf32 pass does not meet (measured ~1e-3 on the env feature).  Every operand is split x = hi + lo (hi = fp16(x)).
//   mode 0 (ENVIDR_ENV_TC_MODE=0): three kind::f16 MMAs per K step, hi*hi + lo*hi + hi*lo, fp32 accumulate; the dropped lo*lo
//       term is below 2^-22 relative.  Tensor work = 3x the algorithmic FLOPs.
//   mode 1 (default; hidden layers whose K is a multiple of 32): the two CORRECTION products only need ~4 significant bits
//       (they are 2^-11 of the main product), so they run as kind::f8f6f4 MMAs on e4m3 operands (K = 32 per instruction, twice
//       the fp16 rate), and the main product hi*hi stays kind::f16:
//           phase A  D  = e4m3(lo_a 2^14) * e4m3(w 2^4)^T + e4m3(a 2^3) * e4m3(lo_w 2^15)^T        (corrections x 2^18)
//           phase B  D  = hi_a * (hi_w 2^3)^T + D * 2^-15     (first main MMA: tcgen05.mma scale-input-d = 15), then accumulate
//           epilogue v  = D * 2^-3 + bias
//       Tensor work = 1 + 1/2 + 1/2 = 2x the algorithmic FLOPs instead of 3x, shared-memory operand bytes unchanged (hi16 +
//       hi8 + lo8 = 4 B per element, as hi16 + lo16).  Error of a correction term: 2^-11 (its size) x 2^-4 (e4m3 rounding of each
//       factor) = 2^-15 of one product term, summed incoherently over K: measured on CPU with torch.float8_e4m3fn against
//       float64 (profiles/fp8_correction_sim.py) RGB L-inf 3e-6 (xavier 256-wide) / 1.2e-5 (shipped trained 160-wide env_net)
//       against 3e-7 for mode 0 and 1e-4 for a single fp16 product.  Layer 0 (K = 72 -> 80, the IDE features) stays mode 0.
#include <math.h>
#include <stdio.h>
#include <cuda.h>
#include <cuda_fp8.h>
#include "common.cuh"
#include "ide_tables.cuh"
#include "tc_common.cuh"
#include "field_tc.cuh"

namespace envidr {

// IDE role: 16 warps.  Rows 64-127 of a tile (reflected direction, per-sample kappa = roughness) get 6 threads each (orders
// m = part mod 6, warps 0-11); rows 0-63 (normal direction, constant kappa = diffuse_kappa_inv) get 2 threads each (warps 12-15):
// with the shipped kappa = 0.64 their bands l >= 8 are attenuated by exp(-36 * 0.64) = 1e-10 and below, so only l <= 4 is
// evaluated there (see ide_nb0 in env_tc_launch) and the rest of those rows stays zero.
constexpr int kTcIdeThreads = 512;
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = (4 + kTcEpiWarps) * 32 + kTcIdeThreads;   // 4 control warps + 8 epilogue warps + 16 IDE warps = 896
constexpr int kTcStagesMax = 6;
constexpr uint32_t kTcRingBytes = 49152;           // weight ring: 3 x 16 KB, a stage = one K step (16) of a 256-wide layer, 2 (hi,lo) x 2 chunks x 256 x 16 B.
                                                   // (6 x 8 KB stages holding the hi / lo halves separately measured SLOWER, run r3_04: 8.02 vs 7.61 ms per
                                                   // frame -- the weight stream is bound by L2 bandwidth, not by bytes in flight: 148 SMs x 16 KB per 460
                                                   // cycles is 82 % of the L2 read bandwidth.  Hence the multicast ring below.)
constexpr uint32_t kTcARegion = 65536;             // 128 rows x 256 K x 2 B
constexpr uint32_t kTcIdeRegion = 20480;           // 128 rows x 80 K x 2 B (IDE features, deg_view <= 5, K padded to 16)
__constant__ IdeTables c_ide_tc;
static int g_ide_tc_deg = 0;
// clock64() timeline of CTA 0 (envidr_debug_env_tc_timeline): slot 0 = number of tiles recorded, then per tile kProfPerTile slots:
//   issuer : 0 wait(ide_full) start, 1 end, 2 layer-0 issued, 3 last layer issued, 4 cycles spent in wait(a_rdy), 5 in wait(full)
//   IDE w12: 8 wait(ide_empty) start, 9 end, 10 operand written
//   epilogue w4: 12 + 2*l wait(acc_ready) end, 13 + 2*l layer epilogue done (l < 4)
//   layer 2 in detail: 24 + c issuer's wait(a_rdy[c]) end (c < 8), 32 issuer committed acc_ready, 33 + i / 37 + i: warp 4 / warp 8 published its
//   i-th chunk of layer 1's epilogue (the A operand of layer 2), 41 warp 8 wait(acc_ready) end of layer 1, 42 issuer: first MMA of layer 2 issued
constexpr uint32_t kProfPerTile = 48;
static unsigned long long* g_prof = nullptr;
static uint32_t g_prof_cap = 0;

// fp32 pair -> packed fp16 (hi) pair and packed fp16 (lo = x - hi) pair
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// mode-1 scales (powers of two, exact): activations hi8 x 2^3, lo8 x 2^14; weights hi8 x 2^4, lo8 x 2^15, main hi16 x 2^3
constexpr float kF8ActHi = 8.0f, kF8ActLo = 16384.0f, kF8WHi = 16.0f, kF8WLo = 32768.0f, kF8WMain = 8.0f, kF8DScale = 0.125f;
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}
// kind::f8f6f4 (e4m3 x e4m3 -> fp32, K = 32) and the kind::f16 MMA that first scales the accumulator by 2^-15
__device__ __forceinline__ void mma_f8_ss_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p, e;\n elect.sync _|e, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n"
                 " @e tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_f16_ss_scale15_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n .reg .pred p, e;\n elect.sync _|e, 0xffffffff;\n setp.ne.b32 p, 1, 0;\n"
                 " @e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 15;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}

// Barrier waits of k_env_tc.  Built with -DENVIDR_TC_WATCHDOG (ENVIDR_TC_WATCHDOG=1 python -m envidr_b200.build --force) a wait that lasts
// longer than ~2 s prints who is stuck where (tag: 1 producer/empty, 2-4 issuer/full|a_rdy (mode 1), 5 issuer/ide_full, 6 issuer/full,
// 7 issuer/a_rdy, 8 epilogue/acc_ready, 9 IDE/ide_empty) and traps instead of hanging the GPU.
__device__ __forceinline__ void wait_tag(uint64_t* bar, uint32_t parity, int tag) {
#ifdef ENVIDR_TC_WATCHDOG
    const unsigned long long t0 = clock64();
    while (!tc::mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ull) {
            if ((threadIdx.x & 31) == 0) printf("k_env_tc stuck: block %d warp %d tag %d parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), tag, parity);
            __trap();
        }
    }
#else
    (void)tag;
    tc::mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void wait_tag_cluster(uint64_t* bar, uint32_t parity, int tag) {
    (void)tag;
    tc::mbar_wait_cluster(bar, parity);
}

// CTAS = 1: one CTA per tile, tcgen05 cta_group::1.
// CTAS = 2: thread-block cluster of two CTAs (one TPC) working as a CTA pair: every MMA is M = 256 (128 rows = one tile per
//   CTA), each CTA streams only ITS half of every weight stage (N/2 rows of B) and the tensor cores read the other half from the
//   peer -- halves the weight bytes each SM pulls from L2 and the B-operand shared-memory reads per MMA.  Measured before
//   this change (profiles/r01_run8_9.md): shared-memory bandwidth was the wall (MMA operand reads + TMA writes + epilogue
//   writes ~ 2.6 MB per tile at 128 B/clk = 20 k cycles vs 14.6 k of tensor-pipe time).  The leader (rank 0) issues; the peer's
//   operand-ready signals go to the leader's mbarriers through shared::cluster arrives, completions come back by multicast
//   commit.
// F8 = true (CTAS = 1 only): layers with E.L[l].f8 run in mode 1 (see the file header); F8 = false: mode 0 everywhere.
// MC > 1 (CTAS = 1 only): thread-block cluster of MC CTAs that work on their own tiles with their own cta_group::1 MMAs and share
//   nothing but the WEIGHT STREAM: every ring stage is fetched from L2 once per cluster -- CTA r loads the r-th 1/MC of the stage and
//   multicasts it into the same ring slot of every CTA of the cluster (cp.async.bulk ... .multicast::cluster, complete_tx on every
//   CTA's full barrier), and a slot is refilled when the MMAs of ALL MC CTAs that read it have retired (tcgen05.commit ...
//   .multicast::cluster onto every CTA's empty barrier, count MC).  Why: all 148 SMs stream the same 608 KB of weight images per
//   tile from L2; at a 256 x 256 layer that is 148 x 16 KB per ~460 cycles = 82 % of the L2 read bandwidth, and the layer ran at the
//   supply rate (7.4 k cycles) instead of the MMA rate (6.1 k).  The CTAs stay coupled only through the 3-stage ring.
// SAVE (training forward, csrc/env_train_tc.cu): every hidden layer's post-ReLU activations [2M, N] (fp32, row = branch * M + sample:
//   the operand of the weight-gradient GEMMs) and their ReLU bit masks [2M, N / 32] go to HBM, and the inverse norm of the un-normalised
//   env feature is left in slot 12 of the feature row (the backward of F.normalize needs it).
template <int CTAS, bool F8, int MC, bool SAVE = false>
__global__ void __launch_bounds__(kTcThreads, 1)
k_env_tc(const TcEnv E, const float* __restrict__ rec, float* __restrict__ feat, const uint32_t* __restrict__ M_dev, uint32_t M_host,
         unsigned long long* __restrict__ prof, uint32_t prof_cap, const __grid_constant__ CUtensorMap tmap,
         const __grid_constant__ CUtensorMap tmap1, const TcSave SV, const int32_t* __restrict__ ridx) {
    constexpr int kTcStages = 3;
    constexpr uint32_t kTcStageBytes = kTcRingBytes / kTcStages;
    constexpr int GROUP = CTAS == 2 ? 2 : MC;             // CTAs per cluster
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA_hi = smem;                                   // hidden-layer A operand (written by the epilogues)
    uint8_t* sA_lo = smem + kTcARegion;
    uint8_t* sA_h8 = sA_lo;                                  // mode 1: e4m3(a 2^3)    [128 x 256] (16-element K chunks), 32 KB
    uint8_t* sA_l8 = sA_lo + kTcARegion / 2;                 // mode 1: e4m3(lo_a 2^14), 32 KB
    uint8_t* sI_hi = smem + 2 * kTcARegion;                  // layer-0 A operand (written by the IDE warps)
    uint8_t* sI_lo = sI_hi + kTcIdeRegion;
    uint8_t* ring = sI_lo + kTcIdeRegion;
    float* s_bias = reinterpret_cast<float*>(ring + kTcRingBytes);         // [8 * 256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + kTcMaxLayers * 256);
    uint64_t* full = bars;                           // [stages]  producer -> issuer (TMA bytes landed)
    uint64_t* empty = bars + kTcStagesMax;           // [stages]  issuer -> producer (MMAs reading the stage retired)
    uint64_t* acc_ready = bars + 2 * kTcStagesMax;   // [2] issuer -> epilogue warps (accumulator buffer b of a layer complete)
    uint64_t* ide_full = acc_ready + 2;              // IDE warps -> issuer (kTcIdeThreads arrivals per CTA)
    uint64_t* ide_empty = acc_ready + 3;             // issuer -> IDE warps (layer-0 MMAs retired)
    uint64_t* a_rdy = acc_ready + 4;                 // [8] epilogue warps -> issuer: 32-column chunk c of the next A operand is in
                                                     //     shared memory (128 arrivals per CTA: the 4 warps that own chunk parity c & 1)
    uint64_t* last_read = a_rdy + 8;                 // epilogue group 1 -> issuer: the last layer's accumulator is in registers (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_rdy + 9);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: role branches stay uniform (UR datapath)
    const uint32_t M = M_dev ? *M_dev : M_host;
    const uint32_t n_tiles = (2 * M + 127) / 128;     // 64 samples x 2 directions per tile
    const int nl = (int)E.n_layers;
    // Issue order (producer, issuer): L0(first tile), then per tile L1 .. L(nl-2), L0(NEXT tile), L(nl-1).  The next tile's layer 0
    // depends on nothing of this tile, so its MMAs fill the tensor pipe while the epilogue drains layer nl-2 (the 16-wide last layer
    // needs that whole epilogue and left the pipe idle for ~3 k cycles per tile, timeline r3_02).  Accumulators: layer l <= nl-2 of
    // tile ti -> buffer base(ti) ^ (l & 1); the last layer -> columns [0, 16) of layer nl-2's buffer (its first MMA waits for chunk 0 of
    // that epilogue, which has read columns [0, 32) by then); the next tile's layer 0 -> the other buffer, hence base alternates
    // from tile to tile when nl is even.  Epilogue order is unchanged (L0 .. L(nl-1) per tile).
    const uint32_t base_flip = (uint32_t)((nl & 1) ^ 1);
    const bool early_l0 = E.reorder != 0;                // 0 (ENVIDR_ENV_TC_REORDER=0): plain order L0 .. L(nl-1) per tile, accumulators alternate per layer
    auto acc_buf = [&](uint32_t ti, int l) -> uint32_t {
        if (!early_l0) return (ti * (uint32_t)nl + (uint32_t)l) & 1u;
        const uint32_t base = (ti * base_flip) & 1u;
        return base ^ (uint32_t)((l == nl - 1 ? nl - 2 : l) & 1);
    };
    // work units: tiles (CTAS = 1) or tile pairs (CTAS = 2: CTA `rank` of cluster u takes tile 2 * unit + rank; a tile past
    // n_tiles is all-invalid rows, computed as zeros and never stored)
    const uint32_t rank = GROUP > 1 ? tc::cluster_ctarank() : 0u;
    const uint32_t unit0 = blockIdx.x / GROUP, unit_step = gridDim.x / GROUP;
    const uint32_t n_units = (n_tiles + GROUP - 1) / GROUP;
    if (unit0 >= n_units) return;                     // nothing to do for this CTA / cluster (tail iterations of the render loop)
    if (blockIdx.x != 0) prof = nullptr;
    auto stamp = [&](uint32_t tile_i, uint32_t slot, unsigned long long v) {
        const uint32_t idx = 1 + tile_i * kProfPerTile + slot;
        if (prof && idx < prof_cap) prof[idx] = v;
    };

    if (tid == 0) {
        for (int i = 0; i < kTcStages; i++) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], MC); }
        tc::mbar_init(&acc_ready[0], 1);
        tc::mbar_init(&acc_ready[1], 1);
        // CTAS = 2: the leader's threads arrive on its barriers; the peer's role groups meet on a named barrier and ONE thread sends a
        // cluster-scope arrive to the leader (+1 below) instead of 128 / 512 remote arrives serialising on the leader's barrier
        const uint32_t fwd = (CTAS == 2 && rank == 0) ? 1u : 0u;
        for (int i = 0; i < 8; i++) tc::mbar_init(&a_rdy[i], 128 + fwd);
        tc::mbar_init(last_read, 128);
        tc::mbar_init(ide_full, kTcIdeThreads + fwd);
        tc::mbar_init(ide_empty, 1);
        tc::mbar_fence_init();
    }
    if (warp == 2) { if (CTAS == 2) tc::tmem_alloc2(tmem_slot, 512); else tc::tmem_alloc(tmem_slot, 512); }
    for (uint32_t i = tid; i < 2 * kTcIdeRegion / 16; i += kTcThreads)       // layer-0 operand starts as zeros: K padding and skipped bands
        reinterpret_cast<uint4*>(sI_hi)[i] = make_uint4(0, 0, 0, 0);
    tc::fence_proxy_async_smem();
    for (uint32_t i = tid; i < (uint32_t)nl * 256; i += kTcThreads) {
        const uint32_t l = i >> 8, c = i & 255;
        s_bias[i] = (c < E.L[l].Np) ? __ldg(E.bias + E.L[l].bias_off + c) : 0.0f;
    }
    tc::tc_fence_before();
    if (GROUP > 1) tc::cluster_sync_all(); else __syncthreads();       // barriers of ALL CTAs of the cluster initialised before any remote arrive
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // operand-ready signals.  Leader (and CTAS = 1): every thread of the role group arrives on the CTA's own barrier.  Peer of a CTA
    // pair: the group (`count` threads) meets on named barrier `bar_id` after its proxy fences, then one thread arrives on the LEADER's
    // barrier with release.cluster -- one hop, no relay warp polling in between.
    auto to_issuer = [&](uint64_t* bar) { return (CTAS == 2 && rank == 1) ? tc::map_to_rank(tc::smem_u32(bar), 0) : tc::smem_u32(bar); };
    auto arrive_issuer = [&](uint32_t addr, uint32_t bar_id, uint32_t count, bool first) {
        if (CTAS == 2 && rank == 1) {
            tc::bar_sync_named(bar_id, count);
            if (first) tc::mbar_arrive_cluster(addr);
        } else {
            asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(addr) : "memory");
        }
    };

    if (warp == 0) {
        // ===================== producer =====================
        uint32_t stage = 0, phase = 0;
        auto stream_layer = [&](int l) {
            {
                // a ring stage carries as many K steps as fit in kTcStageBytes (1 for a 256-wide layer, all 16 for the 16-wide
                // last layer: otherwise that layer is bound by 16 ring round trips of 1 KB each).  CTAS = 2: this CTA's half
                // of B (rows rank * N/2 ...) is a contiguous image of its own, so a stage holds twice the K steps.
                auto stream = [&](const uint8_t* src, uint32_t units, uint32_t ubytes) {
                    const uint32_t uper = max(1u, kTcStageBytes / ubytes);
                    for (uint32_t s = 0; s < units; s += uper) {
                        const uint32_t bytes = min(uper, units - s) * ubytes;
                        wait_tag(&empty[stage], phase ^ 1, 1);
                        if (lane == 0) {
                            if (CTAS == 2) {
                                // each CTA pulls ITS half of the stage as 8 KB boxes of the blob's tensor map; both halves complete_tx
                                // on the leader's full barrier, which the leader arms with the bytes of both
                                const uint32_t lfull = tc::map_to_rank(tc::smem_u32(&full[stage]), 0);
                                if (rank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * bytes);
                                const uint32_t row0 = (uint32_t)((src + (size_t)s * ubytes - E.blob) >> 10);
                                // 8 KB boxes, then 1 KB boxes for the remainder (layer widths that are not multiples of 256)
                                uint32_t b = 0;
                                for (; b + 8192 <= bytes; b += 8192)
                                    tc::tma2_load_2d(tc::smem_u32(ring + stage * kTcStageBytes + b), &tmap, 0, (int32_t)(row0 + (b >> 10)), lfull);
                                for (; b < bytes; b += 1024)
                                    tc::tma2_load_2d(tc::smem_u32(ring + stage * kTcStageBytes + b), &tmap1, 0, (int32_t)(row0 + (b >> 10)), lfull);
                            } else if (MC > 1) {
                                // this CTA arms its own barrier with the whole stage and multicasts ITS 1/MC of it to every CTA
                                const uint32_t part = bytes / MC;
                                tc::mbar_arrive_expect_tx(&full[stage], bytes);
                                tc::bulk_g2s_multicast(ring + stage * kTcStageBytes + rank * part, src + (size_t)s * ubytes + rank * part, part,
                                                       &full[stage], (uint16_t)((1u << MC) - 1));
                            } else if (E.debug & 1u) {
                                tc::mbar_arrive(&full[stage]);              // timing experiment: no weight traffic (results are garbage)
                            } else {
                                tc::mbar_arrive_expect_tx(&full[stage], bytes);
                                tc::bulk_g2s(ring + stage * kTcStageBytes, src + (size_t)s * ubytes, bytes, &full[stage]);
                            }
                        }
                        __syncwarp();
                        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
                    }
                };
                if (F8 && E.L[l].f8) {
                    // mode 1: phase A = Kp/32 chunks of (hi8 | lo8) x Np x 32 B, then phase B = Kp/16 steps of hi16 x Np x 32 B
                    const uint32_t chunks = E.L[l].Kp / 32, cbytes = E.L[l].Np * 64;
                    const uint8_t* src = E.blob + E.L[l].img8_off;
                    stream(src, chunks, cbytes);
                    stream(src + (size_t)chunks * cbytes, E.L[l].Kp / 16, E.L[l].Np * 32);
                } else {
                    const uint32_t ksteps = E.L[l].Kp / 16, kbytes = E.L[l].Np * 64 / CTAS;
                    stream(E.blob + (CTAS == 2 ? E.L[l].img2_off + rank * ksteps * kbytes : E.L[l].img_off), ksteps, kbytes);
                }
            }
        };
        for (uint32_t unit = unit0; unit < n_units; unit += unit_step) {
            if (unit == unit0 || !early_l0) stream_layer(0);
            for (int l = 1; l < nl - 1; l++) stream_layer(l);
            if (early_l0 && unit + unit_step < n_units) stream_layer(0);
            stream_layer(nl - 1);
        }
    } else if (warp == 1 && (CTAS == 1 || rank == 0)) {
        // ===================== MMA issuer (leader CTA) =====================
        // All 32 lanes run this code with warp-uniform values; one elected lane issues (tc::*_w helpers).
        // Accumulators ping-pong between TMEM columns [0,256) and [256,512) from layer to layer, so the epilogue of layer g
        // (which produces the next A operand chunk by chunk) overlaps the MMAs of layer g + 1, which start as soon as
        // chunk 0 is in shared memory.  Hazards: MMA(g+2) reuses the accumulator of layer g; it is issued after the last
        // K step of MMA(g+1), which waited for every chunk of epilogue(g) (or, across tiles, behind chunk 0 of
        // epilogue(g+1), which the same warps run after epilogue(g)).
        uint32_t stage = 0, phase = 0, ide_par = 0, chunk_par = 0;
        const uint32_t ring0 = tc::smem_u32(ring);
        unsigned long long w_a = 0, w_f = 0;
        auto issue_layer = [&](int l, uint32_t ti) {
            {
                const uint32_t ksteps = E.L[l].Kp / 16, Np = E.L[l].Np, Nb = Np / CTAS;     // Nb: rows of B held by one CTA
                const uint32_t idesc = tc::make_idesc_f16(128 * CTAS, Np);
                const uint32_t buf = acc_buf(ti, l);
                const uint32_t d_tmem = tmem + buf * 256u;
                if (F8 && E.L[l].f8) {
                    // ---- mode 1: corrections on e4m3 (phase A, chunk by chunk as the epilogue publishes them), then hi16 * hi16
                    uint64_t da_h8 = tc::make_smem_desc(tc::smem_u32(sA_h8), 2048, 128);
                    uint64_t da_l8 = tc::make_smem_desc(tc::smem_u32(sA_l8), 2048, 128);
                    uint64_t da_m = tc::make_smem_desc(tc::smem_u32(sA_hi), 2048, 128);
                    const uint64_t db0 = tc::make_smem_desc(ring0, Np * 16, 128);
                    const uint32_t chunks = ksteps / 2, cbytes = Np * 64, cper = max(1u, kTcStageBytes / cbytes);
                    for (uint32_t c0 = 0; c0 < chunks; c0 += cper) {
                        const unsigned long long t1 = prof ? clock64() : 0;
                        wait_tag(&full[stage], phase, 2);
                        if (prof) w_f += clock64() - t1;
                        const uint32_t cend = min(chunks, c0 + cper);
                        uint64_t db = tc::desc_advance(db0, stage * kTcStageBytes);
                        for (uint32_t c = c0; c < cend; c++) {
                            const unsigned long long t0 = prof ? clock64() : 0;
                            wait_tag(&a_rdy[c], (chunk_par >> c) & 1u, 3);
                            chunk_par ^= 1u << c;
                            if (prof) w_a += clock64() - t0;
                            tc::tc_fence_after();
                            __syncwarp();
                            mma_f8_ss_w(d_tmem, da_l8, db, idesc, c > 0);                              // lo_a * hi_w
                            mma_f8_ss_w(d_tmem, da_h8, tc::desc_advance(db, Np * 32), idesc, 1);       // hi_a * lo_w
                            da_l8 = tc::desc_advance(da_l8, 4096); da_h8 = tc::desc_advance(da_h8, 4096);
                            db = tc::desc_advance(db, cbytes);
                        }
                        tc::mma_commit_w(&empty[stage]);
                        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
                    }
                    const uint32_t sbytes = Np * 32, sper = max(1u, kTcStageBytes / sbytes);
                    for (uint32_t s0 = 0; s0 < ksteps; s0 += sper) {
                        const unsigned long long t1 = prof ? clock64() : 0;
                        wait_tag(&full[stage], phase, 4);
                        if (prof) w_f += clock64() - t1;
                        const uint32_t kend = min(ksteps, s0 + sper);
                        uint64_t db = tc::desc_advance(db0, stage * kTcStageBytes);
                        for (uint32_t s = s0; s < kend; s++) {
                            tc::tc_fence_after();
                            __syncwarp();
                            if (s == 0) mma_f16_ss_scale15_w(d_tmem, da_m, db, idesc);                 // D = hi_a * hi_w + D * 2^-15
                            else tc::mma_f16_ss_w(d_tmem, da_m, db, idesc, 1);
                            da_m = tc::desc_advance(da_m, 4096);
                            db = tc::desc_advance(db, sbytes);
                        }
                        tc::mma_commit_w(&empty[stage]);
                        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
                    }
                    tc::mma_commit_w(&acc_ready[buf]);
                    if (prof && lane == 0 && l == nl - 1) { stamp(ti, 3, clock64()); stamp(ti, 4, w_a); stamp(ti, 5, w_f); prof[0] = ti + 1; w_a = w_f = 0; }
                    return;
                }
                uint64_t da_hi, da_lo;
                if (l == 0) {
                    if (prof && lane == 0) stamp(ti, 0, clock64());
                    if (CTAS == 2) wait_tag_cluster(ide_full, ide_par, 5); else wait_tag(ide_full, ide_par, 5);
                    ide_par ^= 1;                                            // IDE operand of this tile (pair) is in smem
                    if (prof && lane == 0) stamp(ti, 1, clock64());
                    da_hi = tc::make_smem_desc(tc::smem_u32(sI_hi), 2048, 128);
                    da_lo = tc::make_smem_desc(tc::smem_u32(sI_lo), 2048, 128);
                } else {
                    da_hi = tc::make_smem_desc(tc::smem_u32(sA_hi), 2048, 128);
                    da_lo = tc::make_smem_desc(tc::smem_u32(sA_lo), 2048, 128);
                }
                const uint64_t db0 = tc::make_smem_desc(ring0, Nb * 16, 128);
                const uint32_t lo_off = Nb * 32, kbytes = Nb * 64;
                const uint32_t kper = max(1u, kTcStageBytes / kbytes);
                for (uint32_t s0 = 0; s0 < ksteps; s0 += kper) {
                    const unsigned long long t1 = prof ? clock64() : 0;
                    wait_tag(&full[stage], phase, 6);            // CTAS = 2: both CTAs' halves (complete_tx of both TMA streams)
                    if (prof) w_f += clock64() - t1;
                    const uint32_t kend = min(ksteps, s0 + kper);
                    uint64_t db_hi = tc::desc_advance(db0, stage * kTcStageBytes);
                    for (uint32_t s = s0; s < kend; s++) {
                        if (l > 0 && (s & 1u) == 0) {                        // K steps 2c, 2c+1 read chunk c of the A operand
                            const uint32_t c = s >> 1;
                            const unsigned long long t0 = prof ? clock64() : 0;
                            if (CTAS == 2) wait_tag_cluster(&a_rdy[c], (chunk_par >> c) & 1u, 7);
                            else wait_tag(&a_rdy[c], (chunk_par >> c) & 1u, 7);
                            chunk_par ^= 1u << c;
                            if (prof) w_a += clock64() - t0;
                            if (prof && lane == 0 && l == 2 && c < 8) stamp(ti, 24 + c, clock64());
                        }
                        tc::tc_fence_after();
                        __syncwarp();
                        const uint64_t db_lo = tc::desc_advance(db_hi, lo_off);
                        if (CTAS == 2) {
                            tc::mma2_f16_ss_w(d_tmem, da_hi, db_hi, idesc, s > 0);
                            tc::mma2_f16_ss_w(d_tmem, da_lo, db_hi, idesc, 1);
                            tc::mma2_f16_ss_w(d_tmem, da_hi, db_lo, idesc, 1);
                        } else {
                            tc::mma_f16_ss_w(d_tmem, da_hi, db_hi, idesc, s > 0);
                            if (!(E.debug & 4u)) {                            // timing experiment: hi*hi only
                                tc::mma_f16_ss_w(d_tmem, da_lo, db_hi, idesc, 1);
                                tc::mma_f16_ss_w(d_tmem, da_hi, db_lo, idesc, 1);
                            }
                        }
                        da_hi = tc::desc_advance(da_hi, 4096); da_lo = tc::desc_advance(da_lo, 4096);
                        db_hi = tc::desc_advance(db_hi, kbytes);
                    }
                    // frees the ring slot (in both CTAs) when these MMAs retire
                    if (CTAS == 2) tc::mma2_commit_w(&empty[stage]);
                    else if (MC > 1) tc::mma_commit_mc_w(&empty[stage], (uint16_t)((1u << MC) - 1));
                    else tc::mma_commit_w(&empty[stage]);
                    if (++stage == kTcStages) { stage = 0; phase ^= 1; }
                }
                if (CTAS == 2) {
                    if (l == 0) tc::mma2_commit_w(ide_empty);
                    tc::mma2_commit_w(&acc_ready[buf]);
                } else {
                    if (l == 0) tc::mma_commit_w(ide_empty);      // IDE buffer may be refilled for the next tile
                    tc::mma_commit_w(&acc_ready[buf]);
                    if (prof && lane == 0 && l == 2) stamp(ti, 32, clock64());
                }
                if (prof && lane == 0) {
                    if (l == 0) stamp(ti, 2, clock64());
                    if (l == nl - 1) { stamp(ti, 3, clock64()); stamp(ti, 4, w_a); stamp(ti, 5, w_f); prof[0] = ti + 1; }
                }
                if (l == nl - 1) w_a = w_f = 0;
            }
        };
        uint32_t ti = 0, last_par = 0;
        for (uint32_t unit = unit0; unit < n_units; unit += unit_step, ti++) {
            // the last layer of the previous tile sits in an accumulator that one of this tile's layers overwrites; its epilogue is run by
            // group 1 while group 0 already converts this tile's layer 0 (below), so "chunk 0 is ready" no longer implies "read"
            if (CTAS == 1 && ti > 0) { tc::mbar_wait(last_read, last_par); last_par ^= 1; }
            if (unit == unit0 || !early_l0) issue_layer(0, ti);
            for (int l = 1; l < nl - 1; l++) issue_layer(l, ti);
            if (early_l0 && unit + unit_step < n_units) issue_layer(0, ti + 1);
            issue_layer(nl - 1, ti);
        }
    } else if (warp >= 4 && warp < 4 + kTcEpiWarps) {
        // ===================== epilogue warps =====================
        // 8 warps: quarter = TMEM lane quarter (rows), g = chunk parity of the 32-column chunks this warp converts.  (16 warps on
        // 16-column halves -- the layout that paid off in k_neus_geom_tc -- measured no faster here, run r3_12: the chunk cadence stayed at
        // ~900 cycles per chunk pair, i.e. the accumulator read-out is not bound by the per-warp dependent chain.)
        const uint32_t quarter = warp & 3, g = (warp - 4) >> 2, g4 = g;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t branch = row >> 6;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        uint32_t acc_par = 0, ti = 0;
        const bool pw = prof && tid == 4 * 32;
        const uint32_t a_rdy_addr = to_issuer(&a_rdy[0]);
        for (uint32_t unit = unit0; unit < n_units; unit += unit_step, ti++) {
            const uint32_t tile = unit * GROUP + rank;
            const uint32_t m = tile * 64 + (row & 63);
            const bool valid = m < M;
            for (int l = 0; l < nl; l++) {
                const uint32_t buf = acc_buf(ti, l);
                wait_tag(&acc_ready[buf], (acc_par >> buf) & 1u, 8); acc_par ^= 1u << buf;
                if (pw && l < 4) stamp(ti, 12 + 2 * l, clock64());
                if (prof && l == 1 && tid == 8 * 32) stamp(ti, 41, clock64());
                tc::tc_fence_after();
                const uint32_t acc = tmem + lane_addr + buf * 256u;
                const float* bias = s_bias + l * 256;
                const float dsc = E.L[l].dscale;
                if (l < nl - 1 && F8 && E.L[l + 1].f8) {
                    // next layer runs in mode 1: its A operand = hi16 (main) + e4m3(a 2^3) + e4m3((a - hi16) 2^14)
                    const uint32_t nchunks = E.L[l].N / 32;
                    for (uint32_t cb = g; cb < nchunks; cb += 2) {
                        uint32_t r[32];
                        tc::tmem_ld32(acc + cb * 32, r);
                        tc::tmem_ld_wait();
                        const float4* b4 = reinterpret_cast<const float4*>(bias + cb * 32);
                        #pragma unroll
                        for (int jj = 0; jj < 2; jj++) {                  // 16 columns = one e4m3 K chunk = two fp16 K chunks
                            uint32_t h8[4], l8[4];
                            #pragma unroll
                            for (int j2 = 0; j2 < 2; j2++) {
                                const int j = 2 * jj + j2;
                                const float4 ba = b4[2 * j], bb = b4[2 * j + 1];
                                float v[8];
                                v[0] = fmaxf(fmaf(__uint_as_float(r[8 * j + 0]), dsc, ba.x), 0.f); v[1] = fmaxf(fmaf(__uint_as_float(r[8 * j + 1]), dsc, ba.y), 0.f);
                                v[2] = fmaxf(fmaf(__uint_as_float(r[8 * j + 2]), dsc, ba.z), 0.f); v[3] = fmaxf(fmaf(__uint_as_float(r[8 * j + 3]), dsc, ba.w), 0.f);
                                v[4] = fmaxf(fmaf(__uint_as_float(r[8 * j + 4]), dsc, bb.x), 0.f); v[5] = fmaxf(fmaf(__uint_as_float(r[8 * j + 5]), dsc, bb.y), 0.f);
                                v[6] = fmaxf(fmaf(__uint_as_float(r[8 * j + 6]), dsc, bb.z), 0.f); v[7] = fmaxf(fmaf(__uint_as_float(r[8 * j + 7]), dsc, bb.w), 0.f);
                                uint32_t ph[4];
                                float lo[8];
                                #pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                                    const float2 hf = __half22float2(h);
                                    ph[e] = *reinterpret_cast<const uint32_t*>(&h);
                                    lo[2 * e] = (v[2 * e] - hf.x) * kF8ActLo; lo[2 * e + 1] = (v[2 * e + 1] - hf.y) * kF8ActLo;
                                }
                                *reinterpret_cast<uint4*>(sA_hi + tc::op_off(128, row, cb * 32 + j * 8)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                                h8[2 * j2] = pack_e4m3x4(v[0] * kF8ActHi, v[1] * kF8ActHi, v[2] * kF8ActHi, v[3] * kF8ActHi);
                                h8[2 * j2 + 1] = pack_e4m3x4(v[4] * kF8ActHi, v[5] * kF8ActHi, v[6] * kF8ActHi, v[7] * kF8ActHi);
                                l8[2 * j2] = pack_e4m3x4(lo[0], lo[1], lo[2], lo[3]);
                                l8[2 * j2 + 1] = pack_e4m3x4(lo[4], lo[5], lo[6], lo[7]);
                            }
                            const uint32_t off8 = (cb * 2 + jj) * 2048 + row * 16;      // 16-element K chunk (cb * 32 + jj * 16) / 16
                            *reinterpret_cast<uint4*>(sA_h8 + off8) = make_uint4(h8[0], h8[1], h8[2], h8[3]);
                            *reinterpret_cast<uint4*>(sA_l8 + off8) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
                        }
                        tc::tc_fence_before();
                        tc::fence_proxy_async_smem();
                        arrive_issuer(a_rdy_addr + cb * 8, 1 + g, 128, quarter == 0 && lane == 0);
                    }
                } else if (l < nl - 1) {
                    const uint32_t nchunks = E.L[l].N / 32;
                    for (uint32_t cb = g; cb < nchunks; cb += 2) {
                        uint32_t r[32];
                        tc::tmem_ld32(acc + cb * 32, r);
                        tc::tmem_ld_wait();
                        const float4* b4 = reinterpret_cast<const float4*>(bias + cb * 32);
                        uint32_t mk = 0;
                        #pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float4 ba = b4[2 * j], bb = b4[2 * j + 1];
                            uint32_t ph[4], pl[4];
                            const float v0 = fmaxf(__uint_as_float(r[8 * j + 0]) + ba.x, 0.f), v1 = fmaxf(__uint_as_float(r[8 * j + 1]) + ba.y, 0.f);
                            const float v2 = fmaxf(__uint_as_float(r[8 * j + 2]) + ba.z, 0.f), v3 = fmaxf(__uint_as_float(r[8 * j + 3]) + ba.w, 0.f);
                            const float v4 = fmaxf(__uint_as_float(r[8 * j + 4]) + bb.x, 0.f), v5 = fmaxf(__uint_as_float(r[8 * j + 5]) + bb.y, 0.f);
                            const float v6 = fmaxf(__uint_as_float(r[8 * j + 6]) + bb.z, 0.f), v7 = fmaxf(__uint_as_float(r[8 * j + 7]) + bb.w, 0.f);
                            if (SAVE && valid) {
                                float4* dst = reinterpret_cast<float4*>(SV.act[l] + ((size_t)branch * SV.M + m) * E.L[l].N + cb * 32 + j * 8);
                                dst[0] = make_float4(v0, v1, v2, v3);
                                dst[1] = make_float4(v4, v5, v6, v7);
                                mk |= ((v0 > 0.f ? 1u : 0u) | (v1 > 0.f ? 2u : 0u) | (v2 > 0.f ? 4u : 0u) | (v3 > 0.f ? 8u : 0u) | (v4 > 0.f ? 16u : 0u) |
                                       (v5 > 0.f ? 32u : 0u) | (v6 > 0.f ? 64u : 0u) | (v7 > 0.f ? 128u : 0u)) << (8 * j);
                            }
                            split2(v0, v1, ph[0], pl[0]);
                            split2(v2, v3, ph[1], pl[1]);
                            split2(v4, v5, ph[2], pl[2]);
                            split2(v6, v7, ph[3], pl[3]);
                            const uint32_t off = tc::op_off(128, row, cb * 32 + j * 8);
                            *reinterpret_cast<uint4*>(sA_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(sA_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                        if (SAVE && valid) SV.mask[l][((size_t)branch * SV.M + m) * (E.L[l].N / 32) + cb] = mk;
                        tc::tc_fence_before();
                        tc::fence_proxy_async_smem();
                        arrive_issuer(a_rdy_addr + cb * 8, 1 + g, 128, quarter == 0 && lane == 0);         // chunk cb of the next layer's A operand is ready
                        if (prof && l == 1 && lane == 0 && quarter == 0 && (cb >> 1) < 4) stamp(ti, (g ? 37 : 33) + (cb >> 1), clock64());
                    }
                } else if (CTAS == 1 ? g == 1 : g == 0) {
                    // last layer: 16 columns.  Run by group 1 (one CTA per tile): group 0, which owns chunk 0 of every layer, goes straight on
                    // to the next tile's layer-0 epilogue -- its accumulator has been ready since before this tile's last hidden layer was
                    // drained -- so the next tile's layer 1 starts ~1 k cycles earlier.  Both groups have passed acc_ready of this layer, i.e.
                    // its MMAs no longer read the A operand the next epilogue overwrites.
                    uint32_t r[16];
                    tc::tmem_ld16(acc, r);
                    tc::tmem_ld_wait();
                    if (CTAS == 1) { tc::tc_fence_before(); tc::mbar_arrive(last_read); }      // the accumulator may be overwritten from here on
                    const int Ef = (int)E.E;
                    float f[16], ss = 0.f;
                    #pragma unroll
                    for (int i = 0; i < 16; i++) {
                        f[i] = (i < Ef) ? fmaf(__uint_as_float(r[i]), dsc, bias[i]) : 0.0f;
                        ss += f[i] * f[i];
                    }
                    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);            // F.normalize(eps = 1e-12)
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(feat + (size_t)m * kTcRecFloats + 16 * branch);
                        dst[0] = make_float4(f[0] * inv, f[1] * inv, f[2] * inv, f[3] * inv);
                        dst[1] = make_float4(f[4] * inv, f[5] * inv, f[6] * inv, f[7] * inv);
                        dst[2] = make_float4(f[8] * inv, f[9] * inv, f[10] * inv, f[11] * inv);
                        dst[3] = make_float4(SAVE ? inv : f[12] * inv, f[13] * inv, f[14] * inv, f[15] * inv);      // SAVE: E <= 12 (checked at launch)
                    }
                    tc::tc_fence_before();
                }
                if (pw && l < 4) stamp(ti, 13 + 2 * l, clock64());
            }
        }
    } else if (warp >= 4 + kTcEpiWarps) {
        // ===================== IDE warps: directional encoding of the NEXT tile while the current one is in the MMA pipe ====
        const uint32_t t2 = tid - (4 + kTcEpiWarps) * 32;               // 0 .. 511
        const uint32_t branch = t2 < 384 ? 1u : 0u;      // warp-uniform
        const uint32_t part = branch ? t2 >> 6 : (t2 - 384) >> 6;                // 0..5 (reflected) / 0..1 (normal), warp-uniform
        const uint32_t row = branch ? 64 + (t2 & 63) : (t2 - 384) & 63;
        const uint32_t Kp0 = E.L[0].Kp, P = E.P;
        uint32_t empty_par = 1, ti = 0;                  // first wait passes
        const bool pw = prof && tid == (4 + kTcEpiWarps) * 32;
        const uint32_t ide_full_addr = to_issuer(ide_full);
        for (uint32_t unit = unit0; unit < n_units; unit += unit_step, ti++) {
            const uint32_t tile = unit * GROUP + rank;
            const uint32_t m = tile * 64 + (row & 63);
            const bool valid = m < M;
            float dx = 0.f, dy = 0.f, dz = 1.f, kap = 0.f;
            if (valid) {
                const float* q = rec + (size_t)(ridx ? (uint32_t)ridx[m] : m) * kTcRecFloats;      // ridx: records stay where the geometry pass logged them
                dx = q[22 + 3 * branch]; dy = q[23 + 3 * branch]; dz = q[24 + 3 * branch];
                kap = branch ? q[20] : E.kappa_diffuse;
                if (E.has_rot) {                                  // v @ rot_theta[:3,:3] (renderer.py:160-161, 171-172)
                    const float a0 = dx * E.rot[0] + dy * E.rot[3] + dz * E.rot[6], a1 = dx * E.rot[1] + dy * E.rot[4] + dz * E.rot[7],
                                a2 = dx * E.rot[2] + dy * E.rot[5] + dz * E.rot[8];
                    dx = a0; dy = a1; dz = a2;
                }
            }
            if (pw) stamp(ti, 8, clock64());
            wait_tag(ide_empty, empty_par, 9); empty_par ^= 1;
            if (pw) stamp(ti, 9, clock64());
            if (valid && !(E.debug & 2u)) {                   // debug bit 1: timing experiment without the IDE arithmetic
                // layer-0 K order is interleaved (column 2i = Re_i, 2i+1 = Im_i; the weight image is packed with the same
                // permutation, k_pack_tc `interleave`): one packed fp16x2 store per (hi, lo) instead of four 2-byte stores,
                // and the packed conversion instead of scalar ones
                auto emit = [&](int i, float re, float im) {
                    uint32_t hi, lo;
                    split2(re, im, hi, lo);
                    *reinterpret_cast<uint32_t*>(sI_hi + tc::op_off(128, row, 2 * i)) = hi;
                    *reinterpret_cast<uint32_t*>(sI_lo + tc::op_off(128, row, 2 * i)) = lo;
                };
                const float ls = E.light_scale;
                const uint32_t deg = c_ide_tc.deg;
                if (branch) {                              // reflected direction: all bands, 6 threads per row
                    if (deg == 5) {
                        switch (part) {
                            case 0: ide_eval_emit_static<5, 0, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            case 1: ide_eval_emit_static<5, 1, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            case 2: ide_eval_emit_static<5, 2, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            case 3: ide_eval_emit_static<5, 3, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            case 4: ide_eval_emit_static<5, 4, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            default: ide_eval_emit_static<5, 5, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                        }
                    } else if (deg == 4) {
                        switch (part) {
                            case 0: ide_eval_emit_static<4, 0, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            case 1: ide_eval_emit_static<4, 1, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            case 2: ide_eval_emit_static<4, 2, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            case 3: ide_eval_emit_static<4, 3, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            case 4: ide_eval_emit_static<4, 4, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                            default: ide_eval_emit_static<4, 5, 6>(c_ide_tc, dx, dy, dz, kap, ls, emit); break;
                        }
                    } else {
                        ide_eval_emit(c_ide_tc, dx, dy, dz, kap, ls, emit, (int)part, 6);
                    }
                } else {                                   // normal direction: constant kappa, 2 threads per row
                    if (deg == 5 && E.ide_nb0 == 3) {
                        if (part == 0) ide_eval_emit_static<5, 0, 2, 3>(c_ide_tc, dx, dy, dz, kap, ls, emit);
                        else           ide_eval_emit_static<5, 1, 2, 3>(c_ide_tc, dx, dy, dz, kap, ls, emit);
                    } else if (deg == 4 && E.ide_nb0 == 3) {
                        if (part == 0) ide_eval_emit_static<4, 0, 2, 3>(c_ide_tc, dx, dy, dz, kap, ls, emit);
                        else           ide_eval_emit_static<4, 1, 2, 3>(c_ide_tc, dx, dy, dz, kap, ls, emit);
                    } else {
                        ide_eval_emit(c_ide_tc, dx, dy, dz, kap, ls, emit, (int)part, 2);
                    }
                }
            } else if (part == 0) {
                for (uint32_t k = 0; k < Kp0; k += 8) {
                    *reinterpret_cast<uint4*>(sI_hi + tc::op_off(128, row, k)) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(sI_lo + tc::op_off(128, row, k)) = make_uint4(0, 0, 0, 0);
                }
            }
            tc::fence_proxy_async_smem();
            arrive_issuer(ide_full_addr, 3, kTcIdeThreads, t2 == 0);
            if (pw) stamp(ti, 10, clock64());
        }
    }
    tc::tc_fence_before();
    if (GROUP > 1) tc::cluster_sync_all(); else __syncthreads();       // peers may still read this CTA's B half / write its ring / signal its barriers
    if (warp == 2) { if (CTAS == 2) tc::tmem_dealloc2(tmem, 512); else tc::tmem_dealloc(tmem, 512); }
}

// weight image of one layer: for every K step s: [hi: chunk 0 | chunk 1][lo: chunk 0 | chunk 1], chunk = [Np][8] halfs
// `interleave` = P > 0 (layer 0 only): packed K index k < 2P reads source column (k & 1) * P + (k >> 1), i.e. [Re_0, Im_0, Re_1, ...]
__global__ void k_pack_tc(const float* __restrict__ W, const float* __restrict__ b, uint8_t* __restrict__ img, float* __restrict__ bias,
                          uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np, uint32_t interleave) {
    const uint32_t total = Kp * Np;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t n = i / Kp, k = i - n * Kp;
        const uint32_t ks = (interleave && k < 2 * interleave) ? (k & 1u) * interleave + (k >> 1) : k;
        const float v = (n < N && k < K) ? W[(size_t)n * K + ks] : 0.0f;
        __half h, lo;
        tc::split_f16(v, h, lo);
        const uint32_t s = k >> 4, kk = k & 15;
        const size_t base = (size_t)s * Np * 64 + (kk >> 3) * (Np * 16) + n * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + base) = h;
        *reinterpret_cast<__half*>(img + base + (size_t)Np * 32) = lo;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += gridDim.x * blockDim.x) bias[i] = (b && i < N) ? b[i] : 0.0f;
}

// CTA-pair image of one layer: rank r (rows n in [r * Nb, (r + 1) * Nb), Nb = Np / 2) is a contiguous run of K steps, each
// [hi: chunk 0 | chunk 1][lo: chunk 0 | chunk 1] with chunk = [Nb][8] halfs
__global__ void k_pack_tc2(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np,
                           uint32_t interleave) {
    const uint32_t total = Kp * Np, Nb = Np / 2, ksteps = Kp / 16;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t n = i / Kp, k = i - n * Kp;
        const uint32_t ks = (interleave && k < 2 * interleave) ? (k & 1u) * interleave + (k >> 1) : k;
        const float v = (n < N && k < K) ? W[(size_t)n * K + ks] : 0.0f;
        __half h, lo;
        tc::split_f16(v, h, lo);
        const uint32_t r = n / Nb, j = n - r * Nb, s = k >> 4, kk = k & 15;
        const size_t base = (size_t)r * ksteps * Nb * 64 + (size_t)s * Nb * 64 + (kk >> 3) * (Nb * 16) + j * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + base) = h;
        *reinterpret_cast<__half*>(img + base + (size_t)Nb * 32) = lo;
    }
}

// mode-1 image of one layer (see the file header): phase A chunk c (K in [32c, 32c + 32)) = [hi8: 2 K chunks of 16 x Np x 16 B][lo8: same],
// then phase B step s (K in [16s, 16s + 16)) = [hi16 x 2^3: 2 K chunks of 8 x Np x 16 B]
__global__ void k_pack_tc8(const float* __restrict__ W, uint8_t* __restrict__ img, uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np) {
    const uint32_t total = Kp * Np;
    const size_t phase_b = (size_t)(Kp / 32) * Np * 64;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t n = i / Kp, k = i - n * Kp;
        const float v = (n < N && k < K) ? W[(size_t)n * K + k] : 0.0f;
        const __half h = __float2half_rn(v);
        const float lo = v - __half2float(h);
        const uint32_t c = k >> 5, kk = k & 31;
        const size_t a = (size_t)c * Np * 64 + (kk >> 4) * (Np * 16) + n * 16 + (kk & 15);
        img[a] = (uint8_t)__nv_cvt_float_to_fp8(v * kF8WHi, __NV_SATFINITE, __NV_E4M3);
        img[a + (size_t)Np * 32] = (uint8_t)__nv_cvt_float_to_fp8(lo * kF8WLo, __NV_SATFINITE, __NV_E4M3);
        const uint32_t s = k >> 4, k16 = k & 15;
        const size_t b = phase_b + (size_t)s * Np * 32 + (k16 >> 3) * (Np * 16) + n * 16 + (k16 & 7) * 2;
        *reinterpret_cast<__half*>(img + b) = __float2half_rn(__half2float(h) * kF8WMain);
    }
}

// ENVIDR_ENV_TC_MODE=0 selects three fp16 products per K step everywhere (mode 0); default is mode 1 (e4m3 corrections).
int env_tc_mode() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ENVIDR_ENV_TC_MODE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}

static uint32_t rup(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// Lays the tensor-core images out after `base_bytes` of the packed blob.  Returns false if the env_net shape is outside
// what the tensor-core kernel handles (the FFMA path then has to be used).
bool tc_layout(const envidr_field* f, uint64_t base_bytes, TcEnv* out, uint64_t* total_bytes) {
    TcEnv& t = *out;
    t = TcEnv{};
    if (f->n_env < 2 || f->n_env > (uint32_t)kTcMaxLayers) return false;
    const uint32_t P = (1u << f->ide_degree) - 1 + f->ide_degree;
    uint64_t off = rup((uint32_t)base_bytes, 1024);
    uint32_t boff = 0;
    for (uint32_t i = 0; i < f->n_env; i++) {
        const bool last = (i == f->n_env - 1);
        const uint32_t K = f->env[i].in_dim, N = f->env[i].out_dim;
        if (!last && (N % 32 != 0 || N > 256 || N < 64)) return false;      // >= 64: both epilogue groups own a chunk (see k_env_tc)
        if (last && N > 16) return false;
        if (i > 0 && K != f->env[i - 1].out_dim) return false;
        TcLayer& L = t.L[i];
        L.K = K; L.Kp = rup(K, 16); L.N = N; L.Np = last ? 16 : N;
        if (L.Kp > 256 || (i == 0 && L.Kp > 80)) return false;
        L.img_off = (uint32_t)off;
        off += (uint64_t)(L.Kp / 16) * L.Np * 64;
        L.img2_off = (uint32_t)off;                    // CTA-pair image (same size, rows split by rank)
        off += (uint64_t)(L.Kp / 16) * L.Np * 64;
        L.bias_off = boff;
        boff += L.Np;
        // mode 1 for every layer after the first whose K is a whole number of e4m3 MMAs (K = 32); the image exists either way
        L.f8 = (i > 0 && L.Kp % 32 == 0) ? 1u : 0u;
        L.dscale = 1.0f;
        L.img8_off = (uint32_t)off;
        if (L.f8) off += (uint64_t)(L.Kp / 16) * L.Np * 64;
    }
    if (f->env[0].in_dim != 2 * P) return false;
    off = rup((uint32_t)off, 256);
    const uint64_t bias_bytes_off = off;
    off += (uint64_t)boff * 4;
    t.n_layers = f->n_env; t.P = P; t.E = f->env[f->n_env - 1].out_dim;
    t.reorder = 1;
    t.kappa_diffuse = f->diffuse_kappa_inv; t.light_scale = f->light_intensity_scale;
    // bands of the normal-direction encoding whose attenuation exp(-sigma_l * kappa) is below e^-21 = 7.6e-10 are not evaluated
    // (their features stay exactly zero; the reference computes values of that magnitude, 5 orders below the parity bar):
    // sigma_l = l(l+1)/2 = 1, 3, 10, 36, 136 -> with kappa >= 21/36 only l <= 4 (3 bands) remains
    t.ide_nb0 = (f->diffuse_kappa_inv * 36.0f > 21.0f && f->ide_degree >= 4) ? 3u : f->ide_degree;
    if (f->packed) {
        t.blob = reinterpret_cast<const uint8_t*>(f->packed);
        t.bias = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(f->packed) + bias_bytes_off);
    }
    *total_bytes = off;
    return true;
}

int tc_pack(const envidr_field* f, const TcEnv& t, void* packed, cudaStream_t st) {
    uint8_t* blob = reinterpret_cast<uint8_t*>(packed);
    float* bias = const_cast<float*>(reinterpret_cast<const float*>(blob + (reinterpret_cast<const uint8_t*>(t.bias) - t.blob)));
    for (uint32_t i = 0; i < t.n_layers; i++) {
        const TcLayer& L = t.L[i];
        const uint32_t il = (i == 0) ? t.P : 0u;
        k_pack_tc<<<128, 256, 0, st>>>(f->env[i].weight, f->env[i].bias, blob + L.img_off, bias + L.bias_off, L.K, L.N, L.Kp, L.Np, il);
        k_pack_tc2<<<128, 256, 0, st>>>(f->env[i].weight, blob + L.img2_off, L.K, L.N, L.Kp, L.Np, il);
        if (L.f8) k_pack_tc8<<<128, 256, 0, st>>>(f->env[i].weight, blob + L.img8_off, L.K, L.N, L.Kp, L.Np);
    }
    return check_launch("field_pack_tc");
}

constexpr size_t kTcSmem = 2 * kTcARegion + 2 * kTcIdeRegion + kTcRingBytes + kTcMaxLayers * 256 * sizeof(float) + 256;

// ENVIDR_ENV_TC_CTAS=2 selects the CTA-pair kernel.  Default is the single-CTA kernel: measured on B200 (profiles/r01_run12.md) the
// pair kernel is correct but not faster yet -- the per-tile critical path is the epilogue / operand hand-off chain, not the
// shared-memory bandwidth the pair relieves, and the remote (cluster-scope) arrives lengthen that chain.
static CUtensorMap g_no_tmap{};

// Tensor map of the weight blob for the CTA-pair kernel: the blob as a 2-D array of 1 KB rows (256 x uint32), box = 8 rows = 8 KB (one K
// step of one CTA's half of a 256-wide layer; every image offset and every stage size is a multiple of it).  cuTensorMapEncodeTiled is
// resolved through the runtime (no link-time dependency on a driver symbol version); one map per (blob address, size), cached.
static const CUtensorMap* blob_tensor_map(const TcEnv& t) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static const uint8_t* cached_blob = nullptr;
    static uint64_t cached_rows = 0;
    static CUtensorMap cached[2]{};           // [0]: 8 KB boxes, [1]: 1 KB boxes
    const uint64_t rows = ((uint64_t)(reinterpret_cast<const uint8_t*>(t.bias) - t.blob) + 1023) >> 10;
    if (cached_blob == t.blob && cached_rows == rows) return cached;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled not available"); return nullptr; }
        encode = reinterpret_cast<encode_fn>(fn);
    }
    const cuuint64_t gdim[2] = {256, rows};
    const cuuint64_t gstride[1] = {1024};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < 2; i++) {
        const cuuint32_t box[2] = {256, i == 0 ? 8u : 1u};
        CUresult r = encode(&cached[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(t.blob), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); cached_blob = nullptr; return nullptr; }
    }
    cached_blob = t.blob; cached_rows = rows;
    return cached;
}

static int env_tc_ctas() {
    static int v = 0;
    if (!v) { const char* e = getenv("ENVIDR_ENV_TC_CTAS"); v = (e && e[0] == '2') ? 2 : 1; }
    return v;
}

// ENVIDR_ENV_TC_MULTICAST = 1 | 2 | 4: CTAs per weight-multicast cluster of the default kernel (1 = every CTA streams for itself)
static int env_tc_multicast() {
    static int v = 0;
    if (!v) { const char* e = getenv("ENVIDR_ENV_TC_MULTICAST"); v = e ? atoi(e) : 1; if (v != 1 && v != 2 && v != 4) v = 1; }
    return v;
}

int env_tc_launch(const TcEnv& t_in, uint32_t ide_degree, const float* rec, float* feat, const uint32_t* M_dev, uint32_t M_host, cudaStream_t st,
                  const TcSave* save, const int32_t* ridx) {
    if ((int)ide_degree != g_ide_tc_deg) {
        IdeTables tab;
        if (!ide_build_tables((int)ide_degree, &tab)) return ENVIDR_E_UNSUPPORTED;
        cudaError_t e = cudaMemcpyToSymbol(c_ide_tc, &tab, sizeof(tab));
        if (e != cudaSuccess) { set_error("ide tables: %s", cudaGetErrorString(e)); return (int)e; }
        g_ide_tc_deg = (int)ide_degree;
    }
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_env_tc<1, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_tc<1, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_tc<1, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_tc<1, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_tc<2, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e != cudaSuccess) { set_error("env_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    const uint32_t n_tiles_host = (2 * M_host + 127) / 128;
    if (save) {
        // training forward: the default single-CTA kernel with the SAVE epilogue
        ENVIDR_REQUIRE(!M_dev && t_in.E <= 12 && t_in.n_layers >= 2 && t_in.n_layers <= 4, ENVIDR_E_UNSUPPORTED, "env_net training forward: 2..4 layers, env_feat <= 12");
        static bool attr_s = false;
        if (!attr_s) {
            cudaError_t e = cudaFuncSetAttribute(k_env_tc<1, false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
            if (e != cudaSuccess) { set_error("env_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
            attr_s = true;
        }
        const uint32_t grid_s = min((uint32_t)kSMs, n_tiles_host);
        if (grid_s == 0) return 0;
        k_env_tc<1, false, 1, true><<<grid_s, kTcThreads, kTcSmem, st>>>(t_in, rec, feat, nullptr, M_host, nullptr, 0, g_no_tmap, g_no_tmap, *save, nullptr);
        return check_launch("env_tc(train forward)");
    }
    static int reorder = -1;
    if (reorder < 0) { const char* e = getenv("ENVIDR_ENV_TC_REORDER"); reorder = (e && e[0] == '0') ? 0 : 1; }
    TcEnv t = t_in;
    t.reorder = (uint32_t)reorder;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("ENVIDR_ENV_TC_DEBUG"); dbg = e ? atoi(e) : 0; }      // timing experiments only, see TcEnv::debug
    t.debug = (uint32_t)dbg;
    if (env_tc_ctas() == 2) {
        const CUtensorMap* tmap = blob_tensor_map(t);
        if (!tmap) return ENVIDR_E_UNSUPPORTED;
        uint32_t grid = kSMs & ~1u;                    // whole CTA pairs
        if (!M_dev) grid = min(grid, 2 * ((n_tiles_host + 1) / 2));
        if (grid == 0) return 0;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = kTcSmem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_env_tc<2, false, 1>, t, rec, feat, M_dev, M_host, g_prof, g_prof_cap, tmap[0], tmap[1], TcSave{}, ridx);
        if (e != cudaSuccess) { set_error("env_tc (CTA pair) launch: %s", cudaGetErrorString(e)); return (int)e; }
        return check_launch("env_tc2");
    }
    uint32_t grid = kSMs;
    if (!M_dev) grid = min((uint32_t)kSMs, n_tiles_host);
    if (grid == 0) return 0;
    if (env_tc_mode() == 1) {
        TcEnv t8 = t;
        for (uint32_t i = 0; i < t8.n_layers; i++) if (t8.L[i].f8) t8.L[i].dscale = kF8DScale;
        k_env_tc<1, true, 1><<<grid, kTcThreads, kTcSmem, st>>>(t8, rec, feat, M_dev, M_host, g_prof, g_prof_cap, g_no_tmap, g_no_tmap, TcSave{}, ridx);
    } else if (env_tc_multicast() > 1) {
        // weight-multicast clusters (opt-in; measured no faster, run r3_09): whole clusters only; a cluster whose tiles are all past the end returns at once
        const uint32_t mc = (uint32_t)env_tc_multicast();
        grid = kSMs / mc * mc;
        if (!M_dev) grid = min(grid, (n_tiles_host + mc - 1) / mc * mc);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = kTcSmem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = mc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = mc == 4 ? cudaLaunchKernelEx(&cfg, k_env_tc<1, false, 4>, t, rec, feat, M_dev, M_host, g_prof, g_prof_cap, g_no_tmap, g_no_tmap, TcSave{}, ridx)
                                : cudaLaunchKernelEx(&cfg, k_env_tc<1, false, 2>, t, rec, feat, M_dev, M_host, g_prof, g_prof_cap, g_no_tmap, g_no_tmap, TcSave{}, ridx);
        if (e != cudaSuccess) { set_error("env_tc (multicast cluster) launch: %s", cudaGetErrorString(e)); return (int)e; }
    } else {
        k_env_tc<1, false, 1><<<grid, kTcThreads, kTcSmem, st>>>(t, rec, feat, M_dev, M_host, g_prof, g_prof_cap, g_no_tmap, g_no_tmap, TcSave{}, ridx);
    }
    return check_launch("env_tc");
}

}  // namespace envidr

extern "C" int envidr_debug_env_tc_timeline(uint64_t* buf, uint32_t capacity) {
    envidr::g_prof = reinterpret_cast<unsigned long long*>(buf);
    envidr::g_prof_cap = buf ? capacity : 0;
    return 0;
}
