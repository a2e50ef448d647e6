// tc_common.cuh -- thin inline-PTX layer for the Blackwell tensor-core path (sm_100a):
// mbarrier, 1-D bulk async copy (TMA engine, UBLKCP), tcgen05 alloc / mma / commit / ld, UMMA descriptors.
//
// Operand layout used throughout (no swizzle, K-major, "chunk-major" packing):
//   element (r, k) of an [R x K] fp16 operand lives at byte  (k / 8) * (R * 16) + r * 16 + (k % 8) * 2
// i.e. 8x16-byte core matrices (8 rows x 8 halfs, 128 contiguous bytes), consecutive 8-row groups 128 B apart
// (stride byte offset, SBO) and consecutive 8-element K chunks R*16 B apart (leading byte offset, LBO).
// One tcgen05.mma.kind::f16 consumes K = 16, i.e. two K chunks; the next K step starts 2 * R * 16 bytes further.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace envidr {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// ---- bulk async copy global -> shared (TMA engine, 1-D) ----------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same copy delivered to the same shared-memory offset of every CTA in `cta_mask` of the cluster, complete_tx on the barrier at the
// same offset in each of them
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp, ncols power of 2 >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): no swizzle, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: A,B = fp16, D = fp32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives columns [c, c+32) of TMEM lane (lane_base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---- warp-uniform issue helpers -----------------------------------------------------------------------------
// Called by ALL lanes of a converged warp with warp-uniform operands; one elected lane issues.  Keeping the call
// site free of `if (lane == 0)` lets the compiler hold descriptors in uniform registers instead of wrapping every
// operand in an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence (measured: ~90 SASS instructions per K step before).
__device__ __forceinline__ void mma_f16_ss_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p, e;\n elect.sync _|e, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n"
                 " @e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit_w(uint64_t* bar) {
    asm volatile("{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n"
                 " @e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
// commit that arrives on the barrier at the same offset in every CTA of `cta_mask` (cta_group::1 MMAs of this CTA)
__device__ __forceinline__ void mma_commit_mc_w(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n"
                 " @e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// descriptor of the same operand `bytes` further in shared memory (address field is in 16-byte units)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// ---- thread-block cluster / CTA-pair (cta_group::2) helpers ---------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {          // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_u32, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_u32), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {     // release at cluster scope, possibly remote
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope (remote arrivals)
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {  // same warp id in both CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA over the CTA pair: each CTA supplies its 128 rows of A and its half (N/2 rows) of B from the same shared-memory
// offsets; each CTA's tensor memory receives the accumulator rows of its own A tile.  Issued by the leader CTA only.
__device__ __forceinline__ void mma2_f16_ss_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p, e;\n elect.sync _|e, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n"
                 " @e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (when all prior MMAs of this thread retire) on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void mma2_commit_w(uint64_t* bar) {
    asm volatile("{\n .reg .pred e;\n .reg .b16 m;\n mov.b16 m, 3;\n elect.sync _|e, 0xffffffff;\n"
                 " @e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n}"
                 ::"r"(smem_u32(bar)) : "memory");
}

// ---- CTA-pair weight loads: 2-D tiled TMA whose completion is signalled on the LEADER's mbarrier -------------------
// cp.async.bulk.tensor with .cta_group::2 may complete_tx on a barrier of either CTA of the pair (a plain cp.async.bulk can only signal
// a barrier of the CTA that owns the destination), so the peer's half of a weight stage needs no relay warp.
__device__ __forceinline__ void tma2_load_2d(uint32_t smem_dst_u32, const void* tmap, int32_t c0, int32_t c1, uint32_t mbar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_dst_u32), "l"(tmap), "r"(c0), "r"(c1), "r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_u32(uint32_t bar_u32, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar_u32), "r"(bytes) : "memory");
}
// named barrier over `count` threads (multiple of 32) of this CTA
__device__ __forceinline__ void bar_sync_named(uint32_t id, uint32_t count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// fp32 -> (hi, lo) fp16 pair with hi + lo ~= x to ~2^-22 relative (lo may be subnormal: absolute error <= 3e-8)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
// fp32 pair -> packed fp16 (hi) pair and packed fp16 (lo = x - hi) pair
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// byte offset of element (r, k) inside a chunk-major [R x K] fp16 operand
__host__ __device__ constexpr uint32_t op_off(uint32_t R, uint32_t r, uint32_t k) { return (k >> 3) * (R * 16) + r * 16 + (k & 7) * 2; }

// Epilogue of a hidden layer for one 32-column chunk of one accumulator row: v = relu(D + bias) -> fp16 hi/lo ->
// next A operand (128-row chunk-major buffers dst_hi / dst_lo).  Returns the ReLU mask of the 32 columns.
__device__ __forceinline__ uint32_t hidden_epilogue32(uint32_t taddr, const float* bias32, uint8_t* dst_hi, uint8_t* dst_lo, uint32_t row,
                                                     uint32_t col0) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    tmem_ld_wait();
    uint32_t mk = 0;
    #pragma unroll
    for (int jj = 0; jj < 4; jj++) {
        uint32_t ph[4], pl[4];
        #pragma unroll
        for (int e = 0; e < 4; e++) {
            const int c = jj * 8 + e * 2;
            const float v0 = __uint_as_float(r[c]) + bias32[c], v1 = __uint_as_float(r[c + 1]) + bias32[c + 1];
            mk |= (v0 > 0.f ? 1u : 0u) << c;
            mk |= (v1 > 0.f ? 1u : 0u) << (c + 1);
            split2(fmaxf(v0, 0.f), fmaxf(v1, 0.f), ph[e], pl[e]);
        }
        const uint32_t off = op_off(128, row, col0 + jj * 8);
        *reinterpret_cast<uint4*>(dst_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(dst_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
    return mk;
}
// write 8 consecutive K values of one row of an A operand
__device__ __forceinline__ void store_chunk8(uint8_t* dst_hi, uint8_t* dst_lo, uint32_t row, uint32_t k0, const float (&v)[8]) {
    uint32_t ph[4], pl[4];
    #pragma unroll
    for (int e = 0; e < 4; e++) split2(v[2 * e], v[2 * e + 1], ph[e], pl[e]);
    const uint32_t off = op_off(128, row, k0);
    *reinterpret_cast<uint4*>(dst_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(dst_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

}  // namespace tc
}  // namespace envidr
