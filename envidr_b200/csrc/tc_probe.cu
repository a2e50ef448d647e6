// tc_probe.cu -- minimal tcgen05 GEMM used by the tests to pin the operand-layout / descriptor conventions
// of tc_common.cuh on real hardware:  D[128 x N] (fp32) = A[128 x K] * B[N x K]^T, fp16 operands, one CTA.
#include "common.cuh"
#include "tc_common.cuh"

namespace envidr {

__global__ void __launch_bounds__(128, 1) k_tc_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                    uint32_t N, uint32_t K, uint32_t variant) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sA = smem;                    // [K/8][128][8] halfs
    uint8_t* sB = smem + 128 * K * 2;      // [K/8][N][8] halfs
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    uint32_t ncols = 32;
    while (ncols < N) ncols <<= 1;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
    for (uint32_t i = tid; i < 128 * K; i += 128) {
        const uint32_t r = i / K, k = i % K;
        *reinterpret_cast<__half*>(sA + tc::op_off(128, r, k)) = __float2half_rn(A[i]);
    }
    for (uint32_t i = tid; i < N * K; i += 128) {
        const uint32_t r = i / K, k = i % K;
        *reinterpret_cast<__half*>(sB + tc::op_off(N, r, k)) = __float2half_rn(B[i]);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(128, N);
        for (uint32_t s = 0; s < K / 16; s++) {
            const uint32_t a_addr = tc::smem_u32(sA) + s * 2 * (128 * 16);
            const uint32_t b_addr = tc::smem_u32(sB) + s * 2 * (N * 16);
            uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
            if (variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
            tc::mma_f16_ss(tmem, tc::make_smem_desc(a_addr, a_lbo, a_sbo), tc::make_smem_desc(b_addr, b_lbo, b_sbo), idesc, s > 0);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    const uint32_t row = tid;
    for (uint32_t c = 0; c < N; c += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tmem + ((warp * 32u) << 16) + c, r);
        tc::tmem_ld_wait();
        for (uint32_t j = 0; j < 32 && c + j < N; j++) D[row * N + c + j] = __uint_as_float(r[j]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace envidr

using namespace envidr;

extern "C" int envidr_tc_probe(const float* A, const float* B, float* D, uint32_t N, uint32_t K, uint32_t variant, envidr_stream_t stream) {
    ENVIDR_REQUIRE(A && B && D, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K <= 256 && K % 16 == 0, ENVIDR_E_UNSUPPORTED, "N, K must be multiples of 16, <= 256");
    const size_t smem = (size_t)(128 + N) * K * 2;
    cudaError_t e = cudaFuncSetAttribute(k_tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("tc_probe smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    k_tc_probe<<<1, 128, smem, as_stream(stream)>>>(A, B, D, N, K, variant);
    return check_launch("tc_probe");
}
