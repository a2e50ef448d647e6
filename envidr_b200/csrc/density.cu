// density.cu -- occupancy-grid maintenance: NeRFRenderer.update_extra_state and
// NeRFRenderer.mark_untrained_grid (reference nerf/renderer.py:200-262, 264-352) as four launches around the
// geometry kernel of the render path:
//
//   k_density_points   cell (Morton order) -> jittered query position          (renderer.py:290-301 / 320-331)
//   geometry kernel    hash grid + sdf_net + Laplace density  (field_forward_launch, mode 1: the kernel the march loop uses)
//   k_density_scatter  partial update only: tmp_grid[morton(coords)] = sigma  (renderer.py:335-336)
//   k_density_ema      max(grid * decay, tmp) where both are valid; sum of clamp(grid, 0); the last block to finish turns
//                      the per-block sums into mean_density and the packbits threshold (renderer.py:343-351) -- on the device,
//                      where the reference does a .item() round trip
//   k_density_pack     packbits with the device-side threshold                 (raymarching.cu:267-289)
//
// The reference walks the cells in meshgrid order and scatters through morton3D indices (`tmp_grid[cas, indices] = sigmas`);
// here the full update enumerates the cells in Morton order, so the sigma vector the geometry kernel writes IS tmp_grid and the
// EMA / pack passes stream.  The jitter noise keeps the reference's meshgrid indexing (noise[cas][x*H*H + y*H + z][3]), so a
// caller that draws it with torch.rand in the reference's call order reproduces the reference's query positions bit for bit.
// All of this is HBM/L2-bound streaming except the geometry kernel (1,024 B of L2 gathers + 14 kFLOP per cell).
#include "common.cuh"

namespace envidr {

struct RecCapture;                                                                 // field_tc.cuh (unused here)
int field_forward_launch(const envidr_field* field, const float* xyzs, const float* dirs, const float* r_images,
                         const uint32_t* M_dev, uint32_t M_host, int mode, const envidr_field_out* out, cudaStream_t st,
                         cudaEvent_t* ev, int* ev_recorded, const RecCapture* cap);

namespace {

constexpr int kDBlock = 256;

// query position of one cell: the reference's torch expression, one rounding per torch op (no contraction)
//   xyzs = 2 * coords.float() / (H - 1) - 1;  cas_xyzs = xyzs * (bound - hgs);  cas_xyzs += (rand * 2 - 1) * hgs
// torch's CUDA true-divide by a Python scalar multiplies by the fp32 reciprocal (ATen BinaryDivTrueKernel.cu: a * (1 / b)),
// and the reference runs these lines on the GPU, so that is the arithmetic followed here (rHm1 = 1.0f / (H - 1)).
__device__ __forceinline__ float cell_pos(uint32_t c, float rHm1, float span, float hgs, float u, bool jitter) {
    float x = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c), rHm1), 1.0f);
    x = __fmul_rn(x, span);
    if (jitter) x = __fadd_rn(x, __fmul_rn(__fsub_rn(__fmul_rn(u, 2.0f), 1.0f), hgs));
    return x;
}

// full update: thread t = Morton index of the cell.  coords != nullptr: partial update, thread j = j-th listed cell.
__global__ void __launch_bounds__(kDBlock) k_density_points(uint32_t n, uint32_t H, float span, float hgs, const int32_t* __restrict__ coords,
                                                            const float* __restrict__ noise, float* __restrict__ xyz,
                                                            int32_t* __restrict__ idx) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint32_t cx, cy, cz, j;
    if (coords) {
        cx = (uint32_t)coords[3 * (size_t)t]; cy = (uint32_t)coords[3 * (size_t)t + 1]; cz = (uint32_t)coords[3 * (size_t)t + 2];
        j = t;
        idx[t] = (int32_t)morton3(cx, cy, cz);
    } else {
        cx = compact3(t); cy = compact3(t >> 1); cz = compact3(t >> 2);
        j = (cx * H + cy) * H + cz;                       // the reference's meshgrid (ij) order: the noise stream is indexed by it
    }
    float u0 = 0, u1 = 0, u2 = 0;
    if (noise) { u0 = __ldg(noise + 3 * (size_t)j); u1 = __ldg(noise + 3 * (size_t)j + 1); u2 = __ldg(noise + 3 * (size_t)j + 2); }
    const float Hm1 = __frcp_rn((float)(H - 1));
    xyz[3 * (size_t)t + 0] = cell_pos(cx, Hm1, span, hgs, u0, noise != nullptr);
    xyz[3 * (size_t)t + 1] = cell_pos(cy, Hm1, span, hgs, u1, noise != nullptr);
    xyz[3 * (size_t)t + 2] = cell_pos(cz, Hm1, span, hgs, u2, noise != nullptr);
}

__global__ void __launch_bounds__(kDBlock) k_density_fill(float* __restrict__ p, uint32_t n, float v) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = v;
}

// duplicates: one of the writers wins, as with the reference's index_put (renderer.py:336)
__global__ void __launch_bounds__(kDBlock) k_density_scatter(const int32_t* __restrict__ idx, const float* __restrict__ sigma, uint32_t n,
                                                             float* __restrict__ tmp) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) tmp[idx[t]] = sigma[t];
}

struct DensityStats { float mean, thresh; };

// 4 cells per thread (float4 in, float4 out); block sums in double, combined in block order by the last block: deterministic.
__global__ void __launch_bounds__(kDBlock) k_density_ema(float4* __restrict__ grid, const float4* __restrict__ tmp, uint32_t n4, float decay,
                                                         float density_thresh, double inv_cells, double* __restrict__ partial,
                                                         uint32_t* __restrict__ ticket, DensityStats* __restrict__ stats) {
    __shared__ double red[kDBlock / 32];
    __shared__ bool last;
    double s = 0.0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += gridDim.x * blockDim.x) {
        float4 g = grid[t];
        const float4 m = __ldg(tmp + t);
        // valid_mask = (density_grid >= 0) & (tmp_grid >= 0); grid[valid] = maximum(grid[valid] * decay, tmp[valid])
        if (g.x >= 0.0f && m.x >= 0.0f) g.x = fmaxf(__fmul_rn(g.x, decay), m.x);
        if (g.y >= 0.0f && m.y >= 0.0f) g.y = fmaxf(__fmul_rn(g.y, decay), m.y);
        if (g.z >= 0.0f && m.z >= 0.0f) g.z = fmaxf(__fmul_rn(g.z, decay), m.z);
        if (g.w >= 0.0f && m.w >= 0.0f) g.w = fmaxf(__fmul_rn(g.w, decay), m.w);
        grid[t] = g;
        s += (double)fmaxf(g.x, 0.0f) + (double)fmaxf(g.y, 0.0f) + (double)fmaxf(g.z, 0.0f) + (double)fmaxf(g.w, 0.0f);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < kDBlock / 32; w++) b += red[w];
        partial[blockIdx.x] = b;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {                                              // block-parallel, fixed-order combination of the per-block sums
        __threadfence();
        __shared__ double fin[kDBlock];
        double t = 0.0;
        for (uint32_t b = threadIdx.x; b < gridDim.x; b += kDBlock) t += __ldcg(partial + b);
        fin[threadIdx.x] = t;
        __syncthreads();
        for (int o = kDBlock / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) fin[threadIdx.x] += fin[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const float mean = (float)(fin[0] * inv_cells);     // torch.mean(density_grid.clamp(min=0))
            stats->mean = mean;
            stats->thresh = fminf(mean, density_thresh);         // density_thresh = min(self.mean_density, self.density_thresh)
            *ticket = 0;
        }
    }
}

__global__ void __launch_bounds__(kDBlock) k_density_pack(const float4* __restrict__ grid, uint32_t n_bytes, const DensityStats* __restrict__ stats,
                                                          uint8_t* __restrict__ bitfield) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_bytes) return;
    const float th = stats->thresh;
    const float4 a = __ldg(grid + 2 * (size_t)n), b = __ldg(grid + 2 * (size_t)n + 1);
    uint32_t bits = 0;
    bits |= (a.x > th) ? 1u : 0u;   bits |= (a.y > th) ? 2u : 0u;   bits |= (a.z > th) ? 4u : 0u;   bits |= (a.w > th) ? 8u : 0u;
    bits |= (b.x > th) ? 16u : 0u;  bits |= (b.y > th) ? 32u : 0u;  bits |= (b.z > th) ? 64u : 0u;  bits |= (b.w > th) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// mark_untrained_grid (renderer.py:200-262): a cell is "trained" when its query point projects into at least one camera.
// One thread per (cascade, Morton cell); the poses sit in shared memory.  The arithmetic follows the torch expression
// (fp32, one rounding per op; the 3-term products of `cam_xyzs @ R` are summed left to right without contraction).
constexpr int kMaxPosesSmem = 1024;
__global__ void __launch_bounds__(kDBlock) k_mark_untrained(const float* __restrict__ poses, uint32_t B, float kx, float ky, float span, float hgs2,
                                                            uint32_t H, uint32_t cells, float* __restrict__ grid, int32_t* __restrict__ count_out) {
    extern __shared__ float sp[];                       // [B][12]: R (row-major 3x3) then t
    for (uint32_t i = threadIdx.x; i < B * 12; i += blockDim.x) {
        const uint32_t b = i / 12, k = i % 12;
        sp[i] = k < 9 ? poses[16 * (size_t)b + 4 * (k / 3) + (k % 3)] : poses[16 * (size_t)b + 4 * (k - 9) + 3];
    }
    __syncthreads();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cells) return;
    const float Hm1 = __frcp_rn((float)(H - 1));
    const float wx = cell_pos(compact3(t), Hm1, span, 0.0f, 0.0f, false);
    const float wy = cell_pos(compact3(t >> 1), Hm1, span, 0.0f, 0.0f, false);
    const float wz = cell_pos(compact3(t >> 2), Hm1, span, 0.0f, 0.0f, false);
    int cnt = 0;
    for (uint32_t b = 0; b < B; b++) {
        const float* R = sp + 12 * b;
        const float px = __fsub_rn(wx, R[9]), py = __fsub_rn(wy, R[10]), pz = __fsub_rn(wz, R[11]);
        // cam = p @ R  (c2w rotation applied from the right = world -> camera)
        const float cxm = __fadd_rn(__fadd_rn(__fmul_rn(px, R[0]), __fmul_rn(py, R[3])), __fmul_rn(pz, R[6]));
        const float cym = __fadd_rn(__fadd_rn(__fmul_rn(px, R[1]), __fmul_rn(py, R[4])), __fmul_rn(pz, R[7]));
        const float czm = __fadd_rn(__fadd_rn(__fmul_rn(px, R[2]), __fmul_rn(py, R[5])), __fmul_rn(pz, R[8]));
        const bool mz = czm > 0.0f;
        const bool mx = fabsf(cxm) < __fadd_rn(__fmul_rn(kx, czm), hgs2);
        const bool my = fabsf(cym) < __fadd_rn(__fmul_rn(ky, czm), hgs2);
        cnt += (mz && mx && my) ? 1 : 0;
    }
    if (count_out) count_out[t] += cnt;
    else if (cnt == 0) grid[t] = -1.0f;
}

__global__ void __launch_bounds__(kDBlock) k_mark_apply(const int32_t* __restrict__ count, uint32_t n, float* __restrict__ grid) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && count[t] == 0) grid[t] = -1.0f;
}

uint64_t al256(uint64_t v) { return (v + 255) / 256 * 256; }

struct DWs { uint64_t xyz, sigma, idx, tmp, partial, ticket, total; };
DWs density_ws(uint32_t C, uint32_t H, uint32_t n) {
    DWs w;
    const uint64_t cells = (uint64_t)C * H * H * H;
    uint64_t o = 0;
    w.xyz = o;     o += al256((uint64_t)n * 12);
    w.sigma = o;   o += al256((uint64_t)n * 4);
    w.idx = o;     o += al256((uint64_t)n * 4);
    w.tmp = o;     o += al256(cells * 4);
    w.partial = o; o += al256((uint64_t)kSMs * 8 * 8);
    w.ticket = o;  o += 256;
    w.total = o;
    return w;
}

}  // namespace
}  // namespace envidr

using namespace envidr;

extern "C" {

uint64_t envidr_density_workspace_bytes(uint32_t cascade, uint32_t grid_size, uint32_t n) {
    return density_ws(cascade, grid_size, n).total;
}

int envidr_density_grid_update(const envidr_field* field, float* density_grid, const int32_t* coords, const float* noise, uint32_t n,
                               const envidr_density_opts* opts, uint8_t* bitfield, float* stats, void* workspace,
                               uint64_t workspace_bytes, envidr_stream_t stream) {
    ENVIDR_REQUIRE(field && density_grid && opts && bitfield && stats && workspace, ENVIDR_E_BADARG, "null pointer");
    const uint32_t C = opts->cascade, H = opts->grid_size;
    ENVIDR_REQUIRE(C >= 1 && C <= 8 && H >= 2 && H <= 1024 && (H & (H - 1)) == 0, ENVIDR_E_BADARG, "cascade in 1..8, grid_size a power of two <= 1024");
    const uint32_t cells1 = H * H * H;
    if (!coords) ENVIDR_REQUIRE(n == cells1, ENVIDR_E_BADARG, "full update (coords == NULL) visits n = grid_size^3 cells per cascade");
    ENVIDR_REQUIRE((reinterpret_cast<uintptr_t>(density_grid) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                   ENVIDR_E_BADARG, "density_grid must be 16-byte aligned, workspace 256-byte aligned");
    const DWs w = density_ws(C, H, n);
    ENVIDR_REQUIRE(workspace_bytes >= w.total, ENVIDR_E_WORKSPACE, "workspace too small (envidr_density_workspace_bytes)");
    cudaStream_t st = as_stream(stream);
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
    float* xyz = reinterpret_cast<float*>(base + w.xyz);
    float* sigma = reinterpret_cast<float*>(base + w.sigma);
    int32_t* idx = reinterpret_cast<int32_t*>(base + w.idx);
    float* tmp = reinterpret_cast<float*>(base + w.tmp);
    double* partial = reinterpret_cast<double*>(base + w.partial);
    uint32_t* ticket = reinterpret_cast<uint32_t*>(base + w.ticket);
    const uint64_t cells = (uint64_t)C * cells1;
    cudaMemsetAsync(ticket, 0, 4, st);
    if (coords) k_density_fill<<<ceil_div((uint32_t)cells, kDBlock), kDBlock, 0, st>>>(tmp, (uint32_t)cells, -1.0f);   // tmp_grid = -ones_like
    for (uint32_t cas = 0; cas < C; cas++) {
        const double bound = fmin((double)(1u << cas), (double)opts->bound);   // bound = min(2 ** cas, self.bound)
        const double hgs = bound / H;                                           // half_grid_size = bound / self.grid_size
        if (n == 0) continue;
        k_density_points<<<ceil_div(n, kDBlock), kDBlock, 0, st>>>(n, H, (float)(bound - hgs), (float)hgs,
                                                                  coords ? coords + 3 * (size_t)cas * n : nullptr,
                                                                  noise ? noise + 3 * (size_t)cas * n : nullptr, xyz, idx);
        envidr_field_out fo = {};
        fo.sigma = coords ? sigma : tmp + (size_t)cas * cells1;                 // out.sigma = density * density_scale (renderer.py:303-304)
        int rc = field_forward_launch(field, xyz, xyz, nullptr, nullptr, n, 1, &fo, st, nullptr, nullptr, nullptr);
        if (rc) return rc;
        if (coords) k_density_scatter<<<ceil_div(n, kDBlock), kDBlock, 0, st>>>(idx, sigma, n, tmp + (size_t)cas * cells1);
        g_launches += coords ? 3 : 2;
    }
    const uint32_t n4 = (uint32_t)(cells / 4);
    const uint32_t blocks = min((uint32_t)kSMs * 8, ceil_div(n4, kDBlock));
    k_density_ema<<<blocks, kDBlock, 0, st>>>(reinterpret_cast<float4*>(density_grid), reinterpret_cast<const float4*>(tmp), n4, opts->decay,
                                             opts->density_thresh, 1.0 / (double)cells, partial, ticket,
                                             reinterpret_cast<DensityStats*>(stats));
    const uint32_t n_bytes = (uint32_t)(cells / 8);
    k_density_pack<<<ceil_div(n_bytes, kDBlock), kDBlock, 0, st>>>(reinterpret_cast<const float4*>(density_grid), n_bytes,
                                                                  reinterpret_cast<const DensityStats*>(stats), bitfield);
    g_launches += 2 + (coords ? 1 : 0);
    return check_launch("density_grid_update");
}

int envidr_mark_untrained_grid(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, float bound, uint32_t cascade,
                               uint32_t grid_size, float* density_grid, int32_t* count, envidr_stream_t stream) {
    ENVIDR_REQUIRE(poses && density_grid, ENVIDR_E_BADARG, "null pointer");
    const uint32_t C = cascade, H = grid_size;
    ENVIDR_REQUIRE(C >= 1 && C <= 8 && H >= 2 && H <= 1024 && (H & (H - 1)) == 0, ENVIDR_E_BADARG, "cascade in 1..8, grid_size a power of two <= 1024");
    ENVIDR_REQUIRE(B <= kMaxPosesSmem, ENVIDR_E_UNSUPPORTED, "at most 1024 poses per call (pass count to accumulate over several calls)");
    if (B == 0 && !count) return 0;
    cudaStream_t st = as_stream(stream);
    const uint32_t cells1 = H * H * H;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_mark_untrained, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxPosesSmem * 12 * 4);
        attr = true;
    }
    for (uint32_t cas = 0; cas < C; cas++) {
        const double bnd = fmin((double)(1u << cas), (double)bound);
        const double hgs = bnd / H;
        k_mark_untrained<<<ceil_div(cells1, kDBlock), kDBlock, B * 12 * sizeof(float), st>>>(
            poses, B, (float)((double)cx / (double)fx), (float)((double)cy / (double)fy), (float)(bnd - hgs), (float)(hgs * 2), H, cells1,
            density_grid + (size_t)cas * cells1, count ? count + (size_t)cas * cells1 : nullptr);
        g_launches += 1;
    }
    return check_launch("mark_untrained_grid");
}

int envidr_mark_untrained_apply(const int32_t* count, uint32_t n, float* density_grid, envidr_stream_t stream) {
    ENVIDR_REQUIRE(count && density_grid, ENVIDR_E_BADARG, "null pointer");
    if (n == 0) return 0;
    k_mark_apply<<<ceil_div(n, kDBlock), kDBlock, 0, as_stream(stream)>>>(count, n, density_grid);
    g_launches += 1;
    return check_launch("mark_untrained_apply");
}

}  // extern "C"
