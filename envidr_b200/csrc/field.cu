// field.cu -- fused per-sample field evaluation for sm_100a (exact fp32 path).
//
// One persistent kernel evaluates, for a tile of 128 samples at a time and entirely in shared memory /
// registers, what the reference does with ~70 PyTorch + 2 extension launches per render iteration:
//   hash-grid gather with smoothstep + d/dx  (hashencoder/src/hashencoder.cu:103-254)
//   sdf_net forward                          (nerf/network.py:415-450)
//   normal = d sdf / d xyz, analytic reverse pass instead of autograd.grad (nerf/renderer.py:182-198)
//   Laplace density                          (nerf/network.py:26-44)
//   reflect dir, n.w_o, IDE x2               (nerf/renderer.py:20-39,147-180; ide_encoder.py:98-130)
//   env_net x2 + unitNorm, diffuse_net, color_net (+ renv branch), sigmoid, blend (network.py:524-698)
// HBM traffic per sample is the 24 B of input and the 16-60 B of output; activations never leave the SM.
//
// GEMM scheme (dense_layer): the 128 x K activation tile lives in shared memory (row stride 260 floats);
// each of the 8 warps owns 16 rows, each lane owns columns {lane + 32 j}; the A operand is a warp-wide
// broadcast float4 read, the B operand (transposed weights, staged through shared memory in 16-row
// chunks with double-buffered cp.async) is a conflict-free 128-byte read, so the inner loop is FFMA bound.
// Weights are repacked once (envidr_field_pack) into K-major, column-padded images.
//
// A tensor-core (tcgen05) implementation of the env_net passes lives in field_tc.cu; this file is the
// bit-faithful fp32 path and the numerical reference for it on the device.
#include <math.h>
#include "gridenc.cuh"
#include "ide_tables.cuh"
#include "field_tc.cuh"

namespace envidr {

constexpr int kTile = 128;
constexpr int kThreads = 256;
constexpr int kLd = 260;                 // activation row stride in floats (16-byte aligned rows)
constexpr int kKC = 16;                  // weight rows per staged chunk
constexpr int kWbuf = kKC * 256;         // floats per staging buffer
constexpr int kJacCol = 160;             // activation columns holding dy_dx during the SDF phase
constexpr int kMaxHidden = 64;

// side buffer rows ([field][128], one column per sample)
enum Side {
    S_H = 0,          // 16: last sdf layer output (sdf, geo..., rough_raw, blend_raw)
    S_GEO = 16,       // 12 (up to 15)
    S_N = 32,         // 3 unit normal
    S_WO = 35,        // 3 w_o = -dir
    S_WR = 38,        // 3 reflected dir (rotated)
    S_NE = 41,        // 3 normal used for the diffuse env lookup (rotated)
    S_NDOT = 44, S_ROUGH = 45, S_BLEND = 46, S_SIGMA = 47, S_SDF = 48,
    S_GX = 49,        // 3 raw gradient
    S_FN = 52,        // 12 (up to 16) env feature of n
    S_FR = 68,        // 16 env feature of w_r
    S_CD = 84,        // 3
    S_CS = 87,        // 3
    S_RI = 90,        // 4 r_image (rgb * vis, remapped roughness)
    S_MASK = 94,      // renv mask
    S_CE = 95,        // 3 colour of the inter-reflection branch
    S_FE = 98,        // 16 renv feature
    S_COUNT = 114
};

struct LayerDesc {
    uint32_t K, Kp, N, Np;      // Kp: K rounded up to 4, Np: N rounded up to 32*NJ
    uint32_t wt_off, b_off;     // float offsets into the packed blob ([Kp][Np] image, [Np] bias)
    uint32_t has_bias, pad;
};

struct FieldDev {
    const float* table; const int* offsets; const float* blob;
    uint32_t L, H; float S, bound; int enabled_levels;
    uint32_t n_sdf, n_env, n_diffuse, n_color, n_renv;
    LayerDesc sdf[4], sdf_bwd[4], env[ENVIDR_MAX_LAYERS], diffuse[4], color[4], renv[ENVIDR_MAX_LAYERS];
    uint32_t sdf_row0_off;      // W_last[0, :] of the sdf net
    uint32_t geo_dim, ide_P, ide_Kp;
    float beta, density_scale, rough_bias, rough_act_scale, rough_scale, kappa_diffuse, light_scale, intensity_scale;
    float indir_rough_thresh; int learn_blend; int has_rot; float rot[9];
};

struct FieldOutDev { float *sigma, *rgb, *normal, *sdf, *c_diffuse, *c_specular, *roughness, *grad_x; };

__constant__ IdeTables c_ide_field;
static int g_ide_field_deg = 0;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// out[r, c] = act( sum_k in[r, k] * Wt[k, c] + bias[c] ) for the 128-row tile; see the file header.
// `out` is addressed as out[r * ors + c * ocs]; it may alias `in` (results are held in registers until
// every warp has finished reading).  mask (optional, same addressing as out, may alias out): the result is
// zeroed where mask <= 0 (ReLU derivative).
template <int NJ>
__device__ __forceinline__ void dense_layer(const float* __restrict__ blob, const LayerDesc& ld, const float* in,
                                            float* out, int ors, int ocs, bool relu, const float* mask, float* wbuf) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = warp * 16;
    const int Kp = (int)ld.Kp, Np = NJ * 32;
    const float* __restrict__ wt = blob + ld.wt_off;
    float acc[16][NJ];
    #pragma unroll
    for (int i = 0; i < 16; i++) {
        #pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j] = 0.0f;
    }
    const int nchunks = (Kp + kKC - 1) / kKC;
    auto stage = [&](int c, int buf) {
        const int k0 = c * kKC;
        const int kc = min(kKC, Kp - k0);
        const float4* src = reinterpret_cast<const float4*>(wt + (size_t)k0 * Np);
        float4* dst = reinterpret_cast<float4*>(wbuf + buf * kWbuf);
        const int nvec = kc * Np / 4;
        for (int i = threadIdx.x; i < nvec; i += kThreads) cp_async16(dst + i, src + i);
    };
    stage(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) {
            stage(c + 1, (c + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* wb = wbuf + (c & 1) * kWbuf;
        const int k0 = c * kKC;
        const int kc = min(kKC, Kp - k0);
        for (int kk = 0; kk < kc; kk += 4) {
            float b[4][NJ];
            #pragma unroll
            for (int q = 0; q < 4; q++) {
                #pragma unroll
                for (int j = 0; j < NJ; j++) b[q][j] = wb[(kk + q) * Np + lane + 32 * j];
            }
            #pragma unroll
            for (int i = 0; i < 16; i++) {
                const float4 a = *reinterpret_cast<const float4*>(in + (r0 + i) * kLd + k0 + kk);
                #pragma unroll
                for (int j = 0; j < NJ; j++) {
                    acc[i][j] = fmaf(a.x, b[0][j], acc[i][j]);
                    acc[i][j] = fmaf(a.y, b[1][j], acc[i][j]);
                    acc[i][j] = fmaf(a.z, b[2][j], acc[i][j]);
                    acc[i][j] = fmaf(a.w, b[3][j], acc[i][j]);
                }
            }
        }
        __syncthreads();
    }
    const float* __restrict__ bias = blob + ld.b_off;
    #pragma unroll
    for (int j = 0; j < NJ; j++) {
        const int c = lane + 32 * j;
        if (c < (int)ld.N) {
            const float bv = ld.has_bias ? __ldg(bias + c) : 0.0f;
            #pragma unroll
            for (int i = 0; i < 16; i++) {
                float v = acc[i][j] + bv;
                if (relu) v = fmaxf(v, 0.0f);
                const int o = (r0 + i) * ors + c * ocs;
                if (mask && !(mask[o] > 0.0f)) v = 0.0f;
                out[o] = v;
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void dense_dispatch(const float* __restrict__ blob, const LayerDesc& ld, const float* in, float* out,
                                               int ors, int ocs, bool relu, const float* mask, float* wbuf) {
    switch (ld.Np) {
        case 32:  dense_layer<1>(blob, ld, in, out, ors, ocs, relu, mask, wbuf); break;
        case 64:  dense_layer<2>(blob, ld, in, out, ors, ocs, relu, mask, wbuf); break;
        case 160: dense_layer<5>(blob, ld, in, out, ors, ocs, relu, mask, wbuf); break;
        default:  dense_layer<8>(blob, ld, in, out, ors, ocs, relu, mask, wbuf); break;
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }   // torch.nn.Softplus(beta=1, threshold=20)

// unit-normalise `n` values side[(base+i)*128 + s] in place (F.normalize, eps)
__device__ __forceinline__ void unit_norm(float* side, int base, int n, int s, float eps) {
    float ss = 0.0f;
    for (int i = 0; i < n; i++) { const float v = side[(base + i) * kTile + s]; ss += v * v; }
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    for (int i = 0; i < n; i++) side[(base + i) * kTile + s] *= inv;
}

// Run an MLP stack whose first-layer input is already in act[:, 0:K0).  Hidden activations ping-pong
// between column blocks `colA` and `colB` (in-place is also legal); the last layer (no activation) is
// written to side[(side_base + c) * 128 + r].
__device__ void run_stack(const float* __restrict__ blob, const LayerDesc* layers, int n, float* act, int colA, int colB,
                          float* side, int side_base, float* wbuf) {
    const float* in = act;
    for (int i = 0; i < n; i++) {
        if (i == n - 1) {
            dense_dispatch(blob, layers[i], in, side + side_base * kTile, 1, kTile, false, nullptr, wbuf);
        } else {
            float* o = act + ((i & 1) ? colB : colA);
            dense_dispatch(blob, layers[i], in, o, kLd, 1, true, nullptr, wbuf);
            in = o;
        }
    }
}

// phases: bit 0 = geometry (P1-P4), bit 1 = env_net on the FFMA path (P5), bit 2 = shading heads (P6-P7).
// With the tensor-core env_net (field_tc.cu) the kernel is launched twice: phases = 1 writes a per-sample record
// `rec` [M][32] (geo 0-15, n 16-18, n.w_o 19, roughness 20, blend 21, n_env 22-24, w_r 25-27), k_env_tc turns it into
// unit-normalised env features `feat` [M][32] (f_n 0-15, f_r 16-31), and phases = 4 shades from rec + feat.
constexpr int kRecFloats = 32;
__global__ void __launch_bounds__(kThreads, 1)
k_field(const FieldDev F, const float* __restrict__ xyzs, const float* __restrict__ dirs, const float* __restrict__ r_images,
        const uint32_t* __restrict__ M_dev, uint32_t M_host, int mode, int phases, float* __restrict__ rec, float* __restrict__ feat,
        const FieldOutDev O) {
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                         // [128][260]
    float* wbuf = act + kTile * kLd;           // [2][16*256]
    float* side = wbuf + 2 * kWbuf;            // [S_COUNT][128]
    const uint32_t M = M_dev ? *M_dev : M_host;
    const int tid = threadIdx.x;
    const float* __restrict__ blob = F.blob;
    const int G = (int)F.geo_dim;
    const int n_hidden = (int)F.n_sdf - 1;
    const int Hd = (int)F.sdf[0].N;            // hidden width (<= 64)

    for (uint32_t tile = blockIdx.x; (size_t)tile * kTile < M; tile += gridDim.x) {
        const uint32_t m0 = tile * kTile;
        if (phases & 1) {
        // ---- P1: hash-grid gather: thread = (sample, level parity) ---------------------------------------
        {
            const int s = tid & (kTile - 1), g = tid >> 7;
            const uint32_t m = m0 + s;
            const bool valid = m < M;
            float x01[3] = {0.f, 0.f, 0.f};
            if (valid) {
                #pragma unroll
                for (int d = 0; d < 3; d++) x01[d] = (xyzs[3 * (size_t)m + d] + F.bound) / (2 * F.bound);
            }
            const EncMode em{1, 0, 0};
            for (uint32_t l = g; l < F.L; l += 2) {
                float e0 = 0.f, e1 = 0.f, j[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                Cell<3> cell;
                const bool lvl_on = !(F.enabled_levels > 0 && (int)l >= F.enabled_levels);
                if (valid && lvl_on && cell.setup(em, x01, F.offsets, l, F.S, F.H)) {
                    const float* grid = F.table + (size_t)(uint32_t)F.offsets[l] * 2;
                    float rows[8][2];
                    #pragma unroll
                    for (uint32_t corner = 0; corner < 8; corner++) {
                        uint32_t pl[3];
                        #pragma unroll
                        for (int d = 0; d < 3; d++) pl[d] = cell.pg[d] + ((corner >> d) & 1u);
                        load_row<2>(grid + (size_t)cell_index<3>(em, cell.hashmap_size, cell.resolution, pl) * 2, rows[corner]);
                    }
                    #pragma unroll
                    for (uint32_t corner = 0; corner < 8; corner++) {
                        float wt = 1;
                        #pragma unroll
                        for (int d = 0; d < 3; d++) wt *= ((corner >> d) & 1u) ? cell.w[d] : 1 - cell.w[d];
                        e0 += wt * rows[corner][0];
                        e1 += wt * rows[corner][1];
                    }
                    #pragma unroll
                    for (int gd = 0; gd < 3; gd++) {
                        #pragma unroll
                        for (uint32_t sub = 0; sub < 4; sub++) {
                            float wt = cell.scale;
                            uint32_t corner = 0;
                            #pragma unroll
                            for (int nd = 0; nd < 2; nd++) {
                                const int d = nd >= gd ? nd + 1 : nd;
                                if ((sub >> nd) & 1u) { wt *= cell.w[d]; corner |= 1u << d; }
                                else                  { wt *= 1 - cell.w[d]; }
                            }
                            j[gd * 2 + 0] += wt * (rows[corner | (1u << gd)][0] - rows[corner][0]) * cell.dw[gd];
                            j[gd * 2 + 1] += wt * (rows[corner | (1u << gd)][1] - rows[corner][1]) * cell.dw[gd];
                        }
                    }
                }
                float* a = act + s * kLd;
                a[2 * l] = e0; a[2 * l + 1] = e1;
                #pragma unroll
                for (int q = 0; q < 6; q++) a[kJacCol + 6 * l + q] = j[q];
            }
            // zero-pad encoder columns up to the padded K of the first layer
            if (g == 0) for (uint32_t c = 2 * F.L; c < F.sdf[0].Kp; c++) act[s * kLd + c] = 0.f;
        }
        __syncthreads();
        // ---- P2: sdf_net forward; hidden layer i (1-based) lives at columns 32 + 64 (i-1) --------------------
        {
            const float* in = act;
            for (int i = 0; i < (int)F.n_sdf; i++) {
                if (i == (int)F.n_sdf - 1) {
                    dense_dispatch(blob, F.sdf[i], in, side + S_H * kTile, 1, kTile, false, nullptr, wbuf);
                } else {
                    float* o = act + 32 + kMaxHidden * i;
                    dense_dispatch(blob, F.sdf[i], in, o, kLd, 1, true, nullptr, wbuf);
                    in = o;
                }
            }
        }
        // ---- P3: reverse pass for d sdf / d enc -------------------------------------------------------------
        {
            // g_{n-1} = W_last[0,:] * relu'(h_{n-1}), in place over the last hidden activation
            float* hl = act + 32 + kMaxHidden * (n_hidden - 1);
            const float* row0 = blob + F.sdf_row0_off;
            for (int idx = tid; idx < kTile * Hd; idx += kThreads) {
                const int r = idx / Hd, o = idx - r * Hd;
                float* p = hl + r * kLd + o;
                *p = (*p > 0.0f) ? __ldg(row0 + o) : 0.0f;
            }
            __syncthreads();
            const float* in = hl;
            for (int i = n_hidden - 1; i >= 0; i--) {
                if (i == 0) {
                    dense_dispatch(blob, F.sdf_bwd[0], in, act, kLd, 1, false, nullptr, wbuf);          // d sdf / d enc -> enc columns
                } else {
                    float* o = act + 32 + kMaxHidden * (i - 1);
                    dense_dispatch(blob, F.sdf_bwd[i], in, o, kLd, 1, false, o, wbuf);                  // masked by relu'(h_i), in place
                    in = o;
                }
            }
        }
        // ---- P4: per-sample geometry ------------------------------------------------------------------------
        if (tid < kTile) {
            const int s = tid;
            const uint32_t m = m0 + s;
            const bool valid = m < M;
            const float* a = act + s * kLd;
            float gx = 0.f, gy = 0.f, gz = 0.f;
            for (uint32_t l = 0; l < F.L; l++) {
                const float g0 = a[2 * l], g1 = a[2 * l + 1];
                const float* jq = a + kJacCol + 6 * l;
                gx += g0 * jq[0] + g1 * jq[1];
                gy += g0 * jq[2] + g1 * jq[3];
                gz += g0 * jq[4] + g1 * jq[5];
            }
            const float inv2b = 1.0f / (2 * F.bound);
            gx *= inv2b; gy *= inv2b; gz *= inv2b;
            const float gn = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-10f);
            const float nx = gx / gn, ny = gy / gn, nz = gz / gn;
            const float sdf = side[(S_H + 0) * kTile + s];
            const float sg = (sdf > 0.f) ? 1.f : ((sdf < 0.f) ? -1.f : 0.f);
            const float sigma = (1.0f / F.beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(sdf) / F.beta)) * F.density_scale;
            float ss = 0.f;
            for (int i = 0; i < G; i++) { const float v = side[(S_H + 1 + i) * kTile + s]; ss += v * v; }
            const float ginv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            for (int i = 0; i < G; i++) side[(S_GEO + i) * kTile + s] = side[(S_H + 1 + i) * kTile + s] * ginv;
            float rough = F.rough_act_scale * softplusf_(side[(S_H + 1 + G) * kTile + s] + F.rough_bias) * F.rough_scale;
            const float blend = sigmoidf_(side[(S_H + 2 + G) * kTile + s]);
            float dx = 0.f, dy = 0.f, dz = 1.f;
            if (valid) { dx = dirs[3 * (size_t)m]; dy = dirs[3 * (size_t)m + 1]; dz = dirs[3 * (size_t)m + 2]; }
            const float wox = -dx, woy = -dy, woz = -dz;
            const float ndot = nx * wox + ny * woy + nz * woz;
            float wrx = 2 * ndot * nx - wox, wry = 2 * ndot * ny - woy, wrz = 2 * ndot * nz - woz;
            float nex = nx, ney = ny, nez = nz;
            if (F.has_rot) {   // v @ R
                const float* R = F.rot;
                const float a0 = wrx * R[0] + wry * R[3] + wrz * R[6], a1 = wrx * R[1] + wry * R[4] + wrz * R[7],
                            a2 = wrx * R[2] + wry * R[5] + wrz * R[8];
                wrx = a0; wry = a1; wrz = a2;
                const float b0 = nx * R[0] + ny * R[3] + nz * R[6], b1 = nx * R[1] + ny * R[4] + nz * R[7],
                            b2 = nx * R[2] + ny * R[5] + nz * R[8];
                nex = b0; ney = b1; nez = b2;
            }
            side[(S_N + 0) * kTile + s] = nx; side[(S_N + 1) * kTile + s] = ny; side[(S_N + 2) * kTile + s] = nz;
            side[(S_WR + 0) * kTile + s] = wrx; side[(S_WR + 1) * kTile + s] = wry; side[(S_WR + 2) * kTile + s] = wrz;
            side[(S_NE + 0) * kTile + s] = nex; side[(S_NE + 1) * kTile + s] = ney; side[(S_NE + 2) * kTile + s] = nez;
            side[S_NDOT * kTile + s] = ndot; side[S_ROUGH * kTile + s] = rough; side[S_BLEND * kTile + s] = blend;
            if (valid) {
                if (O.sigma) O.sigma[m] = sigma;
                if (O.sdf) O.sdf[m] = sdf;
                if (O.roughness) O.roughness[m] = rough;
                if (O.normal) { O.normal[3 * (size_t)m] = nx; O.normal[3 * (size_t)m + 1] = ny; O.normal[3 * (size_t)m + 2] = nz; }
                if (O.grad_x) { O.grad_x[3 * (size_t)m] = gx; O.grad_x[3 * (size_t)m + 1] = gy; O.grad_x[3 * (size_t)m + 2] = gz; }
                if (rec && mode != 1) {
                    float* q = rec + (size_t)m * kRecFloats;
                    for (int i = 0; i < G; i++) q[i] = side[(S_GEO + i) * kTile + s];
                    q[16] = nx; q[17] = ny; q[18] = nz; q[19] = ndot; q[20] = rough; q[21] = blend;
                    q[22] = nex; q[23] = ney; q[24] = nez; q[25] = wrx; q[26] = wry; q[27] = wrz;
                }
            }
        }
        __syncthreads();
        }  // phases & 1
        if (mode == 1) continue;   // geometry only (block-uniform)
        if (!(phases & 1)) {       // shading-only launch: reload the per-sample record written by the geometry launch
            if (tid < kTile) {
                const int s = tid;
                const uint32_t m = m0 + s;
                const float* q = rec + (size_t)min(m, M - 1) * kRecFloats;
                for (int i = 0; i < G; i++) side[(S_GEO + i) * kTile + s] = q[i];
                side[(S_N + 0) * kTile + s] = q[16]; side[(S_N + 1) * kTile + s] = q[17]; side[(S_N + 2) * kTile + s] = q[18];
                side[S_NDOT * kTile + s] = q[19]; side[S_ROUGH * kTile + s] = q[20]; side[S_BLEND * kTile + s] = q[21];
            }
            __syncthreads();
        }
        if (!(phases & 6)) continue;

        // ---- P5: env_net on IDE(n, kappa_diffuse) and IDE(w_r, roughness) --------------------------------
        if (phases & 2)
        for (int branch = 0; branch < 2; branch++) {
            if (tid < kTile) {
                const int s = tid;
                const int vb = branch ? S_WR : S_NE;
                const float kap = branch ? side[S_ROUGH * kTile + s] : F.kappa_diffuse;
                float* a = act + s * kLd;
                ide_eval(c_ide_field, side[(vb + 0) * kTile + s], side[(vb + 1) * kTile + s], side[(vb + 2) * kTile + s], kap,
                         F.light_scale, a, 1, a + F.ide_P, 1);
                for (uint32_t c = 2 * F.ide_P; c < F.ide_Kp; c++) a[c] = 0.f;
            }
            __syncthreads();
            const float* in = act;
            for (int i = 0; i < (int)F.n_env; i++) {
                if (i == (int)F.n_env - 1) {
                    dense_dispatch(blob, F.env[i], in, side + (branch ? S_FR : S_FN) * kTile, 1, kTile, false, nullptr, wbuf);
                } else {
                    dense_dispatch(blob, F.env[i], in, act, kLd, 1, true, nullptr, wbuf);    // in place
                    in = act;
                }
            }
        }
        // ---- P6: shading heads ---------------------------------------------------------------------------------
        const int E = (int)F.env[F.n_env - 1].N;       // env feature dim
        if (!(phases & 4)) continue;
        if (tid < kTile) {
            const int s = tid;
            if (phases & 2) {
                unit_norm(side, S_FN, E, s, 1e-12f);
                unit_norm(side, S_FR, E, s, 1e-12f);
            } else {                                    // features come unit-normalised from the tensor-core kernel
                const float* q = feat + (size_t)min(m0 + s, M - 1) * kRecFloats;
                for (int i = 0; i < E; i++) { side[(S_FN + i) * kTile + s] = q[i]; side[(S_FR + i) * kTile + s] = q[16 + i]; }
            }
            float* a = act + s * kLd;
            for (int i = 0; i < G; i++) a[i] = side[(S_GEO + i) * kTile + s];
            for (int i = 0; i < E; i++) a[G + i] = side[(S_FN + i) * kTile + s];
            for (uint32_t c = G + E; c < F.diffuse[0].Kp; c++) a[c] = 0.f;
        }
        __syncthreads();
        run_stack(blob, F.diffuse, (int)F.n_diffuse, act, 64, 128, side, S_CD, wbuf);
        auto load_color_input = [&](int feat_base) {
            if (tid < kTile) {
                const int s = tid;
                float* a = act + s * kLd;
                for (int i = 0; i < G; i++) a[i] = side[(S_GEO + i) * kTile + s];
                for (int i = 0; i < 3; i++) a[G + i] = side[(S_N + i) * kTile + s];
                for (int i = 0; i < E; i++) a[G + 3 + i] = side[(feat_base + i) * kTile + s];
                a[G + 3 + E] = side[S_NDOT * kTile + s];
                for (uint32_t c = G + 4 + E; c < F.color[0].Kp; c++) a[c] = 0.f;
            }
            __syncthreads();
        };
        load_color_input(S_FR);
        run_stack(blob, F.color, (int)F.n_color, act, 64, 128, side, S_CS, wbuf);
        const bool do_renv = (r_images != nullptr) && F.n_renv > 0;     // block-uniform
        if (do_renv) {
            if (tid < kTile) {
                const int s = tid;
                const uint32_t m = m0 + s;
                float r0 = 0.f, r1 = 0.f, r2 = 0.f, vis = 0.f;
                if (m < M) {
                    const float4 ri = *reinterpret_cast<const float4*>(r_images + 4 * (size_t)m);
                    r0 = ri.x; r1 = ri.y; r2 = ri.z; vis = ri.w;
                }
                const float rough = side[S_ROUGH * kTile + s];
                const float rr = sqrtf(rough / F.rough_scale / 0.75f);
                float* a = act + s * kLd;
                a[0] = r0 * vis; a[1] = r1 * vis; a[2] = r2 * vis; a[3] = rr;
                for (uint32_t c = 4; c < F.renv[0].Kp; c++) a[c] = 0.f;
                side[S_MASK * kTile + s] = (rough < F.indir_rough_thresh && vis > 0.9f) ? 1.f : 0.f;
                side[(S_RI + 3) * kTile + s] = rr;
            }
            __syncthreads();
            run_stack(blob, F.renv, (int)F.n_renv, act, 64, 128, side, S_FE, wbuf);
            if (tid < kTile) unit_norm(side, S_FE, E, tid, 1e-12f);
            __syncthreads();
            load_color_input(S_FE);
            run_stack(blob, F.color, (int)F.n_color, act, 64, 128, side, S_CE, wbuf);
        }
        // ---- P7: activations, blend, store ----------------------------------------------------------------
        if (tid < kTile) {
            const int s = tid;
            const uint32_t m = m0 + s;
            if (m < M) {
                float cd[3], cs[3];
                #pragma unroll
                for (int i = 0; i < 3; i++) {
                    cd[i] = sigmoidf_(side[(S_CD + i) * kTile + s]);
                    cs[i] = sigmoidf_(side[(S_CS + i) * kTile + s]);
                }
                if (do_renv && side[S_MASK * kTile + s] > 0.5f) {
                    const float rr = side[(S_RI + 3) * kTile + s];
                    const float bw = F.learn_blend ? 0.98f * side[S_BLEND * kTile + s] : 0.95f * sigmoidf_(80.0f * (rr - 0.18f));
                    #pragma unroll
                    for (int i = 0; i < 3; i++) cs[i] = cs[i] * bw + sigmoidf_(side[(S_CE + i) * kTile + s]) * (1 - bw);
                }
                #pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (O.rgb) O.rgb[3 * (size_t)m + i] = (cd[i] + cs[i]) * F.intensity_scale;
                    if (O.c_diffuse) O.c_diffuse[3 * (size_t)m + i] = cd[i];
                    if (O.c_specular) O.c_specular[3 * (size_t)m + i] = cs[i];
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// weight packing: torch [N][K] -> K-major, padded [Kp][Np] (+ bias [Np])
// ------------------------------------------------------------------------------------------------
__global__ void k_pack(const float* __restrict__ W, const float* __restrict__ b, float* __restrict__ wt, float* __restrict__ bias,
                       uint32_t K, uint32_t N, uint32_t Kp, uint32_t Np, int transpose_src) {
    // transpose_src = 1: W is [N][K] (forward layer); 0: W is already [K][N] (reverse pass uses W as stored)
    const uint32_t total = Kp * Np;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t k = i / Np, n = i - k * Np;
        float v = 0.f;
        if (k < K && n < N) v = transpose_src ? W[(size_t)n * K + k] : W[(size_t)k * N + n];
        wt[i] = v;
    }
    if (bias) {
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += gridDim.x * blockDim.x)
            bias[i] = (b && i < N) ? b[i] : 0.f;
    }
}

static uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
static uint32_t np_for(uint32_t N) { return N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 160 ? 160 : 256)); }

struct Layout {
    FieldDev dev;
    uint64_t floats;
};

// Validate the field description and lay the packed blob out.  Returns 0 or an ENVIDR_E_* code.
static int build_layout(const envidr_field* f, Layout* out) {
    FieldDev& d = out->dev;
    d = FieldDev{};
    ENVIDR_REQUIRE(f, ENVIDR_E_BADARG, "null field");
    ENVIDR_REQUIRE(f->embeddings && f->offsets, ENVIDR_E_BADARG, "null hash grid");
    ENVIDR_REQUIRE(f->level_dim == 2 && f->num_levels >= 1 && f->num_levels <= 16, ENVIDR_E_UNSUPPORTED,
                   "fused field: level_dim must be 2 and num_levels <= 16");
    ENVIDR_REQUIRE(f->n_sdf >= 2 && f->n_sdf <= 3, ENVIDR_E_UNSUPPORTED, "fused field: sdf_net must have 2 or 3 layers");
    ENVIDR_REQUIRE(f->n_env >= 2 && f->n_env <= ENVIDR_MAX_LAYERS, ENVIDR_E_UNSUPPORTED, "fused field: env_net must have 2..8 layers");
    ENVIDR_REQUIRE(f->n_diffuse >= 1 && f->n_diffuse <= 3 && f->n_color >= 1 && f->n_color <= 3, ENVIDR_E_UNSUPPORTED,
                   "fused field: diffuse_net / color_net must have 1..3 layers");
    ENVIDR_REQUIRE(f->n_renv <= ENVIDR_MAX_LAYERS, ENVIDR_E_UNSUPPORTED, "fused field: renv_net too deep");
    ENVIDR_REQUIRE(f->ide_degree >= 1 && f->ide_degree <= 5, ENVIDR_E_UNSUPPORTED, "Only deg_view of at most 5 is numerically stable.");
    const uint32_t G = f->geo_feat_dim;
    ENVIDR_REQUIRE(G >= 1 && G <= 13, ENVIDR_E_UNSUPPORTED, "fused field: geo_feat_dim must be 1..13");
    const uint32_t P = (1u << f->ide_degree) - 1 + f->ide_degree;
    const uint32_t E = f->env[f->n_env - 1].out_dim;
    ENVIDR_REQUIRE(E >= 1 && E <= 16, ENVIDR_E_UNSUPPORTED, "fused field: env_feat_dim must be 1..16");
    // dimension chain checks
    ENVIDR_REQUIRE(f->sdf[0].in_dim == f->num_levels * 2, ENVIDR_E_BADARG, "sdf_net input must be num_levels*level_dim");
    ENVIDR_REQUIRE(f->sdf[f->n_sdf - 1].out_dim >= 3 + G && f->sdf[f->n_sdf - 1].out_dim <= 16, ENVIDR_E_UNSUPPORTED,
                   "fused field: sdf_net must output sdf + geo_feat + roughness + blend (ensemble_mlp)");
    for (uint32_t i = 0; i + 1 < f->n_sdf; i++)
        ENVIDR_REQUIRE(f->sdf[i].out_dim <= kMaxHidden && f->sdf[i].out_dim == f->sdf[0].out_dim, ENVIDR_E_UNSUPPORTED,
                       "fused field: sdf_net hidden width must be uniform and <= 64");
    ENVIDR_REQUIRE(f->env[0].in_dim == 2 * P, ENVIDR_E_BADARG, "env_net input must be the IDE width");
    ENVIDR_REQUIRE(f->diffuse[0].in_dim == G + E, ENVIDR_E_BADARG, "diffuse_net input must be geo_feat + env_feat");
    ENVIDR_REQUIRE(f->color[0].in_dim == G + 3 + E + 1, ENVIDR_E_BADARG, "color_net input must be geo + normal + env_feat + n.v");
    ENVIDR_REQUIRE(f->diffuse[f->n_diffuse - 1].out_dim == 3 && f->color[f->n_color - 1].out_dim == 3, ENVIDR_E_BADARG, "rgb heads must output 3");
    if (f->n_renv) {
        ENVIDR_REQUIRE(f->renv[0].in_dim == 4 && f->renv[f->n_renv - 1].out_dim == E, ENVIDR_E_BADARG, "renv_net must map 4 -> env_feat");
    }
    // hidden activations of the shading heads ping-pong between two 64-column blocks of the tile
    for (uint32_t i = 0; i + 1 < f->n_diffuse; i++) ENVIDR_REQUIRE(f->diffuse[i].out_dim <= 64, ENVIDR_E_UNSUPPORTED, "fused field: diffuse_net hidden width must be <= 64");
    for (uint32_t i = 0; i + 1 < f->n_color; i++) ENVIDR_REQUIRE(f->color[i].out_dim <= 64, ENVIDR_E_UNSUPPORTED, "fused field: color_net hidden width must be <= 64");
    for (uint32_t i = 0; i + 1 < f->n_renv; i++) ENVIDR_REQUIRE(f->renv[i].out_dim <= 64, ENVIDR_E_UNSUPPORTED, "fused field: renv_net hidden width must be <= 64");
    uint64_t off = 0;
    auto lay = [&](const envidr_mlp_layer* src, uint32_t n, LayerDesc* dst, bool reverse) -> int {
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t K = reverse ? src[i].out_dim : src[i].in_dim;
            const uint32_t N = reverse ? src[i].in_dim : src[i].out_dim;
            ENVIDR_REQUIRE(src[i].weight, ENVIDR_E_BADARG, "null layer weight");
            ENVIDR_REQUIRE(K <= 256 && N <= 256, ENVIDR_E_UNSUPPORTED, "fused field: layer widths must be <= 256");
            if (i > 0 && !reverse) ENVIDR_REQUIRE(src[i].in_dim == src[i - 1].out_dim, ENVIDR_E_BADARG, "layer dims do not chain");
            LayerDesc& L = dst[i];
            L.K = K; L.Kp = round_up(K, 4); L.N = N; L.Np = np_for(N);
            L.wt_off = (uint32_t)off; off += (uint64_t)L.Kp * L.Np;
            L.b_off = (uint32_t)off;  off += L.Np;
            L.has_bias = (!reverse && src[i].bias) ? 1 : 0;
        }
        return 0;
    };
    int rc;
    if ((rc = lay(f->sdf, f->n_sdf, d.sdf, false))) return rc;
    if ((rc = lay(f->sdf, f->n_sdf - 1, d.sdf_bwd, true))) return rc;
    if ((rc = lay(f->env, f->n_env, d.env, false))) return rc;
    if ((rc = lay(f->diffuse, f->n_diffuse, d.diffuse, false))) return rc;
    if ((rc = lay(f->color, f->n_color, d.color, false))) return rc;
    if ((rc = lay(f->renv, f->n_renv, d.renv, false))) return rc;
    d.sdf_row0_off = (uint32_t)off; off += kMaxHidden;
    out->floats = off;
    d.table = f->embeddings; d.offsets = f->offsets;
    d.L = f->num_levels; d.H = f->base_resolution; d.S = f->log2_per_level_scale; d.bound = f->bound;
    d.enabled_levels = f->enabled_levels;
    d.n_sdf = f->n_sdf; d.n_env = f->n_env; d.n_diffuse = f->n_diffuse; d.n_color = f->n_color; d.n_renv = f->n_renv;
    d.geo_dim = G; d.ide_P = P; d.ide_Kp = round_up(2 * P, 4);
    d.beta = f->beta; d.density_scale = f->density_scale; d.rough_bias = f->roughness_bias;
    d.rough_act_scale = f->roughness_act_scale; d.rough_scale = f->roughness_scale; d.kappa_diffuse = f->diffuse_kappa_inv;
    d.light_scale = f->light_intensity_scale; d.intensity_scale = f->intensity_scale;
    d.indir_rough_thresh = f->indir_roughness_thresh; d.learn_blend = f->learn_indir_blend;
    d.has_rot = f->has_env_rot;
    for (int i = 0; i < 9; i++) d.rot[i] = f->env_rot[i];
    d.blob = reinterpret_cast<const float*>(f->packed);
    return 0;
}

static int ensure_ide_field(uint32_t deg) {
    if ((int)deg == g_ide_field_deg) return 0;
    IdeTables t;
    if (!ide_build_tables((int)deg, &t)) return ENVIDR_E_UNSUPPORTED;
    cudaError_t e = cudaMemcpyToSymbol(c_ide_field, &t, sizeof(t));
    if (e != cudaSuccess) { set_error("ide tables: %s", cudaGetErrorString(e)); return (int)e; }
    g_ide_field_deg = (int)deg;
    return 0;
}

constexpr size_t kFieldSmem = (size_t)(kTile * kLd + 2 * kWbuf + S_COUNT * kTile) * sizeof(float);

// timing hook (bench.py): when non-null, ev[0]/ev[1] are recorded around the dominant kernel of this call
// mode 0: full field; 1: geometry only (sigma, normal); 2: geometry only + geometry records captured through `cap`
int field_forward_launch(const envidr_field* field, const float* xyzs, const float* dirs, const float* r_images,
                         const uint32_t* M_dev, uint32_t M_host, int mode, const envidr_field_out* out, cudaStream_t st,
                         cudaEvent_t* ev, int* ev_recorded, const RecCapture* cap) {
    if (ev_recorded) *ev_recorded = 0;
    if (mode == 2) ENVIDR_REQUIRE(cap && cap->rec && field->precision == 1, ENVIDR_E_BADARG, "record capture needs the tensor-core field");
    Layout lay;
    int rc = build_layout(field, &lay);
    if (rc) return rc;
    const uint64_t simt_bytes = lay.floats * sizeof(float);
    ENVIDR_REQUIRE(field->packed && field->packed_bytes >= simt_bytes, ENVIDR_E_WORKSPACE,
                   "field->packed missing or too small (call envidr_field_pack)");
    if (!M_dev && M_host == 0) return 0;               // empty batch: nothing to do (pointers may be NULL)
    ENVIDR_REQUIRE(xyzs && dirs && out, ENVIDR_E_BADARG, "null pointer");
    if ((rc = ensure_ide_field(field->ide_degree))) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_field, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFieldSmem);
        if (e != cudaSuccess) { set_error("field smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    FieldOutDev O{out->sigma, out->rgb, out->normal, out->sdf, out->c_diffuse, out->c_specular, out->roughness, out->grad_x};
    uint32_t grid = kSMs;
    if (!M_dev) grid = min((uint32_t)kSMs, ceil_div(M_host, kTile));
    if (field->precision != 1) {
        if (ev) cudaEventRecord(ev[0], st);
        k_field<<<grid, kThreads, kFieldSmem, st>>>(lay.dev, xyzs, dirs, r_images, M_dev, M_host, mode, mode == 1 ? 1 : 7, nullptr, nullptr, O);
        if (ev) { cudaEventRecord(ev[1], st); if (ev_recorded) *ev_recorded = 1; }
        return check_launch("field_forward");
    }
    // tensor-core path: geometry (k_geom_tc) -> record, env_net (k_env_tc) -> features, shading heads (k_field phases = 4)
    TcEnv tcenv;
    TcGeom tcgeom;
    TcShade tcshade;
    uint64_t total = 0, total2 = 0, total3 = 0;
    ENVIDR_REQUIRE(tc_layout(field, simt_bytes, &tcenv, &total), ENVIDR_E_UNSUPPORTED,
                   "precision=1: env_net shape outside the tensor-core kernel (hidden widths must be multiples of 32, <= 256; env_feat <= 16)");
    const bool geom_tc = geom_tc_layout(field, total, &tcgeom, &total2);
    if (!geom_tc) total2 = total;
    const bool shade_tc = shade_tc_layout(field, total2, &tcshade, &total3);
    if (!shade_tc) total3 = total2;
    ENVIDR_REQUIRE(field->packed_bytes >= total3, ENVIDR_E_WORKSPACE,
                   "field->packed too small for the tensor-core images (envidr_field_pack_bytes)");
    float* rec = nullptr;
    float* feat = nullptr;
    if (mode == 2) {
        ENVIDR_REQUIRE(geom_tc, ENVIDR_E_UNSUPPORTED, "record capture needs the tensor-core geometry kernel");
        return geom_tc_launch(tcgeom, xyzs, dirs, M_dev, M_host, 0, cap->rec, out, st, cap->base_dev, cap->cap);
    }
    if (mode != 1) {
        const uint64_t cap = M_dev ? field->scratch_samples : M_host;
        ENVIDR_REQUIRE(field->scratch && field->scratch_samples >= cap && cap > 0, ENVIDR_E_WORKSPACE,
                       "precision=1 needs field->scratch (256 B per sample)");
        rec = reinterpret_cast<float*>(field->scratch);
        feat = rec + (size_t)field->scratch_samples * kRecFloats;
    }
    if (geom_tc) {
        if ((rc = geom_tc_launch(tcgeom, xyzs, dirs, M_dev, M_host, mode, rec, out, st))) return rc;
    } else {
        k_field<<<grid, kThreads, kFieldSmem, st>>>(lay.dev, xyzs, dirs, r_images, M_dev, M_host, mode, 1, rec, feat, O);
    }
    if (mode == 1) return check_launch("field_forward(tc, geometry)");
    if (ev) cudaEventRecord(ev[0], st);
    rc = env_tc_launch(tcenv, field->ide_degree, rec, feat, M_dev, M_host, st);
    if (ev) { cudaEventRecord(ev[1], st); if (ev_recorded) *ev_recorded = 1; }
    if (rc) return rc;
    if (shade_tc) return shade_tc_launch(tcshade, rec, feat, r_images, M_dev, M_host, out, st);
    k_field<<<grid, kThreads, kFieldSmem, st>>>(lay.dev, xyzs, dirs, r_images, M_dev, M_host, mode, 4, rec, feat, O);
    return check_launch("field_forward(tc)");
}

}  // namespace envidr

using namespace envidr;

extern "C" {

uint64_t envidr_field_pack_bytes(const envidr_field* field) {
    Layout lay;
    if (build_layout(field, &lay)) return 0;
    TcEnv t;
    TcGeom g;
    TcShade sh;
    uint64_t total = 0, total2 = 0, total3 = 0;
    if (!tc_layout(field, lay.floats * sizeof(float), &t, &total)) return lay.floats * sizeof(float);
    if (!geom_tc_layout(field, total, &g, &total2)) total2 = total;                  // FFMA + env_net (+ sdf_net) (+ heads) images
    if (!shade_tc_layout(field, total2, &sh, &total3)) total3 = total2;
    return total3;
}

int envidr_field_pack(const envidr_field* field, void* packed, uint64_t packed_bytes, envidr_stream_t stream) {
    Layout lay;
    int rc = build_layout(field, &lay);
    if (rc) return rc;
    ENVIDR_REQUIRE(packed && packed_bytes >= lay.floats * sizeof(float), ENVIDR_E_WORKSPACE, "packed buffer too small");
    float* blob = reinterpret_cast<float*>(packed);
    cudaStream_t st = as_stream(stream);
    auto pack = [&](const envidr_mlp_layer* src, uint32_t n, const LayerDesc* dsc, bool reverse) {
        for (uint32_t i = 0; i < n; i++) {
            const LayerDesc& L = dsc[i];
            // forward: W [N][K] -> [Kp][Np] (transpose); reverse: W [out][in] is already [K=out][N=in]
            const uint32_t srcK = reverse ? src[i].out_dim : src[i].in_dim, srcN = reverse ? src[i].in_dim : src[i].out_dim;
            k_pack<<<64, 256, 0, st>>>(src[i].weight, reverse ? nullptr : src[i].bias, blob + L.wt_off, blob + L.b_off, srcK, srcN, L.Kp,
                                       L.Np, reverse ? 0 : 1);
        }
    };
    pack(field->sdf, field->n_sdf, lay.dev.sdf, false);
    pack(field->sdf, field->n_sdf - 1, lay.dev.sdf_bwd, true);
    pack(field->env, field->n_env, lay.dev.env, false);
    pack(field->diffuse, field->n_diffuse, lay.dev.diffuse, false);
    pack(field->color, field->n_color, lay.dev.color, false);
    pack(field->renv, field->n_renv, lay.dev.renv, false);
    // row 0 of the last sdf layer ( d sdf / d h_last )
    const envidr_mlp_layer& last = field->sdf[field->n_sdf - 1];
    cudaMemsetAsync(blob + lay.dev.sdf_row0_off, 0, kMaxHidden * sizeof(float), st);
    cudaMemcpyAsync(blob + lay.dev.sdf_row0_off, last.weight, last.in_dim * sizeof(float), cudaMemcpyDeviceToDevice, st);
    rc = check_launch("field_pack");
    if (rc) return rc;
    envidr_field tmp = *field;
    tmp.packed = packed; tmp.packed_bytes = packed_bytes;
    TcEnv t;
    TcGeom g;
    TcShade sh;
    uint64_t total = 0, total2 = 0, total3 = 0;
    if (tc_layout(&tmp, lay.floats * sizeof(float), &t, &total) && packed_bytes >= total) {
        if ((rc = tc_pack(field, t, packed, st))) return rc;
        if (geom_tc_layout(&tmp, total, &g, &total2) && packed_bytes >= total2) { if ((rc = geom_tc_pack(field, g, packed, st))) return rc; }
        else total2 = total;
        if (shade_tc_layout(&tmp, total2, &sh, &total3) && packed_bytes >= total3) return shade_tc_pack(field, sh, packed, st);
    }
    return 0;
}

int envidr_field_forward(const envidr_field* field, const float* xyzs, const float* dirs, const float* r_images, uint32_t M, int mode,
                         const envidr_field_out* out, envidr_stream_t stream) {
    cudaEvent_t* ev = (mode != 1 && field && field->precision == 1) ? timing_acquire() : nullptr;
    int recorded = 0;
    ENVIDR_REQUIRE(mode == 0 || mode == 1, ENVIDR_E_BADARG, "mode must be 0 or 1");
    const int rc = field_forward_launch(field, xyzs, dirs, r_images, nullptr, M, mode, out, as_stream(stream), ev, &recorded, nullptr);
    if (ev && recorded) timing_commit();
    if (rc == 0 && M > 0) g_launches += (field->precision == 1 && mode != 1) ? 3 : 1;
    return rc;
}

int envidr_field_forward_records(const envidr_field* field, const float* rec, const float* r_images, uint32_t M,
                                 const envidr_field_out* out, envidr_stream_t stream) {
    return envidr_field_forward_records_indexed(field, rec, nullptr, r_images, M, out, stream);
}

int envidr_field_forward_records_indexed(const envidr_field* field, const float* rec, const int32_t* rec_index, const float* r_images, uint32_t M,
                                         const envidr_field_out* out, envidr_stream_t stream) {
    ENVIDR_REQUIRE(field && out, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(field->precision == 1, ENVIDR_E_UNSUPPORTED, "field_forward_records: tensor-core field only (precision = 1)");
    if (M == 0) return 0;
    ENVIDR_REQUIRE(rec && out->rgb, ENVIDR_E_BADARG, "null pointer");
    Layout lay;
    int rc = build_layout(field, &lay);
    if (rc) return rc;
    TcEnv tcenv; TcGeom tcgeom; TcShade tcshade;
    uint64_t total = 0, total2 = 0, total3 = 0;
    ENVIDR_REQUIRE(tc_layout(field, lay.floats * sizeof(float), &tcenv, &total), ENVIDR_E_UNSUPPORTED, "env_net outside the tensor-core kernel");
    if (!geom_tc_layout(field, total, &tcgeom, &total2)) total2 = total;
    ENVIDR_REQUIRE(shade_tc_layout(field, total2, &tcshade, &total3), ENVIDR_E_UNSUPPORTED, "shading heads outside the tensor-core kernel");
    ENVIDR_REQUIRE(field->packed && field->packed_bytes >= total3, ENVIDR_E_WORKSPACE, "field->packed too small");
    ENVIDR_REQUIRE(field->scratch && field->scratch_samples * 2 >= M, ENVIDR_E_WORKSPACE, "field->scratch: 128 B per sample needed");
    cudaStream_t st = as_stream(stream);
    float* feat = reinterpret_cast<float*>(field->scratch);
    cudaEvent_t* ev = timing_acquire();
    if (ev) cudaEventRecord(ev[0], st);
    if (field->rec_unrotated && field->has_env_rot) {
        tcenv.has_rot = 1;
        for (int i = 0; i < 9; i++) tcenv.rot[i] = field->env_rot[i];
    }
    rc = env_tc_launch(tcenv, field->ide_degree, rec, feat, nullptr, M, st, nullptr, rec_index);
    if (ev) { cudaEventRecord(ev[1], st); timing_commit(); }
    if (rc) return rc;
    g_launches += 2;
    return shade_tc_launch(tcshade, rec, feat, r_images, nullptr, M, out, st, rec_index);
}

}  // extern "C"
