// epilogue.cu -- the steps either side of the training branch:
//
//   k_get_rays      nerf/utils.py:110-209 get_rays: pixel index -> ray origin / normalised direction (one thread per ray)
//   k_loss_mark     cuda_ray.py:173-182: flag the last sample of every valid ray
//   k_loss_fwd      Trainer.train_step's loss terms (utils.py:661-662 colour, :712-717 mask BCE, :735-747 back-sdf,
//                   :762-776 Cauchy, :793-798 eikonal) over run_cuda's outputs, including the auxiliary block that feeds them
//                   (cuda_ray.py:173-211: roll-by-one differences, point mask); block sums in fp64, combined in block order
//                   by the last block (deterministic) -> terms[8] on the device
//   k_loss_bwd      the gradients of the weighted total w.r.t. image, weights_sum, sdfs, sdf_gradients in one pass
//
// The reference runs ~60 torch kernels for these (masks, rolls, boolean-index gathers with their device->host syncs, reductions
// and the autograd mirror of each).  Here: 2 launches forward, 1 backward, no synchronisation; every sample / ray is read once
// per pass (HBM-bound streaming, 40 B per sample + 32 B per ray forward).
#include "common.cuh"

namespace envidr {
namespace {

constexpr int kEBlock = 256;

// ---- get_rays ---------------------------------------------------------------------------------------------------------
// i = w + 0.5, j = h + 0.5; xs = (i - cx) / fx; ys = (j - cy) / fy; d = (xs, ys, 1) / |.|; rays_d = d @ R^T; rays_o = t.
// `x / python_scalar` on the GPU is x * fp32(1 / scalar) in torch (see density.cu); rfx / rfy are those reciprocals.
__global__ void __launch_bounds__(kEBlock) k_get_rays(const float* __restrict__ poses, uint32_t B, uint32_t N, uint32_t W, float cx, float cy,
                                                      float rfx, float rfy, const int64_t* __restrict__ inds, float* __restrict__ rays_o,
                                                      float* __restrict__ rays_d) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * N) return;
    const uint32_t b = t / N, n = t - b * N;
    const uint32_t pix = inds ? (uint32_t)inds[n] : n;
    const float i = (float)(pix % W) + 0.5f, j = (float)(pix / W) + 0.5f;
    const float xs = __fmul_rn(__fsub_rn(i, cx), rfx), ys = __fmul_rn(__fsub_rn(j, cy), rfy);
    const float nrm = __fsqrt_rn(fmaf(xs, xs, fmaf(ys, ys, 1.0f)));
    const float dx = __fdiv_rn(xs, nrm), dy = __fdiv_rn(ys, nrm), dz = __fdiv_rn(1.0f, nrm);
    const float* P = poses + 16 * (size_t)b;
    float* o = rays_o + 3 * (size_t)t;
    float* d = rays_d + 3 * (size_t)t;
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        d[k] = fmaf(dz, P[4 * k + 2], fmaf(dy, P[4 * k + 1], __fmul_rn(dx, P[4 * k])));
        o[k] = P[4 * k + 3];
    }
}

// ---- loss epilogue -------------------------------------------------------------------------------------------------------
struct LossArgs {
    const float *image, *weights_sum, *gt_rgb, *gt_mask, *sdfs, *sdf_grad, *weights, *deltas, *beta;
    const int32_t* rays;
    uint8_t* last;              // [M] 1 = last sample of a valid ray
    uint32_t N, M, n_rays;
    int color_l1, use_aux, backsdf_mean;
    float color_w, mask_w, cauchy_w, eikonal_w, backsdf_w, backsdf_thresh;
};

enum { T_TOTAL = 0, T_COLOR, T_MASK, T_CAUCHY, T_EIKONAL, T_BACKSDF, T_AUXCOUNT, T_DENOM, T_N };
enum { S_COLOR = 0, S_MASK, S_CAUCHY, S_EIKONAL, S_BACK, S_COUNT, S_MW, S_N };

__global__ void __launch_bounds__(kEBlock) k_loss_mark(const int32_t* __restrict__ rays, uint32_t n_rays, uint32_t M, uint8_t* __restrict__ last) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rays) return;
    const int32_t off = rays[3 * n + 1], cnt = rays[3 * n + 2];
    // ray_valid = (rays[:,2] > 0) * (rays[:,1] + rays[:,2] < M);  point_mask[start + count - 1] = False
    if (cnt > 0 && (int64_t)off + cnt < (int64_t)M) {
        int64_t e = (int64_t)off + cnt - 1;
        if (e < 0) e += M;                                   // torch negative indexing (offset -1 of a dropped ray cannot occur here)
        last[e] = 1;
    }
}

// point mask of sample i (cuda_ray.py:173-190) and the back-sdf ingredients; returns false when the sample is masked out
__device__ __forceinline__ bool aux_point(const LossArgs& A, uint32_t i, float& relsdf, float& dist) {
    const uint32_t ip = (i + 1 == A.M) ? 0u : i + 1;        // torch.roll(x, -1, 0)
    const float d0 = A.deltas[2 * (size_t)ip], d1 = A.deltas[2 * (size_t)ip + 1];
    const bool ok = !A.last[i] && d0 > 0.0f && d1 > 0.0f && d1 < __fmul_rn(1.2f, d0);
    relsdf = __fsub_rn(A.sdfs[ip], A.sdfs[i]);
    dist = d1;
    return ok;
}

// d/ds of w * s^2 / (max(dist, 5e-4)^2 + s^2) for a sample inside the back-sdf mask, else 0
__device__ __forceinline__ float backsdf_coeff(const LossArgs& A, uint32_t i) {
    float s, dist;
    if (!aux_point(A, i, s, dist)) return 0.0f;
    const float w = A.weights[i];
    if (!(w > A.backsdf_thresh) || !(s > 0.0f)) return 0.0f;
    const float dc = fmaxf(dist, 5e-4f);
    const float D = dc * dc, q = D + s * s;
    return w * (2.0f * s * D) / (q * q);
}

__device__ __forceinline__ void block_reduce_store(double* v, int n, double* partial /* [gridDim.x][S_N] */) {
    __shared__ double red[kEBlock / 32][S_N];
    for (int k = 0; k < n; k++) {
        double s = v[k];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < n) {
        double b = 0.0;
        for (int w = 0; w < kEBlock / 32; w++) b += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * S_N + threadIdx.x] = b;
    }
}

__global__ void __launch_bounds__(kEBlock) k_loss_fwd(const LossArgs A, double* __restrict__ partial, uint32_t* __restrict__ ticket,
                                                      float* __restrict__ terms) {
    double acc[S_N] = {0, 0, 0, 0, 0, 0, 0};
    const uint32_t stride = gridDim.x * blockDim.x;
    const float beta = *A.beta;
    // per ray: colour, mask
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < A.N; n += stride) {
        float c = 0.0f;
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            const float d = A.image[3 * (size_t)n + k] - A.gt_rgb[3 * (size_t)n + k];
            c += A.color_l1 ? fabsf(d) : d * d;
        }
        acc[S_COLOR] += (double)c;
        if (A.mask_w > 0.0f && A.gt_mask) {
            // F.binary_cross_entropy(weights_sum.clip(1e-3, 1 - 1e-3), alpha_mask); torch clamps the logs at -100
            const float x = fminf(fmaxf(A.weights_sum[n], 1e-3f), 1.0f - 1e-3f), y = A.gt_mask[n];
            acc[S_MASK] += (double)(-(y * fmaxf(logf(x), -100.0f) + (1.0f - y) * fmaxf(log1pf(-x), -100.0f)));
        }
    }
    // per sample: Cauchy (masked by the point mask when the auxiliary block is on), eikonal, back-sdf
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.M; i += stride) {
        bool in_aux = true;
        float s = 0.0f, dist = 0.0f;
        if (A.use_aux) in_aux = aux_point(A, i, s, dist);
        if (in_aux) {
            acc[S_COUNT] += 1.0;
            if (A.cauchy_w > 0.0f) {
                // reg = 0.5 + 0.5 * sign(sdf) * expm1(-|sdf| / beta) (density_func with alpha = 1); log1p((1 - reg)^2 * 16)
                const float sd = A.sdfs[i];
                const float sg = (sd > 0.0f) - (sd < 0.0f);
                const float reg = 0.5f + 0.5f * sg * expm1f(-fabsf(sd) / beta);
                const float u = 1.0f - reg;
                acc[S_CAUCHY] += (double)log1pf(u * u * 16.0f);
            }
            if (A.use_aux && A.backsdf_w > 0.0f) {
                const float w = A.weights[i];
                if (w > A.backsdf_thresh && s > 0.0f) {
                    const float dc = fmaxf(dist, 5e-4f);
                    const float ss = s * s;
                    acc[S_BACK] += (double)(w * (ss / (dc * dc + ss)));
                    acc[S_MW] += (double)w;
                }
            }
        }
        if (A.eikonal_w > 0.0f && A.sdf_grad) {
            const float gx = A.sdf_grad[3 * (size_t)i], gy = A.sdf_grad[3 * (size_t)i + 1], gz = A.sdf_grad[3 * (size_t)i + 2];
            const float e = sqrtf(gx * gx + gy * gy + gz * gz) - 1.0f;
            acc[S_EIKONAL] += (double)(e * e);
        }
    }
    block_reduce_store(acc, S_N, partial);
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    __shared__ double fin[S_N][kEBlock];
    if (is_last) {                                           // block-parallel, fixed-order combination of the per-block sums
        __threadfence();
        double t[S_N] = {0, 0, 0, 0, 0, 0, 0};
        for (uint32_t b = threadIdx.x; b < gridDim.x; b += kEBlock)
            for (int k = 0; k < S_N; k++) t[k] += __ldcg(partial + (size_t)b * S_N + k);
        for (int k = 0; k < S_N; k++) fin[k][threadIdx.x] = t[k];
        __syncthreads();
        for (int o = kEBlock / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o)
                for (int k = 0; k < S_N; k++) fin[k][threadIdx.x] += fin[k][threadIdx.x + o];
            __syncthreads();
        }
    }
    if (is_last && threadIdx.x == 0) {
        double tot[S_N];
        for (int k = 0; k < S_N; k++) tot[k] = fin[k][0];
        const double color = tot[S_COLOR] / (3.0 * A.N);
        const double mask = (A.mask_w > 0.0f && A.gt_mask) ? tot[S_MASK] / A.N : 0.0;
        const double cnt = tot[S_COUNT];
        const double cauchy = (A.cauchy_w > 0.0f && cnt > 0) ? 0.25 * tot[S_CAUCHY] / cnt : 0.0;
        const double eik = (A.eikonal_w > 0.0f && A.sdf_grad && A.M) ? tot[S_EIKONAL] / A.M : 0.0;
        const double denom = A.backsdf_mean ? 1.0 + tot[S_MW] : 1.0;
        const double back = (A.use_aux && A.backsdf_w > 0.0f) ? tot[S_BACK] / denom : 0.0;
        terms[T_COLOR] = (float)color; terms[T_MASK] = (float)mask; terms[T_CAUCHY] = (float)cauchy; terms[T_EIKONAL] = (float)eik;
        terms[T_BACKSDF] = (float)back; terms[T_AUXCOUNT] = (float)cnt; terms[T_DENOM] = (float)denom;
        terms[T_TOTAL] = (float)(A.color_w * color + A.mask_w * mask + A.cauchy_w * cauchy + A.eikonal_w * eik + A.backsdf_w * back);
        *ticket = 0;
    }
}

__global__ void __launch_bounds__(kEBlock) k_loss_bwd(const LossArgs A, const float* __restrict__ terms, const float* __restrict__ grad_loss,
                                                      float* __restrict__ d_image, float* __restrict__ d_ws, float* __restrict__ d_sdfs,
                                                      float* __restrict__ d_sdf_grad) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const float go = grad_loss ? *grad_loss : 1.0f;
    const float beta = *A.beta;
    const float cnt = terms[T_AUXCOUNT], denom = terms[T_DENOM];
    const float kc = go * A.color_w / (3.0f * (float)A.N);
    const float km = go * A.mask_w / (float)A.N;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < A.N; n += stride) {
        if (d_image) {
            #pragma unroll
            for (int k = 0; k < 3; k++) {
                const float d = A.image[3 * (size_t)n + k] - A.gt_rgb[3 * (size_t)n + k];
                d_image[3 * (size_t)n + k] = A.color_l1 ? kc * (float)((d > 0.0f) - (d < 0.0f)) : kc * 2.0f * d;
            }
        }
        if (d_ws) {
            float g = 0.0f;
            if (A.mask_w > 0.0f && A.gt_mask) {
                const float w = A.weights_sum[n], y = A.gt_mask[n];
                if (w >= 1e-3f && w <= 1.0f - 1e-3f)          // clamp passes the gradient on [min, max]
                    g = km * (w - y) / fmaxf((1.0f - w) * w, 1e-12f);
            }
            d_ws[n] = g;
        }
    }
    const float kca = (A.cauchy_w > 0.0f && cnt > 0.0f) ? go * A.cauchy_w * 0.25f / cnt : 0.0f;
    const float kb = (A.use_aux && A.backsdf_w > 0.0f) ? go * A.backsdf_w / denom : 0.0f;
    const float ke = (A.eikonal_w > 0.0f && A.M) ? go * A.eikonal_w / (float)A.M : 0.0f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.M; i += stride) {
        if (d_sdfs) {
            float g = 0.0f;
            bool in_aux = true;
            float s, dist;
            if (A.use_aux) in_aux = aux_point(A, i, s, dist);
            if (in_aux && kca != 0.0f) {
                const float sd = A.sdfs[i];
                const float sg = (sd > 0.0f) - (sd < 0.0f);
                const float ex = expf(-fabsf(sd) / beta);
                const float reg = 0.5f + 0.5f * sg * (ex - 1.0f);
                const float u = 1.0f - reg;
                // d log1p(16 u^2) / d reg = -32 u / (1 + 16 u^2);  d reg / d sdf = -0.5 * sign^2 / beta * exp(-|sdf| / beta)
                g += kca * (-32.0f * u / (1.0f + 16.0f * u * u)) * (-0.5f * sg * sg / beta * ex);
            }
            if (kb != 0.0f) {
                // relsdf_i = sdf[i+1] - sdf[i]: sample i receives -c_i and +c_{i-1}
                const uint32_t im = i == 0 ? A.M - 1 : i - 1;
                g += kb * (backsdf_coeff(A, im) - backsdf_coeff(A, i));
            }
            d_sdfs[i] = g;
        }
        if (d_sdf_grad) {
            float gx = 0.0f, gy = 0.0f, gz = 0.0f;
            if (ke != 0.0f && A.sdf_grad) {
                const float x = A.sdf_grad[3 * (size_t)i], y = A.sdf_grad[3 * (size_t)i + 1], z = A.sdf_grad[3 * (size_t)i + 2];
                const float nr = sqrtf(x * x + y * y + z * z);
                if (nr > 0.0f) {                               // torch: the subgradient of norm at 0 is 0
                    const float c = ke * 2.0f * (nr - 1.0f) / nr;
                    gx = c * x; gy = c * y; gz = c * z;
                }
            }
            d_sdf_grad[3 * (size_t)i] = gx; d_sdf_grad[3 * (size_t)i + 1] = gy; d_sdf_grad[3 * (size_t)i + 2] = gz;
        }
    }
}

uint64_t al256(uint64_t v) { return (v + 255) / 256 * 256; }
constexpr uint32_t kLossBlocks = kSMs * 4;

int fill_args(const envidr_loss_in* in, const envidr_loss_opts* o, void* workspace, uint64_t bytes, LossArgs* A) {
    ENVIDR_REQUIRE(in && o && workspace, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(o->N == 0 || (in->image && in->gt_rgb), ENVIDR_E_BADARG, "image / gt_rgb missing");
    ENVIDR_REQUIRE(in->beta, ENVIDR_E_BADARG, "beta (device scalar) missing");
    const bool aux = o->backsdf_w > 0.0f;
    ENVIDR_REQUIRE(o->M == 0 || in->sdfs, ENVIDR_E_BADARG, "sdfs missing");
    if (aux) ENVIDR_REQUIRE(in->weights && in->deltas && in->rays, ENVIDR_E_BADARG, "back-sdf term needs weights, deltas and rays");
    if (o->mask_w > 0.0f && in->gt_mask) ENVIDR_REQUIRE(in->weights_sum, ENVIDR_E_BADARG, "weights_sum missing");
    ENVIDR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && bytes >= envidr_train_loss_workspace_bytes(o->M), ENVIDR_E_WORKSPACE,
                   "workspace: 256-byte aligned, envidr_train_loss_workspace_bytes(M)");
    A->image = in->image; A->weights_sum = in->weights_sum; A->gt_rgb = in->gt_rgb; A->gt_mask = in->gt_mask; A->sdfs = in->sdfs;
    A->sdf_grad = in->sdf_gradients; A->weights = in->weights; A->deltas = in->deltas; A->beta = in->beta; A->rays = in->rays;
    A->last = reinterpret_cast<uint8_t*>(workspace) + al256((uint64_t)kLossBlocks * S_N * 8) + 256;
    A->N = o->N; A->M = o->M; A->n_rays = o->n_rays;
    A->color_l1 = o->color_l1; A->use_aux = aux; A->backsdf_mean = o->backsdf_mean;
    A->color_w = o->color_w; A->mask_w = o->mask_w; A->cauchy_w = o->cauchy_w; A->eikonal_w = o->eikonal_w; A->backsdf_w = o->backsdf_w;
    A->backsdf_thresh = o->backsdf_thresh;
    return 0;
}

}  // namespace
}  // namespace envidr

using namespace envidr;

extern "C" {

int envidr_get_rays(const float* poses, uint32_t B, const float* intrinsics4, uint32_t H, uint32_t W, const int64_t* inds, uint32_t N,
                    float* rays_o, float* rays_d, envidr_stream_t stream) {
    ENVIDR_REQUIRE(poses && intrinsics4 && rays_o && rays_d, ENVIDR_E_BADARG, "null pointer");
    ENVIDR_REQUIRE(inds || N == H * W, ENVIDR_E_BADARG, "inds == NULL means all H * W pixels");
    if (B == 0 || N == 0) return 0;
    ENVIDR_REQUIRE((uint64_t)B * N < 0xffffffffull, ENVIDR_E_BADARG, "too many rays");
    // host scalars as torch derives them: (i - cx) / fx  ->  (i - fp32(cx)) * fp32(1 / fp32(fx))
    const float fx = intrinsics4[0], fy = intrinsics4[1], cx = intrinsics4[2], cy = intrinsics4[3];
    k_get_rays<<<ceil_div(B * N, kEBlock), kEBlock, 0, as_stream(stream)>>>(poses, B, N, W, cx, cy, 1.0f / fx, 1.0f / fy, inds, rays_o, rays_d);
    g_launches += 1;
    return check_launch("get_rays");
}

uint64_t envidr_train_loss_workspace_bytes(uint32_t M) {
    return al256((uint64_t)kLossBlocks * S_N * 8) + 256 + al256(M);
}

int envidr_train_loss_forward(const envidr_loss_in* in, const envidr_loss_opts* opts, float* terms, void* workspace, uint64_t workspace_bytes,
                              envidr_stream_t stream) {
    LossArgs A;
    int rc = fill_args(in, opts, workspace, workspace_bytes, &A);
    if (rc) return rc;
    ENVIDR_REQUIRE(terms, ENVIDR_E_BADARG, "terms missing");
    cudaStream_t st = as_stream(stream);
    double* partial = reinterpret_cast<double*>(workspace);
    uint32_t* ticket = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + al256((uint64_t)kLossBlocks * S_N * 8));
    cudaMemsetAsync(ticket, 0, 4, st);
    if (A.use_aux) {
        cudaMemsetAsync(A.last, 0, A.M, st);
        if (A.n_rays) k_loss_mark<<<ceil_div(A.n_rays, kEBlock), kEBlock, 0, st>>>(A.rays, A.n_rays, A.M, A.last);
        g_launches += 1;
    }
    const uint32_t work = A.M > A.N ? A.M : A.N;
    const uint32_t blocks = work == 0 ? 1 : min(kLossBlocks, ceil_div(work, kEBlock));
    k_loss_fwd<<<blocks, kEBlock, 0, st>>>(A, partial, ticket, terms);
    g_launches += 1;
    return check_launch("train_loss_forward");
}

int envidr_train_loss_backward(const envidr_loss_in* in, const envidr_loss_opts* opts, const float* terms, const float* grad_loss,
                               float* d_image, float* d_weights_sum, float* d_sdfs, float* d_sdf_gradients, void* workspace,
                               uint64_t workspace_bytes, envidr_stream_t stream) {
    LossArgs A;
    int rc = fill_args(in, opts, workspace, workspace_bytes, &A);
    if (rc) return rc;
    ENVIDR_REQUIRE(terms, ENVIDR_E_BADARG, "terms missing (run envidr_train_loss_forward on the same workspace first)");
    const uint32_t work = A.M > A.N ? A.M : A.N;
    if (work == 0) return 0;
    k_loss_bwd<<<min(kLossBlocks, ceil_div(work, kEBlock)), kEBlock, 0, as_stream(stream)>>>(A, terms, grad_loss, d_image, d_weights_sum, d_sdfs,
                                                                                            d_sdf_gradients);
    g_launches += 1;
    return check_launch("train_loss_backward");
}

}  // extern "C"
