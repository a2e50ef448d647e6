"""Dense layer of the training branch on tensor cores (csrc/linear_tc.cu).

`linear_tc(x, W, b, relu)` = relu?(x @ W.T + b) with fp32 tensors and fp32-level accuracy (fp16 hi/lo split operands, three
tcgen05.mma per K step, fp32 accumulation); the reference runs these layers as cuBLAS fp32 GEMMs behind nn.Linear
(nerf/network.py:527-698).  Backward: the data gradient is the same kernel on the image of W^T, the weight gradient stays a
tensor-core kernel k_wgrad_tc (dY^T X) and the bias gradient comes out of the operand-scale pass.  `linear_tc` is once differentiable: used for
the env / colour / diffuse / renv MLPs; sdf_net, whose normals need a double backward (nerf/renderer.py:182-198), uses the any-order family
`mm_nt / mm_nn / mm_tn` at the end of this file (same kernels).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from ._lib import check, lib, ptr, stream


def _image(W: torch.Tensor) -> torch.Tensor:
    """Packed fp16 hi/lo operand image of W [N, K] (device uint8 tensor, 16-byte aligned)."""
    N, K = W.shape
    nbytes = lib().envidr_linear_tc_image_bytes(N, K)
    if nbytes == 0:
        raise ValueError(f"linear_tc supports 1 <= N, K <= 256 (got N={N}, K={K})")
    img = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    check(lib().envidr_linear_tc_pack(ptr(W), N, K, ptr(img), stream()), "linear_tc_pack")
    return img


def _run(x: torch.Tensor, img: torch.Tensor, bias, N: int, relu: bool) -> torch.Tensor:
    M, K = x.shape
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    check(lib().envidr_linear_tc(ptr(x), M, K, ptr(img), ptr(bias), N, int(relu), ptr(y), stream()), "linear_tc")
    return y


def _pow2_scales(gy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Device tensor {s_gy, s_x, 1 / (s_gy s_x)}: powers of two that bring each operand's largest magnitude into [2^13, 2^14)."""
    m = torch.stack([gy.abs().amax(), x.abs().amax()]).clamp_min(1e-30)
    s = torch.exp2(13.0 - torch.floor(torch.log2(m)))
    return torch.cat([s, (1.0 / (s[0] * s[1])).reshape(1)]).contiguous()


WGRAD_TC = True      # weight gradients through csrc/linear_tc.cu::k_wgrad_tc (False: cuBLAS fp32 dY^T X)


WGRAD_ATOMIC = True  # CTAs add their partial dW into one [N, K] tensor with vector atomics (False: [grid, N, K] partials + torch sum)


def wgrad_tc(gy: torch.Tensor, x: torch.Tensor, variant: int = 0, with_bias: bool = False):
    """dW [N, K] = gy^T x on tensor cores (gy [M, N], x [M, K], CUDA fp32 contiguous).
    with_bias: also return the bias gradient gy.sum(0) -> (dW, db); it comes out of the same pass over gy that finds the operand
    scales when N is a power of two in 4..256 (otherwise a torch column sum)."""
    M, N = gy.shape
    K = x.shape[1]
    if not WGRAD_ATOMIC:
        P = lib().envidr_wgrad_tc_partials(M)
        partial = torch.empty(P, N, K, dtype=torch.float32, device=gy.device)
        sc = _pow2_scales(gy, x)
        check(lib().envidr_wgrad_tc(ptr(gy), ptr(x), M, N, K, ptr(sc), ptr(partial), variant, stream()), "wgrad_tc")
        dW = partial.sum(0)
        return (dW, gy.sum(0)) if with_bias else dW
    fused_bias = with_bias and 4 <= N <= 256 and (N & (N - 1)) == 0 and gy.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0
    buf = torch.zeros(8 + N * K + (N if fused_bias else 0), dtype=torch.float32, device=gy.device)     # scales8 + dW (+ db): one zero-fill
    sc, dW = buf[:8], buf[8:8 + N * K].view(N, K)
    db = buf[8 + N * K:] if fused_bias else None
    check(lib().envidr_pow2_scales(ptr(gy), gy.numel(), ptr(x), x.numel(), ptr(sc), N, ptr(db), stream()), "pow2_scales")
    check(lib().envidr_wgrad_tc(ptr(gy), ptr(x), M, N, K, ptr(sc), ptr(dW), variant | 2, stream()), "wgrad_tc")
    if with_bias:
        return dW, (db if fused_bias else gy.sum(0))
    return dW


class _linear_tc(Function):
    @staticmethod
    def forward(ctx, x, W, b, relu):
        x2 = x.detach().float().contiguous()
        Wc = W.detach().float().contiguous()
        bc = None if b is None else b.detach().float().contiguous()
        y = _run(x2, _image(Wc), bc, Wc.shape[0], relu)
        ctx.save_for_backward(x2, Wc, y if relu else None)
        ctx.relu, ctx.has_bias = relu, b is not None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x2, Wc, y = ctx.saved_tensors
        gy = gy.float().contiguous()
        if ctx.relu:
            gy = torch.ops.aten.threshold_backward(gy, y, 0.0)          # gy * (y > 0) in one kernel instead of a compare and a multiply
        gx = gW = gb = None
        if ctx.needs_input_grad[0]:
            gx = _run(gy, _image(Wc.t().contiguous()), None, Wc.shape[1], False)       # dY W = dY (W^T)^T
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            if WGRAD_TC:
                r = wgrad_tc(gy, x2, with_bias=want_b)
                gW, gb = r if want_b else (r, None)
            else:
                gW = gy.t() @ x2
        if want_b and gb is None:
            gb = gy.sum(0)
        return gx, gW, gb, None


def linear_tc(x: torch.Tensor, W: torch.Tensor, b=None, relu: bool = False) -> torch.Tensor:
    """x [M, K] (CUDA fp32), W [N, K], b [N] or None -> [M, N]."""
    return _linear_tc.apply(x, W, b, relu)


# ---------------------------------------------------------------------------------------------------------------------
# Differentiable-to-any-order family for sdf_net: the normals are d sdf / d x taken with create_graph=True (renderer.py:182-198), so
# the loss differentiates THROUGH the first backward of these layers.  Three products, each one's backward written with the other
# two (all on the tensor-core kernels above), close under differentiation:
#     nt(a [M,K], b [N,K]) = a b^T  [M,N]      d a = nn(g, b)      d b = tn(g, a)
#     nn(a [M,N], b [N,K]) = a b    [M,K]      d a = nt(g, b)      d b = tn(a, g)
#     tn(a [M,N], b [M,K]) = a^T b  [N,K]      d a = nt(b, g)      d b = nn(a, g)
# M is the sample dimension (large); the other operand of nt / nn and the result of tn are weight-sized (<= 256 x 256).
# ---------------------------------------------------------------------------------------------------------------------

class _mm_nt(Function):
    @staticmethod
    def forward(ctx, a, b):
        a2, b2 = a.detach().float().contiguous(), b.detach().float().contiguous()
        ctx.save_for_backward(a, b)
        return _run(a2, _image(b2), None, b2.shape[0], False)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return (mm_nn(g, b) if ctx.needs_input_grad[0] else None), (mm_tn(g, a) if ctx.needs_input_grad[1] else None)


class _mm_nn(Function):
    @staticmethod
    def forward(ctx, a, b):
        a2, b2 = a.detach().float().contiguous(), b.detach().float().contiguous()
        ctx.save_for_backward(a, b)
        return _run(a2, _image(b2.t().contiguous()), None, b2.shape[1], False)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return (mm_nt(g, b) if ctx.needs_input_grad[0] else None), (mm_tn(a, g) if ctx.needs_input_grad[1] else None)


class _mm_tn(Function):
    @staticmethod
    def forward(ctx, a, b):
        a2, b2 = a.detach().float().contiguous(), b.detach().float().contiguous()
        ctx.save_for_backward(a, b)
        return wgrad_tc(a2, b2)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return (mm_nt(b, g) if ctx.needs_input_grad[0] else None), (mm_nn(a, g) if ctx.needs_input_grad[1] else None)


def mm_nt(a, b):
    return _mm_nt.apply(a, b)


def mm_nn(a, b):
    return _mm_nn.apply(a, b)


def mm_tn(a, b):
    return _mm_tn.apply(a, b)


def linear_tc_nd(x: torch.Tensor, W: torch.Tensor, b=None) -> torch.Tensor:
    """x W^T + b on the tensor-core kernels, differentiable to any order (no fused ReLU: torch's relu supplies its own
    double backward)."""
    y = mm_nt(x, W)
    return y if b is None else y + b
