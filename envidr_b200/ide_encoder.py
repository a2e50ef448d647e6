"""Integrated Directional Encoding (mirrors reference ide_encoder/ide_encoder.py:57-130).

Same module interface: buffers mat / ml_array / pow_level / sigma (non-persistent), output_dim,
forward(xyz, roughness=0).  The reference has no native code for IDE (about 20 small torch kernels and an
`isnan().any()` host sync per call, ~60 more in the backward); here forward and backward are one CUDA kernel
each (envidr_ide_encode_forward / _backward) behind an autograd.Function, so gradients reach the normals and
the roughness exactly as in the reference.  The torch formulation is kept for CPU tensors (it is what the
tests compare the kernels with) and for double backward through the encoding, which no shipped loss needs.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from ._lib import check, lib, ptr, stream


def _gen_binom(a, k):
    return float(np.prod(a - np.arange(k))) / math.factorial(k)


def _sph_harm_coeff(l, m, k):
    legendre = ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m)
                * _gen_binom(0.5 * (l + k + m - 1.0), l))
    return math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * math.pi * math.factorial(l + m))) * legendre


def get_ml_array(deg_view):
    ml_list = [(m, 2 ** i) for i in range(deg_view) for m in range(2 ** i + 1)]
    return np.array(ml_list).T


# the reference module's public helper names (ide_encoder.py:5-41; demo.ipynb imports the module wholesale)
def generalized_binomial_coeff(a, k):
    return _gen_binom(a, k)


def assoc_legendre_coeff(l, m, k):
    return ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m) * _gen_binom(0.5 * (l + k + m - 1.0), l))


def sph_harm_coeff(l, m, k):
    return _sph_harm_coeff(l, m, k)


class _ide_encode(torch.autograd.Function):
    """out [B, 2P] = IDE(dirs [B,3], kappa_inv [B] or scalar); once differentiable w.r.t. dirs and the kappa array."""

    @staticmethod
    def forward(ctx, dirs, kappa, deg_view, out_dim):
        d = dirs.detach().float().contiguous()
        B = d.shape[0]
        out = torch.empty(B, out_dim, dtype=torch.float32, device=d.device)
        if torch.is_tensor(kappa):
            k = kappa.detach().float().reshape(-1).contiguous()
            assert k.numel() == B, "roughness must be [..., 1] matching xyz"
            check(lib().envidr_ide_encode_forward(ptr(d), ptr(k), 0.0, B, deg_view, 1.0, ptr(out), stream()), "ide_encode_forward")
            ctx.save_for_backward(d, k)
            ctx.kappa_scalar = 0.0
        else:
            check(lib().envidr_ide_encode_forward(ptr(d), None, float(kappa), B, deg_view, 1.0, ptr(out), stream()), "ide_encode_forward")
            ctx.save_for_backward(d)
            ctx.kappa_scalar = float(kappa)
        ctx.deg_view = deg_view
        ctx.kappa_shape = tuple(kappa.shape) if torch.is_tensor(kappa) else None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        saved = ctx.saved_tensors
        d = saved[0]
        k = saved[1] if len(saved) > 1 else None
        B = d.shape[0]
        grad = grad.float().contiguous()
        gd = torch.empty_like(d)
        gk = torch.empty(B, dtype=torch.float32, device=d.device) if k is not None else None
        check(lib().envidr_ide_encode_backward(ptr(d), ptr(k), ctx.kappa_scalar, B, ctx.deg_view, 1.0, ptr(grad), ptr(gd), ptr(gk), stream()),
              "ide_encode_backward")
        return gd, (gk.view(ctx.kappa_shape) if gk is not None else None), None, None


class IntegratedDirEncoder(nn.Module):
    def __init__(self, input_dim=3, deg_view=4):
        super().__init__()
        self.deg_view = deg_view
        if deg_view > 5:
            raise ValueError("Only deg_view of at most 5 is numerically stable.")
        ml_array = get_ml_array(deg_view)
        l_max = 2 ** (deg_view - 1)
        mat = np.zeros((l_max + 1, ml_array.shape[1]))
        for i, (m, l) in enumerate(ml_array.T):
            for k in range(l - m + 1):
                mat[k, i] = _sph_harm_coeff(l, m, k)
        sigma = 0.5 * ml_array[1, :] * (ml_array[1, :] + 1)
        self.register_buffer("mat", torch.Tensor(mat), False)
        self.register_buffer("ml_array", torch.Tensor(ml_array), False)
        self.register_buffer("pow_level", torch.arange(l_max + 1), False)
        self.register_buffer("sigma", torch.Tensor(sigma), False)
        self.output_dim = (2 ** deg_view - 1 + deg_view) * 2

    def forward(self, xyz, roughness=0, **kwargs):
        if not xyz.is_cuda:
            return self._forward_torch(xyz, roughness)
        needs_grad = torch.is_grad_enabled() and (xyz.requires_grad or (torch.is_tensor(roughness) and roughness.requires_grad))
        if not needs_grad:
            return self._forward_kernel(xyz, roughness)
        prefix = xyz.shape[:-1]
        kap = roughness
        if torch.is_tensor(kap):
            kap = kap.reshape(-1) if kap.numel() > 1 else float(kap)
        out = _ide_encode.apply(xyz.reshape(-1, 3), kap, self.deg_view, self.output_dim)
        return out.reshape(*prefix, self.output_dim)

    def _forward_kernel(self, xyz, roughness):
        prefix = xyz.shape[:-1]
        d = xyz.detach().float().reshape(-1, 3).contiguous()
        B = d.shape[0]
        out = torch.empty(B, self.output_dim, dtype=torch.float32, device=d.device)
        if torch.is_tensor(roughness) and roughness.numel() > 1:
            k = roughness.detach().float().reshape(-1).contiguous()
            assert k.numel() == B, "roughness must be [..., 1] matching xyz"
            check(lib().envidr_ide_encode_forward(ptr(d), ptr(k), 0.0, B, self.deg_view, 1.0, ptr(out), stream()), "ide_encode_forward")
        else:
            check(lib().envidr_ide_encode_forward(ptr(d), None, float(roughness), B, self.deg_view, 1.0, ptr(out), stream()),
                  "ide_encode_forward")
        return out.reshape(*prefix, self.output_dim)

    def _forward_torch(self, xyz, roughness):
        kappa_inv = roughness
        x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
        y = y + torch.logical_and(x == 0, y == 0)
        vmz = z ** self.pow_level
        m_max = int(self.ml_array[0].max().item())
        re, im = [torch.ones_like(x)], [torch.zeros_like(x)]
        for _ in range(m_max):
            re, im = re + [re[-1] * x - im[-1] * y], im + [re[-1] * y + im[-1] * x]
        idx = self.ml_array[0].long()
        re, im = torch.cat(re, -1)[..., idx], torch.cat(im, -1)[..., idx]
        zc = torch.matmul(vmz, self.mat)
        att = torch.exp(-self.sigma * kappa_inv)
        return torch.cat([re * zc * att, im * zc * att], dim=-1)
