"""Integrated Directional Encoding (mirrors reference ide_encoder/ide_encoder.py:57-130).

Same module interface: buffers mat / ml_array / pow_level / sigma (non-persistent), output_dim,
forward(xyz, roughness=0).  The reference has no native code for IDE (about 20 small torch kernels and an
`isnan().any()` host sync per call); here the no-grad CUDA forward is one kernel
(envidr_ide_encode_forward).  When autograd needs the encoding (training), the torch formulation below is
used so gradients reach the normals and the roughness exactly as in the reference.
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
        needs_grad = torch.is_grad_enabled() and (xyz.requires_grad or (torch.is_tensor(roughness) and roughness.requires_grad))
        if xyz.is_cuda and not needs_grad:
            return self._forward_kernel(xyz, roughness)
        return self._forward_torch(xyz, roughness)

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
