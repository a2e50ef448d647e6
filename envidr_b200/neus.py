"""NeuSDensity (reference nerf/network.py:46-102) on libenvidr_b200 (csrc/neus.cu): the opacity of the `use_neus_sdf` configs
(BASELINE config 4), one kernel forward and one backward instead of ~25 elementwise torch kernels each way.

Same constructor, attributes and call signature as the reference module:
    alpha = NeuSDensity(init_val)(sdf, dirs, dists, gradients, cos_anneal_ratio=1.0)
sdf [M]; dirs [M,3]; dists: float or [M]; gradients [M,3] or None; alpha [M], to be composited with input_alpha=True
(raymarching.composite_rays / composite_rays_train).  Differentiable w.r.t. sdf, gradients and the `variance` parameter.
"""
from __future__ import annotations

import torch
from torch import nn

from ._lib import check, lib, ptr, stream

SQRT3 = 3 ** 0.5


class _neus_alpha(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, gradients, variance, dirs, dists, cos_anneal_ratio):
        if not sdf.is_cuda:
            raise RuntimeError("envidr_b200.neus: CUDA tensors only (no CPU fallback)")
        f = lambda t: None if t is None else t.detach().float().contiguous()
        sdf_c, grad_c, dirs_c = f(sdf).reshape(-1), f(gradients), f(dirs).reshape(-1, 3)
        M = sdf_c.shape[0]
        if grad_c is not None:
            grad_c = grad_c.reshape(-1, 3)
            assert grad_c.shape[0] == M
        var_c = variance.detach().float().reshape(1).contiguous()
        if torch.is_tensor(dists):
            d_t, d_s = f(dists).expand(M).contiguous() if dists.numel() != M else f(dists).reshape(-1), 0.0
        else:
            d_t, d_s = None, float(dists)
        alpha = torch.empty(M, dtype=torch.float32, device=sdf.device)
        check(lib().envidr_neus_alpha_forward(ptr(sdf_c), ptr(dirs_c), ptr(grad_c), ptr(d_t), d_s, ptr(var_c), float(cos_anneal_ratio), M,
                                              ptr(alpha), stream()), "neus_alpha_forward")
        ctx.keep = (sdf_c, grad_c, dirs_c, d_t, d_s, var_c, float(cos_anneal_ratio), sdf.shape, None if gradients is None else gradients.shape,
                    variance.shape)
        return alpha.view(sdf.shape)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        sdf_c, grad_c, dirs_c, d_t, d_s, var_c, r, sdf_shape, grad_shape, var_shape = ctx.keep
        M = sdf_c.shape[0]
        dev = sdf_c.device
        g = g.float().contiguous().reshape(-1)
        need = ctx.needs_input_grad
        g_sdf = torch.empty(M, dtype=torch.float32, device=dev) if need[0] else None
        g_grad = torch.empty(M, 3, dtype=torch.float32, device=dev) if (need[1] and grad_c is not None) else None
        g_var = torch.zeros(1, dtype=torch.float32, device=dev) if need[2] else None
        ws = torch.empty(int(lib().envidr_neus_workspace_bytes()), dtype=torch.uint8, device=dev)
        check(lib().envidr_neus_alpha_backward(ptr(g), ptr(sdf_c), ptr(dirs_c), ptr(grad_c), ptr(d_t), d_s, ptr(var_c), r, M, ptr(g_sdf),
                                               ptr(g_grad), ptr(g_var), ptr(ws), ws.numel(), stream()), "neus_alpha_backward")
        return (None if g_sdf is None else g_sdf.view(sdf_shape), None if g_grad is None else g_grad.view(grad_shape),
                None if g_var is None else g_var.view(var_shape), None, None, None)


class NeuSDensity(nn.Module):
    def __init__(self, init_val, base_steps=1024, neus_n_detach=False):
        super().__init__()
        self.variance = nn.Parameter(torch.tensor(init_val))
        self.scale = 1.0
        self.base_steps = base_steps
        self.current_steps = base_steps
        self.base_dist = 2 * SQRT3 / base_steps
        self.neus_n_detach = neus_n_detach

    def update_scale(self, steps):           # a no-op in the reference as well (`if False and ...`, network.py:56-64)
        pass

    def get_variance(self):
        return self.variance

    def forward(self, sdf, dirs, dists, gradients, cos_anneal_ratio=1.0):
        if gradients is not None and self.neus_n_detach:
            gradients = gradients.detach()
        return _neus_alpha.apply(sdf, gradients, self.variance, dirs, dists, cos_anneal_ratio)
