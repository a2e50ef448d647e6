"""Ray generation and loss epilogue of a training step on libenvidr_b200 (csrc/epilogue.cu).

    get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1)    mirrors nerf.utils.get_rays (utils.py:110-209):
        same arguments, same result dict (rays_o, rays_d, inds[, inds_coarse]); the pixel choice uses the reference's torch calls
        in its order (randint / multinomial + rand), the ray arithmetic is one kernel instead of ~15.
    train_loss(outputs, gt_rgb, gt_mask, beta, cfg)                          the loss terms Trainer.train_step applies to
        run_cuda's training outputs (utils.py:661-808) + the auxiliary block that feeds them (cuda_ray.py:173-211), as ONE
        autograd.Function: forward = 2 launches, backward = 1, no host synchronisation (the reference's boolean-mask gathers
        each synchronise).  Returns the weighted total (differentiable w.r.t. image, weights_sum, sdfs, sdf_gradients) and the
        individual terms.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Dict, Optional

import numpy as np
import torch

from ._lib import check, lib, ptr, stream


# ---------------------------------------------------------------------------------------------
# get_rays
# ---------------------------------------------------------------------------------------------

@torch.no_grad()
def get_rays(poses: torch.Tensor, intrinsics, H: int, W: int, N: int = -1, error_map=None, patch_size: int = 1) -> Dict[str, torch.Tensor]:
    device = poses.device
    if not poses.is_cuda:
        raise RuntimeError("envidr_b200.epilogue.get_rays: CUDA tensors only (no CPU fallback)")
    poses = poses.float().contiguous()
    B = int(poses.shape[0])
    results = {}
    inds = None
    if N > 0:
        N = min(N, H * W)
        if patch_size > 1:                                                              # utils.py:136-153
            num_patch = N // (patch_size ** 2)
            inds_x = torch.randint(0, H - patch_size, size=[num_patch], device=device)
            inds_y = torch.randint(0, W - patch_size, size=[num_patch], device=device)
            ii = torch.stack([inds_x, inds_y], dim=-1)
            pi, pj = torch.meshgrid(torch.arange(patch_size, device=device), torch.arange(patch_size, device=device), indexing="ij")
            offsets = torch.stack([pi.reshape(-1), pj.reshape(-1)], dim=-1)
            ii = (ii.unsqueeze(1) + offsets.unsqueeze(0)).view(-1, 2)
            inds = ii[:, 0] * W + ii[:, 1]
        elif error_map is None:                                                         # utils.py:155-158
            inds = torch.randint(0, H * W, size=[N], device=device)
        else:                                                                           # utils.py:175-186 (B == 1 in the reference)
            inds_coarse = torch.multinomial(error_map.to(device), N, replacement=False)
            inds_x, inds_y = inds_coarse // 128, inds_coarse % 128
            sx, sy = H / 128, W / 128
            inds_x = (inds_x * sx + torch.rand(B, N, device=device) * sx).long().clamp(max=H - 1)
            inds_y = (inds_y * sy + torch.rand(B, N, device=device) * sy).long().clamp(max=W - 1)
            inds2 = inds_x * W + inds_y
            results["inds_coarse"] = inds_coarse
            if B != 1:
                raise RuntimeError("get_rays: error_map sampling is per pose; the kernel shares inds across poses (B must be 1)")
            inds = inds2[0]
        inds = inds.long().contiguous()
        n = int(inds.shape[0])
        results["inds"] = inds.expand([B, n]) if error_map is None else inds2
    else:
        n = H * W
    rays_o = torch.empty(B, n, 3, dtype=torch.float32, device=device)
    rays_d = torch.empty(B, n, 3, dtype=torch.float32, device=device)
    intr = (ctypes.c_float * 4)(*[float(v) for v in intrinsics])
    check(lib().envidr_get_rays(ptr(poses), B, intr, H, W, ptr(inds), n, ptr(rays_o), ptr(rays_d), stream()), "get_rays")
    results["rays_o"] = rays_o
    results["rays_d"] = rays_d
    return results


# ---------------------------------------------------------------------------------------------
# loss epilogue
# ---------------------------------------------------------------------------------------------

class LossIn(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("image", "weights_sum", "gt_rgb", "gt_mask", "sdfs", "sdf_gradients", "weights", "deltas",
                                              "rays", "beta")]


class LossOpts(ctypes.Structure):
    _fields_ = [("N", ctypes.c_uint32), ("M", ctypes.c_uint32), ("n_rays", ctypes.c_uint32), ("color_l1", ctypes.c_int32),
                ("backsdf_mean", ctypes.c_int32), ("color_w", ctypes.c_float), ("mask_w", ctypes.c_float), ("cauchy_w", ctypes.c_float),
                ("eikonal_w", ctypes.c_float), ("backsdf_w", ctypes.c_float), ("backsdf_thresh", ctypes.c_float)]


@dataclasses.dataclass
class LossConfig:
    """The opt.* fields Trainer.train_step reads for these terms (defaults: configs/scenes/toaster.ini at its first schedule stage)."""
    color_l1: bool = True            # color_l1_loss -> L1Loss, else MSELoss (main_nerf.py:84-86)
    color_w: float = 1.0             # color_loss_weight
    mask_w: float = 1.0              # mask_loss_weight
    cauchy_w: float = 0.1            # cauchy_loss_weight
    eikonal_w: float = 0.001         # eikonal_loss_weight
    backsdf_w: float = 5e-3          # backsdf_loss_weight (0 also switches the auxiliary block off)
    backsdf_thresh: float = 0.01
    backsdf_mean: bool = False       # backsdf_mode != 'sum'


TERM_NAMES = ("total", "color", "mask", "cauchy", "eikonal", "backsdf", "aux_points", "backsdf_denom")


class _TrainLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, weights_sum, sdfs, sdf_gradients, gt_rgb, gt_mask, weights, deltas, rays, beta, cfg: LossConfig):
        f = lambda t: None if t is None else t.detach().float().contiguous()
        image, weights_sum, sdfs, sdf_gradients, gt_rgb, gt_mask, weights, deltas = [f(t) for t in (image, weights_sum, sdfs, sdf_gradients,
                                                                                                 gt_rgb, gt_mask, weights, deltas)]
        if not image.is_cuda:
            raise RuntimeError("envidr_b200.epilogue.train_loss: CUDA tensors only (no CPU fallback)")
        dev = image.device
        image = image.view(-1, 3)
        N, M = int(image.shape[0]), int(sdfs.shape[0])
        rays = None if rays is None else rays.int().contiguous()
        beta = beta.detach().float().reshape(1).contiguous() if torch.is_tensor(beta) else torch.tensor([float(beta)], device=dev)
        li = LossIn(*[0 if t is None else t.data_ptr() for t in (image, weights_sum, gt_rgb, gt_mask, sdfs, sdf_gradients, weights, deltas,
                                                               rays, beta)])
        lo = LossOpts(N, M, 0 if rays is None else int(rays.shape[0]), int(cfg.color_l1), int(cfg.backsdf_mean), cfg.color_w, cfg.mask_w,
                      cfg.cauchy_w, cfg.eikonal_w, cfg.backsdf_w, cfg.backsdf_thresh)
        ws = torch.empty(int(lib().envidr_train_loss_workspace_bytes(M)), dtype=torch.uint8, device=dev)
        terms = torch.empty(8, dtype=torch.float32, device=dev)
        check(lib().envidr_train_loss_forward(ctypes.byref(li), ctypes.byref(lo), ptr(terms), ptr(ws), ws.numel(), stream()), "train_loss_forward")
        ctx.keep = (image, weights_sum, sdfs, sdf_gradients, gt_rgb, gt_mask, weights, deltas, rays, beta, ws, terms)
        ctx.li, ctx.lo = li, lo
        ctx.shapes = (N, M)
        ctx.mark_non_differentiable(terms)
        return terms[0].clone(), terms

    @staticmethod
    def backward(ctx, grad_total, _grad_terms):
        image, weights_sum, sdfs, sdf_gradients, *_rest, ws, terms = ctx.keep
        N, M = ctx.shapes
        dev = image.device
        need = ctx.needs_input_grad
        z = lambda shape, on: torch.empty(shape, dtype=torch.float32, device=dev) if on else None
        d_image = z((N, 3), need[0])
        d_ws = z((N,), need[1] and weights_sum is not None)
        d_sdfs = z((M,), need[2])
        d_grad = z((M, 3), need[3] and sdf_gradients is not None)
        g = grad_total.detach().float().reshape(1).contiguous()
        check(lib().envidr_train_loss_backward(ctypes.byref(ctx.li), ctypes.byref(ctx.lo), ptr(terms), ptr(g), ptr(d_image), ptr(d_ws),
                                               ptr(d_sdfs), ptr(d_grad), ptr(ws), ws.numel(), stream()), "train_loss_backward")
        return d_image, d_ws, d_sdfs, d_grad, None, None, None, None, None, None, None


def train_loss(image, weights_sum, sdfs, sdf_gradients, gt_rgb, gt_mask, weights, deltas, rays, beta, cfg: Optional[LossConfig] = None):
    """-> (total loss [scalar tensor, differentiable], terms dict of 0-dim tensors: color, mask, cauchy, eikonal, backsdf, ...).
    image [N,3] / [1,N,3], weights_sum [N], sdfs [M], sdf_gradients [M,3] or None, gt_rgb [N,3], gt_mask [N] or None,
    weights [M], deltas [M,2], rays [n,3] int32 (needed when cfg.backsdf_w > 0), beta: tensor or float."""
    cfg = cfg or LossConfig()
    shape = image.shape
    total, terms = _TrainLoss.apply(image.reshape(-1, 3), None if weights_sum is None else weights_sum.reshape(-1), sdfs, sdf_gradients,
                                    gt_rgb.reshape(-1, 3), None if gt_mask is None else gt_mask.reshape(-1), weights, deltas, rays, beta, cfg)
    del shape
    return total, {k: terms[i] for i, k in enumerate(TERM_NAMES)}
