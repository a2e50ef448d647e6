"""Training branch of the volumetric-render hot path on this library's CUDA operators.

Mirrors, for the shipped scene family (hashgrid_diff + SDF + env / diffuse / colour / renv MLPs):
  * run_cuda, `self.training` branch                nerf/render_func/cuda_ray.py:64-168
  * NeRFNetwork.forward_geometry / forward_sigma    nerf/network.py:381-522 (LaplaceDensity :26-44)
  * NeRFRenderer.compute_normal                     nerf/renderer.py:182-198 (autograd.grad, create_graph=True)
  * get_color_mlp_extra_params / forward_color      nerf/renderer.py:147-180, nerf/network.py:524-698
The operators underneath are the library's own kernels, all differentiable the way the reference's are:
`raymarching.march_rays_train` (count -> scan -> write, deterministic slots), `hashencoder.hash_encode` (forward with
dy_dx, backward, second-order backward -- the normals feed the colour MLPs and the eikonal term),
`raymarching.composite_rays_train` (warp-scan forward / backward), `raymarching.get_scatter_idx`, and the
IDE encoder (forward / backward kernels).  The dense layers of the env / colour / diffuse / renv MLPs run on tensor cores
(`linear_tc`: forward and data gradient, csrc/linear_tc.cu, same fp16 hi/lo split as the inference kernels); their weight
gradients and sdf_net (double backward for the normals) stay cuBLAS fp32.

`TrainOps` exists so that tests can drive the same glue on the CPU with the oracle's operators; the default is the CUDA
library and there is no fallback: without it every call raises.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .field import FieldParams
from .render import RenderConfig

STACKS = ("sdf", "env", "diffuse", "color", "renv")


@dataclasses.dataclass
class TrainOps:
    near_far_from_aabb: Callable
    march_rays_train: Callable
    composite_rays_train: Callable
    get_scatter_idx: Callable
    hash_encode: Callable
    ide_encode: Callable          # (dirs [M,3], kappa_inv tensor-or-float, deg_view) -> [M, 2P]


_default_ops: Optional[TrainOps] = None


def default_ops() -> TrainOps:
    """The library's CUDA operators (loads libenvidr_b200.so; raises if it is missing)."""
    global _default_ops
    if _default_ops is None:
        from . import hashencoder, ide_encoder, raymarching
        encoders: Dict[int, nn.Module] = {}

        def ide(dirs, kappa_inv, deg):
            if deg not in encoders:
                encoders[deg] = ide_encoder.IntegratedDirEncoder(deg_view=deg).to(dirs.device)
            return encoders[deg].to(dirs.device)(dirs, kappa_inv)

        _default_ops = TrainOps(raymarching.near_far_from_aabb, raymarching.march_rays_train, raymarching.composite_rays_train,
                                raymarching.get_scatter_idx, hashencoder.hash_encode, ide)
    return _default_ops


class TrainableField(nn.Module):
    """Parameters of the per-sample field as leaves: `embeddings`, `beta`, `<stack>_w.<i>` / `<stack>_b.<i>`.
    `frozen` lists the stacks that do not train (toaster.ini: frozen_mlps = [specular, diffuse], network.py:785-796)."""

    def __init__(self, fp: FieldParams, frozen: Sequence[str] = ("diffuse", "color"), *, detach_normal: bool = False,
                 normal_anneal_ratio: float = 1.0):
        super().__init__()
        self.detach_normal = detach_normal                     # opt.detach_normal       (renderer.py:191)
        self.normal_anneal_ratio = normal_anneal_ratio         # opt.normal_anneal_ratio (renderer.py:193-196)
        self.cfg = {f.name: getattr(fp, f.name) for f in dataclasses.fields(fp)
                    if f.name not in ("embeddings", "offsets", "_packed", "_scratch") + STACKS}
        self.register_buffer("offsets", fp.offsets.clone().int())
        self.embeddings = nn.Parameter(fp.embeddings.detach().clone().float())
        self.beta = nn.Parameter(torch.tensor(float(fp.beta), device=fp.embeddings.device))
        self.n_layers: Dict[str, int] = {}
        for name in STACKS:
            st = getattr(fp, name)
            self.n_layers[name] = 0 if st is None else len(st)
            for i, (W, b) in enumerate(st or []):
                self.register_parameter(f"{name}_w{i}", nn.Parameter(W.detach().clone().float(), requires_grad=name not in frozen))
                # mlp_bias=False stacks carry no bias (FieldParams / the checkpoint loader allow None)
                self.register_parameter(f"{name}_b{i}", None if b is None else
                                        nn.Parameter(b.detach().clone().float(), requires_grad=name not in frozen))

    # ------------------------------------------------------------------------------------------------
    def stack(self, name: str) -> List:
        return [(getattr(self, f"{name}_w{i}"), getattr(self, f"{name}_b{i}")) for i in range(self.n_layers[name])]

    tc_linear: bool = True      # env / colour / diffuse / renv layers through csrc/linear_tc.cu on CUDA tensors
    fused_heads: bool = True    # diffuse / colour / renv heads: one forward + one backward kernel each (env_train.mlp_fused)
    fused_env: bool = True      # env_net (both evaluations, IDE included) as one forward + one backward kernel (env_train.py)
    tc_sdf: bool = True         # sdf_net through the any-order-differentiable nt / nn / tn family of linear_tc.py (False: torch layers)

    def mlp(self, name: str, x: torch.Tensor) -> torch.Tensor:
        layers = self.stack(name)
        if self.fused_heads and self.tc_linear and name != "sdf" and x.is_cuda and x.dtype == torch.float32 and x.shape[0] > 0:
            # the whole head as one forward kernel and one backward kernel (envidr_b200/env_train.py::mlp_fused)
            from . import env_train
            if env_train.mlp_supported([W for W, _ in layers]):
                return env_train.mlp_fused(x, layers)
        if self.tc_linear and name != "sdf" and x.is_cuda and x.dtype == torch.float32:
            from .linear_tc import linear_tc
            for i, (W, b) in enumerate(layers):
                x = linear_tc(x, W, b, relu=(i != len(layers) - 1))
            return x
        if self.tc_linear and self.tc_sdf and name == "sdf" and x.is_cuda and x.dtype == torch.float32:
            from .linear_tc import linear_tc_nd                       # differentiable to any order (the normals' double backward)
            for i, (W, b) in enumerate(layers):
                x = linear_tc_nd(x, W, b)
                if i != len(layers) - 1:
                    x = F.relu(x)
            return x
        for i, (W, b) in enumerate(layers):
            x = F.linear(x, W, b)
            if i != len(layers) - 1:
                x = F.relu(x)
        return x

    def get_beta(self) -> torch.Tensor:
        """LaplaceDensity.get_beta (network.py:39-44): clamp with a straight-through gradient."""
        b = self.beta
        return b + (torch.clamp(b.detach(), self.cfg["beta_min"], self.cfg["beta_max"]) - b.detach())

    @staticmethod
    def laplace_density(sdf: torch.Tensor, beta, alpha=None) -> torch.Tensor:
        alpha = 1 / beta if alpha is None else alpha
        return alpha * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))            # network.py:32-37

    def export(self) -> FieldParams:
        """Current weights as FieldParams (for the fused inference path; call .pack() on the result)."""
        kw = dict(self.cfg)
        kw["beta"] = float(self.beta.detach())
        for name in STACKS:
            kw[name] = [(W.detach().clone(), None if b is None else b.detach().clone()) for W, b in self.stack(name)] \
                if self.n_layers[name] else None
        return FieldParams(embeddings=self.embeddings.detach().clone(), offsets=self.offsets.clone(), **kw)

    # ------------------------------------------------------------------------------------------------
    def forward_sigma(self, ops: TrainOps, xyzs: torch.Tensor, eikonal: bool = True):
        """forward_geometry + compute_normal + density.  xyzs [M,3] must require grad."""
        c = self.cfg
        bound = float(c["bound"])
        x01 = (xyzs + bound) / (2 * bound)                                                    # hashgrid.py:161
        enc = ops.hash_encode(x01, self.embeddings, self.offsets, c["per_level_scale"], c["base_resolution"], True)
        if c["enabled_levels"] > 0:                                                           # network.py:390-393
            L = self.offsets.shape[0] - 1
            mask = torch.zeros(L, 2, dtype=enc.dtype, device=enc.device)
            mask[: c["enabled_levels"]] = 1
            enc = enc * mask.reshape(-1)
        h = self.mlp("sdf", enc)
        G = int(c["geo_feat_dim"])
        sdf = h[:, 0]
        geo = F.normalize(h[:, 1:1 + G], dim=-1)                                              # geo_feat_act = unitNorm
        roughness = c["roughness_act_scale"] * F.softplus(h[:, 1 + G:2 + G] + c["roughness_bias"]) * c["roughness_scale"]
        blend = torch.sigmoid(h[:, 2 + G:3 + G])
        grad_x = torch.autograd.grad(sdf, xyzs, torch.ones_like(sdf), retain_graph=True, create_graph=True)[0]
        # compute_normal (renderer.py:182-198): opt.detach_normal, opt.normal_anneal_ratio (blend with the radial direction)
        n_src = grad_x.detach() if getattr(self, "detach_normal", False) else grad_x
        normals = F.normalize(n_src, dim=-1, eps=1e-10)
        ratio = float(getattr(self, "normal_anneal_ratio", 1.0))
        if ratio < 1:
            normals = normals * ratio + (1 - ratio) * F.normalize(xyzs.detach(), dim=-1, eps=1e-10)
            normals = F.normalize(normals, dim=-1, eps=1e-10)
        sigma = self.laplace_density(sdf, self.get_beta()) * c["density_scale"]
        return sdf, sigma, geo, normals, (grad_x if eikonal else None), roughness, blend

    def forward_color(self, ops: TrainOps, geo, dirs, normals, roughness, blend, r_images=None):
        c = self.cfg
        w_o = -dirs
        n_dot = (normals * w_o).sum(-1, keepdim=True)
        w_r = 2 * n_dot * normals - w_o                                                       # renderer.py:20-39
        deg = int(c["ide_degree"])
        lis = c["light_intensity_scale"]
        M = normals.shape[0]
        f_n = None
        if self.fused_env and self.tc_linear and ops is default_ops() and normals.is_cuda and normals.dtype == torch.float32 and M > 0:
            # IDE x 2 -> env_net x 2 -> unit norm as ONE forward kernel, and one fused backward kernel (envidr_b200/env_train.py)
            from . import env_train
            layers = self.stack("env")
            if env_train.supported([W for W, _ in layers], deg):
                f_n, f_r = env_train.env_features(normals, w_r, roughness, layers, deg, c["diffuse_kappa_inv"], lis)
        if f_n is None:
            w_r_enc = ops.ide_encode(w_r, roughness, deg) * lis
            n_enc = ops.ide_encode(normals, c["diffuse_kappa_inv"], deg) * lis
            # env_net sees both direction sets in ONE batch (rows [0, M) = normal, [M, 2M) = reflected): half the layer launches of the
            # reference's two calls (network.py:527-541, 589-607), twice the rows per weight-gradient GEMM; same per-row arithmetic
            f_both = F.normalize(self.mlp("env", torch.cat([n_enc, w_r_enc], 0)), dim=-1)
            f_n, f_r = f_both[:M], f_both[M:]
        c_d = torch.sigmoid(self.mlp("diffuse", torch.cat([geo, f_n], -1)))
        hh = torch.cat([geo, normals], -1)
        x_s = torch.cat([hh, f_r, n_dot], -1)
        if r_images is not None and self.n_layers["renv"]:                                    # network.py:612-659, 682-690
            mask = roughness.squeeze(-1) < c["indir_roughness_thresh"]
            ri = r_images
            if ri.shape[-1] == 4:
                vis = ri[:, 3]
                ri = ri[:, :3] * vis[:, None]
                mask = mask & (vis > 0.9)
            rr = torch.sqrt(roughness / c["roughness_scale"] / 0.75)
            bw = 0.98 * blend if c["learn_indir_blend"] else 0.95 * torch.sigmoid(80 * (rr - 0.18))
            f_e = F.normalize(self.mlp("renv", torch.cat([ri, rr], -1)), dim=-1)
            # the two colour evaluations (reflected-light feature, inter-reflection feature) as one batch of 2M rows
            c_both = torch.sigmoid(self.mlp("color", torch.cat([x_s, torch.cat([hh, f_e, n_dot], -1)], 0)))
            c_s, c_e = c_both[:M], c_both[M:]
            c_s = torch.where(mask[:, None], c_s * bw + c_e * (1 - bw), c_s)
        else:
            c_s = torch.sigmoid(self.mlp("color", x_s))
        return (c_d + c_s) * c["intensity_scale"]


_aabb_cache: Dict[tuple, torch.Tensor] = {}


def _aabb_tensor(cfg: RenderConfig, device) -> torch.Tensor:
    """The box as a device tensor, built once per (box, device): a host->device copy per step would be a sync, and is not
    allowed inside CUDA-graph capture."""
    key = (tuple(float(v) for v in cfg.aabb6()), str(device))
    t = _aabb_cache.get(key)
    if t is None:
        t = torch.tensor(key[0], dtype=torch.float32, device=device)
        _aabb_cache[key] = t
    return t


def render_train(field: TrainableField, bitfield: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, cfg: RenderConfig, *,
                 bg_color=1.0, perturb: bool = False, force_all_rays: bool = True, mean_count: int = -1,
                 step_counter: Optional[torch.Tensor] = None, early_stop_steps: int = -1, r_images: Optional[torch.Tensor] = None,
                 geometry_only: bool = False, ops: Optional[TrainOps] = None, ret_weights: bool = False) -> Dict[str, torch.Tensor]:
    """run_cuda, training branch (cuda_ray.py:64-168): march (no grad) -> field -> composite; everything returned is
    differentiable w.r.t. the field parameters exactly as in the reference (image, weights_sum, sdfs, sdf_gradients, ...)."""
    ops = ops or default_ops()
    rays_o = rays_o.contiguous().view(-1, 3)
    rays_d = rays_d.contiguous().view(-1, 3)
    aabb = _aabb_tensor(cfg, rays_o.device)
    nears, fars = ops.near_far_from_aabb(rays_o, rays_d, aabb, cfg.min_near)
    with torch.no_grad():
        xyzs, dirs, deltas, rays = ops.march_rays_train(rays_o, rays_d, cfg.bound, bitfield, cfg.cascade, cfg.grid_size, nears, fars,
                                                        step_counter, mean_count, perturb, 128, force_all_rays, cfg.dt_gamma,
                                                        cfg.max_steps, early_stop_steps)
    xyzs = xyzs.detach().requires_grad_(True)
    if r_images is not None:                                                                  # cuda_ray.py:96-99
        idx = ops.get_scatter_idx(rays, rays.new_zeros(xyzs.shape[0])).long()
        r_images = r_images.view(-1, r_images.shape[-1])[idx]
    sdf, sigma, geo, normals, sdf_grad, roughness, blend = field.forward_sigma(ops, xyzs)
    out: Dict[str, torch.Tensor] = {}
    if geometry_only:                                                                         # cuda_ray.py:121-127
        ws, depth, nimg, _ = ops.composite_rays_train(sigma, normals, deltas, rays, cfg.T_thresh, False, cfg.input_alpha)
        out["normal_image"] = F.normalize(nimg, dim=-1)
        image = None
    else:
        rgbs = field.forward_color(ops, geo, dirs, normals, roughness, blend, r_images)
        ws, depth, image, weights = ops.composite_rays_train(sigma, rgbs, deltas, rays, cfg.T_thresh, ret_weights, cfg.input_alpha)
        if ret_weights:
            out["weights"] = weights.detach()             # no gradient path: the composite backward ignores grad_weights (raymarching.py:291)
        image = image + (1 - ws).unsqueeze(-1) * bg_color
    depth = (depth + nears) * (depth != 0)
    out.update(image=image, depth=depth, weights_sum=ws, sigmas=sigma, sdfs=sdf, roughness=roughness, sdf_gradients=sdf_grad,
               xyzs=xyzs, dirs=dirs, deltas=deltas, rays=rays)
    return out


def loss_epilogue(field: TrainableField, out: Dict[str, torch.Tensor], gt_rgb: torch.Tensor, gt_mask: torch.Tensor, *, color_w=1.0,
                  mask_w=1.0, cauchy_w=0.1, eikonal_w=0.01, backsdf_w=0.0, backsdf_thresh=0.01, fused: Optional[bool] = None) -> torch.Tensor:
    """The Trainer.train_step terms (nerf/utils.py:661-808) that reach every output of the branch under toaster.ini:
    colour L1 (:661-662), mask BCE (:712-717), Cauchy (:762-776), eikonal (:793-798), and with backsdf_w > 0 the back-sdf term
    (:735-747) over run_cuda's auxiliary block (cuda_ray.py:173-211; render_train(..., ret_weights=True) supplies the weights).
    On CUDA tensors this is the fused epilogue kernel (envidr_b200.epilogue.train_loss: 2 launches forward, 1 backward); the
    torch formulation below serves the CPU checks that run this module's glue on the oracle's operators (tests/test_train_cpu.py)."""
    if fused is None:
        fused = out["image"].is_cuda
    if fused:
        from . import epilogue
        cfg = epilogue.LossConfig(color_l1=True, color_w=color_w, mask_w=mask_w, cauchy_w=cauchy_w, eikonal_w=eikonal_w, backsdf_w=backsdf_w,
                                  backsdf_thresh=backsdf_thresh)
        total, _ = epilogue.train_loss(out["image"], out["weights_sum"], out["sdfs"], out["sdf_gradients"], gt_rgb, gt_mask, out.get("weights"),
                                       out["deltas"], out["rays"], field.get_beta().detach(), cfg)
        return total
    if backsdf_w > 0:
        raise NotImplementedError("the back-sdf term exists in the fused epilogue only")
    loss = color_w * (out["image"] - gt_rgb).abs().mean(-1).mean()
    loss = loss + mask_w * F.binary_cross_entropy(out["weights_sum"].clip(1e-3, 1.0 - 1e-3), gt_mask)
    reg = field.laplace_density(out["sdfs"], field.get_beta().detach(), 1)
    loss = loss + cauchy_w * (1.0 / 4.0 * torch.log1p((1 - reg) ** 2 * 16.0)).mean()
    loss = loss + eikonal_w * ((out["sdf_gradients"].norm(p=2, dim=-1) - 1) ** 2).mean()
    return loss


class GraphedTrainStep:
    """One training step (render_train -> loss_epilogue -> backward) captured in a CUDA graph.

    The reference's steady-state train step has static shapes -- `mean_count` fixes the sample buffer, rays that do not
    fit are dropped by the march (raymarching.py:213-216, cuda_ray.py:64-79) -- and no host synchronisation, but issues ~600
    small kernels from Python; on B200 the step is then bound by the host (measured: 13.8 ms of CPU for 9.4 ms of GPU work).
    Capturing the whole forward + backward once and replaying it removes the host from the loop.  Gradients are left in
    `param.grad` (static tensors owned by the graph); the optimizer step stays outside, as in the reference trainer."""

    def __init__(self, field: TrainableField, bitfield: torch.Tensor, cfg: RenderConfig, n_rays: int, mean_count: int, *,
                 with_r_images: bool = True, perturb: bool = True, bg_color=1.0, loss_kwargs: Optional[dict] = None, warmup: int = 3):
        dev = bitfield.device
        self.field, self.n_rays = field, n_rays
        f32 = dict(dtype=torch.float32, device=dev)
        self.rays_o, self.rays_d = torch.zeros(n_rays, 3, **f32), torch.zeros(n_rays, 3, **f32)
        self.rays_d[:, 2] = 1.0
        self.gt_rgb, self.gt_mask = torch.zeros(n_rays, 3, **f32), torch.zeros(n_rays, **f32)
        self.r_images = torch.zeros(n_rays, 4, **f32) if with_r_images else None
        self.counter = torch.zeros(2, dtype=torch.int32, device=dev)
        lk = loss_kwargs or {}

        def body():
            for p in field.parameters():
                p.grad = None
            self.counter.zero_()
            out = render_train(field, bitfield, self.rays_o, self.rays_d, cfg, bg_color=bg_color, perturb=perturb, force_all_rays=False,
                               mean_count=mean_count, step_counter=self.counter, r_images=self.r_images,
                               ret_weights=lk.get("backsdf_w", 0.0) > 0)
            loss = loss_epilogue(field, out, self.gt_rgb, self.gt_mask, **lk)
            loss.backward()
            return loss.detach(), out["image"].detach(), out["weights_sum"].detach()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.image, self.weights_sum = body()

    def __call__(self, rays_o, rays_d, gt_rgb, gt_mask, r_images=None):
        """Copies the step's inputs into the graph's static buffers and replays; returns the (static) loss tensor."""
        self.rays_o.copy_(rays_o.view(-1, 3), non_blocking=True)
        self.rays_d.copy_(rays_d.view(-1, 3), non_blocking=True)
        self.gt_rgb.copy_(gt_rgb.view(-1, 3), non_blocking=True)
        self.gt_mask.copy_(gt_mask.view(-1), non_blocking=True)
        if self.r_images is not None and r_images is not None:
            self.r_images.copy_(r_images.view(-1, 4), non_blocking=True)
        self.graph.replay()
        return self.loss
