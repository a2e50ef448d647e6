"""The reference's CUDA path on this GPU -- TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/oracle.py header).

What nexuslrf/ENVIDR itself executes for an inference frame, restated around the reference's OWN kernels:
  * kernels: oracle/_ref/_raymarching.so and _hashencoder.so = the unmodified reference sources compiled for sm_100a by
    oracle/build_ref.py (raymarching/src/raymarching.cu, hashencoder/src/hashencoder.cu);
  * Python around them restated from the reference: the march / composite wrappers (raymarching/raymarching.py:316-394),
    the hash_encode autograd.Function (hashencoder/hashgrid.py:17-107), the MLP / IDE / normal glue (nerf/network.py,
    nerf/renderer.py:147-198; shared with oracle/train_oracle.py, torch ops -> cuBLAS fp32 on the GPU), the inference while-loop
    with its host-side compaction and `get_normal_image=True` (nerf/render_func/cuda_ray.py:238-359, as Trainer.eval_step
    calls it, utils.py:857-859) and the three-pass indirect-reflection render (nerf/renderer.py:439-513).
Used by tests/test_gpu_refpath.py (our fused path against the reference's kernels end to end) and by bench.py as the
`gpu_reference` figure: the "reference rays/sec on the same B200" the north star's >= 10x target refers to.
The IDE here is the oracle's real-arithmetic formulation (fewer launches than the reference's complex pow): the figure is
conservative in the reference's favour."""
from __future__ import annotations

import math
import os
import sys
from typing import Dict, Optional

import numpy as np
import torch
from torch.autograd import Function

from . import train_oracle as TO

_HERE = os.path.dirname(os.path.abspath(__file__))
_mods: Dict[str, object] = {}


def ref_module(name: str):
    """Import oracle/_ref/<name>.so (raises ImportError when it was not built / shipped)."""
    if name not in _mods:
        d = os.path.join(_HERE, "_ref")
        if d not in sys.path:
            sys.path.insert(0, d)
        _mods[name] = __import__(name)
    return _mods[name]


def available() -> bool:
    try:
        ref_module("_raymarching"); ref_module("_hashencoder")
        return True
    except Exception:
        return False


class _hash_encode(Function):
    """hashgrid.py:17-86 (forward with dy_dx, backward to inputs and embeddings); first order is all inference needs."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs):
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, embeddings.shape[1]
        S, H = np.log2(per_level_scale), base_resolution
        inputs, embeddings = inputs.contiguous(), embeddings.contiguous()
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=inputs.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=inputs.dtype) if calc_grad_inputs else \
            torch.empty(1, device=inputs.device, dtype=inputs.dtype)
        ref_module("_hashencoder").hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H, calc_grad_inputs)
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, calc = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_inputs, grad_embeddings = torch.zeros_like(inputs), torch.zeros_like(embeddings)
        ref_module("_hashencoder").hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc, dy_dx,
                                                        grad_inputs)
        return (grad_inputs if calc else None), grad_embeddings, None, None, None, None


def _encode(x01, emb, offsets, pls, H, calc):
    return _hash_encode.apply(x01, emb, offsets, pls, H, calc)


class RefField:
    """Field state on the device as the reference model holds it (fp32 tensors)."""

    def __init__(self, P: Dict, device):
        self.P = dict(P)
        self.P["offsets"] = torch.from_numpy(np.asarray(P["offsets"], np.int32)).to(device)
        self.th = TO.params_from_dict(P, torch.float32, device=device)
        self.device = device


def run_cuda_infer(F_: RefField, bitfield, rays_o, rays_d, *, cascade=1, grid_size=128, min_near=0.2, dt_gamma=0.0, max_steps=1024,
                   T_thresh=1e-4, bg_color=1.0, r_images=None, geometry_only=False, stats: Optional[dict] = None):
    """cuda_ray.py:238-359 with get_normal_image=True, on the reference's kernels."""
    R = ref_module("_raymarching")
    P, th, dev = F_.P, F_.th, F_.device
    rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
    N = rays_o.shape[0]
    bound = float(P["bound"])
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    nears, fars = torch.empty(N, **f32), torch.empty(N, **f32)
    R.near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
    ws, depth, image = torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(N, 3, **f32)
    n_ws, n_depth, n_img = torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(N, 3, **f32)
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    n_alive_list, rays_t, n_t = alive.clone(), nears.clone(), nears.clone()
    step = samples = iters = 0
    while step < max_steps:
        n_alive = alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        align = 128
        ri = None
        if r_images is not None:
            ri = r_images[alive.long()][:, None, :].expand(-1, n_step, -1).reshape(-1, r_images.shape[-1])
            align = -1
        M = n_alive * n_step                                               # raymarching.py:339-343
        if align > 0:
            M += align - (M % align)
        xyzs, dirs, deltas = torch.zeros(M, 3, **f32), torch.zeros(M, 3, **f32), torch.zeros(M, 2, **f32)
        noises = torch.zeros(n_alive, **f32)
        R.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, cascade, grid_size, bitfield, nears, fars,
                     xyzs, dirs, deltas, noises)
        iters += 1
        if stats is not None:
            samples += int((deltas[:, 0] > 0).sum())
        with torch.enable_grad():
            xyzs.requires_grad = True
            sdf, sigma, geo, normals, _, rough, blend = TO.forward_sigma(th, P, xyzs, eikonal=False, encode=_encode)
            sigma = sigma.detach()
            if geometry_only:
                R.composite_rays(n_alive, n_step, T_thresh, 1, 0, alive, rays_t, sigma, normals.detach().contiguous(), deltas, ws, depth, n_img)
            else:
                rgbs = TO.forward_color(th, P, geo, dirs, normals, rough, blend, ri).detach().contiguous()
                R.composite_rays(n_alive, n_step, T_thresh, 1, 0, alive, rays_t, sigma, rgbs, deltas, ws, depth, image)
                R.composite_rays(n_alive, n_step, T_thresh, 1, 0, n_alive_list, n_t, sigma, normals.detach().contiguous(), deltas, n_ws,
                                 n_depth, n_img)
                n_alive_list = n_alive_list[n_alive_list >= 0]
        alive = alive[alive >= 0]
        step += n_step
    if stats is not None:
        stats.update(samples=samples, iterations=iters)
    n_img = torch.nn.functional.normalize(n_img, dim=-1, eps=1e-10)
    if geometry_only:
        return dict(image=None, depth=depth, weights_sum=ws, normal_image=n_img)
    return dict(image=image + (1 - ws).unsqueeze(-1) * bg_color, depth=depth, weights_sum=ws, normal_image=n_img)


def render(F_: RefField, bitfield, rays_o, rays_d, *, indir_ref=False, indir_max_steps=1024, bg_color=1.0, min_near=0.2, max_steps=1024,
           T_thresh=1e-4, stats: Optional[list] = None):
    """NeRFRenderer.render, cuda_ray path (renderer.py:364-531): one pass or the three passes of :439-513."""
    rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
    N = rays_o.shape[0]

    def run(o, d, **k):
        st = {} if stats is not None else None
        r = run_cuda_infer(F_, bitfield, o, d, max_steps=k.pop("max_steps", max_steps), T_thresh=T_thresh, stats=st, **k)
        if stats is not None:
            stats.append(st)
        return r

    if not indir_ref:
        res = run(rays_o, rays_d, bg_color=bg_color, min_near=min_near)
    else:
        dt = 2 * math.sqrt(3) / indir_max_steps
        geo = run(rays_o, rays_d, geometry_only=True, min_near=min_near)
        normals, depth, ws = geo["normal_image"], geo["depth"] - dt, geo["weights_sum"]
        ref_mask = (depth != 0) & (ws > 0.9)
        ray_mask = (depth != 0) & (ws > 0.3)
        ref_o = rays_o + depth[:, None] * rays_d
        w_o = -rays_d
        ref_d = 2 * (w_o * normals).sum(-1, keepdim=True) * normals - w_o
        sec = run(ref_o[ref_mask], ref_d[ref_mask], bg_color=0.0, min_near=dt * 2, max_steps=indir_max_steps)
        ref_image = torch.cat([sec["image"], sec["weights_sum"][:, None]], -1)
        r_img = ref_image.new_zeros(int(ray_mask.sum()), 4)
        r_img[ref_mask[ray_mask]] = ref_image
        main = run(rays_o[ray_mask], rays_d[ray_mask], bg_color=0.0, r_images=r_img, min_near=min_near)
        img = normals.new_zeros(N, 3); img[ray_mask] = main["image"]
        wsf = normals.new_zeros(N); wsf[ray_mask] = main["weights_sum"]
        res = dict(image=(torch.zeros_like(img) + bg_color) * (1 - wsf[:, None]) + img, weights_sum=wsf, depth=depth, normal_image=normals)
    w = res["weights_sum"][:, None]
    res["normal_image"] = res["normal_image"] * w + (1 - w)
    return res


@torch.no_grad()
def update_extra_state(F_: RefField, density_grid, density_bitfield, *, cascade=1, grid_size=128, decay=0.95, density_thresh=0.01,
                       S=128, noise=None, iter_density=0, full_update=False):
    """NeRFRenderer.update_extra_state (renderer.py:264-352) as the reference executes it: torch ops + its own morton3D /
    packbits kernels + its hash encoder kernel + torch fp32 MLPs, including the mean().item() round trip.
    density_grid [C, H^3] is updated in place; returns (mean_density, density_bitfield).
    noise: list (one per cascade) of [H^3, 3] tensors in meshgrid order replacing torch.rand_like (full update only)."""
    R = ref_module("_raymarching")
    P, th, dev = F_.P, F_.th, F_.device
    H = grid_size
    bound_model = float(P["bound"])

    def morton3D(coords):
        coords = coords.int().contiguous()
        idx = torch.empty(coords.shape[0], dtype=torch.int32, device=dev)
        R.morton3D(coords, coords.shape[0], idx)
        return idx

    def morton3D_invert(indices):
        indices = indices.int().contiguous()
        c = torch.empty(indices.shape[0], 3, dtype=torch.int32, device=dev)
        R.morton3D_invert(indices, indices.shape[0], c)
        return c

    def density(x):                                                                      # network.py:713-716 under no_grad
        x01 = (x + bound_model) / (2 * bound_model)
        enc = _encode(x01, th["embeddings"], P["offsets"], P["per_level_scale"], P["base_resolution"], False)
        if P.get("enabled_levels", -1) > 0:
            L = P["offsets"].shape[0] - 1
            mask = torch.zeros(L, 2, dtype=enc.dtype, device=dev)
            mask[: P["enabled_levels"]] = 1
            enc = enc * mask.reshape(-1)
        sdf = TO._mlp(enc, TO._stack(th, "sdf"))[:, 0]
        return TO.laplace_density(sdf, TO.get_beta(th, P))

    tmp_grid = -torch.ones_like(density_grid)
    if iter_density < 16 or full_update:
        X = torch.arange(H, dtype=torch.int32, device=dev).split(S)
        for xs in X:
            for ys in X:
                for zs in X:
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
                    coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                    indices = morton3D(coords).long()
                    xyzs = 2 * coords.float() / (H - 1) - 1
                    for cas in range(cascade):
                        bound = min(2 ** cas, bound_model)
                        half_grid_size = bound / H
                        cas_xyzs = xyzs * (bound - half_grid_size)
                        u = torch.rand_like(cas_xyzs) if noise is None else noise[cas]
                        cas_xyzs += (u * 2 - 1) * half_grid_size
                        sigmas = density(cas_xyzs).reshape(-1).detach()
                        sigmas *= P.get("density_scale", 1.0)
                        tmp_grid[cas, indices] = sigmas
    else:
        N = H ** 3 // 4
        for cas in range(cascade):
            coords = torch.randint(0, H, (N, 3), device=dev)
            indices = morton3D(coords).long()
            occ_indices = torch.nonzero(density_grid[cas] > 0).squeeze(-1)
            rand_mask = torch.randint(0, occ_indices.shape[0], [N], dtype=torch.long, device=dev)
            occ_indices = occ_indices[rand_mask]
            occ_coords = morton3D_invert(occ_indices)
            indices = torch.cat([indices, occ_indices], dim=0)
            coords = torch.cat([coords, occ_coords], dim=0)
            xyzs = 2 * coords.float() / (H - 1) - 1
            bound = min(2 ** cas, bound_model)
            half_grid_size = bound / H
            cas_xyzs = xyzs * (bound - half_grid_size)
            cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half_grid_size
            sigmas = density(cas_xyzs).reshape(-1).detach()
            sigmas *= P.get("density_scale", 1.0)
            tmp_grid[cas, indices] = sigmas
    valid_mask = (density_grid >= 0) & (tmp_grid >= 0)
    density_grid[valid_mask] = torch.maximum(density_grid[valid_mask] * decay, tmp_grid[valid_mask])
    mean_density = torch.mean(density_grid.clamp(min=0)).item()
    thresh = min(mean_density, density_thresh)
    N8 = density_grid.numel() // 8
    R.packbits(density_grid.contiguous(), N8, thresh, density_bitfield)
    return mean_density, density_bitfield
