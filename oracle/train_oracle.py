"""CPU oracle of the TRAIN branch of the hot path -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).

Restates, on torch-CPU autograd (float64 MLP arithmetic) over the C restatements of the reference kernels:
  * run_cuda, training branch                       nerf/render_func/cuda_ray.py:64-168
  * NeRFNetwork.forward_geometry / forward_sigma    nerf/network.py:381-522, LaplaceDensity :26-44
  * NeRFRenderer.compute_normal (autograd.grad, create_graph)    nerf/renderer.py:182-198
  * get_color_mlp_extra_params / reflect_dir        nerf/renderer.py:20-39, 147-180
  * NeRFNetwork.forward_color                       nerf/network.py:524-698
  * hash_encode autograd.Function with second-order backward     hashencoder/hashgrid.py:17-107
  * composite_rays_train autograd.Function          raymarching/raymarching.py:249-310
and the loss terms of Trainer.train_step that touch every output of the branch (nerf/utils.py:661-808: colour L1,
mask BCE, Cauchy, eikonal), used only to drive gradients through all outputs in the parity tests.

Parity status: the torch glue is pinned through oracle.field_forward (same formulas, golden vectors from the reference's
network.py, tests/test_oracle_golden.py::test_field_glue_matches_reference_network) by
tests/test_train_cpu.py::test_train_oracle_forward_equals_pinned_field_oracle; the backward is torch autograd of that forward.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
from torch.autograd import Function

from . import oracle as O

SQRT3 = 3 ** 0.5


def _np32(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


# ----------------------------------------------------------------------------------------------------
# hash_encode with second-order backward (hashgrid.py:17-107) on the C restatement of the kernels
# ----------------------------------------------------------------------------------------------------

class _HashEncode(Function):
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs):
        out, dy_dx = O.hash_encode_forward(_np32(inputs), _np32(embeddings), offsets, per_level_scale, base_resolution, calc_grad_inputs)
        L, B, C = out.shape
        ctx.save_for_backward(inputs, embeddings)
        ctx.aux = (offsets, per_level_scale, base_resolution, dy_dx, calc_grad_inputs)
        return torch.from_numpy(np.ascontiguousarray(out.transpose(1, 0, 2).reshape(B, L * C))).to(inputs.dtype)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings = ctx.saved_tensors
        offsets, pls, H, dy_dx, calc = ctx.aux
        B = inputs.shape[0]
        L, C = offsets.shape[0] - 1, embeddings.shape[1]
        g = grad.reshape(B, L, C).permute(1, 0, 2).contiguous()                       # hashgrid.py:64
        grad_inputs, grad_embeddings = _HashEncodeBackward.apply(g, inputs, embeddings, offsets, pls, H, dy_dx)
        return (grad_inputs if calc else None), grad_embeddings, None, None, None, None


class _HashEncodeBackward(Function):
    @staticmethod
    def forward(ctx, grad, inputs, embeddings, offsets, pls, H, dy_dx):
        g_emb, g_in = O.hash_encode_backward(_np32(grad), _np32(inputs), _np32(embeddings), offsets, pls, H, dy_dx)
        ctx.save_for_backward(grad, inputs, embeddings)
        ctx.aux = (offsets, pls, H, dy_dx)
        return torch.from_numpy(g_in).to(inputs.dtype), torch.from_numpy(g_emb).to(embeddings.dtype)

    @staticmethod
    def backward(ctx, grad_grad_inputs, grad_grad_embeddings):
        # hashgrid.py:88-104: d(grad_inputs)/d(grad) and d(grad_inputs)/d(embeddings); the dependence of grad_embeddings on
        # `grad` (linear) is NOT propagated by the reference either (its kernel ignores grad_grad_embeddings).
        grad, inputs, embeddings = ctx.saved_tensors
        offsets, pls, H, dy_dx = ctx.aux
        if dy_dx is None:
            return torch.zeros_like(grad), None, torch.zeros_like(embeddings), None, None, None, None
        gg, g2 = O.hash_encode_second_backward(_np32(grad), _np32(inputs), _np32(embeddings), offsets, pls, H, dy_dx, _np32(grad_grad_inputs))
        return torch.from_numpy(gg).to(grad.dtype), None, torch.from_numpy(g2).to(embeddings.dtype), None, None, None, None


def hash_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False):
    return _HashEncode.apply(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs)


# ----------------------------------------------------------------------------------------------------
# composite_rays_train (raymarching.py:249-310) on the C restatement
# ----------------------------------------------------------------------------------------------------

class _CompositeTrain(Function):
    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh, ret_weights, input_alpha, accum_deltas):
        N = rays.shape[0]
        ws, depth, image, weights = O.composite_rays_train_forward(_np32(sigmas), _np32(rgbs), deltas, rays, N, T_thresh, ret_weights,
                                                                   input_alpha, accum_deltas)
        ctx.save_for_backward(sigmas, rgbs)
        ctx.aux = (deltas, rays, ws, depth, image, T_thresh, input_alpha, accum_deltas)
        dt = sigmas.dtype
        w = torch.from_numpy(weights).to(dt) if ret_weights else torch.zeros(0, dtype=dt)
        return torch.from_numpy(ws).to(dt), torch.from_numpy(depth).to(dt), torch.from_numpy(image).to(dt), w

    @staticmethod
    def backward(ctx, g_ws, g_depth, g_image, g_weights):
        sigmas, rgbs = ctx.saved_tensors
        deltas, rays, ws, depth, image, T_thresh, input_alpha, accum_deltas = ctx.aux
        gs, gr = O.composite_rays_train_backward(_np32(g_ws), _np32(g_image), _np32(g_depth), _np32(sigmas), _np32(rgbs), deltas, rays,
                                                 ws, image, depth, T_thresh, input_alpha, accum_deltas)
        return torch.from_numpy(gs).to(sigmas.dtype), torch.from_numpy(gr).to(rgbs.dtype), None, None, None, None, None, None


def composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh=1e-4, ret_weights=False, input_alpha=False, accum_deltas=True):
    return _CompositeTrain.apply(sigmas, rgbs, deltas, rays, T_thresh, ret_weights, input_alpha, accum_deltas)


# ----------------------------------------------------------------------------------------------------
# differentiable field (network.py / renderer.py)
# ----------------------------------------------------------------------------------------------------

STACKS = ("sdf", "env", "diffuse", "color", "renv")


def params_from_dict(P: Dict, dtype=torch.float64, frozen=("diffuse", "color"), device="cpu") -> Dict[str, torch.Tensor]:
    """Leaf tensors {embeddings, beta, <stack>.<i>.weight/bias}; toaster.ini freezes the colour and diffuse MLPs
    (frozen_mlps = [specular, diffuse], network.py:785-796)."""
    th: Dict[str, torch.Tensor] = {}
    th["embeddings"] = torch.from_numpy(P["embeddings"]).to(device=device, dtype=dtype).requires_grad_(True)
    th["beta"] = torch.tensor(float(P["beta"]), dtype=dtype, device=device, requires_grad=True)
    for name in STACKS:
        if P.get(name) is None:
            continue
        for i, (W, b) in enumerate(P[name]):
            th[f"{name}.{i}.weight"] = torch.from_numpy(W).to(device=device, dtype=dtype).requires_grad_(name not in frozen)
            th[f"{name}.{i}.bias"] = torch.from_numpy(b).to(device=device, dtype=dtype).requires_grad_(name not in frozen)
    return th


def _stack(th, name) -> List:
    out, i = [], 0
    while f"{name}.{i}.weight" in th:
        out.append((th[f"{name}.{i}.weight"], th[f"{name}.{i}.bias"]))
        i += 1
    return out


def _mlp(x, layers):
    for i, (W, b) in enumerate(layers):
        x = torch.nn.functional.linear(x, W, b)
        if i != len(layers) - 1:
            x = torch.relu(x)
    return x


def _unit(x, eps):
    return torch.nn.functional.normalize(x, dim=-1, eps=eps)


def get_beta(th, P):
    """LaplaceDensity.get_beta (network.py:39-44): clamp with a straight-through gradient."""
    b = th["beta"]
    return b + (torch.clamp(b.detach(), P["beta_min"], P["beta_max"]) - b.detach())


def laplace_density(sdf, beta, alpha=None):
    alpha = 1 / beta if alpha is None else alpha
    return alpha * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))              # network.py:32-37


def forward_sigma(th, P, xyzs, eikonal=True, encode=None):
    """forward_geometry + compute_normal + LaplaceDensity.  xyzs must require grad.  `encode` replaces the C-oracle hash
    encoder (oracle/ref_cuda.py passes the reference's own CUDA kernel)."""
    bound = float(P["bound"])
    x01 = (xyzs + bound) / (2 * bound)                                                    # hashgrid.py:161
    enc = (encode or hash_encode)(x01, th["embeddings"], P["offsets"], P["per_level_scale"], P["base_resolution"], True)
    L = P["offsets"].shape[0] - 1
    if P.get("enabled_levels", -1) > 0:                                                   # network.py:390-393
        mask = torch.zeros(L, 2, dtype=enc.dtype, device=enc.device)
        mask[: P["enabled_levels"]] = 1
        enc = enc * mask.reshape(-1)
    h = _mlp(enc, _stack(th, "sdf"))
    G = int(P["geo_feat_dim"])
    sdf = h[:, 0]
    geo = _unit(h[:, 1:1 + G], 1e-12)
    roughness = P["roughness_act_scale"] * torch.nn.functional.softplus(h[:, 1 + G:2 + G] + P["roughness_bias"]) * P["roughness_scale"]
    blend = torch.sigmoid(h[:, 2 + G:3 + G])
    grad_x = torch.autograd.grad(sdf, xyzs, torch.ones_like(sdf), retain_graph=True, create_graph=True)[0]   # renderer.py:182-198
    normals = _unit(grad_x, 1e-10)
    sigma = laplace_density(sdf, get_beta(th, P)) * P.get("density_scale", 1.0)
    return sdf, sigma, geo, normals, (grad_x if eikonal else None), roughness, blend


def ide(xyz, kappa_inv, deg):
    return O.ide_encode(xyz, kappa_inv, deg)


def forward_color(th, P, geo, dirs, normals, roughness, blend, r_images=None):
    w_o = -dirs
    n_dot = (normals * w_o).sum(-1, keepdim=True)
    w_r = 2 * n_dot * normals - w_o
    deg = int(P["ide_degree"])
    lis = P.get("light_intensity_scale", 1.0)
    w_r_enc = ide(w_r, roughness, deg) * lis
    n_enc = ide(normals, P["diffuse_kappa_inv"], deg) * lis
    env = _stack(th, "env")
    f_n = _unit(_mlp(n_enc, env), 1e-12)
    c_d = torch.sigmoid(_mlp(torch.cat([geo, f_n], -1), _stack(th, "diffuse")))
    f_r = _unit(_mlp(w_r_enc, env), 1e-12)
    hh = torch.cat([geo, normals], -1)
    color = _stack(th, "color")
    c_s = torch.sigmoid(_mlp(torch.cat([hh, f_r, n_dot], -1), color))
    if r_images is not None and P.get("renv") is not None:                                  # network.py:612-659, 682-690
        mask = roughness.squeeze(-1) < P["indir_roughness_thresh"]
        ri = r_images
        if ri.shape[-1] == 4:
            vis = ri[:, 3]
            ri = ri[:, :3] * vis[:, None]
            mask = mask & (vis > 0.9)
        rr = torch.sqrt(roughness / P["roughness_scale"] / 0.75)
        bw = 0.98 * blend if P.get("learn_indir_blend", False) else 0.95 * torch.sigmoid(80 * (rr - 0.18))
        f_e = _unit(_mlp(torch.cat([ri, rr], -1), _stack(th, "renv")), 1e-12)
        c_e = torch.sigmoid(_mlp(torch.cat([hh, f_e, n_dot], -1), color))
        c_s = torch.where(mask[:, None], c_s * bw + c_e * (1 - bw), c_s)
    return (c_d + c_s) * P.get("intensity_scale", 1.0)


# ----------------------------------------------------------------------------------------------------
# run_cuda, training branch (cuda_ray.py:64-168)
# ----------------------------------------------------------------------------------------------------

def render_train(th, P, rays_o, rays_d, bitfield, *, cascade=1, grid_size=128, min_near=0.2, aabb=None, dt_gamma=0.0,
                 max_steps=1024, T_thresh=1e-4, bg_color=1.0, early_stop_steps=-1, r_images=None, geometry_only=False,
                 noises=None, dtype=torch.float64) -> Dict[str, torch.Tensor]:
    rays_o = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3)
    rays_d = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
    N = rays_o.shape[0]
    bound = float(P["bound"])
    if aabb is None:
        aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = O.near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    xyzs, dirs, deltas, rays, counter = O.march_rays_train(rays_o, rays_d, bound, bitfield, cascade, grid_size, nears, fars, N * max_steps,
                                                           noises=noises, dt_gamma=dt_gamma, max_steps=max_steps,
                                                           early_stop_steps=early_stop_steps)
    M = int(counter[0])
    M += 128 - M % 128            # raymarching.py:235-241 with align = 128 (cuda_ray.py:79): zero-padded samples ride along
    xyzs, dirs, deltas = xyzs[:M], dirs[:M], deltas[:M]
    x = torch.from_numpy(xyzs).to(dtype).requires_grad_(True)
    d = torch.from_numpy(dirs).to(dtype)
    ri = None
    if r_images is not None:                                                               # cuda_ray.py:96-99
        idx = O.get_scatter_idx(rays, M)
        ri = torch.from_numpy(np.asarray(r_images, np.float32)[idx]).to(dtype)
    sdf, sigma, geo, normals, sdf_grad, roughness, blend = forward_sigma(th, P, x)
    out: Dict[str, torch.Tensor] = {}
    if geometry_only:                                                                      # cuda_ray.py:121-127
        ws, depth, nimg, _ = composite_rays_train(sigma, normals, deltas, rays, T_thresh, False)
        out["normal_image"] = _unit(nimg, 1e-12)
        image = None
    else:
        rgbs = forward_color(th, P, geo, d, normals, roughness, blend, ri)
        ws, depth, image, _ = composite_rays_train(sigma, rgbs, deltas, rays, T_thresh, False)
        image = image + (1 - ws).unsqueeze(-1) * bg_color
    depth = (depth + torch.from_numpy(nears).to(dtype)) * (depth != 0)
    out.update(image=image, depth=depth, weights_sum=ws, sigmas=sigma, sdfs=sdf, roughness=roughness, sdf_gradients=sdf_grad,
               xyzs=x, dirs=d, deltas=deltas, rays=rays, num_samples=M)
    return out


def loss_epilogue(th, P, out, gt_rgb, gt_mask, *, color_w=1.0, mask_w=1.0, cauchy_w=0.1, eikonal_w=0.01) -> torch.Tensor:
    """The loss terms of Trainer.train_step (nerf/utils.py:661-808) active for toaster.ini that reach every output of the
    branch: colour L1 (:661-662, color_l1_loss), mask BCE (:712-717), Cauchy (:762-776), eikonal (:793-798)."""
    dt = out["image"].dtype
    gt_rgb, gt_mask = torch.as_tensor(gt_rgb, dtype=dt), torch.as_tensor(gt_mask, dtype=dt)
    loss = color_w * (out["image"] - gt_rgb).abs().mean(-1).mean()
    loss = loss + mask_w * torch.nn.functional.binary_cross_entropy(out["weights_sum"].clip(1e-3, 1.0 - 1e-3), gt_mask)
    reg = laplace_density(out["sdfs"], get_beta(th, P).detach(), 1)
    loss = loss + cauchy_w * (1.0 / 4.0 * torch.log1p((1 - reg) ** 2 * 16.0)).mean()
    loss = loss + eikonal_w * ((out["sdf_gradients"].norm(p=2, dim=-1) - 1) ** 2).mean()
    return loss


def train_step(P: Dict, rays_o, rays_d, bitfield, gt_rgb, gt_mask, *, dtype=torch.float64, r_images=None, **kw):
    """One forward + backward.  Returns (loss, grads {name: float32 numpy}, out)."""
    th = params_from_dict(P, dtype)
    out = render_train(th, P, rays_o, rays_d, bitfield, r_images=r_images, dtype=dtype, **kw)
    loss = loss_epilogue(th, P, out, gt_rgb, gt_mask)
    leaves = {k: v for k, v in th.items() if v.requires_grad}
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    g = {k: (None if gr is None else gr.detach().to(torch.float32).numpy()) for k, gr in zip(leaves.keys(), grads)}
    return float(loss.detach()), g, out


# --------------------------------------------------------------------------------------
# ray generation + loss epilogue: the steps either side of the training branch
# --------------------------------------------------------------------------------------

def get_rays(poses, intrinsics, H: int, W: int, inds=None, dtype=torch.float64):
    """nerf/utils.py:110-209 for given pixel indices (`inds` [N] flat h * W + w, None = all pixels): pixel centres + 0.5,
    normalised directions rotated by the camera-to-world matrix.  poses [B,4,4].  -> rays_o, rays_d [B,N,3]."""
    poses = torch.as_tensor(np.asarray(poses), dtype=dtype)
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    inds = torch.arange(H * W) if inds is None else torch.as_tensor(np.asarray(inds)).long()
    i = (inds % W).to(dtype) + 0.5
    j = (inds // W).to(dtype) + 0.5
    zs = torch.ones_like(i)
    directions = torch.stack(((i - cx) / fx * zs, (j - cy) / fy * zs, zs), dim=-1)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_d = directions[None] @ poses[:, :3, :3].transpose(-1, -2)
    rays_o = poses[:, None, :3, 3].expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()


def aux_point_mask(deltas, rays, M: int):
    """The point mask of run_cuda's auxiliary block (cuda_ray.py:173-190): every sample but the last of each valid ray, whose
    successor in the buffer is a real, contiguous continuation of the same ray.  deltas [M,2], rays [N,3] torch tensors."""
    point_mask = torch.ones(M, dtype=torch.bool)
    ray_valid = (rays[:, 2] > 0) * (rays[:, 1] + rays[:, 2] < M)
    end = rays[ray_valid, 1] + rays[ray_valid, 2] - 1
    point_mask[end.long()] = False
    ds = torch.roll(deltas, -1, 0)
    delta_mask = (ds[:, 0] > 0) * (ds[:, 1] > 0)
    cont_mask = ds[:, 1] < 1.2 * ds[:, 0]
    return point_mask * delta_mask * cont_mask, ds


def train_loss(image, weights_sum, gt_rgb, gt_mask, sdfs, sdf_gradients, weights, deltas, rays, beta: float, *, color_l1=True,
               color_w=1.0, mask_w=1.0, cauchy_w=0.1, eikonal_w=0.001, backsdf_w=5e-3, backsdf_thresh=0.01, backsdf_mean=False,
               dtype=torch.float64):
    """Trainer.train_step's loss terms for the shipped scene configs (nerf/utils.py:661-662 colour, :712-717 mask BCE, :735-747
    back-sdf, :762-776 Cauchy, :793-798 eikonal) on the outputs of run_cuda's training branch including its auxiliary block
    (cuda_ray.py:173-211: relsdf / sdf_weights / sdf_dist and the MASKED sdfs the Cauchy term then sees).  A term with weight 0
    is skipped as in the reference; backsdf_w == 0 also switches the auxiliary block off (Cauchy then runs over all M samples).
    The compositing weights carry no gradient (the reference's composite backward ignores grad_weights, raymarching.py:291).
    Returns (terms dict of floats incl. 'total', grads dict: image, weights_sum, sdfs, sdf_gradients as float32 numpy)."""
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
    image, weights_sum, sdfs, sdf_gradients = [t(a).clone().requires_grad_(True) for a in (image, weights_sum, sdfs, sdf_gradients)]
    gt_rgb, gt_mask, weights, deltas = t(gt_rgb), t(gt_mask), t(weights), t(deltas)
    rays = torch.as_tensor(np.asarray(rays)).long()
    M = sdfs.shape[0]
    terms = {}
    diff = image - gt_rgb
    color = (diff.abs() if color_l1 else diff ** 2).mean(-1).mean()
    loss = color_w * color
    terms["color"] = color
    if mask_w > 0:
        lm = torch.nn.functional.binary_cross_entropy(weights_sum.clip(1e-3, 1.0 - 1e-3), gt_mask)
        loss = loss + mask_w * lm
        terms["mask"] = lm
    sd = sdfs
    if backsdf_w > 0:
        pm, ds = aux_point_mask(deltas, rays, M)
        relsdf = (torch.roll(sdfs, -1, dims=0) - sdfs)[pm]
        w, dist, sd = weights[pm], ds[pm, 1], sdfs[pm]
        mask = (w > backsdf_thresh) * (relsdf > 0)
        s_sq = relsdf[mask] ** 2
        mw = w[mask]
        r_cos_sq = s_sq / (dist[mask].clamp(min=5e-4) ** 2 + s_sq)
        denom = (1 + mw.sum()) if backsdf_mean else 1
        lb = (mw * r_cos_sq).sum() / denom
        loss = loss + backsdf_w * lb
        terms["backsdf"] = lb
    if cauchy_w > 0:
        reg = laplace_density(sd, torch.tensor(float(beta), dtype=dtype), 1)
        lc = 1.0 / 4.0 * torch.log1p((1 - reg) ** 2 * 16.0).mean()
        loss = loss + cauchy_w * lc
        terms["cauchy"] = lc
    if eikonal_w > 0:
        le = ((sdf_gradients.norm(p=2, dim=-1) - 1) ** 2).mean()
        loss = loss + eikonal_w * le
        terms["eikonal"] = le
    terms["total"] = loss
    leaves = [image, weights_sum, sdfs, sdf_gradients]
    g = torch.autograd.grad(loss, leaves, allow_unused=True)
    f32 = lambda x, like: (torch.zeros_like(like) if x is None else x).detach().to(torch.float32).numpy()
    grads = dict(image=f32(g[0], image), weights_sum=f32(g[1], weights_sum), sdfs=f32(g[2], sdfs), sdf_gradients=f32(g[3], sdf_gradients))
    return {k: float(v.detach()) for k, v in terms.items()}, grads
