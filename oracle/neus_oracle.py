"""CPU restatement of BASELINE config 4 (NeuS-style geometry, no hash grid) -- TEST / BASELINE INFRASTRUCTURE ONLY
(see oracle/oracle.py header; nothing under envidr_b200/ may import this).

Restated reference code:
  * FreqEncoder column order (freqencoder/src/freqencoder.cu:47-56): [x | sin(2^0 x) | cos(2^0 x) | sin(2^1 x) | ...], D columns per block;
    here in differentiable torch (float64) -- checked against the C restatement orc_freq_encode_forward in tests/test_oracle_golden.py
  * forward_geometry with geometric_init / skip_layers (nerf/network.py:154-222, 415-448): weight-normed layers (folded),
    Softplus(beta=100) between them, h = cat([h, x]) / sqrt(2) in front of a skip layer, head = (sdf, unitNorm(geo_feat), roughness, blend)
  * compute_normal (nerf/renderer.py:182-198): autograd.grad of sdf w.r.t. xyz, F.normalize(eps=1e-10)
  * NeuSDensity (network.py:69-102) -> alpha, composited with input_alpha (oracle.composite_rays)
  * forward_color (network.py:524-698): the same block as oracle.field_forward, on this geometry
Pinned to the reference's own NeRFNetwork by tests/golden/neus_field.npz (tests/golden/make_golden.py::gen_neus_field).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import oracle as O


def freq_encode(x: torch.Tensor, degree: int) -> torch.Tensor:
    out = [x]
    for f in range(degree):
        out += [torch.sin(x * (2.0 ** f)), torch.cos(x * (2.0 ** f))]
    return torch.cat(out, -1)


def geometry(P: Dict, xyzs, dirs, dists, dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """P: sdf = [(W, b)] effective (weight norm folded) numpy weights, skip_layers, multires, geo_feat_dim, variance,
    cos_anneal_ratio, roughness_* scalars.  Returns torch tensors (dtype)."""
    x = torch.as_tensor(np.asarray(xyzs), dtype=dtype).clone().requires_grad_(True)
    d = torch.as_tensor(np.asarray(dirs), dtype=dtype)
    enc = freq_encode(x, int(P["multires"]))
    h = enc
    layers = P["sdf"]
    for l, (W, b) in enumerate(layers):
        if l in P["skip_layers"]:
            h = torch.cat([h, enc], -1) / math.sqrt(2.0)
        h = h @ torch.as_tensor(W, dtype=dtype).T + torch.as_tensor(b, dtype=dtype)
        if l != len(layers) - 1:
            h = torch.nn.functional.softplus(h, beta=100)
    G = int(P["geo_feat_dim"])
    sdf = h[:, 0]
    grad_x = torch.autograd.grad(sdf.sum(), x)[0]
    normals = grad_x / grad_x.norm(dim=-1, keepdim=True).clamp_min(1e-10)
    geo = h[:, 1:1 + G]
    geo = geo / geo.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    roughness = P["roughness_act_scale"] * torch.nn.functional.softplus(h[:, 1 + G:2 + G] + P["roughness_bias"]) * P["roughness_scale"]
    blend = torch.sigmoid(h[:, 2 + G:3 + G]) if h.shape[1] > 2 + G else None
    dist = torch.as_tensor(np.asarray(dists), dtype=dtype) if not np.isscalar(dists) else float(dists)
    inv_s = float(np.clip(np.exp(float(P["variance"]) * 10.0), 1e-6, 1e6))
    true_cos = (d * normals).sum(-1)
    r = float(P.get("cos_anneal_ratio", 1.0))
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - r) + torch.relu(-true_cos) * r)
    nxt, prv = sdf + iter_cos * dist * 0.5, sdf - iter_cos * dist * 0.5
    pc, nc = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
    alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0)
    det = lambda t: None if t is None else t.detach()
    return dict(sdf=det(sdf), grad_x=det(grad_x), normal=det(normals), geo=det(geo), roughness=det(roughness), blend=det(blend),
                sigma=det(alpha), dirs=d)


def forward_color(P: Dict, g: Dict[str, torch.Tensor], r_images=None, env_rot_radian: Optional[float] = None, dtype=torch.float64,
                  ide_dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """network.py:524-698 + renderer.py:147-180 on the outputs of `geometry` (the block of oracle.field_forward)."""
    geo, normals, roughness, blend, d = g["geo"], g["normal"], g["roughness"], g["blend"], g["dirs"]
    w_o = -d
    n_dot = (normals * w_o).sum(-1, keepdim=True)
    w_r = 2 * n_dot * normals - w_o
    n_env = normals
    if env_rot_radian is not None:
        R = torch.from_numpy(O.rot_theta3(env_rot_radian)).to(dtype)
        w_r, n_env = w_r @ R, normals @ R
    deg = int(P["ide_degree"])
    lis = P.get("light_intensity_scale", 1.0)
    w_r_enc = O.ide_encode(w_r.to(ide_dtype), roughness.to(ide_dtype), deg).to(dtype) * lis
    n_enc = O.ide_encode(n_env.to(ide_dtype), P["diffuse_kappa_inv"], deg).to(dtype) * lis
    f_n = O._unit(O._mlp(n_enc, P["env"]), 1e-12)
    c_d = torch.sigmoid(O._mlp(torch.cat([geo, f_n], -1), P["diffuse"]))
    f_r = O._unit(O._mlp(w_r_enc, P["env"]), 1e-12)
    hh = torch.cat([geo, normals], -1)
    c_s = torch.sigmoid(O._mlp(torch.cat([hh, f_r, n_dot], -1), P["color"]))
    if r_images is not None and P.get("renv") is not None:
        ri = torch.from_numpy(np.asarray(r_images, np.float32)).to(dtype)
        mask = roughness.squeeze(-1) < P["indir_roughness_thresh"]
        vis = ri[:, 3]
        ri = ri[:, :3] * vis[:, None]
        mask = mask & (vis > 0.9)
        rr = torch.sqrt(roughness / P["roughness_scale"] / 0.75)
        bw = 0.98 * blend if P.get("learn_indir_blend", False) else 0.95 * torch.sigmoid(80 * (rr - 0.18))
        f_e = O._unit(O._mlp(torch.cat([ri, rr], -1), P["renv"]), 1e-12)
        c_e = torch.sigmoid(O._mlp(torch.cat([hh, f_e, n_dot], -1), P["color"]))
        c_s = torch.where(mask[:, None], c_s * bw + c_e * (1 - bw), c_s)
    return dict(rgb=(c_d + c_s) * P.get("intensity_scale", 1.0), c_diffuse=c_d, c_specular=c_s)


def field_forward(P: Dict, xyzs, dirs, dists, r_images=None, env_rot_radian=None, dtype=torch.float64) -> Dict[str, np.ndarray]:
    g = geometry(P, xyzs, dirs, dists, dtype)
    c = forward_color(P, g, r_images, env_rot_radian, dtype)
    f32 = lambda t: None if t is None else t.to(torch.float32).numpy()
    return dict(sdf=f32(g["sdf"]), sigma=f32(g["sigma"]), normal=f32(g["normal"]), grad_x=f32(g["grad_x"]), geo_feat=f32(g["geo"]),
                roughness=f32(g["roughness"]), blend=f32(g["blend"]), rgb=f32(c["rgb"]), c_diffuse=f32(c["c_diffuse"]), c_specular=f32(c["c_specular"]))


def render_rays(P: Dict, rays_o, rays_d, bitfield, *, cascade=1, grid_size=128, min_near=0.2, dt_gamma=0.0, max_steps=1024, T_thresh=1e-4,
                bg_color=1.0, env_rot_radian=None, dtype=torch.float32, stats: Optional[dict] = None) -> Dict[str, np.ndarray]:
    """run_cuda inference loop (cuda_ray.py:238-359) for the NeuS field: the oracle's march / composite C restatements with
    input_alpha=True, geometry + colour per iteration as the reference does."""
    rays_o, rays_d = O._c32(rays_o).reshape(-1, 3), O._c32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    bound = float(P.get("bound", 1.0))
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = O.near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    n_ws, n_depth, n_img = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    alive = np.arange(N, dtype=np.int32)
    n_alive_list, rays_t, n_t = alive.copy(), nears.copy(), nears.copy()
    step = iters = samples = 0
    while step < max_steps:
        n_alive = alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas, _ = O.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, bound, bitfield, cascade, grid_size, nears, fars,
                                             align=128, dt_gamma=dt_gamma, max_steps=max_steps)
        out = field_forward(P, xyzs, dirs, deltas[:, 0], None, env_rot_radian, dtype)
        samples += int((deltas[:, 0] > 0).sum())
        O.composite_rays(n_alive, n_step, alive, rays_t, out["sigma"], out["rgb"], deltas, ws, depth, image, T_thresh=T_thresh, input_alpha=True)
        O.composite_rays(n_alive, n_step, n_alive_list, n_t, out["sigma"], out["normal"], deltas, n_ws, n_depth, n_img, T_thresh=T_thresh,
                         input_alpha=True)
        n_alive_list = n_alive_list[n_alive_list >= 0]
        alive = alive[alive >= 0]
        step += n_step
        iters += 1
    if stats is not None:
        stats.update(samples=samples, iterations=iters)
    nn = np.maximum(np.linalg.norm(n_img, axis=-1, keepdims=True), 1e-10)
    return dict(image=image + (1 - ws)[:, None] * np.float32(bg_color), depth=depth, weights_sum=ws, normal_image=n_img / nn)
