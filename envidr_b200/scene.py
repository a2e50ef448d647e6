"""Seeded synthetic stand-ins for the benchmark scenes (there is no dataset / checkpoint access):
cameras and rays restated from the reference's data path, and an analytic "toaster-like" field whose hash
grid + MLP weights are *constructed* (not trained) so that both the CUDA path and the CPU oracle load the
same arrays.

Restated reference code (the caller side of the hot path):
    get_rays              nerf/utils.py:110-209 (full-image branch: pixel centres + 0.5, normalised dirs)
    pose_spherical        nerf/sph_loader.py:67-76
    nerf_matrix_to_ngp    nerf/provider.py:32-40
    HashEncoder offsets   hashencoder/hashgrid.py:130-146
    net_init xavier       nerf/net_init.py (xavier_uniform weights, zero bias)
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

from .field import FieldParams

SQRT3 = 3 ** 0.5


# ---------------------------------------------------------------------------------------------
# cameras / rays
# ---------------------------------------------------------------------------------------------

def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    trans_t = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], np.float64)
    p = phi_deg / 180.0 * np.pi
    rot_phi = np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]])
    t = theta_deg / 180.0 * np.pi
    rot_th = np.array([[np.cos(t), 0, -np.sin(t), 0], [0, 1, 0, 0], [np.sin(t), 0, np.cos(t), 0], [0, 0, 0, 1]])
    c2w = rot_th @ (rot_phi @ trans_t)
    return np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float64) @ c2w


def nerf_matrix_to_ngp(pose: np.ndarray, scale: float = 0.33, offset=(0, 0, 0)) -> np.ndarray:
    return np.array([
        [pose[1, 0], -pose[1, 1], -pose[1, 2], pose[1, 3] * scale + offset[0]],
        [pose[2, 0], -pose[2, 1], -pose[2, 2], pose[2, 3] * scale + offset[1]],
        [pose[0, 0], -pose[0, 1], -pose[0, 2], pose[0, 3] * scale + offset[2]],
        [0, 0, 0, 1]], dtype=np.float32)


def intrinsics_from_fov(W: int, H: int, camera_angle_x: float) -> np.ndarray:
    fl = W / (2 * np.tan(camera_angle_x / 2))
    return np.array([fl, fl, W / 2, H / 2], np.float64)


def get_rays(pose: np.ndarray, intrinsics: np.ndarray, H: int, W: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Full-image rays, [H*W, 3] each (torch fp32, CPU)."""
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    pose_t = torch.from_numpy(np.asarray(pose, np.float32))[None]
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t().reshape(1, H * W) + 0.5
    j = j.t().reshape(1, H * W) + 0.5
    zs = torch.ones_like(i)
    directions = torch.stack(((i - cx) / fx * zs, (j - cy) / fy * zs, zs), dim=-1)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_d = directions @ pose_t[:, :3, :3].transpose(-1, -2)
    rays_o = pose_t[..., :3, 3][..., None, :].expand_as(rays_d)
    return rays_o.reshape(-1, 3).contiguous(), rays_d.reshape(-1, 3).contiguous()


def camera_rays(W: int, H: int, theta_deg: float = 40.0, phi_deg: float = -30.0, radius: float = 4.0, scale: float = 0.65,
                camera_angle_x: float = 0.6911112070083618):
    pose = nerf_matrix_to_ngp(pose_spherical(theta_deg, phi_deg, radius), scale=scale)
    return get_rays(pose, intrinsics_from_fov(W, H, camera_angle_x), H, W)


# ---------------------------------------------------------------------------------------------
# analytic scene
# ---------------------------------------------------------------------------------------------

def analytic_sdf(xyz: np.ndarray, scale: float = 0.65) -> np.ndarray:
    """Rounded box (0.55 x 0.35 x 0.3) united with a sphere r=0.25 at (0.3, 0.25, 0), all times `scale`
    (the synthetic toaster of SURVEY.md 8d).  xyz [..., 3] float64 -> signed distance."""
    p = np.asarray(xyz, np.float64) / scale
    half = np.array([0.55, 0.35, 0.30])
    r = 0.08
    q = np.abs(p) - (half - r)
    box = np.linalg.norm(np.maximum(q, 0.0), axis=-1) + np.minimum(q.max(axis=-1), 0.0) - r
    sph = np.linalg.norm(p - np.array([0.3, 0.25, 0.0]), axis=-1) - 0.25
    return np.minimum(box, sph) * scale


def hash_offsets(num_levels=16, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048, input_dim=3):
    per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        offsets.append(offset)
        offset += min(2 ** log2_hashmap_size, res ** input_dim)
    offsets.append(offset)
    return np.array(offsets, np.int32), float(per_level_scale)


def _xavier(rng: np.random.Generator, out_dim: int, in_dim: int) -> np.ndarray:
    a = math.sqrt(6.0 / (in_dim + out_dim))
    return rng.uniform(-a, a, size=(out_dim, in_dim)).astype(np.float32)


def _mlp(rng, dims, bias_scale=0.0):
    layers = []
    for i in range(len(dims) - 1):
        W = _xavier(rng, dims[i + 1], dims[i])
        b = (rng.standard_normal(dims[i + 1]) * bias_scale).astype(np.float32)
        layers.append((W, b))
    return layers


def make_synthetic_field(seed: int = 0, *, hidden_dim_env: int = 256, ide_degree: int = 5, scene_scale: float = 0.65,
                         bound: float = 1.0, beta: float = 0.01, num_levels: int = 16, log2_hashmap_size: int = 19,
                         desired_resolution: int = 2048, with_renv: bool = True, device="cpu", precision: str = "fp32") -> FieldParams:
    """toaster.ini dimensions (hash L16/C2/base16/2048/T19, sdf 32-64-64-15, env IDE-256-256-256-12, diffuse 24-32-3,
    color 28-64-64-3, renv 4-64-64-64-12) with weights built so that the SDF head equals the mean of the
    smoothstep-interpolated analytic SDF stored in channel 0 of the dense levels; everything else is seeded random.
    precision: the synthetic field defaults to the exact FFMA path ("fp32") because the parity tests compare it with the fp64
    oracle at 1e-6..1e-5; the benchmark and the drop-in entries use "tc" (FieldParams' own default)."""
    rng = np.random.default_rng(seed)
    offsets, pls = hash_offsets(num_levels, 16, log2_hashmap_size, desired_resolution * bound)
    T = int(offsets[-1])
    emb = rng.uniform(-0.3, 0.3, size=(T, 2)).astype(np.float32)
    S = np.float32(np.log2(pls))
    dense_levels = []
    for l in range(num_levels):
        scale = np.float32(np.exp2(np.float32(l) * S) * np.float32(16) - np.float32(1.0))
        res = int(np.ceil(scale)) + 1
        if res ** 3 > 2 ** log2_hashmap_size:
            break
        dense_levels.append(l)
        idx = np.arange(res)
        x01 = idx.astype(np.float64) / float(scale)
        gx, gy, gz = np.meshgrid(x01, x01, x01, indexing="ij")
        pts = np.stack([gx, gy, gz], -1) * 2 * bound - bound
        sd = analytic_sdf(pts, scene_scale)
        lin = (idx[:, None, None] + idx[None, :, None] * res + idx[None, None, :] * res * res).reshape(-1)
        emb[offsets[l] + lin, 0] = sd.reshape(-1).astype(np.float32)
    nd = len(dense_levels)
    G, E, Hd = 12, 12, 64
    in_dim = num_levels * 2
    sdf = _mlp(rng, [in_dim, Hd, Hd, 1 + G + 2], bias_scale=0.05)
    W1, b1 = sdf[0]; W2, b2 = sdf[1]; W3, b3 = sdf[2]
    # units 0/1 of each hidden layer carry +u / -u with u = mean over dense levels of channel 0
    W1[0:2] = 0; b1[0:2] = 0
    for l in dense_levels:
        W1[0, 2 * l] = 1.0 / nd
        W1[1, 2 * l] = -1.0 / nd
    W2[0:2] = 0; b2[0:2] = 0
    W2[0, 0] = 1.0; W2[1, 1] = 1.0
    W2[2:, 0:2] = 0                       # keep the other units independent of the SDF carrier
    W3[0] = 0; b3[0] = 0
    W3[0, 0] = 1.0; W3[0, 1] = -1.0
    W3[1:, 0:2] = 0
    P = 2 ** ide_degree - 1 + ide_degree
    env = _mlp(rng, [2 * P, hidden_dim_env, hidden_dim_env, hidden_dim_env, E])
    diffuse = _mlp(rng, [G + E, 32, 3])
    color = _mlp(rng, [G + 3 + E + 1, 64, 64, 3])
    color[-1] = (color[-1][0], color[-1][1] - np.float32(np.log(3)))        # network.py:363-364
    renv = _mlp(rng, [4, 64, 64, 64, E]) if with_renv else None
    tt = lambda layers: None if layers is None else [(torch.from_numpy(W.copy()), torch.from_numpy(b.copy())) for W, b in layers]
    fp = FieldParams(embeddings=torch.from_numpy(emb), offsets=torch.from_numpy(offsets), per_level_scale=pls, base_resolution=16,
                     bound=bound, sdf=tt(sdf), env=tt(env), diffuse=tt(diffuse), color=tt(color), renv=tt(renv), geo_feat_dim=G,
                     ide_degree=ide_degree, beta=beta, precision=precision)
    return fp.to(device) if str(device) != "cpu" else fp


def make_bitfield(scene_scale: float = 0.65, bound: float = 1.0, grid_size: int = 128, margin_cells: float = 2.5) -> np.ndarray:
    """Occupancy bit field (cascade 1) in the reference's layout: bit (morton3D(i,j,k) % 8) of byte
    (morton3D(i,j,k) / 8) is set when the analytic SDF at the cell centre is below a margin -- what
    update_extra_state + packbits converge to for an SDF field (interior and a thin shell occupied)."""
    H = grid_size
    c = ((np.arange(H) + 0.5) / H * 2 - 1) * bound
    gx, gy, gz = np.meshgrid(c, c, c, indexing="ij")
    sd = analytic_sdf(np.stack([gx, gy, gz], -1), scene_scale)
    occ = sd < margin_cells * (2 * bound / H)

    def spread(v):
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v

    i = np.arange(H, dtype=np.uint64)
    m = (spread(i)[:, None, None] | (spread(i)[None, :, None] << 1) | (spread(i)[None, None, :] << 2)).astype(np.int64)
    flat = np.zeros(H ** 3, np.uint8)
    flat[m.reshape(-1)] = occ.reshape(-1)
    return np.packbits(flat.reshape(-1, 8), axis=-1, bitorder="little").reshape(-1)


# ---------------------------------------------------------------------------------------------
# BASELINE config 4: NeuS-style geometry without a hash grid (materials.ini + use_neus_sdf, encoding_pos=frequency,
# multires 6, geometric_init, 8 x 256 layers, skip_layers [4])
# ---------------------------------------------------------------------------------------------

def make_neus_weights(seed: int = 0, *, hidden_dim: int = 256, num_layers: int = 8, skip_layers=(4,), multires: int = 6,
                      geo_feat_dim: int = 12, radius: float = 0.5):
    """Effective (weight norm folded) weights of the geometry network in the state geometric_init leaves it in
    (nerf/network.py:195-216: the SDF of a sphere of radius `geo_init_bias`), with the feature rows of the head seeded random so that
    geo_feat / roughness / blend vary over space.  Returns [(W [out, in], b [out])] as float32 numpy arrays."""
    rng = np.random.default_rng(seed)
    in_dim = 3 + 2 * 3 * multires
    layers = []
    for l in range(num_layers):
        k = in_dim if l == 0 else hidden_dim
        if l == num_layers - 1:
            n = 1 + geo_feat_dim + 2
        elif l + 1 in skip_layers:
            n = hidden_dim - in_dim
        else:
            n = hidden_dim
        if l == num_layers - 1:
            W = rng.normal(np.sqrt(np.pi) / np.sqrt(k), 1e-4, size=(n, k))
            b = np.full(n, -radius)
            W[1:] = rng.normal(0.0, 1.0 / np.sqrt(k), size=(n - 1, k))
            b[1:] = rng.normal(0.0, 0.3, size=n - 1)
        elif l == 0:
            W = np.zeros((n, k))
            W[:, :3] = rng.normal(0.0, np.sqrt(2) / np.sqrt(n), size=(n, 3))
            b = np.zeros(n)
        elif l in skip_layers:
            W = rng.normal(0.0, np.sqrt(2) / np.sqrt(n), size=(n, k))
            W[:, -(in_dim - 3):] = 0.0
            b = np.zeros(n)
        else:
            W = rng.normal(0.0, np.sqrt(2) / np.sqrt(n), size=(n, k))
            b = np.zeros(n)
        layers.append((W.astype(np.float32), b.astype(np.float32)))
    return layers


def make_sphere_bitfield(radius: float = 0.5, bound: float = 1.0, grid_size: int = 128, margin_cells: float = 2.5) -> np.ndarray:
    """Occupancy bit field of the geometric-init sphere (interior and a thin shell), reference layout (see make_bitfield)."""
    H = grid_size
    c = ((np.arange(H) + 0.5) / H * 2 - 1) * bound
    gx, gy, gz = np.meshgrid(c, c, c, indexing="ij")
    occ = (np.sqrt(gx * gx + gy * gy + gz * gz) - radius) < margin_cells * (2 * bound / H)

    def spread(v):
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v

    i = np.arange(H, dtype=np.uint64)
    m = (spread(i)[:, None, None] | (spread(i)[None, :, None] << 1) | (spread(i)[None, None, :] << 2)).astype(np.int64)
    flat = np.zeros(H ** 3, np.uint8)
    flat[m.reshape(-1)] = occ.reshape(-1)
    return np.packbits(flat.reshape(-1, 8), axis=-1, bitorder="little").reshape(-1)


def make_neus_field(seed: int = 0, *, hidden_dim_env: int = 256, ide_degree: int = 5, variance: float = 0.6, radius: float = 0.5,
                    device="cpu", **kw):
    """Config-4 field: geometry network of make_neus_weights + the rendering MLPs of make_synthetic_field(seed)."""
    from .neus_field import NeusField
    base = make_synthetic_field(seed, hidden_dim_env=hidden_dim_env, ide_degree=ide_degree, num_levels=2, log2_hashmap_size=10, desired_resolution=32)
    layers = make_neus_weights(seed, radius=radius, **kw)
    dev = torch.device(device)
    sdf = [(torch.from_numpy(W).to(dev), torch.from_numpy(b).to(dev)) for W, b in layers]
    shading = NeusField.shading_params(base.env, base.diffuse, base.color, base.renv, dev, geo_feat_dim=12, ide_degree=ide_degree)
    return NeusField(sdf=sdf, skip_layers=tuple(kw.get("skip_layers", (4,))), multires=int(kw.get("multires", 6)),
                     variance=torch.tensor([variance], dtype=torch.float32, device=dev), shading=shading, geo_feat_dim=12)


def neus_to_oracle(nf) -> dict:
    """Plain numpy / scalar dict of a NeusField for oracle.neus_oracle (tests / CPU baseline only)."""
    P = nf.shading.to_oracle()
    P.update(sdf=[(W.detach().cpu().numpy(), b.detach().cpu().numpy()) for W, b in nf.sdf], skip_layers=tuple(nf.skip_layers), multires=nf.multires,
             variance=float(nf.variance.reshape(-1)[0]), cos_anneal_ratio=nf.cos_anneal_ratio, geo_feat_dim=nf.geo_feat_dim)
    return P


# ---------------------------------------------------------------------------------------------
# Fitted variant of the synthetic toaster (SURVEY.md 8d recipe): the geometry is LEARNED through this library's operators instead of
# being written into the dense levels, and the occupancy bit field comes from the field itself through update_extra_state.
# ---------------------------------------------------------------------------------------------

def analytic_sdf_torch(xyz: torch.Tensor, scale: float = 0.65) -> torch.Tensor:
    """analytic_sdf on a torch tensor [..., 3] (any device / dtype)."""
    p = xyz / scale
    half = torch.tensor([0.55, 0.35, 0.30], dtype=p.dtype, device=p.device)
    r = 0.08
    q = p.abs() - (half - r)
    box = q.clamp_min(0).norm(dim=-1) + q.max(dim=-1).values.clamp_max(0) - r
    sph = (p - torch.tensor([0.3, 0.25, 0.0], dtype=p.dtype, device=p.device)).norm(dim=-1) - 0.25
    return torch.minimum(box, sph) * scale


def fit_synthetic_field(seed: int = 0, *, device="cuda", steps: int = 1000, batch: int = 1 << 18, lr_grid: float = 1e-2, lr_mlp: float = 2e-3,
                        density_thresh: float = 10.0, density_updates: int = 17, **field_kw):
    """The recipe of SURVEY.md 8d on the GPU: (1) hash table re-initialised to U(-1e-4, 1e-4) (hashgrid.py:104-106) and sdf_net to
    Xavier weights, then `steps` Adam steps (betas (0.9, 0.99), eps 1e-15: main_nerf.py:150) on an L1 loss between sdf_net(hash_encode(x))[0]
    and the analytic SDF over `batch` points per step (half uniform in the box, half within 0.1 of the surface), through the library's
    hash_encode forward / backward kernels; (2) 1 + 16 update_extra_state calls (renderer.py:264-352) from an empty grid -> bit field.
    Everything else (rendering MLPs, head biases) is the constructed scene's.  Returns (FieldParams on `device`, bitfield uint8 tensor, info)."""
    from . import density, hashencoder
    dev = torch.device(device)
    fp = make_synthetic_field(seed, **field_kw).to(dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    emb = ((torch.rand(fp.embeddings.shape, generator=g, device=dev) * 2 - 1) * 1e-4).requires_grad_(True)
    layers = []
    for W, b in fp.sdf:
        a = math.sqrt(6.0 / (W.shape[0] + W.shape[1]))
        layers.append((((torch.rand(W.shape, generator=g, device=dev) * 2 - 1) * a).requires_grad_(True), torch.zeros_like(b).requires_grad_(True)))
    offsets = fp.offsets.to(dev).int()
    opt = torch.optim.Adam([{"params": [emb], "lr": lr_grid}, {"params": [t for Wb in layers for t in Wb], "lr": lr_mlp}], betas=(0.9, 0.99), eps=1e-15)
    bound = float(fp.bound)

    def sdf_of(x):
        h = hashencoder.hash_encode((x + bound) / (2 * bound), emb, offsets, fp.per_level_scale, fp.base_resolution, False)
        for i, (W, b) in enumerate(layers):
            h = torch.nn.functional.linear(h, W, b)
            if i != len(layers) - 1:
                h = torch.relu(h)
        return h[:, 0]

    def draw(n):
        x = (torch.rand(4 * n, 3, generator=g, device=dev) * 2 - 1) * 0.8 * bound
        sd = analytic_sdf_torch(x)
        near = torch.nonzero(sd.abs() < 0.1).flatten()[: n // 2]
        return torch.cat([x[near], x[-(n - near.numel()):]], 0)

    last = 0.0
    for it in range(steps):
        x = draw(batch)
        loss = (sdf_of(x) - analytic_sdf_torch(x)).abs().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if it == steps - 1:
            last = float(loss.detach())
    with torch.no_grad():
        xt = draw(1 << 20)
        err = (sdf_of(xt) - analytic_sdf_torch(xt)).abs()
        info = {"steps": steps, "batch": batch, "train_loss_last": last, "sdf_l1_mean": float(err.mean()), "sdf_l1_p99": float(torch.quantile(err[:1 << 18], 0.99))}
    fitted = FieldParams(**{**{f: getattr(fp, f) for f in fp.__dataclass_fields__ if not f.startswith("_")},
                            "embeddings": emb.detach().contiguous(), "sdf": [(W.detach().contiguous(), b.detach().contiguous()) for W, b in layers]})
    fitted.precision = "tc"
    fitted.pack()
    dg = density.DensityGrid(bound=bound, density_thresh=density_thresh, grid_size=128, device=dev)
    for _ in range(density_updates):
        dg.update_extra_state(fitted, decay=0.95, full_update=True)
    bits = dg.density_bitfield
    occ = int(torch.tensor([bin(v).count("1") for v in range(256)], device=dev)[bits.long()].sum())
    info.update(occupied_cells=occ, occupied_fraction=occ / float(128 ** 3), mean_density=float(dg.mean_density), density_thresh=density_thresh,
                density_updates=density_updates)
    return fitted, bits, info
