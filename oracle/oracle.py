"""CPU oracle for the ENVIDR volumetric-render hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module (tests/golden/make_golden.py injects its operators into the REFERENCE's
own Python to produce the golden vectors; oracle/ref_cuda.py -- the reference's kernels around
the reference's host logic -- is additionally timed as a baseline by bench.py's gpu_reference /
density_update keys and the profiles/*_bench.py scripts).  The product (envidr_b200/) never
imports anything from here; it fails loudly when its CUDA library is missing instead of
falling back (tests/test_cabi.py::test_product_does_not_import_oracle).

Two layers:
  * ctypes bindings to oracle/envidr_oracle.c (march / composite / grid encoders / freq / SH),
    each restating one reference CUDA kernel (file:line cited in the C source);
  * a torch-CPU restatement of the PyTorch part of the path: IDE
    (ide_encoder/ide_encoder.py:5-130), the SDF / env / diffuse / colour MLPs and their glue
    (nerf/network.py:26-44, 381-698, nerf/renderer.py:20-39, 147-198) and the inference
    loop (nerf/render_func/cuda_ray.py:238-359).

Parity status: the reference ships no tests for this path.  The C layer is pinned on the GPU
box against the reference's own kernels rebuilt for sm_100a (oracle/_ref, tests/test_gpu_ops.py);
the torch layer is pinned against tests/golden/*.npz, produced by importing the reference's
Python modules in the build container (tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "envidr_oracle.c")
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

c_f = ctypes.POINTER(ctypes.c_float)
c_i = ctypes.POINTER(ctypes.c_int32)
c_u8 = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/_build/liboracle.so (seconds)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               _SRC, "-o", _SO, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_f)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_i)


def _u8(a):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(c_u8)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


u32 = ctypes.c_uint32
cf = ctypes.c_float

# --------------------------------------------------------------------------------------
# raymarching.* restatements
# --------------------------------------------------------------------------------------

def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _c32(rays_o).reshape(-1, 3), _c32(rays_d).reshape(-1, 3), _c32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_f(rays_o), _f(rays_d), _f(aabb), u32(N), cf(min_near), _f(nears), _f(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _c32(rays_o).reshape(-1, 3), _c32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().orc_sph_from_ray(_f(rays_o), _f(rays_d), cf(radius), u32(N), _f(coords))
    return coords


def morton3D(coords):
    coords = np.ascontiguousarray(coords, np.int32)
    out = np.empty(coords.shape[0], np.int32)
    lib().orc_morton3D(_i(coords), u32(coords.shape[0]), _i(out))
    return out


def morton3D_invert(indices):
    indices = np.ascontiguousarray(indices, np.int32)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_i(indices), u32(indices.shape[0]), _i(out))
    return out


def packbits(grid, thresh):
    grid = _c32(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().orc_packbits(_f(grid), u32(N), cf(thresh), _u8(out))
    return out


def get_scatter_idx(rays, M):
    rays = np.ascontiguousarray(rays, np.int32)
    out = np.zeros(M, np.int32)
    lib().orc_get_scatter_idx(_i(rays), u32(rays.shape[0]), _i(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, M, noises=None,
                     dt_gamma=0.0, max_steps=1024, early_stop_steps=-1):
    """Deterministic (ray-order) version of raymarching.march_rays_train; returns
    xyzs[M,3], dirs[M,3], deltas[M,2], rays[N,3], counter[2]."""
    rays_o, rays_d = _c32(rays_o).reshape(-1, 3), _c32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    noises = np.zeros(N, np.float32) if noises is None else _c32(noises)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays, counter = np.zeros((N, 3), np.int32), np.zeros(2, np.int32)
    es = max_steps if early_stop_steps <= 0 else early_stop_steps
    lib().orc_march_rays_train(_f(rays_o), _f(rays_d), _u8(bitfield), cf(bound), cf(dt_gamma), u32(max_steps), u32(es),
                               u32(N), u32(C), u32(H), u32(M), _f(_c32(nears)), _f(_c32(fars)),
                               _f(xyzs), _f(dirs), _f(deltas), _i(rays), _i(counter), _f(noises))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, N_out, T_thresh=1e-4, ret_weights=True,
                                 input_alpha=False, accum_deltas=True):
    sigmas, rgbs, deltas = _c32(sigmas), _c32(rgbs), _c32(deltas)
    rays = np.ascontiguousarray(rays, np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.zeros(N_out, np.float32), np.zeros(N_out, np.float32), np.zeros((N_out, 3), np.float32)
    weights = np.zeros(M, np.float32) if ret_weights else None
    lib().orc_composite_rays_train_forward(_f(sigmas), _f(rgbs), _f(deltas), _i(rays), u32(M), u32(N), cf(T_thresh),
                                           u32(int(accum_deltas)), u32(int(input_alpha)), _f(ws), _f(depth), _f(image),
                                           _f(weights) if ret_weights else None)
    return ws, depth, image, weights


def composite_rays_train_backward(grad_ws, grad_image, grad_depth, sigmas, rgbs, deltas, rays, ws, image, depth,
                                  T_thresh=1e-4, input_alpha=False, accum_deltas=True):
    sigmas, rgbs, deltas = _c32(sigmas), _c32(rgbs), _c32(deltas)
    rays = np.ascontiguousarray(rays, np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gr = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().orc_composite_rays_train_backward(_f(_c32(grad_ws)), _f(_c32(grad_image)), _f(_c32(grad_depth)), _f(sigmas), _f(rgbs),
                                            _f(deltas), _i(rays), _f(_c32(ws)), _f(_c32(image)), _f(_c32(depth)),
                                            u32(M), u32(N), cf(T_thresh), _f(gs), _f(gr),
                                            u32(int(accum_deltas)), u32(int(input_alpha)))
    return gs, gr


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars,
               align=-1, noises=None, dt_gamma=0.0, max_steps=1024):
    """raymarching.march_rays (raymarching.py:316-367): zero-filled fixed slots; also returns per-slot counts."""
    rays_o, rays_d = _c32(rays_o).reshape(-1, 3), _c32(rays_d).reshape(-1, 3)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    noises = np.zeros(n_alive, np.float32) if noises is None else _c32(noises)
    counts = np.zeros(n_alive, np.int32)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _i(np.ascontiguousarray(rays_alive, np.int32)), _f(_c32(rays_t)),
                         _f(rays_o), _f(rays_d), cf(bound), cf(dt_gamma), u32(max_steps), u32(C), u32(H), _u8(bitfield),
                         _f(_c32(nears)), _f(_c32(fars)), _f(xyzs), _f(dirs), _f(deltas), _f(noises), _i(counts))
    return xyzs, dirs, deltas, counts


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image,
                   T_thresh=1e-4, input_alpha=False, accum_deltas=True):
    """In place on rays_alive / rays_t / weights_sum / depth / image (all C-contiguous numpy arrays)."""
    lib().orc_composite_rays(u32(n_alive), u32(n_step), cf(T_thresh), u32(int(accum_deltas)), u32(int(input_alpha)),
                             _i(rays_alive), _f(rays_t), _f(_c32(sigmas)), _f(_c32(rgbs)), _f(_c32(deltas)),
                             _f(weights_sum), _f(depth), _f(image))


# --------------------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------------------

def hash_offsets(input_dim=3, num_levels=16, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048,
                 per_level_scale=2.0):
    """Level table of HashEncoder (hashencoder/hashgrid.py:110-146): float64 numpy geometry, no padding."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        offsets.append(offset)
        offset += min(2 ** log2_hashmap_size, res ** input_dim)
    offsets.append(offset)
    return np.array(offsets, np.int32), float(per_level_scale)


def grid_offsets(input_dim=3, num_levels=16, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048,
                 per_level_scale=2.0, align_corners=False):
    """Level table of GridEncoder (gridencoder/grid.py:92-121): (res+1)^D cells, padded to a multiple of 8."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(2 ** log2_hashmap_size, (res if align_corners else res + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, np.int32), float(per_level_scale)


def set_level_scales(scales=None):
    """Install (or clear, with None) device-measured per-level scales; see orc_set_level_scales in the C source."""
    if scales is None:
        lib().orc_set_level_scales(None, 0)
    else:
        a = _c32(scales)
        lib().orc_set_level_scales(_f(a), ctypes.c_int(a.shape[0]))


def hash_encode_forward(inputs, emb, offsets, per_level_scale, H, calc_grad_inputs=False):
    """inputs in [0,1]; returns outputs [L,B,C] and dy_dx [B, L*D*C] (or None) like hash_encode_forward."""
    inputs, emb = _c32(inputs), _c32(emb)
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    out = np.empty((L, B, C), np.float32)
    dy_dx = np.empty((B, L * D * C), np.float32) if calc_grad_inputs else None
    rc = lib().orc_hash_encode_forward(_f(inputs), _f(emb), _i(offsets), _f(out), u32(B), u32(D), u32(C), u32(L), cf(S), u32(H),
                                       ctypes.c_int(int(calc_grad_inputs)), _f(dy_dx) if calc_grad_inputs else None)
    assert rc == 0
    return out, dy_dx


def hash_encode_backward(grad, inputs, emb, offsets, per_level_scale, H, dy_dx=None):
    grad, inputs, emb = _c32(grad), _c32(inputs), _c32(emb)
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    g_emb = np.zeros_like(emb)
    g_in = np.zeros_like(inputs)
    rc = lib().orc_hash_encode_backward(_f(grad), _f(inputs), _f(emb), _i(offsets), _f(g_emb), u32(B), u32(D), u32(C), u32(L),
                                        cf(S), u32(H), ctypes.c_int(int(dy_dx is not None)),
                                        _f(_c32(dy_dx)) if dy_dx is not None else None, _f(g_in))
    assert rc == 0
    return g_emb, g_in


def hash_encode_second_backward(grad, inputs, emb, offsets, per_level_scale, H, dy_dx, grad_grad_inputs):
    grad, inputs, emb = _c32(grad), _c32(inputs), _c32(emb)
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    gg = np.zeros_like(grad)
    g2 = np.zeros_like(emb)
    rc = lib().orc_hash_encode_second_backward(_f(grad), _f(inputs), _f(emb), _i(offsets), u32(B), u32(D), u32(C), u32(L),
                                               cf(S), u32(H), ctypes.c_int(1), _f(_c32(dy_dx)), _f(_c32(grad_grad_inputs)),
                                               _f(gg), _f(g2))
    assert rc == 0
    return gg, g2


def grid_encode_forward(inputs, emb, offsets, per_level_scale, H, calc_grad_inputs=False, gridtype=0, align_corners=False):
    inputs, emb = _c32(inputs), _c32(emb)
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    out = np.empty((L, B, C), np.float32)
    dy_dx = np.empty((B, L * D * C), np.float32) if calc_grad_inputs else None
    rc = lib().orc_grid_encode_forward(_f(inputs), _f(emb), _i(offsets), _f(out), u32(B), u32(D), u32(C), u32(L), cf(S), u32(H),
                                       _f(dy_dx) if calc_grad_inputs else None, u32(gridtype), ctypes.c_int(int(align_corners)))
    assert rc == 0
    return out, dy_dx


def grid_encode_backward(grad, inputs, emb, offsets, per_level_scale, H, dy_dx=None, gridtype=0, align_corners=False):
    grad, inputs, emb = _c32(grad), _c32(inputs), _c32(emb)
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    g_emb, g_in = np.zeros_like(emb), np.zeros_like(inputs)
    rc = lib().orc_grid_encode_backward(_f(grad), _f(inputs), _f(emb), _i(offsets), _f(g_emb), u32(B), u32(D), u32(C), u32(L),
                                        cf(S), u32(H), _f(_c32(dy_dx)) if dy_dx is not None else None, _f(g_in),
                                        u32(gridtype), ctypes.c_int(int(align_corners)))
    assert rc == 0
    return g_emb, (g_in if dy_dx is not None else None)


def freq_encode_forward(inputs, degree):
    inputs = _c32(inputs)
    B, D = inputs.shape
    C = D + 2 * D * degree
    out = np.empty((B, C), np.float32)
    lib().orc_freq_encode_forward(_f(inputs), u32(B), u32(D), u32(degree), u32(C), _f(out))
    return out


def freq_encode_backward(grad, outputs, D, degree):
    grad, outputs = _c32(grad), _c32(outputs)
    B, C = outputs.shape
    g = np.zeros((B, D), np.float32)
    lib().orc_freq_encode_backward(_f(grad), _f(outputs), u32(B), u32(D), u32(degree), u32(C), _f(g))
    return g


def sh_encode_forward(inputs, degree, calc_grad_inputs=False):
    inputs = _c32(inputs)
    B, D = inputs.shape
    out = np.empty((B, degree * degree), np.float32)
    dy_dx = np.empty((B, D * degree * degree), np.float32) if calc_grad_inputs else None
    lib().orc_sh_encode_forward(_f(inputs), _f(out), u32(B), u32(D), u32(degree), _f(dy_dx) if calc_grad_inputs else None)
    return out, dy_dx


def sh_encode_backward(grad, inputs, degree, dy_dx):
    grad, inputs = _c32(grad), _c32(inputs)
    B, D = inputs.shape
    g = np.zeros((B, D), np.float32)
    lib().orc_sh_encode_backward(_f(grad), _f(inputs), u32(B), u32(D), u32(degree), _f(_c32(dy_dx)), _f(g))
    return g


# --------------------------------------------------------------------------------------
# IDE (ide_encoder/ide_encoder.py:5-130), restated
# --------------------------------------------------------------------------------------

def _gen_binom(a: float, k: int) -> float:
    return float(np.prod(a - np.arange(k))) / math.factorial(k)


def _sph_harm_coeff(l: int, m: int, k: int) -> float:
    """Coefficient of z^k in the (l, m) spherical harmonic's z-polynomial (ide_encoder.py:5-42)."""
    legendre = ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m)
                * _gen_binom(0.5 * (l + k + m - 1.0), l))
    return math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * math.pi * math.factorial(l + m))) * legendre


def ide_tables(deg_view: int, exact: bool = False):
    """(ml[2,P] int, mat[l_max+1,P] float32, sigma[P] float32); ide_encoder.py:45-96.
    exact=True keeps the coefficients in float64 (the reference rounds them to fp32 when it registers the buffer, which
    by itself moves the l = 16 band by up to ~6e-4: |coeff| ~ 1e5 x 6e-8)."""
    if deg_view > 5:
        raise ValueError("Only deg_view of at most 5 is numerically stable.")
    ml = [(m, 2 ** i) for i in range(deg_view) for m in range(2 ** i + 1)]
    l_max = 2 ** (deg_view - 1)
    mat = np.zeros((l_max + 1, len(ml)))
    for i, (m, l) in enumerate(ml):
        for k in range(l - m + 1):
            mat[k, i] = _sph_harm_coeff(l, m, k)
    ml_a = np.array(ml).T
    sigma = 0.5 * ml_a[1] * (ml_a[1] + 1)
    if exact:
        return ml_a.astype(np.int32), mat, sigma.astype(np.float64)
    return ml_a.astype(np.int32), mat.astype(np.float32), sigma.astype(np.float32)


def ide_encode(xyz: torch.Tensor, kappa_inv, deg_view: int, exact_tables: bool = False) -> torch.Tensor:
    """IntegratedDirEncoder.forward (ide_encoder.py:98-130) with repeated complex products instead of
    complex pow.  dtype follows xyz (float32 or float64); the tables are the fp32-rounded ones of the reference unless
    exact_tables (then, in float64, the result is the mathematically exact encoding)."""
    ml, mat, sigma = ide_tables(deg_view, exact_tables)
    dt = xyz.dtype
    x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
    y = y + ((x == 0) & (y == 0)).to(dt)                      # "avoid 0 + 0j exponentiation"
    l_max = mat.shape[0] - 1
    vmz = torch.cat([z ** k if k else torch.ones_like(z) for k in range(l_max + 1)], dim=-1)
    re, im = [torch.ones_like(x)], [torch.zeros_like(x)]
    for _ in range(int(ml[0].max())):
        re, im = re + [re[-1] * x - im[-1] * y], im + [re[-1] * y + im[-1] * x]
    re = torch.cat([re[m] for m in ml[0]], dim=-1)
    im = torch.cat([im[m] for m in ml[0]], dim=-1)
    zc = vmz @ torch.from_numpy(mat).to(device=xyz.device, dtype=dt)
    if not torch.is_tensor(kappa_inv):
        kappa_inv = torch.tensor(float(kappa_inv), dtype=dt, device=xyz.device)
    att = torch.exp(-torch.from_numpy(sigma).to(device=xyz.device, dtype=dt) * kappa_inv.to(dt))
    return torch.cat([re * zc * att, im * zc * att], dim=-1)


# --------------------------------------------------------------------------------------
# field (SDF + rendering MLPs), restated from nerf/network.py + nerf/renderer.py
# --------------------------------------------------------------------------------------

def _mlp(x: torch.Tensor, layers: Sequence[Tuple[np.ndarray, Optional[np.ndarray]]]) -> torch.Tensor:
    """Linear+ReLU stack, no activation after the last layer (network.py:415-421, 555-558, 592-595, 672-675)."""
    for i, (W, b) in enumerate(layers):
        x = x @ torch.from_numpy(W).to(x.dtype).T
        if b is not None:
            x = x + torch.from_numpy(b).to(x.dtype)
        if i != len(layers) - 1:
            x = torch.relu(x)
    return x


def _unit(x: torch.Tensor, eps: float) -> torch.Tensor:
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)   # F.normalize


def rot_theta3(th: float) -> np.ndarray:
    """Upper-left 3x3 of rot_theta (nerf/utils.py, same matrix as demo.ipynb cell 5)."""
    return np.array([[np.cos(th), 0, -np.sin(th)], [0, 1, 0], [np.sin(th), 0, np.cos(th)]], np.float32)


def field_forward(P: Dict, xyzs: np.ndarray, dirs: np.ndarray, r_images: Optional[np.ndarray] = None,
                  env_rot_radian: Optional[float] = None, dtype=torch.float64,
                  enc_override: Optional[Tuple[np.ndarray, np.ndarray]] = None,
                  ide_dtype=torch.float64) -> Dict[str, np.ndarray]:
    """forward_sigma + get_color_mlp_extra_params + forward_color for the shipped scene configuration
    (hashgrid_diff, ensemble_mlp, unitNorm features, IDE reflected-dir encoding, diffuse_with_env concat,
    wo_viewdir, normal_with_mlp identity, n_dot_viewdir).  P is a plain dict of numpy arrays/scalars
    (envidr_b200.field.FieldParams.to_oracle()).  Returns float32 numpy arrays.
    enc_override = (enc[M,F], d enc / d xyz [M,F,3]) replaces the hash encoder (used to pin this function
    against the reference's network.py driven through a differentiable stand-in encoder).
    ide_dtype: the reference evaluates IDE in fp32 as a power-basis Vandermonde product, which for the
    l = 16 band (deg_view 5) cancels catastrophically (|coeff| ~ 1e5): its fp32 result is ~5e-4 off the
    exact value.  float32 restates the reference's arithmetic; float64 (default) gives the exact encoding, which
    is what the CUDA kernel computes (stable recurrence, see csrc/ide_tables.cuh)."""
    M = xyzs.shape[0]
    bound = float(P["bound"])
    lvl_mask = None
    if enc_override is None:
        x01 = (np.asarray(xyzs, np.float32) + np.float32(bound)) / np.float32(2 * bound)      # hashgrid.py:161
        enc, dy_dx = hash_encode_forward(x01, P["embeddings"], P["offsets"], P["per_level_scale"], P["base_resolution"], True)
        L, C = enc.shape[0], enc.shape[2]
        enc = torch.from_numpy(np.ascontiguousarray(enc.transpose(1, 0, 2).reshape(M, L * C))).to(dtype)
        # dy_dx [M,L,D,C] w.r.t. x01 -> [M, L*C, D] w.r.t. xyz
        jac = torch.from_numpy(dy_dx.reshape(M, L, 3, C)).to(dtype).permute(0, 1, 3, 2).reshape(M, L * C, 3) / (2 * bound)
        if P.get("enabled_levels", -1) > 0:                                                 # network.py:390-393
            lvl_mask = torch.zeros(L, C, dtype=dtype)
            lvl_mask[: P["enabled_levels"]] = 1
            lvl_mask = lvl_mask.reshape(-1)
            enc = enc * lvl_mask
    else:
        enc = torch.from_numpy(np.asarray(enc_override[0])).to(dtype)
        jac = torch.from_numpy(np.asarray(enc_override[1])).to(dtype)
    # sdf_net forward, keeping the ReLU masks for the analytic gradient (network.py:415-421)
    h, acts = enc, []
    sdf_layers = P["sdf"]
    for i, (W, b) in enumerate(sdf_layers):
        h = h @ torch.from_numpy(W).to(dtype).T + torch.from_numpy(b).to(dtype)
        if i != len(sdf_layers) - 1:
            acts.append(h > 0)
            h = torch.relu(h)
    G = int(P["geo_feat_dim"])
    sdf = h[:, 0]
    geo = _unit(h[:, 1:1 + G], 1e-12)                                                       # network.py:431-435
    rough_raw = h[:, 1 + G:2 + G]
    roughness = P["roughness_act_scale"] * torch.nn.functional.softplus(rough_raw + P["roughness_bias"]) * P["roughness_scale"]
    blend = torch.sigmoid(h[:, 2 + G:3 + G]) if h.shape[1] > 2 + G else None
    # normal = d sdf / d xyz (renderer.py:182-198): reverse pass through the ReLU stack and dy_dx
    g = torch.from_numpy(sdf_layers[-1][0][0:1]).to(dtype).expand(M, -1)
    for i in range(len(sdf_layers) - 2, -1, -1):
        g = (g * acts[i].to(dtype)) @ torch.from_numpy(sdf_layers[i][0]).to(dtype)
    if lvl_mask is not None:
        g = g * lvl_mask
    grad_x = torch.einsum("mf,mfd->md", g, jac)
    normals = _unit(grad_x, 1e-10)
    beta = min(max(float(P["beta"]), float(P["beta_min"])), float(P["beta_max"]))         # network.py:39-44
    sigma = (1.0 / beta) * (0.5 + 0.5 * torch.sign(sdf) * torch.expm1(-sdf.abs() / beta))   # network.py:32-37
    sigma = sigma * P.get("density_scale", 1.0)
    d = torch.from_numpy(np.asarray(dirs, np.float32)).to(dtype)
    w_o = -d
    n_dot = (normals * w_o).sum(-1, keepdim=True)
    w_r = 2 * n_dot * normals - w_o                                                         # renderer.py:20-39
    n_env = normals
    if env_rot_radian is not None:                                                          # renderer.py:160-161,171-172
        R = torch.from_numpy(rot_theta3(env_rot_radian)).to(dtype)
        w_r = w_r @ R
        n_env = normals @ R
    deg = int(P["ide_degree"])
    lis = P.get("light_intensity_scale", 1.0)
    w_r_enc = ide_encode(w_r.to(ide_dtype), roughness.to(ide_dtype), deg).to(dtype) * lis
    n_enc = ide_encode(n_env.to(ide_dtype), P["diffuse_kappa_inv"], deg).to(dtype) * lis
    f_n = _unit(_mlp(n_enc, P["env"]), 1e-12)                                               # network.py:527-541
    c_d = torch.sigmoid(_mlp(torch.cat([geo, f_n], -1), P["diffuse"]))                      # :555-572
    f_r = _unit(_mlp(w_r_enc, P["env"]), 1e-12)                                             # :589-607
    hh = torch.cat([geo, normals], -1)
    c_s = torch.sigmoid(_mlp(torch.cat([hh, f_r, n_dot], -1), P["color"]))                  # :664-679
    if r_images is not None and P.get("renv") is not None:                                  # :612-659, 682-690
        ri = torch.from_numpy(np.asarray(r_images, np.float32)).to(dtype)
        mask = roughness.squeeze(-1) < P["indir_roughness_thresh"]
        if ri.shape[-1] == 4:
            vis = ri[:, 3]
            ri = ri[:, :3] * vis[:, None]
            mask = mask & (vis > 0.9)
        rr = torch.sqrt(roughness / P["roughness_scale"] / 0.75)
        if P.get("learn_indir_blend", False):
            bw = 0.98 * blend
        else:
            bw = 0.95 * torch.sigmoid(80 * (rr - 0.18))
        f_e = _unit(_mlp(torch.cat([ri, rr], -1), P["renv"]), 1e-12)
        c_e = torch.sigmoid(_mlp(torch.cat([hh, f_e, n_dot], -1), P["color"]))
        c_s = torch.where(mask[:, None], c_s * bw + c_e * (1 - bw), c_s)
    rgb = (c_d + c_s) * P.get("intensity_scale", 1.0)
    f32 = lambda t: t.to(torch.float32).numpy()
    return dict(sdf=f32(sdf), sigma=f32(sigma), geo_feat=f32(geo), normal=f32(normals), grad_x=f32(grad_x),
                roughness=f32(roughness), rgb=f32(rgb), c_diffuse=f32(c_d), c_specular=f32(c_s),
                n_dot_w_o=f32(n_dot), w_r_enc=f32(w_r_enc), n_env_enc=f32(n_enc),
                blend=None if blend is None else f32(blend))


# --------------------------------------------------------------------------------------
# inference loop (nerf/render_func/cuda_ray.py:238-359), restated on the CPU
# --------------------------------------------------------------------------------------

def render_rays(P: Dict, rays_o, rays_d, bitfield, *, cascade=1, grid_size=128, min_near=0.2, aabb=None,
                dt_gamma=0.0, max_steps=1024, T_thresh=1e-4, bg_color=1.0, env_rot_radian=None,
                r_images=None, geometry_only=False, dtype=torch.float64, stats: Optional[dict] = None):
    """Returns dict(image[N,3], depth[N], weights_sum[N], normal_image[N,3]) as float32 numpy.
    Follows the reference schedule exactly: n_step = max(min(N // n_alive, 8), 1), one march +
    field + composite per iteration, ordered compaction of the alive list."""
    rays_o, rays_d = _c32(rays_o).reshape(-1, 3), _c32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    bound = float(P["bound"])
    if aabb is None:
        aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    n_ws, n_depth, n_img = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    alive = np.arange(N, dtype=np.int32)
    rays_t = nears.copy()
    n_alive_n, n_t = alive.copy(), nears.copy()
    step, n_samples, n_iters = 0, 0, 0
    while step < max_steps:
        n_alive = alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas, counts = march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, bound, bitfield,
                                                cascade, grid_size, nears, fars, align=-1, dt_gamma=dt_gamma,
                                                max_steps=max_steps)
        n_samples += int(counts.sum()); n_iters += 1
        real = deltas[:, 0] > 0
        idx = np.nonzero(real)[0]
        sig = np.zeros(xyzs.shape[0], np.float32)
        rgb = np.zeros((xyzs.shape[0], 3), np.float32)
        nrm = np.zeros((xyzs.shape[0], 3), np.float32)
        if idx.size:
            ri = None
            if r_images is not None:
                ri = np.repeat(np.asarray(r_images)[alive], n_step, axis=0)[idx]
            out = field_forward(P, xyzs[idx], dirs[idx], r_images=ri, env_rot_radian=env_rot_radian, dtype=dtype)
            sig[idx], rgb[idx], nrm[idx] = out["sigma"], out["rgb"], out["normal"]
        alive_n = alive.copy(); t_n = rays_t.copy()
        composite_rays(n_alive, n_step, alive_n, t_n, sig, nrm, deltas, n_ws, n_depth, n_img, T_thresh)
        if not geometry_only:
            composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, deltas, ws, depth, image, T_thresh)
        else:
            alive, rays_t = alive_n, t_n
        alive = alive[alive >= 0]
        step += n_step
    if stats is not None:
        stats.update(samples=n_samples, iterations=n_iters)
    nrm_img = n_img / np.maximum(np.linalg.norm(n_img, axis=-1, keepdims=True), 1e-10)
    if geometry_only:
        return dict(image=None, depth=n_depth, weights_sum=n_ws, normal_image=nrm_img.astype(np.float32))
    image = image + (1 - ws)[:, None] * np.float32(bg_color)
    return dict(image=image.astype(np.float32), depth=depth, weights_sum=ws, normal_image=nrm_img.astype(np.float32))


def render(P: Dict, rays_o, rays_d, bitfield, *, indir_ref=False, indir_max_steps=1024, bg_color=1.0, min_near=0.2,
           max_steps=1024, T_thresh=1e-4, env_rot_radian=None, obj_aabb=None, dtype=torch.float64, stats: Optional[list] = None, **kw):
    """NeRFRenderer.render for the cuda_ray inference path (nerf/renderer.py:364-531): one pass, or the three passes of
    the indirect-reflection scheme (:439-513).  normal_image is post-processed as in :529-530."""
    rays_o, rays_d = _c32(rays_o).reshape(-1, 3), _c32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    common = dict(max_steps=max_steps, T_thresh=T_thresh, env_rot_radian=env_rot_radian, dtype=dtype, **kw)

    def run(o, d, **k):
        st = {}
        r = render_rays(P, o, d, bitfield, stats=st, **{**common, **k})
        if stats is not None:
            stats.append(st)
        return r

    if not indir_ref:
        res = run(rays_o, rays_d, bg_color=bg_color, min_near=min_near)
    else:
        dt = np.float32(2 * math.sqrt(3) / indir_max_steps)
        geo = run(rays_o, rays_d, geometry_only=True, min_near=min_near)
        normals = geo["normal_image"]
        depth = geo["depth"] - dt
        ws = geo["weights_sum"]
        ref_mask = (depth != 0) & (ws > 0.9)
        ray_mask = (depth != 0) & (ws > 0.3)
        ref_o = rays_o + depth[:, None] * rays_d
        w_o = -rays_d
        ref_d = 2 * np.sum(w_o * normals, -1, keepdims=True) * normals - w_o
        if obj_aabb is not None:
            ob = np.asarray(obj_aabb, np.float32)
            ref_mask &= (ref_o > ob[:3]).all(-1) & (ref_o < ob[3:]).all(-1)
        sec = run(ref_o[ref_mask], ref_d[ref_mask], bg_color=0.0, min_near=float(dt * 2), max_steps=indir_max_steps)
        ref_image = np.concatenate([sec["image"], sec["weights_sum"][:, None]], -1)
        r_img = np.zeros((int(ray_mask.sum()), 4), np.float32)
        r_img[ref_mask[ray_mask]] = ref_image
        main = run(rays_o[ray_mask], rays_d[ray_mask], bg_color=0.0, r_images=r_img, min_near=min_near)
        img = np.zeros((N, 3), np.float32); img[ray_mask] = main["image"]
        wsf = np.zeros(N, np.float32); wsf[ray_mask] = main["weights_sum"]
        res = dict(image=(np.zeros((N, 3), np.float32) + np.float32(bg_color)) * (1 - wsf[:, None]) + img, weights_sum=wsf,
                   depth=depth, normal_image=normals)
    w = res["weights_sum"][:, None]
    res["normal_image"] = res["normal_image"] * w + (1 - w)
    return res


# --------------------------------------------------------------------------------------
# occupancy-grid maintenance (nerf/renderer.py:200-352), restated on the CPU  
# --------------------------------------------------------------------------------------

def forward_sigma(P: Dict, xyzs: np.ndarray, dtype=torch.float32, enc_fn=None) -> np.ndarray:
    """NeRFNetwork.density(x)['sigma'] for the shipped scene family (network.py:381-448, 497-522, 32-44): hash encoding ->
    sdf_net -> Laplace density.  fp32 by default (what the reference computes).  enc_fn(xyz tensor) -> [M, F] replaces the hash
    encoder (stand-in encoder of tests/golden/make_golden.py).  Returns sigma [M] float32, NOT yet times density_scale."""
    M = xyzs.shape[0]
    bound = float(P["bound"])
    if enc_fn is None:
        x01 = (np.asarray(xyzs, np.float32) + np.float32(bound)) / np.float32(2 * bound)      # hashgrid.py:161
        enc, _ = hash_encode_forward(x01, P["embeddings"], P["offsets"], P["per_level_scale"], P["base_resolution"], False)
        L, C = enc.shape[0], enc.shape[2]
        enc = torch.from_numpy(np.ascontiguousarray(enc.transpose(1, 0, 2).reshape(M, L * C))).to(dtype)
        if P.get("enabled_levels", -1) > 0:                                                 # network.py:390-393
            m = torch.zeros(L, C, dtype=dtype)
            m[: P["enabled_levels"]] = 1
            enc = enc * m.reshape(-1)
    else:
        enc = enc_fn(torch.from_numpy(np.asarray(xyzs, np.float32))).to(dtype)
    sdf = _mlp(enc, P["sdf"])[:, 0]
    beta = min(max(float(P["beta"]), float(P["beta_min"])), float(P["beta_max"]))         # network.py:39-44
    sigma = (1.0 / beta) * (0.5 + 0.5 * torch.sign(sdf) * torch.expm1(-sdf.abs() / beta))   # network.py:32-37
    return sigma.to(torch.float32).numpy()


CUDA_SCALAR_DIV = False      # True: x / s evaluated as x * fp32(1 / s), torch's CUDA kernel for division by a Python scalar


def _div_scalar(x: torch.Tensor, s: float) -> torch.Tensor:
    """torch `tensor / python_scalar`: an IEEE division on the CPU (what the golden vectors were made with), a multiplication by
    the fp32 reciprocal on CUDA (ATen BinaryDivTrueKernel.cu) -- where the reference actually runs these lines.  GPU parity
    tests set CUDA_SCALAR_DIV."""
    if CUDA_SCALAR_DIV:
        return x * float(np.float32(1.0) / np.float32(s))
    return x / s


def density_cell_positions(coords: np.ndarray, noise: Optional[np.ndarray], bound_c: float, H: int) -> np.ndarray:
    """renderer.py:290-301 / 320-331 in fp32, one rounding per torch op: xyzs = 2 * coords.float() / (H - 1) - 1;
    cas_xyzs = xyzs * (bound - half_grid_size); cas_xyzs += (rand * 2 - 1) * half_grid_size."""
    hgs = bound_c / H
    x = torch.from_numpy(np.asarray(coords)).float()
    x = _div_scalar(2 * x, H - 1) - 1
    x = x * (bound_c - hgs)
    if noise is not None:
        x = x + (torch.from_numpy(np.asarray(noise, np.float32)) * 2 - 1) * hgs
    return x.numpy()


def update_extra_state(P: Dict, density_grid: np.ndarray, *, cascade: int = 1, grid_size: int = 128, decay: float = 0.95,
                       density_thresh: float = 0.01, density_scale: float = 1.0, noise: Optional[np.ndarray] = None,
                       coords: Optional[np.ndarray] = None, enc_fn=None, dtype=torch.float32):
    """NeRFRenderer.update_extra_state (renderer.py:264-352), density-grid half.
    density_grid [C, H^3] fp32 (Morton order), returned updated (copy).
    coords None: full update -- every cell, meshgrid (ij) order as the reference's custom_meshgrid(xs, ys, zs) with S = 128;
    noise [C, H^3, 3] in that order (None: no jitter).  coords [C, n, 3] int: partial update with noise [C, n, 3].
    Returns (density_grid, mean_density, density_thresh_used, bitfield uint8 [C*H^3/8], tmp_grid)."""
    H = grid_size
    grid = np.array(density_grid, np.float32).reshape(cascade, H ** 3).copy()
    tmp = -np.ones_like(grid)
    if coords is None:
        r = np.arange(H, dtype=np.int32)
        cc = np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3)
    for cas in range(cascade):
        bound_c = min(2 ** cas, float(P["bound"]))
        c = cc if coords is None else np.asarray(coords[cas], np.int32)
        idx = morton3D(c).astype(np.int64)
        xyz = density_cell_positions(c, None if noise is None else noise[cas], bound_c, H)
        sig = forward_sigma(P, xyz, dtype=dtype, enc_fn=enc_fn) * np.float32(density_scale)
        tmp[cas, idx] = sig                                                          # duplicates: last writer wins (numpy)
    valid = (grid >= 0) & (tmp >= 0)                                                 # renderer.py:343-345
    grid[valid] = np.maximum(grid[valid] * np.float32(decay), tmp[valid])
    mean = float(torch.mean(torch.from_numpy(grid).clamp(min=0)).item())
    th = min(mean, density_thresh)
    return grid, mean, th, packbits(grid.reshape(-1), th), tmp


def mark_untrained_grid(poses: np.ndarray, intrinsic, *, bound: float = 1.0, cascade: int = 1, grid_size: int = 128,
                        dtype=np.float64):
    """NeRFRenderer.mark_untrained_grid (renderer.py:200-262): per-cell count of the cameras that see the cell, Morton order
    [C, H^3] int32, plus the smallest margin of any of the three comparisons per cell (cells whose margin is at rounding
    level are ambiguous between two correct fp32 evaluations)."""
    H = grid_size
    fx, fy, cx, cy = [float(v) for v in intrinsic]
    r = np.arange(H, dtype=np.int32)
    cc = np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3)
    idx = morton3D(cc).astype(np.int64)
    world = (_div_scalar(2 * torch.from_numpy(cc).float(), H - 1) - 1).numpy()
    count = np.zeros((cascade, H ** 3), np.int32)
    margin = np.full((cascade, H ** 3), np.inf)
    poses = np.asarray(poses, np.float32)
    for cas in range(cascade):
        bound_c = min(2 ** cas, bound)
        hgs = bound_c / H
        w = (world * np.float32(bound_c - hgs)).astype(dtype)
        cnt = np.zeros(H ** 3, np.int32)
        mg = np.full(H ** 3, np.inf)
        for b in range(poses.shape[0]):
            cam = (w - poses[b, :3, 3].astype(dtype)) @ poses[b, :3, :3].astype(dtype)
            lim_x = cx / fx * cam[:, 2] + hgs * 2
            lim_y = cy / fy * cam[:, 2] + hgs * 2
            m = (cam[:, 2] > 0) & (np.abs(cam[:, 0]) < lim_x) & (np.abs(cam[:, 1]) < lim_y)
            cnt += m
            mg = np.minimum(mg, np.minimum(np.abs(cam[:, 2]), np.minimum(np.abs(np.abs(cam[:, 0]) - lim_x), np.abs(np.abs(cam[:, 1]) - lim_y))))
        count[cas, idx] = cnt
        margin[cas, idx] = mg
    return count, margin


# --------------------------------------------------------------------------------------
# optimizer step: torch.optim.Adam as the reference configures it (main_nerf.py:150)
# --------------------------------------------------------------------------------------

def _fma32(a, b, c):
    """fp32 fused multiply-add emulated through fp64 (the product of two fp32 values is exact in fp64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def adam_step(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, beta1: float = 0.9, beta2: float = 0.99, eps: float = 1e-15):
    """One torch.optim.Adam step (torch/optim/adam.py::_single_tensor_adam, no weight decay / amsgrad / maximize; the third-party
    arithmetic of the reference's training loop, nerf/utils.py:1079-1087) on fp32 numpy arrays, with the fused multiply-adds the
    CUDA kernels of torch contract to.  step: 1-based.  Returns (param, exp_avg, exp_avg_sq)."""
    f = np.float32
    p, g, m, v = [np.asarray(a, np.float32) for a in (param, grad, exp_avg, exp_avg_sq)]
    m = _fma32(np.full_like(g, f(1 - beta1)), g - m, m)                                  # exp_avg.lerp_(grad, 1 - beta1)
    v = _fma32(np.full_like(g, f(1 - beta2)), g * g, v * f(beta2))                      # mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    step_size = lr / bias_correction1
    denom = np.sqrt(v) / f(bias_correction2 ** 0.5) + f(eps)
    p = _fma32(np.full_like(g, f(-step_size)), m / denom, p)                            # addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v


# --------------------------------------------------------------------------------------
# NeuS-style opacity (nerf/network.py:46-102), SURVEY.md 8 a-6
# --------------------------------------------------------------------------------------

def neus_alpha(sdf, dirs, dists, gradients, variance: float, cos_anneal_ratio: float = 1.0, grad_alpha=None, dtype=torch.float64):
    """NeuSDensity.forward restated (float64 by default).  Returns alpha [M] float32, and with grad_alpha [M] also the gradients of
    sum(grad_alpha * alpha) w.r.t. (sdf, gradients, variance) by autograd."""
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a), dtype=dtype)
    sdf_t, g_t = t(sdf).clone().requires_grad_(True), t(gradients)
    if g_t is not None:
        g_t = g_t.clone().requires_grad_(True)
    var = torch.tensor(float(variance), dtype=dtype, requires_grad=True)
    d = t(dirs)
    dist = t(dists) if not np.isscalar(dists) else float(dists)
    inv_s = torch.exp(var * 10.0).clip(1e-6, 1e6)
    if g_t is not None:
        true_cos = (d * g_t).sum(-1, keepdim=True)
        iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + torch.relu(-true_cos) * cos_anneal_ratio)
        nxt = sdf_t + iter_cos.squeeze(-1) * dist * 0.5
        prv = sdf_t - iter_cos.squeeze(-1) * dist * 0.5
    else:
        nxt = sdf_t - dist * 0.5
        prv = sdf_t + dist * 0.5
    pc, nc = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
    alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0)
    out = alpha.detach().to(torch.float32).numpy()
    if grad_alpha is None:
        return out
    leaves = [sdf_t, var] + ([g_t] if g_t is not None else [])
    gr = torch.autograd.grad((alpha * t(grad_alpha)).sum(), leaves)
    f32 = lambda x: x.detach().to(torch.float32).numpy()
    return out, f32(gr[0]), (f32(gr[2]) if g_t is not None else None), float(gr[1])
