"""Drop-in `_backend` objects: the five pybind11 extension modules of the reference, re-exposed with
identical function names, argument order and in-place output conventions, but implemented by the C ABI
of libenvidr_b200.so (include/envidr_b200.h).

Reference modules replaced:
    raymarching._ext._raymarching   raymarching/src/bindings.cpp:5-20
    hashencoder._ext._hashencoder   hashencoder/src/bindings.cpp
    gridencoder._ext._gridencoder   gridencoder/src/bindings.cpp
    shencoder._ext._shencoder       shencoder/src/bindings.cpp
    freqencoder._ext._freqencoder   freqencoder/src/bindings.cpp

`install_into_sys_modules()` registers them under the reference's import names so that the reference's own
Python wrappers (raymarching/raymarching.py:10 `from raymarching._ext import _raymarching as _backend`, ...)
bind to these kernels without any source change.

Error behaviour follows the reference: encoder entry points validate device / contiguity / dtype and raise
RuntimeError (TORCH_CHECK there); unsupported C / D raise RuntimeError; the raymarching entries do not validate
(the reference defines its CHECK macros but never uses them) beyond what the C ABI needs to be memory-safe.
"""
from __future__ import annotations

import sys
import types

import torch

from . import _lib
from ._lib import check, lib, ptr, stream


def _chk_cuda_contig(**tensors):
    for name, t in tensors.items():
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous tensor")


def _chk_float(**tensors):
    for name, t in tensors.items():
        if t is None:
            continue
        if t.dtype not in (torch.float32, torch.float16, torch.float64):
            raise RuntimeError(f"{name} must be a floating tensor")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name}: libenvidr_b200 implements the fp32 path of the shipped configs; got {t.dtype}")


def _chk_int(**tensors):
    for name, t in tensors.items():
        if t.dtype != torch.int32:
            raise RuntimeError(f"{name} must be an int tensor")


class _Raymarching:
    """raymarching/src/raymarching.h:7-18"""

    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
        check(lib().envidr_near_far_from_aabb(ptr(rays_o), ptr(rays_d), ptr(aabb), N, min_near, ptr(nears), ptr(fars), stream()),
              "near_far_from_aabb")

    @staticmethod
    def sph_from_ray(rays_o, rays_d, radius, N, coords):
        check(lib().envidr_sph_from_ray(ptr(rays_o), ptr(rays_d), radius, N, ptr(coords), stream()), "sph_from_ray")

    @staticmethod
    def morton3D(coords, N, indices):
        check(lib().envidr_morton3D(ptr(coords), N, ptr(indices), stream()), "morton3D")

    @staticmethod
    def morton3D_invert(indices, N, coords):
        check(lib().envidr_morton3D_invert(ptr(indices), N, ptr(coords), stream()), "morton3D_invert")

    @staticmethod
    def packbits(grid, N, density_thresh, bitfield):
        check(lib().envidr_packbits(ptr(grid), N, density_thresh, ptr(bitfield), stream()), "packbits")

    @staticmethod
    def get_scatter_idx(rays, N, idx_map):
        check(lib().envidr_get_scatter_idx(ptr(rays), N, int(idx_map.numel()), ptr(idx_map), stream()), "get_scatter_idx")

    @staticmethod
    def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, early_stop_steps, N, C, H, M, nears, fars,
                         xyzs, dirs, deltas, rays, counter, noises):
        check(lib().envidr_march_rays_train(ptr(rays_o), ptr(rays_d), ptr(grid), bound, dt_gamma, max_steps, early_stop_steps,
                                            N, C, H, M, ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays),
                                            ptr(counter), ptr(noises), stream()), "march_rays_train")

    @staticmethod
    def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, accum_deltas, input_alpha,
                                     weights_sum, depth, image, weights):
        w = ptr(weights) if weights is not None and weights.numel() > 0 else None      # raymarching.cu:707
        check(lib().envidr_composite_rays_train_forward(ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), M, N, T_thresh,
                                                        accum_deltas, input_alpha, ptr(weights_sum), ptr(depth), ptr(image), w,
                                                        stream()), "composite_rays_train_forward")

    @staticmethod
    def composite_rays_train_backward(grad_weights_sum, grad_image, grad_depth, sigmas, rgbs, deltas, rays, weights_sum, image,
                                      depth, M, N, T_thresh, grad_sigmas, grad_rgbs, accum_deltas, input_alpha):
        check(lib().envidr_composite_rays_train_backward(ptr(grad_weights_sum), ptr(grad_image), ptr(grad_depth), ptr(sigmas),
                                                         ptr(rgbs), ptr(deltas), ptr(rays), ptr(weights_sum), ptr(image), ptr(depth),
                                                         M, N, T_thresh, ptr(grad_sigmas), ptr(grad_rgbs), accum_deltas, input_alpha,
                                                         stream()), "composite_rays_train_backward")

    @staticmethod
    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars,
                   xyzs, dirs, deltas, noises):
        check(lib().envidr_march_rays(n_alive, n_step, ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), bound, dt_gamma,
                                      max_steps, C, H, ptr(grid), ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas),
                                      ptr(noises), stream()), "march_rays")

    @staticmethod
    def composite_rays(n_alive, n_step, T_thresh, accum_deltas, input_alpha, rays_alive, rays_t, sigmas, rgbs, deltas,
                       weights_sum, depth, image):
        check(lib().envidr_composite_rays(n_alive, n_step, T_thresh, accum_deltas, input_alpha, ptr(rays_alive), ptr(rays_t),
                                          ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(weights_sum), ptr(depth), ptr(image), stream()),
              "composite_rays")


class _Hashencoder:
    """hashencoder/src/hashencoder.h:13-15"""

    @staticmethod
    def hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx):
        _chk_cuda_contig(inputs=inputs, embeddings=embeddings, offsets=offsets, outputs=outputs, dy_dx=dy_dx)
        _chk_float(inputs=inputs, embeddings=embeddings, outputs=outputs, dy_dx=dy_dx)
        _chk_int(offsets=offsets)
        check(lib().envidr_hash_encode_forward(ptr(inputs), ptr(embeddings), ptr(offsets), ptr(outputs), B, D, C, L, float(S), H,
                                               int(bool(calc_grad_inputs)), ptr(dy_dx), stream()), "hash_encode_forward")

    @staticmethod
    def hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx, grad_inputs):
        _chk_cuda_contig(grad=grad, inputs=inputs, embeddings=embeddings, offsets=offsets, grad_embeddings=grad_embeddings, dy_dx=dy_dx,
                         grad_inputs=grad_inputs)
        _chk_float(grad=grad, inputs=inputs, embeddings=embeddings, grad_embeddings=grad_embeddings, dy_dx=dy_dx, grad_inputs=grad_inputs)
        _chk_int(offsets=offsets)
        check(lib().envidr_hash_encode_backward(ptr(grad), ptr(inputs), ptr(embeddings), ptr(offsets), ptr(grad_embeddings), B, D, C, L,
                                                float(S), H, int(bool(calc_grad_inputs)), ptr(dy_dx), ptr(grad_inputs), stream()),
              "hash_encode_backward")

    @staticmethod
    def hash_encode_second_backward(grad, inputs, embeddings, offsets, B, D, C, L, S, H, calc_grad_inputs, dy_dx, grad_grad_inputs,
                                    grad_grad, grad2_embeddings):
        _chk_cuda_contig(grad=grad, inputs=inputs, embeddings=embeddings, offsets=offsets, dy_dx=dy_dx, grad_grad_inputs=grad_grad_inputs,
                         grad_grad=grad_grad, grad2_embeddings=grad2_embeddings)
        _chk_float(grad=grad, inputs=inputs, embeddings=embeddings, dy_dx=dy_dx, grad_grad_inputs=grad_grad_inputs, grad_grad=grad_grad,
                   grad2_embeddings=grad2_embeddings)
        _chk_int(offsets=offsets)
        check(lib().envidr_hash_encode_second_backward(ptr(grad), ptr(inputs), ptr(embeddings), ptr(offsets), B, D, C, L, float(S), H,
                                                       int(bool(calc_grad_inputs)), ptr(dy_dx), ptr(grad_grad_inputs), ptr(grad_grad),
                                                       ptr(grad2_embeddings), stream()), "hash_encode_second_backward")


class _Gridencoder:
    """gridencoder/src/gridencoder.h:12-13"""

    @staticmethod
    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners):
        _chk_cuda_contig(inputs=inputs, embeddings=embeddings, offsets=offsets, outputs=outputs)
        _chk_float(inputs=inputs, embeddings=embeddings, outputs=outputs, dy_dx=dy_dx)
        _chk_int(offsets=offsets)
        check(lib().envidr_grid_encode_forward(ptr(inputs), ptr(embeddings), ptr(offsets), ptr(outputs), B, D, C, L, float(S), H,
                                               ptr(dy_dx), gridtype, int(bool(align_corners)), stream()), "grid_encode_forward")

    @staticmethod
    def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype,
                             align_corners):
        _chk_cuda_contig(grad=grad, inputs=inputs, embeddings=embeddings, offsets=offsets, grad_embeddings=grad_embeddings)
        _chk_float(grad=grad, inputs=inputs, embeddings=embeddings, grad_embeddings=grad_embeddings, dy_dx=dy_dx, grad_inputs=grad_inputs)
        _chk_int(offsets=offsets)
        check(lib().envidr_grid_encode_backward(ptr(grad), ptr(inputs), ptr(embeddings), ptr(offsets), ptr(grad_embeddings), B, D, C, L,
                                                float(S), H, ptr(dy_dx), ptr(grad_inputs), gridtype, int(bool(align_corners)), stream()),
              "grid_encode_backward")


class _Shencoder:
    """shencoder/src/shencoder.h:9-10"""

    @staticmethod
    def sh_encode_forward(inputs, outputs, B, D, degree, dy_dx):
        _chk_cuda_contig(inputs=inputs, outputs=outputs, dy_dx=dy_dx)
        _chk_float(inputs=inputs, outputs=outputs, dy_dx=dy_dx)
        check(lib().envidr_sh_encode_forward(ptr(inputs), ptr(outputs), B, D, degree, ptr(dy_dx), stream()), "sh_encode_forward")

    @staticmethod
    def sh_encode_backward(grad, inputs, B, D, degree, dy_dx, grad_inputs):
        _chk_cuda_contig(grad=grad, inputs=inputs, dy_dx=dy_dx, grad_inputs=grad_inputs)
        _chk_float(grad=grad, inputs=inputs, dy_dx=dy_dx, grad_inputs=grad_inputs)
        check(lib().envidr_sh_encode_backward(ptr(grad), ptr(inputs), B, D, degree, ptr(dy_dx), ptr(grad_inputs), stream()),
              "sh_encode_backward")


class _Freqencoder:
    """freqencoder/src/freqencoder.h:7-10"""

    @staticmethod
    def freq_encode_forward(inputs, B, D, deg, C, outputs):
        _chk_cuda_contig(inputs=inputs, outputs=outputs)
        _chk_float(inputs=inputs, outputs=outputs)
        check(lib().envidr_freq_encode_forward(ptr(inputs), B, D, deg, C, ptr(outputs), stream()), "freq_encode_forward")

    @staticmethod
    def freq_encode_backward(grad, outputs, B, D, deg, C, grad_inputs):
        _chk_cuda_contig(grad=grad, outputs=outputs, grad_inputs=grad_inputs)
        _chk_float(grad=grad, outputs=outputs, grad_inputs=grad_inputs)
        check(lib().envidr_freq_encode_backward(ptr(grad), ptr(outputs), B, D, deg, C, ptr(grad_inputs), stream()), "freq_encode_backward")


_raymarching = _Raymarching()
_hashencoder = _Hashencoder()
_gridencoder = _Gridencoder()
_shencoder = _Shencoder()
_freqencoder = _Freqencoder()


def install_into_sys_modules():
    """Make `from <pkg>._ext import _<pkg> as _backend` in the reference's wrappers resolve to this library."""
    table = {"raymarching": _raymarching, "hashencoder": _hashencoder, "gridencoder": _gridencoder, "shencoder": _shencoder,
             "freqencoder": _freqencoder}
    for pkg, backend in table.items():
        ext = types.ModuleType(f"{pkg}._ext")
        setattr(ext, f"_{pkg}", backend)
        sys.modules[f"{pkg}._ext"] = ext
        sys.modules[f"{pkg}._ext._{pkg}"] = backend
    # wrappers that were imported before install(): `_backend` is a module global looked up at call time
    for pkg, modname in (("raymarching", "raymarching.raymarching"), ("hashencoder", "hashencoder.hashgrid"), ("gridencoder", "gridencoder.grid"),
                         ("shencoder", "shencoder.sphere_harmonics"), ("freqencoder", "freqencoder.freq")):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "_backend"):
            mod._backend = table[pkg]
