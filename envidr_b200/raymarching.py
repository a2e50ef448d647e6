"""raymarching.* operator surface (mirrors reference raymarching/raymarching.py), backed by libenvidr_b200.

Same function names, argument order, defaults and return conventions as the reference module so that
nerf/renderer.py and nerf/render_func/cuda_ray.py work on top of it unchanged:
near_far_from_aabb, sph_from_ray, morton3D, morton3D_invert, packbits, get_scatter_idx,
march_rays_train, composite_rays_train, march_rays, composite_rays.
Differences (all behaviour-preserving):
  * march_rays_train assigns sample offsets by an exclusive scan in ray order (deterministic) instead
    of the reference's atomic-order slots (raymarching.cu:434-435); per-ray counts / samples are identical.
  * no torch.cuda.empty_cache() after the first epochs (raymarching.py:242) -- it only serialises the device.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .backend import _raymarching as _backend


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


class _near_far_from_aabb(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """raymarching.py:19-49.  rays_o/d [N,3], aabb [6] -> nears [N], fars [N]."""
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = _f32(rays_o).contiguous().view(-1, 3)
        rays_d = _f32(rays_d).contiguous().view(-1, 3)
        aabb = _f32(aabb).contiguous().to(rays_o.device)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _backend.near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, radius):
        """raymarching.py:52-81.  -> coords [N,2] in [-1,1]."""
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = _f32(rays_o).contiguous().view(-1, 3)
        rays_d = _f32(rays_d).contiguous().view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _backend.sph_from_ray(rays_o, rays_d, radius, N, coords)
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """raymarching.py:84-104.  coords [N,3] int32 in [0,128) -> indices [N] int32."""
        if not coords.is_cuda: coords = coords.cuda()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        _backend.morton3D(coords.int().contiguous(), N, indices)
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        """raymarching.py:106-126."""
        if not indices.is_cuda: indices = indices.cuda()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        _backend.morton3D_invert(indices.int().contiguous(), N, coords)
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    def forward(ctx, grid, thresh, bitfield=None):
        """raymarching.py:129-155.  grid [C, H^3] -> bitfield uint8 [C*H^3/8]."""
        if not grid.is_cuda: grid = grid.cuda()
        grid = _f32(grid).contiguous()
        C, H3 = grid.shape[0], grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        _backend.packbits(grid, N, thresh, bitfield)
        return bitfield


packbits = _packbits.apply


class _get_scatter_idx(Function):
    @staticmethod
    def forward(ctx, rays, source):
        """raymarching.py:157-164."""
        N = rays.shape[0]
        _backend.get_scatter_idx(rays, N, source)
        return source


get_scatter_idx = _get_scatter_idx.apply


class _march_rays_train(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False,
                align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024, early_stop_steps=-1):
        """raymarching.py:170-246.  Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3]."""
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        if not density_bitfield.is_cuda: density_bitfield = density_bitfield.cuda()
        rays_o = _f32(rays_o).contiguous().view(-1, 3)
        rays_d = _f32(rays_d).contiguous().view(-1, 3)
        density_bitfield = density_bitfield.contiguous()
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        dev = rays_o.device
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=rays_o.dtype, device=dev) if perturb else torch.zeros(N, dtype=rays_o.dtype, device=dev)
        early_stop_steps = max_steps if early_stop_steps <= 0 else early_stop_steps
        _backend.march_rays_train(rays_o, rays_d, density_bitfield, bound, dt_gamma, max_steps, early_stop_steps, N, C, H, M,
                                  _f32(nears).contiguous(), _f32(fars).contiguous(), xyzs, dirs, deltas, rays, step_counter, noises)
        if force_all_rays or mean_count <= 0:
            m = step_counter[0].item()          # D2H copy, as in the reference (first epochs only)
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4, ret_weights=False, input_alpha=False, accum_deltas=True):
        """raymarching.py:249-310.  Returns weights_sum [N], depth [N], image [N,3], weights [M] (or empty)."""
        sigmas = _f32(sigmas).contiguous()
        rgbs = _f32(rgbs).contiguous()
        deltas = _f32(deltas).contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        accum_deltas, input_alpha = int(accum_deltas), int(input_alpha)
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        weights = torch.zeros_like(sigmas) if ret_weights else torch.zeros(0, device=sigmas.device)
        _backend.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, accum_deltas, input_alpha, weights_sum, depth,
                                              image, weights)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.aux_flags = (accum_deltas, input_alpha)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, depth, image, weights

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_depth, grad_image, grad_weights):
        grad_weights_sum = _f32(grad_weights_sum).contiguous()
        grad_image = _f32(grad_image).contiguous()
        grad_depth = _f32(grad_depth).contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        accum_deltas, input_alpha = ctx.aux_flags
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _backend.composite_rays_train_backward(grad_weights_sum, grad_image, grad_depth, sigmas, rgbs, deltas, rays, weights_sum, image,
                                               depth, M, N, T_thresh, grad_sigmas, grad_rgbs, accum_deltas, input_alpha)
        return grad_sigmas, grad_rgbs, None, None, None, None, None, None


composite_rays_train = _composite_rays_train.apply


class _march_rays(Function):
    @staticmethod
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
                perturb=False, dt_gamma=0, max_steps=1024):
        """raymarching.py:316-367.  Returns xyzs, dirs [n_alive*n_step (+pad), 3], deltas [.., 2]."""
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = _f32(rays_o).contiguous().view(-1, 3)
        rays_d = _f32(rays_d).contiguous().view(-1, 3)
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        dev = rays_o.device
        # only the alignment tail needs the zero fill: the kernel writes every slot of [0, n_alive*n_step)
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
        noises = torch.rand(n_alive, dtype=rays_o.dtype, device=dev) if perturb else torch.zeros(n_alive, dtype=rays_o.dtype, device=dev)
        _backend.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, density_bitfield,
                            near, far, xyzs, dirs, deltas, noises)
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2,
                input_alpha=False, accum_deltas=True):
        """raymarching.py:370-394.  In place on rays_alive, rays_t, weights_sum, depth, image."""
        _backend.composite_rays(n_alive, n_step, T_thresh, int(accum_deltas), int(input_alpha), rays_alive, rays_t,
                                _f32(sigmas).contiguous(), _f32(rgbs).contiguous(), _f32(deltas).contiguous(), weights_sum, depth, image)
        return tuple()


composite_rays = _composite_rays.apply
