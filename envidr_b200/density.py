"""Occupancy-grid maintenance: NeRFRenderer.update_extra_state / mark_untrained_grid (reference nerf/renderer.py:200-358)
on libenvidr_b200 (csrc/density.cu).  SURVEY.md 8 f-1: the caller on the input side of the march.

Two ways in, same code:
  * DensityGrid            the extra state NeRFRenderer keeps for cuda_ray (renderer.py:104-122) as a small object, for use
                           without the reference model (tests, bench, scene building);
  * install(model)         binds update_extra_state / mark_untrained_grid of a live reference NeRFNetwork instance to this
                           module (same names, arguments, attributes written), reading the field weights from the model
                           on every call (they change while training).

What the reference does per update: two 2 M-row torch pipelines (meshgrid -> morton3D -> jitter -> self.density() -> index_put)
plus mean().item() and packbits.  Here: cell enumeration in Morton order, the geometry kernel of the render path, one
EMA + mean pass, one pack pass; mean_density stays on the device (read back lazily).

Randomness: the jitter (and, for the partial update, the visited cells) is drawn with the torch ops of the reference in the
reference's call order, so a run seeded like the reference visits the same points (S >= grid_size, the default, is assumed
for that statement: the reference draws its jitter block by block).
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream
from .field import FieldParams


class DensityOpts(ctypes.Structure):
    _fields_ = [("bound", ctypes.c_float), ("cascade", ctypes.c_uint32), ("grid_size", ctypes.c_uint32), ("decay", ctypes.c_float),
                ("density_thresh", ctypes.c_float)]


class DensityGrid:
    """cuda_ray extra state of NeRFRenderer (renderer.py:70-76, 104-122)."""

    def __init__(self, bound: float = 1.0, density_thresh: float = 0.01, grid_size: int = 128, device="cuda"):
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = grid_size
        self.density_thresh = density_thresh
        self.density_grid = torch.zeros(self.cascade, grid_size ** 3, device=device)
        self.density_bitfield = torch.zeros(self.cascade * grid_size ** 3 // 8, dtype=torch.uint8, device=device)
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter = torch.zeros(16, 2, dtype=torch.int32, device=device)
        self.mean_count = 0
        self.local_step = 0

    def mark_untrained_grid(self, poses, intrinsic, S=64):
        return mark_untrained_grid(self, poses, intrinsic, S)

    def update_extra_state(self, field: FieldParams, decay=0.95, S=128, full_update=False, **kw):
        return update_extra_state(self, field, decay, S, full_update, **kw)


_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = str(device)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


@torch.no_grad()
def mark_untrained_grid(state, poses, intrinsic, S=64):
    """renderer.py:200-262: density_grid[cells no camera sees] = -1.  poses [B,4,4] camera-to-world, intrinsic (fx, fy, cx, cy).
    Returns the number of marked cells (the reference prints it)."""
    if isinstance(poses, np.ndarray):
        poses = torch.from_numpy(poses)
    dev = state.density_grid.device
    poses = poses.to(device=dev, dtype=torch.float32).contiguous()
    B = int(poses.shape[0])
    fx, fy, cx, cy = [float(v) for v in intrinsic]
    C, H = int(state.cascade), int(state.grid_size)
    grid = state.density_grid
    assert grid.is_contiguous() and grid.dtype == torch.float32
    count = torch.zeros(C * H ** 3, dtype=torch.int32, device=dev)
    for head in range(0, max(B, 1), 1024):
        chunk = poses[head:head + 1024]
        check(lib().envidr_mark_untrained_grid(ptr(chunk), int(chunk.shape[0]), fx, fy, cx, cy, float(state.bound), C, H, ptr(grid),
                                               ptr(count), stream()), "mark_untrained_grid")
    check(lib().envidr_mark_untrained_apply(ptr(count), count.numel(), ptr(grid), stream()), "mark_untrained_apply")
    return count


def draw_partial_cells(state):
    """The cells a partial update visits (renderer.py:310-324), drawn with the reference's torch calls in its order:
    per cascade H^3/4 uniform cells + H^3/4 cells resampled from the occupied ones.  -> int32 [C, H^3/2, 3], noise [C, H^3/2, 3]."""
    from . import raymarching
    H, C = int(state.grid_size), int(state.cascade)
    dev = state.density_grid.device
    N = H ** 3 // 4
    coords_all, noise_all = [], []
    for cas in range(C):
        coords = torch.randint(0, H, (N, 3), device=dev)
        occ_indices = torch.nonzero(state.density_grid[cas] > 0).squeeze(-1)
        rand_mask = torch.randint(0, occ_indices.shape[0], [N], dtype=torch.long, device=dev)
        occ_coords = raymarching.morton3D_invert(occ_indices[rand_mask])
        coords = torch.cat([coords.int(), occ_coords.int()], dim=0)
        coords_all.append(coords)
        noise_all.append(torch.rand(2 * N, 3, device=dev))                  # torch.rand_like(cas_xyzs)
    return torch.stack(coords_all).contiguous(), torch.stack(noise_all).contiguous()


@torch.no_grad()
def update_extra_state(state, field: FieldParams, decay=0.95, S=128, full_update=False, *, noise: Optional[torch.Tensor] = None,
                       coords: Optional[torch.Tensor] = None, sync: bool = True):
    """renderer.py:264-358.  state: a DensityGrid or a reference NeRFRenderer (attributes density_grid, density_bitfield,
    cascade, grid_size, bound, density_thresh, iter_density, mean_density, step_counter, mean_count, local_step).
    field: the packed FieldParams whose density the grid tracks (density_scale is part of it).
    noise / coords: override the random draws (tests); noise=False switches the jitter off.
    sync=False leaves mean_density on the device (state.density_stats[0]) instead of the reference's .item()."""
    C, H = int(state.cascade), int(state.grid_size)
    dev = state.density_grid.device
    grid = state.density_grid
    assert grid.is_contiguous() and grid.dtype == torch.float32 and grid.numel() == C * H ** 3
    if field._packed is None:
        field.pack()
    if coords is None and not (state.iter_density < 16 or full_update):
        coords, drawn = draw_partial_cells(state)
        if noise is None:
            noise = drawn
    if coords is None:
        n = H ** 3
        if noise is None:
            noise = torch.stack([torch.rand(n, 3, device=dev) for _ in range(C)])      # one rand_like per cascade (S >= H)
    else:
        coords = coords.to(device=dev, dtype=torch.int32).contiguous().view(C, -1, 3)
        n = int(coords.shape[1])
        if noise is None:
            noise = torch.stack([torch.rand(n, 3, device=dev) for _ in range(C)])
    if noise is False:
        noise = None
    if noise is not None:
        noise = noise.to(device=dev, dtype=torch.float32).contiguous()
        assert noise.numel() == C * n * 3
    opts = DensityOpts(float(state.bound), C, H, float(decay), float(state.density_thresh))
    nbytes = int(lib().envidr_density_workspace_bytes(C, H, n))
    ws = _workspace(nbytes, dev)
    stats = getattr(state, "density_stats", None)
    if stats is None or stats.device != dev:
        stats = torch.zeros(2, dtype=torch.float32, device=dev)
        state.density_stats = stats
    bitfield = state.density_bitfield
    assert bitfield.dtype == torch.uint8 and bitfield.numel() == C * H ** 3 // 8 and bitfield.is_contiguous()
    f = field.cstruct()
    check(lib().envidr_density_grid_update(ctypes.byref(f), ptr(grid), ptr(coords), ptr(noise), n, ctypes.byref(opts), ptr(bitfield),
                                           ptr(stats), ptr(ws), ws.numel(), stream()), "density_grid_update")
    if sync:
        state.mean_density = float(stats[0].item())
    state.iter_density += 1
    # step counter (renderer.py:354-358)
    total_step = min(16, state.local_step)
    if total_step > 0:
        state.mean_count = int(state.step_counter[:total_step, 0].sum().item() / total_step)
    state.local_step = 0


def install(model):
    """Bind update_extra_state / mark_untrained_grid of a reference NeRFNetwork instance (nerf/renderer.py) to this module.
    The call signatures stay the reference's; the field weights are re-read from the model on every update."""
    import types

    def _update(self, decay=0.95, S=128, full_update=False):
        if not self.cuda_ray:
            return
        fp = FieldParams.from_reference_model(self)
        fp.precision = "tc"
        update_extra_state(self, fp.pack(), decay, S, full_update)

    def _mark(self, poses, intrinsic, S=64):
        if not self.cuda_ray:
            return
        count = mark_untrained_grid(self, poses, intrinsic, S)
        print(f'[mark untrained grid] {(count == 0).sum()} from {self.grid_size ** 3 * self.cascade}')

    model.update_extra_state = types.MethodType(_update, model)
    model.mark_untrained_grid = types.MethodType(_mark, model)
    return model
