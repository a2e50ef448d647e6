"""Multi-GPU rendering: rays are independent, so the path shards with NO data-path collective; the only
exchange is the final assembly of the image (one all-gather), as SURVEY.md 8e lays out.

The reference has no reachable multi-GPU path (its DDP scaffold is never activated, nerf/utils.py:360-402,
1353-1371), so there is no reference operator to mirror here; the partitioning follows the cost structure of the
path: pixels on the object cost 10-100x more samples than background pixels, hence rays are dealt to ranks in
interleaved 8x8-pixel tiles, round-robin, not in contiguous row blocks.

All model state (48.8 MB hash table, bit field, < 1 MB of MLP weights) is replicated on every rank.
One process per GPU; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is only plumbing.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist

TILE = 8


_idx_cache: Dict[tuple, torch.Tensor] = {}


def tile_shard_indices(H: int, W: int, rank: int, world_size: int, tile: int = TILE) -> torch.Tensor:
    """Flat pixel indices (row-major, int64) of the rays owned by `rank`: tiles of tile x tile pixels are numbered
    row-major and tile t belongs to rank t % world_size.  Ragged borders are handled (partial tiles)."""
    key = (H, W, rank, world_size, tile)
    if key not in _idx_cache:
        if len(_idx_cache) > 64:
            _idx_cache.clear()
        _idx_cache[key] = _tile_shard_indices(H, W, rank, world_size, tile)
    return _idx_cache[key]


def _tile_shard_indices(H: int, W: int, rank: int, world_size: int, tile: int) -> torch.Tensor:
    ty = torch.arange(H) // tile
    tx = torch.arange(W) // tile
    tiles_x = (W + tile - 1) // tile
    tid = ty[:, None] * tiles_x + tx[None, :]
    return torch.nonzero((tid % world_size == rank).reshape(-1)).flatten()


_dev_cache: Dict[tuple, torch.Tensor] = {}


def _dev_idx(H: int, W: int, rank: int, world_size: int, device) -> torch.Tensor:
    key = (H, W, rank, world_size, str(device))
    if key not in _dev_cache:
        if len(_dev_cache) > 64:
            _dev_cache.clear()
        _dev_cache[key] = tile_shard_indices(H, W, rank, world_size).to(device)
    return _dev_cache[key]


def shard_sizes(H: int, W: int, world_size: int, tile: int = TILE):
    return [int(tile_shard_indices(H, W, r, world_size, tile).numel()) for r in range(world_size)]


_inv_cache: Dict[tuple, torch.Tensor] = {}


def _gather_index(H: int, W: int, world_size: int, device) -> torch.Tensor:
    """inv [H*W] int64: pixel p lives at row inv[p] of the all-gathered [world_size * n_max, C] buffer (rank r's rows start at
    r * n_max, in the order of tile_shard_indices).  The frame is assembled with ONE index_select."""
    key = (H, W, world_size, str(device))
    if key not in _inv_cache:
        if len(_inv_cache) > 16:
            _inv_cache.clear()
        sizes = shard_sizes(H, W, world_size)
        n_max = max(sizes)
        inv = torch.empty(H * W, dtype=torch.int64)
        for r in range(world_size):
            idx = tile_shard_indices(H, W, r, world_size)
            inv[idx] = r * n_max + torch.arange(idx.numel())
        _inv_cache[key] = inv.to(device)
    return _inv_cache[key]


def _deinterleave(out: torch.Tensor, inv: torch.Tensor) -> torch.Tensor:
    """full[p] = out[inv[p]].  On CUDA with whole-float4 rows: envidr_gather_rows (torch's index_select moves the 82 MB of a 1600x1600 frame
    at 62 GB/s: 1.33 ms against 0.12 ms for the all-gather itself, run r3_27); otherwise (gloo tests on the CPU) index_select."""
    if out.is_cuda and out.dtype == torch.float32 and out.shape[1] % 4 == 0 and out.is_contiguous():
        from ._lib import check, lib, ptr, stream
        key = ("i32", inv.data_ptr())
        idx32 = _inv_cache.get(key)
        if idx32 is None:
            idx32 = inv.to(torch.int32)
            _inv_cache[key] = idx32
        full = torch.empty(inv.shape[0], out.shape[1], dtype=torch.float32, device=out.device)
        check(lib().envidr_gather_rows(ptr(out), ptr(idx32), int(inv.shape[0]), int(out.shape[1]), ptr(full), stream()), "gather_rows")
        return full
    return out.index_select(0, inv)


def render_sharded(render_fn: Callable[[torch.Tensor, torch.Tensor], Dict[str, torch.Tensor]], rays_o: torch.Tensor, rays_d: torch.Tensor,
                   H: int, W: int, keys=("image", "depth", "weights_sum", "normal_image"), group=None, presharded: bool = False) -> Dict[str, torch.Tensor]:
    """Strong-scaling render of ONE frame: every rank renders its interleaved tiles with `render_fn(rays_o, rays_d)` and
    one all-gather of the packed per-ray outputs ([n_r, C] fp32, C = 8 for rgb+depth+ws+normal = 32 B/ray) assembles the full
    frame on every rank.  rays_o / rays_d are the full [H*W, 3] tensors (replicated inputs), or with presharded=True this rank's rows."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if presharded:                                                   # rays_o / rays_d are already this rank's tile_shard_indices rows
        res = render_fn(rays_o, rays_d)
    else:
        idx = _dev_idx(H, W, rank, ws, rays_o.device)
        res = render_fn(rays_o[idx].contiguous(), rays_d[idx].contiguous())
    cols = []
    for k in keys:
        v = res[k]
        cols.append(v.reshape(v.shape[0], -1).float())
    packed = torch.cat(cols, -1).contiguous()
    C = packed.shape[1]
    if ws == 1:
        out = packed
    else:
        n_max = max(shard_sizes(H, W, ws))
        if packed.shape[0] != n_max:                                 # ragged frame sizes only: equal contributions are sent as they are
            buf = packed.new_zeros(n_max, C)
            buf[: packed.shape[0]] = packed
            packed = buf
        out = packed.new_empty(ws * n_max, C)
        dist.all_gather_into_tensor(out, packed, group=group)        # the one collective of the path
    full = _deinterleave(out, _gather_index(H, W, ws, out.device))   # de-interleave the tiles: one gather kernel
    outd, c0 = {}, 0
    for k in keys:
        w = res[k].reshape(res[k].shape[0], -1).shape[1]
        outd[k] = full[:, c0:c0 + w].reshape(H * W, *res[k].shape[1:])
        c0 += w
    return outd


def gather_frames(frame: torch.Tensor, group=None) -> torch.Tensor:
    """Weak-scaling sweep (one frame / light rotation per rank): all-gather the [N, C] frames -> [world, N, C]."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return frame[None]
    ws = dist.get_world_size(group)
    out = frame.new_empty(ws, *frame.shape)
    dist.all_gather_into_tensor(out.view(ws * frame.shape[0], *frame.shape[1:]), frame.contiguous(), group=group)
    return out


def allreduce_gradients(params, group=None, average: bool = True) -> None:
    """Data-parallel training (optional): every rank runs the training step on its share of the ray batch, then the
    gradients of the replicated state (hash table 48.8 MB dense + < 1 MB of MLP weights) are summed / averaged with ONE all-reduce
    over a flat buffer -- the only exchange step of data-parallel training; the optimizer step stays local and identical on all ranks.
    The reference has no reachable multi-GPU training path (its DDP scaffold is never activated, nerf/utils.py:360-402)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)
