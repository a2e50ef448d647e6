"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: interleaved-tile ray sharding and the all-gather that
assembles the frame.  The render function is a deterministic stand-in (no CUDA here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from envidr_b200 import dist as edist


def test_tile_sharding_is_a_partition_and_balanced():
    for H, W, ws in ((800, 800, 8), (37, 53, 4), (64, 64, 3), (5, 3, 2)):
        parts = [edist.tile_shard_indices(H, W, r, ws) for r in range(ws)]
        allidx = torch.cat(parts)
        assert allidx.numel() == H * W and torch.equal(torch.sort(allidx).values, torch.arange(H * W))
        assert edist.shard_sizes(H, W, ws) == [p.numel() for p in parts]
        if H * W >= 64 * 64 * 4:
            sizes = torch.tensor([p.numel() for p in parts], dtype=torch.float64)
            assert sizes.max() / sizes.min() < 1.05
    # 8x8 tiles: the first 8 pixels of row 0 belong to rank 0, the next 8 to rank 1
    p0 = edist.tile_shard_indices(16, 32, 0, 2)
    assert p0[:8].tolist() == list(range(8)) and 8 not in p0.tolist()


def _fake_render(o, d):
    n = o.shape[0]
    return dict(image=o * 2 + d, depth=o[:, 0] + 1, weights_sum=d[:, 1] * 3, normal_image=d - o)


def _worker(rank, world, port, H, W, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    o, d = torch.randn(H * W, 3, generator=g), torch.randn(H * W, 3, generator=g)
    out = edist.render_sharded(_fake_render, o, d, H, W)
    ref = _fake_render(o, d)
    ok = all(torch.equal(out[k], ref[k]) for k in ref)
    frames = edist.gather_frames(torch.full((7, 3), float(rank)))
    ok = ok and frames.shape == (world, 7, 3) and all(float(frames[r].mean()) == r for r in range(world))
    torch.save(ok, os.path.join(tmp, f"ok{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("H,W", [(40, 56), (37, 19)])
def test_render_sharded_gloo_world2(tmp_path, H, W):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, H, W, str(tmp_path)), nprocs=2, join=True)
    assert all(torch.load(os.path.join(tmp_path, f"ok{r}.pt")) for r in range(2))


def test_single_process_paths():
    o, d = torch.randn(24 * 24, 3), torch.randn(24 * 24, 3)
    out = edist.render_sharded(_fake_render, o, d, 24, 24)
    assert torch.equal(out["image"], o * 2 + d)
    assert edist.gather_frames(torch.ones(4, 3)).shape == (1, 4, 3)


def _dp_worker(rank, world, port, tmp):
    """Data-parallel step: every rank differentiates the loss of ITS half of the batch; after allreduce_gradients both hold the
    gradient of the full-batch mean loss and an identical optimizer step keeps the replicas in sync."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    W1, b1, unused = torch.randn(5, 3, generator=g), torch.randn(5, generator=g), torch.randn(2, generator=g)
    x, y = torch.randn(8, 3, generator=g), torch.randn(8, 5, generator=g)
    params = [W1.clone().requires_grad_(True), b1.clone().requires_grad_(True), unused.clone().requires_grad_(True)]
    sl = slice(rank * 4, rank * 4 + 4)
    ((x[sl] @ params[0].T + params[1] - y[sl]) ** 2).mean().backward()
    edist.allreduce_gradients(params)
    full = [W1.clone().requires_grad_(True), b1.clone().requires_grad_(True)]
    ((x @ full[0].T + full[1] - y) ** 2).mean().backward()
    ok = all(torch.allclose(p.grad, f.grad, atol=1e-6) for p, f in zip(params, full)) and params[2].grad is None
    opt = torch.optim.Adam(params[:2], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    opt.step()
    gathered = [torch.zeros_like(params[0]) for _ in range(world)]
    dist.all_gather(gathered, params[0].detach())
    ok = ok and torch.equal(gathered[0], gathered[1])
    torch.save(ok, os.path.join(tmp, f"dp{rank}.pt"))
    dist.destroy_process_group()


def test_allreduce_gradients_gloo_world2(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(torch.load(os.path.join(tmp_path, f"dp{r}.pt")) for r in range(2))
    p = torch.ones(3, requires_grad=True)
    p.grad = torch.ones(3)
    edist.allreduce_gradients([p])                       # single process: a no-op
    assert torch.equal(p.grad, torch.ones(3))
