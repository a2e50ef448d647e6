#!/usr/bin/env python
"""Ray generation and loss epilogue at the training-step size (4,096 rays, 70 k samples) and get_rays at
800x800: our kernels against the same steps written with the reference's torch ops on the same GPU (oracle/train_oracle.py's
restatement in fp32 on CUDA = what Trainer.train_step + run_cuda's auxiliary block execute).  Median over bursts; one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from envidr_b200 import epilogue, scene


def timed(fn, reps=10, warm=3, burst=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(burst):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b) / burst)
    return float(np.median(ms))


def torch_get_rays(poses, intrinsics, H, W):
    """nerf/utils.py:110-209, full-image branch, on the GPU."""
    device = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = intrinsics
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=device), torch.linspace(0, H - 1, H, device=device), indexing="ij")
    i = i.t().reshape([1, H * W]).expand([B, H * W]) + 0.5
    j = j.t().reshape([1, H * W]).expand([B, H * W]) + 0.5
    zs = torch.ones_like(i)
    directions = torch.stack(((i - cx) / fx * zs, (j - cy) / fy * zs, zs), dim=-1)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_d = directions @ poses[:, :3, :3].transpose(-1, -2)
    rays_o = poses[..., :3, 3][..., None, :].expand_as(rays_d)
    return rays_o, rays_d


def torch_loss(image, ws, gt, mask, sdfs, sgrad, weights, deltas, rays, beta, cfg):
    """Trainer.train_step's terms + run_cuda's auxiliary block with the reference's torch ops (fp32, CUDA), forward + backward."""
    M = sdfs.shape[0]
    loss = cfg.color_w * (image - gt).abs().mean(-1).mean()
    loss = loss + cfg.mask_w * torch.nn.functional.binary_cross_entropy(ws.clip(1e-3, 1.0 - 1e-3), mask)
    point_mask = torch.ones_like(sdfs, dtype=torch.bool)
    ray_valid = (rays[:, 2] > 0) * (rays[:, 1] + rays[:, 2] < M)
    start = rays[ray_valid, 1]
    end = start + rays[ray_valid, 2] - 1
    point_mask[end.long()] = False
    ds = torch.roll(deltas, -1, 0)
    point_mask = point_mask * ((ds[:, 0] > 0) * (ds[:, 1] > 0)) * (ds[:, 1] < 1.2 * ds[:, 0])
    relsdf = (torch.roll(sdfs, -1, dims=0) - sdfs)[point_mask]
    w, dist, sd = weights[point_mask], ds[point_mask, 1], sdfs[point_mask]
    m = (w > cfg.backsdf_thresh) * (relsdf > 0)
    s_sq = relsdf[m] ** 2
    loss = loss + cfg.backsdf_w * (w[m] * (s_sq / (dist[m].clamp(min=5e-4) ** 2 + s_sq))).sum()
    reg = 0.5 + 0.5 * sd.sign() * torch.expm1(-sd.abs() / beta)
    loss = loss + cfg.cauchy_w * 0.25 * torch.log1p((1 - reg) ** 2 * 16.0).mean()
    loss = loss + cfg.eikonal_w * ((sgrad.norm(p=2, dim=-1) - 1) ** 2).mean()
    return loss


def measure(dev):
    out = {}
    H = W = 800
    pose = torch.from_numpy(scene.nerf_matrix_to_ngp(scene.pose_spherical(40.0, -30.0, 4.0), scale=0.65))[None].to(dev)
    intr = scene.intrinsics_from_fov(W, H, 0.6911112070083618)
    out["get_rays_800_ms"] = timed(lambda: epilogue.get_rays(pose, intr, H, W))
    out["get_rays_800_torch_ms"] = timed(lambda: torch_get_rays(pose, intr, H, W))
    out["get_rays_GBps"] = H * W * 24 / (out["get_rays_800_ms"] * 1e-3) / 1e9
    g = torch.Generator().manual_seed(0)
    N, M = 4096, 70_000
    cnt = torch.randint(0, 34, (N,), generator=g)
    off = torch.cumsum(cnt, 0) - cnt
    total = int(cnt.sum())
    rays = torch.stack([torch.arange(N), off, cnt], -1).int().to(dev)
    deltas = torch.zeros(M, 2); deltas[:total, 0] = 0.0034; deltas[:total, 1] = 0.0034
    deltas = deltas.to(dev)
    mk = lambda *s: torch.rand(*s, generator=g).to(dev)
    image, ws, sdfs, sgrad = mk(N, 3).requires_grad_(True), mk(N).requires_grad_(True), ((mk(M) - 0.5) * 0.05).requires_grad_(True), mk(M, 3).requires_grad_(True)
    gt, mask, weights = mk(N, 3), (mk(N) > 0.5).float(), mk(M) * 0.05
    beta = torch.tensor(0.015, device=dev)
    cfg = epilogue.LossConfig()

    def ours():
        for t in (image, ws, sdfs, sgrad):
            t.grad = None
        total_, _ = epilogue.train_loss(image, ws, sdfs, sgrad, gt, mask, weights, deltas, rays, beta, cfg)
        total_.backward()

    def ref():
        for t in (image, ws, sdfs, sgrad):
            t.grad = None
        torch_loss(image, ws, gt, mask, sdfs, sgrad, weights, deltas, rays, beta, cfg).backward()

    out["loss_fwd_bwd_ms"] = timed(ours)
    out["loss_fwd_bwd_torch_ms"] = timed(ref)
    out["loss_speedup"] = out["loss_fwd_bwd_torch_ms"] / out["loss_fwd_bwd_ms"]
    out["get_rays_speedup"] = out["get_rays_800_torch_ms"] / out["get_rays_800_ms"]
    out["what"] = "4,096 rays / 70,000 samples, colour L1 + mask BCE + back-sdf + Cauchy + eikonal, forward + backward (host-inclusive: bursts of 10)"
    return out


if __name__ == "__main__":
    print(json.dumps(measure(torch.device("cuda:0"))))
