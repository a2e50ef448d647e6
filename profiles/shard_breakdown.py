#!/usr/bin/env python
"""Where one sharded 1600x1600 frame goes at N ranks: per-rank render time (CUDA events around render.render), the packing + all-gather +
de-interleave, and the spread over the ranks.  torchrun --nproc-per-node N profiles/shard_breakdown.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from envidr_b200 import dist as D, render, scene  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
W = 1600
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5); fp.precision = "tc"; fp = fp.to(dev).pack()
bf = torch.from_numpy(scene.make_bitfield()).to(dev)
ro, rd = scene.camera_rays(W, W)
idx = D.tile_shard_indices(W, W, rank, world).to(dev)
o, d = ro.to(dev)[idx].contiguous(), rd.to(dev)[idx].contiguous()
cfg = render.RenderConfig(indir_ref=True)
marks = []


def fn(a, b):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = render.render(fp, bf, a, b, cfg, bg_color=1.0, get_normal_image=True); e1.record()
    marks.append((e0, e1))
    return r


flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for _ in range(4):
    D.render_sharded(fn, o, d, W, W, presharded=True)
marks.clear()
tot = []
for i in range(8):
    flush.fill_(1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); D.render_sharded(fn, o, d, W, W, presharded=True); b.record()
    torch.cuda.synchronize()
    tot.append(a.elapsed_time(b))
ren = [x.elapsed_time(y) for x, y in marks]
t = torch.tensor([sorted(tot)[4], sorted(ren)[4]], device=dev)
allt = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allt, t)
else:
    allt = [t]
if rank == 0:
    print(f"world {world}: per rank (frame ms, render ms): " + ", ".join(f"({float(x[0]):.2f}, {float(x[1]):.2f})" for x in allt))
    print(f"   max frame {max(float(x[0]) for x in allt):.2f} ms, max render {max(float(x[1]) for x in allt):.2f} ms, min render {min(float(x[1]) for x in allt):.2f} ms")
if world > 1:
    dist.destroy_process_group()
