#!/usr/bin/env python
"""CUDA-event time of the pieces of dist.render_sharded's frame assembly (pack, all_gather_into_tensor, torch index_select, slicing) on synthetic
per-ray outputs of a 1600x1600 frame (run r3_27b: the index_select was 1.33 ms of 1.5; it is envidr_gather_rows now).
torchrun --nproc-per-node N profiles/gather_pieces.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from envidr_b200 import dist as D
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = 1600
n_r = W * W // world
res = {"image": torch.rand(n_r, 3, device=dev), "depth": torch.rand(n_r, device=dev), "weights_sum": torch.rand(n_r, device=dev), "normal_image": torch.rand(n_r, 3, device=dev)}
keys = ("image", "depth", "weights_sum", "normal_image")
inv = D._gather_index(W, W, world, dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    e = [ev() for _ in range(5)]
    e[0].record()
    cols = [res[k].reshape(n_r, -1).float() for k in keys]
    packed = torch.cat(cols, -1).contiguous()
    e[1].record()
    out = packed.new_empty(world * n_r, 8)
    dist.all_gather_into_tensor(out, packed)
    e[2].record()
    full = out.index_select(0, inv)
    e[3].record()
    outd, c0 = {}, 0
    for k in keys:
        w = res[k].reshape(n_r, -1).shape[1]
        outd[k] = full[:, c0:c0 + w].reshape(W * W, *res[k].shape[1:]); c0 += w
    e[4].record()
    torch.cuda.synchronize()
    if rank == 0 and it >= 3:
        print(f"world {world}: pack {e[0].elapsed_time(e[1]):.3f} all_gather {e[1].elapsed_time(e[2]):.3f} index_select {e[2].elapsed_time(e[3]):.3f} slice {e[3].elapsed_time(e[4]):.3f} ms")
dist.destroy_process_group()
