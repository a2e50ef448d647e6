#!/usr/bin/env python
"""torch.profiler table of one train step (run_cuda train branch fwd+bwd, 4096 rays).  GPU box: python profiles/train_profile.py"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from envidr_b200 import render, scene, train  # noqa: E402

dev = torch.device("cuda:0")
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
bf = scene.make_bitfield()
ro, rd = scene.camera_rays(800, 800)
print(bench.train_step_bench(fp, bf, ro, rd, dev, 4096))
g = torch.Generator().manual_seed(0)
sel = torch.randperm(ro.shape[0], generator=g)[:4096]
o, d = ro[sel].to(dev), rd[sel].to(dev)
field = train.TrainableField(fp.to(dev))
bft = torch.from_numpy(bf).to(dev)
cfg = render.RenderConfig()
gt = torch.rand(4096, 3, device=dev); gm = (torch.rand(4096, device=dev) > 0.5).float(); ri = torch.rand(4096, 4, device=dev)


def step():
    for p in field.parameters():
        p.grad = None
    out = train.render_train(field, bft, o, d, cfg, r_images=ri, perturb=True)
    train.loss_epilogue(field, out, gt, gm).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=60))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.elapsed_us() for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
small = [e for e in ev if e.time_range.elapsed_us() < 10]
print(f"eager step: {len(ev)} GPU kernels/copies, busy {busy / 1e3:.2f} ms, span {span / 1e3:.2f} ms; {len(small)} of them < 10 us "
      f"({sum(e.time_range.elapsed_us() for e in small) / 1e3:.2f} ms)")
