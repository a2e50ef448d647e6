#!/usr/bin/env python
"""Per-kernel GPU time of ONE live 800x800 3-pass frame (torch.profiler CUDA activities, warm caches, real overlap) and the
idle time between kernels (frame span - sum of kernel durations): what the ncu launch list cannot show.
GPU box: python profiles/frame_kernels.py > gpurun_out/frame_kernels.md"""
import collections
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from envidr_b200 import render, scene  # noqa: E402

dev = torch.device("cuda:0")
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
fp.precision = "tc"
fp = fp.to(dev).pack()
bf = torch.from_numpy(scene.make_bitfield()).to(dev)
ro, rd = scene.camera_rays(800, 800)
ro, rd = ro.to(dev), rd.to(dev)
cfg = render.RenderConfig(indir_ref=True)
for _ in range(30):
    render.render(fp, bf, ro, rd, cfg, bg_color=1.0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        render.render(fp, bf, ro, rd, cfg, bg_color=1.0)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e.name.split("(")[0][:64]
    agg[n][0] += 1
    agg[n][1] += e.time_range.elapsed_us()
span = ev[-1].time_range.end - ev[0].time_range.start
busy = sum(v[1] for v in agg.values())
print(f"3 frames: span {span / 3e3:.2f} ms / frame, kernel time {busy / 3e3:.2f} ms / frame, idle {100 * (1 - busy / span):.1f} %\n")
print("| kernel | launches / frame | ms / frame | share of span |\n|---|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
    print(f"| `{n}` | {c / 3:.1f} | {t / 3e3:.3f} | {100 * t / span:.1f}% |")
