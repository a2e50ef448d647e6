#!/usr/bin/env python
"""Where one 800x800 3-pass frame goes: CUDA-event time of each envidr_render_rays pass vs the whole render() call
(the difference is the torch glue between the passes: masks, boolean indexing, scatter).  GPU box: python profiles/pass_breakdown.py"""
import sys

import torch

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
orig = render.render_rays
marks = []


def timed_render_rays(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(*a, **k)
    e1.record()
    marks.append((e0, e1, a[2].shape[0]))
    return r


for it in range(6):
    if it == 3:
        render.render_rays = timed_render_rays
    marks.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    out = render.render(fp, bf, ro, rd, cfg, bg_color=1.0)
    t1.record()
    torch.cuda.synchronize()
    if it >= 3:
        per = [(a.elapsed_time(b), n) for a, b, n in marks]
        print(f"frame {t0.elapsed_time(t1):.2f} ms | passes " + ", ".join(f"{ms:.2f} ms ({n} rays)" for ms, n in per)
              + f" | glue {t0.elapsed_time(t1) - sum(ms for ms, _ in per):.2f} ms")
