#!/usr/bin/env python
"""CUDA-event time of the phases of one 800x800 three-pass frame (default schedule): the two geometry loops (render_rays), the log gathers
(prepare_from_log), the two shading batches (shade_prepared = env_net + heads + compositor), and what is left (torch glue, host
synchronisations).  GPU box: python profiles/frame_phases.py [W]"""
import sys

import torch

sys.path.insert(0, ".")
from envidr_b200 import render, scene  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 800
dev = torch.device("cuda:0")
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
fp.precision = "tc"
fp = fp.to(dev).pack()
bf = torch.from_numpy(scene.make_bitfield()).to(dev)
ro, rd = scene.camera_rays(W, W)
ro, rd = ro.to(dev), rd.to(dev)
cfg = render.RenderConfig(indir_ref=True)
marks = []


def wrap(name):
    orig = getattr(render, name)

    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **k)
        e1.record()
        marks.append((name, e0, e1))
        return r
    setattr(render, name, f)


for it in range(8):
    if it == 4:
        for n in ("render_rays", "prepare_from_log", "shade_prepared", "last_stats"):
            wrap(n)
    marks.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    out = render.render(fp, bf, ro, rd, cfg, bg_color=1.0)
    t1.record()
    torch.cuda.synchronize()
    if it >= 4:
        per = [(n, a.elapsed_time(b)) for n, a, b in marks]
        tot = t0.elapsed_time(t1)
        print(f"frame {tot:.2f} ms | " + ", ".join(f"{n} {ms:.2f}" for n, ms in per) + f" | rest {tot - sum(ms for _, ms in per):.2f} ms")
