#!/usr/bin/env python
"""n_step cap / secondary floor of the logged geometry passes: frame time, iterations, marched samples and max |d image| against the default, for the
800x800 frame and for one rank's share of the 1600x1600 frame at 8 ranks (run r3_22).  GPU box: python profiles/batching_experiment.py"""
import sys, torch
sys.path.insert(0, ".")
from envidr_b200 import dist as D, render, scene
dev = torch.device("cuda:0")
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5); fp.precision = "tc"; fp = fp.to(dev).pack()
bf = torch.from_numpy(scene.make_bitfield()).to(dev)
def run(W, ws, **kw):
    ro, rd = scene.camera_rays(W, W); ro, rd = ro.to(dev), rd.to(dev)
    cfg = render.RenderConfig(indir_ref=True, **kw)
    idx = D.tile_shard_indices(W, W, 0, ws).to(dev)
    o, d = ro[idx].contiguous(), rd[idx].contiguous()
    st = []
    for _ in range(4):
        st = []
        res = render.render(fp, bf, o, d, cfg, bg_color=1.0, stats=st)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); render.render(fp, bf, o, d, cfg, bg_color=1.0); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2], st, res["image"]
for W, ws in ((800, 1), (1600, 8)):
    base_img = None
    for kw in ({}, {"logged_n_step_cap": 16}, {"secondary_n_step_floor": 8}, {"logged_n_step_cap": 16, "secondary_n_step_floor": 8}):
        ms, st, img = run(W, ws, **kw)
        if base_img is None: base_img = img
        print(f"W={W} ws={ws} {kw}: {ms:.2f} ms; passes " + "; ".join(f"{s.get('iterations')} it / {s.get('samples')}" for s in st) + f"; max|d image| vs default {float((img - base_img).abs().max()):.2e}")
