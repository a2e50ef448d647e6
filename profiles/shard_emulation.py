#!/usr/bin/env python
"""One rank's share of the 1600x1600 config-5 frame on ONE GPU, for world sizes 1 / 2 / 4 / 8: what the per-rank time of the sharded frame is
without the all-gather (strong-scaling ceiling of bench.py --gpus N).  GPU box: python profiles/shard_emulation.py [W]"""
import sys

import torch

sys.path.insert(0, ".")
from envidr_b200 import dist as D, render, scene  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1600
dev = torch.device("cuda:0")
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
fp.precision = "tc"
fp = fp.to(dev).pack()
bf = torch.from_numpy(scene.make_bitfield()).to(dev)
ro, rd = scene.camera_rays(W, W)
ro, rd = ro.to(dev), rd.to(dev)
cfg = render.RenderConfig(indir_ref=True)
base = None
for ws in (1, 2, 4, 8):
    idx = D.tile_shard_indices(W, W, 0, ws).to(dev)
    o, d = ro[idx].contiguous(), rd[idx].contiguous()
    st = []
    for _ in range(4):
        st = []
        render.render(fp, bf, o, d, cfg, bg_color=1.0, stats=st)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); render.render(fp, bf, o, d, cfg, bg_color=1.0); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    base = base or ms
    print(f"world {ws}: rank-0 shard {o.shape[0]} rays, {ms:.2f} ms (ideal {base / ws:.2f}, efficiency {base / ws / ms:.2f}); passes: "
          + "; ".join(f"{s.get('iterations')} it / {s.get('samples')} samples" for s in st))
