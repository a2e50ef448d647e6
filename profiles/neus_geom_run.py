#!/usr/bin/env python
"""One batch of the fused NeuS geometry kernel (csrc/neus_geom_tc.cu) for ncu / timing: python profiles/neus_geom_run.py [M] [reps]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from envidr_b200 import scene  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 148 * 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
nf = scene.make_neus_field(0).to(dev).pack()
rng = np.random.default_rng(0)
u = rng.standard_normal((M, 3)); u /= np.linalg.norm(u, axis=-1, keepdims=True)
x = torch.from_numpy((u * rng.uniform(0.2, 0.7, (M, 1))).astype(np.float32)).to(dev)
for _ in range(2):
    nf._geometry_fused(x)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, b in ev:
    a.record(); nf._geometry_fused(x); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]
flop = (sum(2.0 * W.shape[0] * W.shape[1] for W, _ in nf.sdf) + sum(2.0 * W.shape[0] * W.shape[1] for W, _ in nf.sdf[:-1])) * M
print(f"M={M} tiles/SM={M / 128 / 148:.1f} {ms:.3f} ms  {ms * 1e3 / (M / 128 / 148):.1f} us per tile  {flop / ms / 1e9:.1f} TFLOP/s algorithmic")
