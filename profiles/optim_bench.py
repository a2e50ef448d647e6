#!/usr/bin/env python
"""One optimizer step over the toaster-dims trainable state (hash table 12.2 M floats + sdf / env / colour / diffuse / renv MLPs):
FusedAdam (one launch, csrc/optim.cu) against torch.optim.Adam (foreach, what the reference runs) + optimizer.zero_grad(set_to_none=False).
CUDA events, median.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from envidr_b200 import scene
from envidr_b200.optim import FusedAdam


def timed(fn, reps=10, warm=5, burst=20):
    """ms per call over `burst` back-to-back calls between two CUDA events (the host runs ahead of the device when it can, so
    this is max(host, device) time per call), median of `reps` bursts."""
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


def measure(dev, fp_cpu):
    fp = fp_cpu.to(dev)
    tensors = [fp.embeddings]
    for st in fp.stacks().values():
        if st is not None:
            for W, b in st:
                tensors += [W] + ([b] if b is not None else [])
    n = sum(t.numel() for t in tensors)
    out = {"parameters": n, "tensors": len(tensors)}

    def setup(cls, **kw):
        ps = [t.detach().clone().requires_grad_(True) for t in tensors]
        for p in ps:
            p.grad = torch.randn_like(p) * 1e-3
        return ps, cls([{"params": ps[:1], "lr": 1e-2}, {"params": ps[1:], "lr": 1e-3}], betas=(0.9, 0.99), eps=1e-15, **kw)

    ps, opt = setup(FusedAdam, zero_grad=True)
    out["fused_adam_ms"] = timed(opt.step)
    ps, opt = setup(torch.optim.Adam)

    def torch_step():
        opt.step()
        opt.zero_grad(set_to_none=False)
    out["torch_adam_foreach_ms"] = timed(torch_step)
    try:
        ps, opt = setup(torch.optim.Adam, fused=True)
        out["torch_adam_fused_ms"] = timed(torch_step)
    except Exception as e:
        out["torch_adam_fused_error"] = repr(e)[:100]
    out["algorithmic_bytes_per_param"] = 32                      # p, g, m, v read; p, m, v, g written
    out["fused_adam_GBps"] = n * 32 / (out["fused_adam_ms"] * 1e-3) / 1e9
    out["speedup_vs_torch_foreach"] = out["torch_adam_foreach_ms"] / out["fused_adam_ms"]
    return out


if __name__ == "__main__":
    print(json.dumps(measure(torch.device("cuda:0"), scene.make_synthetic_field(0))))
