#!/usr/bin/env python
"""CUDA-event breakdown of the replayed main pass (800x800, 3 passes).  GPU box: python profiles/replay_breakdown.py"""
import sys

import torch

sys.path.insert(0, ".")
from envidr_b200 import _lib, render, scene  # noqa: E402

dev = torch.device("cuda:0")
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
fp.precision = "tc"
fp = fp.to(dev).pack()
bf = torch.from_numpy(scene.make_bitfield()).to(dev)
ro, rd = scene.camera_rays(800, 800)
ro, rd = ro.to(dev), rd.to(dev)
cfg = render.RenderConfig(indir_ref=True)
marks = []
orig_forward = type(fp).forward
orig_check = render.check


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def forward(self, *a, **k):
    a0 = ev()
    r = orig_forward(self, *a, **k)
    marks.append(("field.forward M=%d" % a[0].shape[0], a0, ev()))
    return r


type(fp).forward = forward
for it in range(5):
    marks.clear()
    lib = _lib.lib()
    lib.envidr_render_timing(1)
    t0 = ev()
    out = render.render(fp, bf, ro, rd, cfg, bg_color=1.0)
    t1 = ev()
    torch.cuda.synchronize()
    import ctypes
    fms, fl = ctypes.c_float(), ctypes.c_uint32()
    lib.envidr_render_field_time(ctypes.byref(fms), ctypes.byref(fl))
    lib.envidr_render_timing(0)
    if it >= 3:
        print(f"frame {t0.elapsed_time(t1):.2f} ms; env_tc total {fms.value:.2f} ms over {fl.value} launches; "
              + "; ".join(f"{n}: {a.elapsed_time(b):.2f} ms" for n, a, b in marks))
