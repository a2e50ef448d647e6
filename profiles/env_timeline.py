#!/usr/bin/env python
"""Device-side timeline of k_env_tc (CTA 0): python profiles/env_timeline.py [M]   (GPU box)
Prints, per tile, the issuer / IDE / epilogue timestamps in cycles relative to the tile's first event."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from envidr_b200 import _lib, scene  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64 * 6
dev = torch.device("cuda:0")
fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
fp.precision = "tc"
fp = fp.to(dev).pack()
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.uniform(-0.6, 0.6, (M, 3)).astype(np.float32)).to(dev)
d = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
for _ in range(2):
    fp.forward(x, d)
PT = 48
cap = 1 + PT * 64
buf = torch.zeros(cap, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.envidr_debug_env_tc_timeline(ctypes.c_void_p(buf.data_ptr()), cap)
fp.forward(x, d)
torch.cuda.synchronize()
lib.envidr_debug_env_tc_timeline(None, 0)
b = buf.cpu().numpy()
n = int(b[0])
print(f"M={M} tiles recorded for CTA 0: {n}")
t00 = None
for t in range(n):
    r = b[1 + t * PT: 1 + (t + 1) * PT]
    if t00 is None:
        t00 = min(v for v in (r[0], r[8]) if v > 0)
    rel = lambda v: int(v - t00) if v > 0 else -1
    print(f"tile {t}: issuer wait_ide {rel(r[0])}->{rel(r[1])} L0 issued {rel(r[2])} last issued {rel(r[3])} | in wait(a_rdy) {int(r[4])} wait(full) {int(r[5])}"
          f" | IDE wait_empty {rel(r[8])}->{rel(r[9])} done {rel(r[10])}"
          f" | epi " + " ".join(f"L{l}:{rel(r[12 + 2 * l])}->{rel(r[13 + 2 * l])}" for l in range(4)))
    print("        layer 2 detail: epi(L1) wake w4 %d w8 %d | chunks published w4 %s w8 %s | issuer a_rdy[c] passed %s | acc_ready committed %d"
          % (rel(r[14]), rel(r[41]), [rel(v) for v in r[33:37]], [rel(v) for v in r[37:41]], [rel(v) for v in r[24:32]], rel(r[32])))
