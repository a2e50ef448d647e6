#!/usr/bin/env python
"""Operator-level micro-benchmarks: every kernel behind the reference's pybind operator surface, ours against the
reference's own kernel (oracle/_ref/*.so = unmodified reference sources rebuilt for sm_100a) with IDENTICAL arguments, against the
HBM roofline computed from the algorithmic bytes per unit of SURVEY 8d.

    python profiles/op_bench.py [log2 M ...]  > gpurun_out/op_bench.md          (GPU box; default sizes 2^18, 2^20, 2^22)

Inputs as SURVEY 8d prescribes: xyz ~ U[-1,1]^3 (seed 0), embeddings U(-1e-4, 1e-4) (reference init), dirs normalised N(0,1)^3, composite:
sigma = softplus(N(0,1)) * 50, delta0 = 2 sqrt(3) / 1024, ray lengths ~ Poisson(32).  CUDA events, bursts of 5, median of 7, outputs
preallocated (the operator surface's allocation / zero-fill conventions are the caller's in both implementations)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from envidr_b200 import backend as B
from envidr_b200 import scene

PEAK_GBPS = 6550.0
try:
    PEAK_GBPS = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", PEAK_GBPS))
except Exception:
    pass


def timed(fn, reps=7, burst=5, warm=2):
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


def ref_mods():
    try:
        from oracle import ref_cuda
        return {n: ref_cuda.ref_module(n) for n in ("_raymarching", "_hashencoder", "_gridencoder", "_freqencoder", "_shencoder")}
    except Exception as e:
        print(f"(reference kernels unavailable: {e!r})", file=sys.stderr)
        return {}


def main():
    dev = torch.device("cuda:0")
    sizes = [int(a) for a in sys.argv[1:]] or [18, 20, 22]
    R = ref_mods()
    g = torch.Generator().manual_seed(0)
    offsets_np, pls = scene.hash_offsets()
    offsets = torch.from_numpy(offsets_np).to(dev)
    T = int(offsets_np[-1])
    emb = ((torch.rand(T, 2, generator=g) * 2 - 1) * 1e-4).to(dev)
    L, C, D, H = 16, 2, 3, 16
    S = float(np.log2(pls))
    rows = []

    def add(op, M, unit_bytes, fn_ours, fn_ref):
        ms = timed(fn_ours)
        ms_ref = timed(fn_ref) if fn_ref is not None else None
        gbps = M * unit_bytes / (ms * 1e-3) / 1e9
        rows.append(dict(op=op, M=M, bytes_per_unit=unit_bytes, ms=ms, GBps=gbps, frac_hbm=gbps / PEAK_GBPS, ref_ms=ms_ref,
                         speedup=None if ms_ref is None else ms_ref / ms))

    for lg in sizes:
        M = 1 << lg
        x = torch.rand(M, 3, generator=g).to(dev)                        # encoder inputs live in [0, 1] (hashgrid.py:161)
        out = torch.empty(L, M, C, device=dev)
        dydx = torch.empty(M, L * D * C, device=dev)
        one = torch.empty(1, device=dev)
        hb, hr = B._hashencoder, R.get("_hashencoder")
        add("hash_encode_forward", M, 12 + 1024 + 128, lambda: hb.hash_encode_forward(x, emb, offsets, out, M, D, C, L, S, H, False, one),
            None if hr is None else lambda: hr.hash_encode_forward(x, emb, offsets, out, M, D, C, L, S, H, False, one))
        add("hash_encode_forward + dy_dx", M, 12 + 1024 + 128 + 384, lambda: hb.hash_encode_forward(x, emb, offsets, out, M, D, C, L, S, H, True, dydx),
            None if hr is None else lambda: hr.hash_encode_forward(x, emb, offsets, out, M, D, C, L, S, H, True, dydx))
        grad = torch.randn(L, M, C, generator=g).to(dev)
        gemb, gin = torch.zeros_like(emb), torch.zeros_like(x)
        add("hash_encode_backward (table + inputs)", M, 128 + 12 + 1024 + 384 + 12,
            lambda: hb.hash_encode_backward(grad, x, emb, offsets, gemb, M, D, C, L, S, H, True, dydx, gin),
            None if hr is None else lambda: hr.hash_encode_backward(grad, x, emb, offsets, gemb, M, D, C, L, S, H, True, dydx, gin))
        ggi = torch.randn(M, 3, generator=g).to(dev)
        gg, g2e = torch.zeros(L, M, C, device=dev), torch.zeros_like(emb)
        add("hash_encode_second_backward", M, 128 + 12 + 1024 + 384 + 12 + 128 + 1024,
            lambda: hb.hash_encode_second_backward(grad, x, emb, offsets, M, D, C, L, S, H, True, dydx, ggi, gg, g2e),
            None if hr is None else lambda: hr.hash_encode_second_backward(grad, x, emb, offsets, M, D, C, L, S, H, True, dydx, ggi, gg, g2e))
        # freq / SH
        xf = (torch.rand(M, 3, generator=g) * 2 - 1).to(dev)
        fo = torch.empty(M, 39, device=dev)
        fb, fr = B._freqencoder, R.get("_freqencoder")
        add("freq_encode_forward (deg 6)", M, 12 + 156, lambda: fb.freq_encode_forward(xf, M, 3, 6, 39, fo),
            None if fr is None else lambda: fr.freq_encode_forward(xf, M, 3, 6, 39, fo))
        dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1).to(dev)
        so = torch.empty(M, 16, device=dev)
        sb, sr = B._shencoder, R.get("_shencoder")
        add("sh_encode_forward (deg 4)", M, 12 + 64, lambda: sb.sh_encode_forward(dirs, so, M, 3, 4, None),
            None if sr is None else lambda: sr.sh_encode_forward(dirs, so, M, 3, 4, None))
        # composite (train): ray lengths ~ Poisson(32)
        n_rays = max(1, M // 32)
        cnt = torch.poisson(torch.full((n_rays,), 32.0), generator=g).clamp(0, 1024).int()
        cnt = torch.minimum(cnt, torch.tensor(1024))
        off = (torch.cumsum(cnt, 0) - cnt).int()
        tot = int(cnt.sum())
        Mc = max(tot, 1)
        rays = torch.stack([torch.arange(n_rays, dtype=torch.int32), off, cnt], -1).contiguous().to(dev)
        sig = (torch.nn.functional.softplus(torch.randn(Mc, generator=g)) * 50).to(dev)
        rgb = torch.rand(Mc, 3, generator=g).to(dev)
        dl = torch.full((Mc, 2), 2 * 3 ** 0.5 / 1024).to(dev)
        ws, dp, im, wt = torch.empty(n_rays, device=dev), torch.empty(n_rays, device=dev), torch.empty(n_rays, 3, device=dev), torch.empty(0, device=dev)
        rb, rr = B._raymarching, R.get("_raymarching")
        add("composite_rays_train_forward", Mc, 24 + 20 * n_rays / Mc,
            lambda: rb.composite_rays_train_forward(sig, rgb, dl, rays, Mc, n_rays, 1e-4, 1, 0, ws, dp, im, wt),
            None if rr is None else lambda: rr.composite_rays_train_forward(sig, rgb, dl, rays, Mc, n_rays, 1e-4, 1, 0, ws, dp, im, wt))
        gws, gim, gdp = torch.rand(n_rays, device=dev), torch.rand(n_rays, 3, device=dev), torch.zeros(n_rays, device=dev)
        gs, gr = torch.zeros(Mc, device=dev), torch.zeros(Mc, 3, device=dev)
        add("composite_rays_train_backward", Mc, 24 + 16 + 40 * n_rays / Mc,
            lambda: rb.composite_rays_train_backward(gws, gim, gdp, sig, rgb, dl, rays, ws, im, dp, Mc, n_rays, 1e-4, gs, gr, 1, 0),
            None if rr is None else lambda: rr.composite_rays_train_backward(gws, gim, gdp, sig, rgb, dl, rays, ws, im, dp, Mc, n_rays, 1e-4, gs, gr, 1, 0))
    # march (train) over the synthetic scene, packbits
    bf = torch.from_numpy(scene.make_bitfield()).to(dev)
    ro, rd = scene.camera_rays(800, 800)
    ro, rd = ro.to(dev), rd.to(dev)
    N = ro.shape[0]
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1], device=dev)
    nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
    rb, rr = B._raymarching, R.get("_raymarching")
    ro16, rd16 = scene.camera_rays(1600, 1600)
    ro16, rd16 = ro16.to(dev), rd16.to(dev)
    N16 = ro16.shape[0]
    nears16, fars16 = torch.empty(N16, device=dev), torch.empty(N16, device=dev)
    add("near_far_from_aabb (1600x1600)", N16, 24 + 8, lambda: rb.near_far_from_aabb(ro16, rd16, aabb, N16, 0.2, nears16, fars16),
        None if rr is None else lambda: rr.near_far_from_aabb(ro16, rd16, aabb, N16, 0.2, nears16, fars16))
    rb.near_far_from_aabb(ro, rd, aabb, N, 0.2, nears, fars)
    Mm = 8 << 20
    xyzs, dirs2, deltas = torch.empty(Mm, 3, device=dev), torch.empty(Mm, 3, device=dev), torch.empty(Mm, 2, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    noises = torch.zeros(N, device=dev)

    def march(mod):
        counter.zero_()
        mod.march_rays_train(ro, rd, bf, 1.0, 0.0, 1024, 1024, N, 1, 128, Mm, nears, fars, xyzs, dirs2, deltas, rays, counter, noises)
    march(rb)
    n_samples = int(counter[0].item())
    add(f"march_rays_train (800x800 rays -> {n_samples} samples)", n_samples, 32 + 36 * N / n_samples, lambda: march(rb),
        None if rr is None else lambda: march(rr))
    cells = 8 * 128 ** 3
    grid = torch.rand(cells, device=dev)
    bits = torch.empty(cells // 8, dtype=torch.uint8, device=dev)
    add("packbits (8 cascades x 128^3)", cells, 4.125, lambda: rb.packbits(grid, cells // 8, 0.5, bits),
        None if rr is None else lambda: rr.packbits(grid, cells // 8, 0.5, bits))
    # the training-step size of march_rays_train (4,096 random rays of the frame)
    sel = torch.randperm(N, generator=g)[:4096].to(dev)
    ro4, rd4, n4, f4 = ro[sel].contiguous(), rd[sel].contiguous(), nears[sel].contiguous(), fars[sel].contiguous()
    rays4 = torch.empty(4096, 3, dtype=torch.int32, device=dev)

    def march4(mod):
        counter.zero_()
        mod.march_rays_train(ro4, rd4, bf, 1.0, 0.0, 1024, 1024, 4096, 1, 128, Mm, n4, f4, xyzs, dirs2, deltas, rays4, counter, noises)
    march4(rb)
    n4s = int(counter[0].item())
    add(f"march_rays_train (4,096 rays -> {n4s} samples)", n4s, 32 + 36 * 4096 / max(n4s, 1), lambda: march4(rb),
        None if rr is None else lambda: march4(rr))

    print(f"| operator | units | algorithmic B / unit | ours ms | ours GB/s (of {PEAK_GBPS:.0f} measured HBM) | reference kernel ms | speed-up |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for r in rows:
        ref = "n/a" if r["ref_ms"] is None else f"{r['ref_ms']:.3f}"
        sp = "" if r["speedup"] is None else f"{r['speedup']:.2f}x"
        print(f"| {r['op']} | {r['M']:,} | {r['bytes_per_unit']:.0f} | {r['ms']:.3f} | {r['GBps']:.0f} ({100 * r['frac_hbm']:.0f} %) | {ref} | {sp} |")
    print("\n```json\n" + json.dumps(rows) + "\n```")


if __name__ == "__main__":
    main()
