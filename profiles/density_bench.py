#!/usr/bin/env python
"""Timing of one occupancy-grid update on the GPU: ours (csrc/density.cu through envidr_b200.density) against
the reference's own path (oracle/ref_cuda.update_extra_state: its kernels + torch ops + the mean().item() sync), same field,
CUDA events, median of `reps`.  Also mark_untrained_grid for 100 cameras.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from envidr_b200 import density, scene

H = 128
ALG_BYTES_PER_CELL = 12 + 12 + 12 + 1024 + 4 + 4 + 4 + 4 + 4 + 0.125   # noise, xyz w/r, hash gather, tmp w/r, grid r/w, pack read, bit


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def measure(dev, fp_cpu, with_reference=True):
    out = {}
    noise = torch.rand(1, H ** 3, 3, device=dev)
    for prec in ("tc", "fp32"):
        fp_cpu.precision = prec
        fp = fp_cpu.to(dev).pack()
        st = density.DensityGrid(device=dev)
        out[f"full_update_ms_{prec}"] = timed(lambda: st.update_extra_state(fp, full_update=True, noise=noise, sync=False))
        out[f"full_update_ms_{prec}_with_rand_and_item"] = timed(lambda: st.update_extra_state(fp, full_update=True))
        st.iter_density = 16
        out[f"partial_update_ms_{prec}"] = timed(lambda: st.update_extra_state(fp))
    poses = np.stack([scene.nerf_matrix_to_ngp(scene.pose_spherical(th, -30.0, 4.0), scale=0.65) for th in np.linspace(0, 360, 100, endpoint=False)])
    intr = scene.intrinsics_from_fov(800, 800, 0.69)
    poses_d = torch.from_numpy(poses).to(dev)
    out["mark_untrained_ms_100_poses"] = timed(lambda: st.mark_untrained_grid(poses_d, intr))
    cells = H ** 3
    out["cells"] = cells
    out["algorithmic_bytes_per_cell"] = ALG_BYTES_PER_CELL
    out["full_update_GBps_tc"] = cells * ALG_BYTES_PER_CELL / (out["full_update_ms_tc"] * 1e-3) / 1e9
    try:
        from oracle import ref_cuda
        if with_reference and ref_cuda.available():
            F_ = ref_cuda.RefField(fp_cpu.to_oracle(), dev)
            grid_r = torch.zeros(1, H ** 3, device=dev)
            bits_r = torch.zeros(H ** 3 // 8, dtype=torch.uint8, device=dev)
            out["reference_full_update_ms"] = timed(lambda: ref_cuda.update_extra_state(F_, grid_r, bits_r, iter_density=0), reps=10)
            out["reference_partial_update_ms"] = timed(lambda: ref_cuda.update_extra_state(F_, grid_r, bits_r, iter_density=16), reps=10)
            out["speedup_full"] = out["reference_full_update_ms"] / out["full_update_ms_tc_with_rand_and_item"]
            out["speedup_partial"] = out["reference_partial_update_ms"] / out["partial_update_ms_tc"]
    except Exception as e:
        out["reference_error"] = repr(e)[:200]
    return out


def main():
    print(json.dumps(measure(torch.device("cuda:0"), scene.make_synthetic_field(0))))


if __name__ == "__main__":
    main()
