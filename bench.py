#!/usr/bin/env python
"""bench.py -- rays/sec of the ENVIDR volumetric-render hot path at 800x800 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--width 800]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full 800x800 frame (640,000 primary rays) of the synthetic toaster-dimension scene
(hash L16/C2/T19, sdf 32-64-64-15, env IDE(deg 5)-256-256-256-12, diffuse, colour, renv), rendered through
NeRFRenderer.render's inference path with use_renv + indir_ref (geometry pass, reflected secondary rays, main pass).
  value  : rays/s with the rays resident in HBM (CUDA events around the K steps, max over ranks)
  e2e    : rays/s through the public API with HOST buffers: rays copied from pinned host memory every step and the
           image copied back to pinned host memory inside the timed region
  N > 1  : weak scaling -- every rank renders one frame of a light-rotation sweep (env_rot = rank * 360/N degrees,
           BASELINE config 5) and one NCCL all-gather assembles the N frames on every rank inside the timed region
--impl reference: the oracle's CPU port of the same path (the reference has no CPU implementation of march / hash /
composite; its CUDA extensions cannot run without a GPU), all host threads, on a bounded strided sample of the
same rays; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE = 651_008            # toaster dims, forward (sdf + normal + env x2 + diffuse + colour)
FLOP_RENV_EXTRA = 18_432 + 12_160     # renv_net + second colour evaluation (main pass with r_images)
FLOP_GEOMETRY = 14_208 + 12_416 + 192  # sdf fwd + reverse pass + jacobian contraction (geometry-only pass)
FLOP_ENV = 610_304                    # env_net 72-256-256-256-12 evaluated twice per sample


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=800)
    ap.add_argument("--no-indir", action="store_true", help="single pass (BASELINE config 2 style)")
    ap.add_argument("--cpu-sample", type=int, default=12288, help="rays in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip timing the reference's own CUDA path (oracle/ref_cuda.py)")
    ap.add_argument("--no-train", action="store_true", help="skip the auxiliary train-step (BASELINE config 3) measurement")
    ap.add_argument("--sec-floor", type=int, default=None, help="RenderConfig.secondary_n_step_floor override (experiments)")
    ap.add_argument("--no-defer", action="store_true", help="shade inside the iterative loops (RenderConfig.defer_shading / defer_secondary_shading off)")
    ap.add_argument("--no-extra-warmup", action="store_true", help="profiling runs (ncu --launch-skip counts on exactly W warm-up frames)")
    ap.add_argument("--no-density", action="store_true", help="skip the auxiliary occupancy-grid update measurement (SURVEY 8 f-1)")
    ap.add_argument("--train-rays", type=int, default=4096)
    ap.add_argument("--sweep-width", type=int, default=1600, help="frame width of the relight sweep (BASELINE config 5)")
    ap.add_argument("--sweep-rotations", type=int, default=4, help="light rotations per timed sweep (5-degree steps of the 72-step sweep)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-5 measurements at N = 1")
    ap.add_argument("--config", default="toaster", choices=["toaster", "neus"], help="neus: BASELINE config 4 (NeuS geometry, no hash grid)")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"],
                    help="env_net arithmetic: tc = tcgen05 tensor cores with fp16 hi/lo split operands (default), fp32 = FFMA path")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (wall clock of the timed region); the sampler is started a little earlier because
        nvidia-smi needs a few hundred ms to produce its first line.  If the region was shorter than the sampling period, the
        samples nearest to it (taken under the same load: the untimed frames right before it) are used and counted as such."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= (t1 or t) + 0.15)]
        inside = len(rows)
        if not rows and self.rows:
            rows = [r for _, r in sorted(self.rows, key=lambda x: abs(x[0] - t0))[:3]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_inside_timed_region": inside}


def workload(args):
    from envidr_b200 import scene
    W = H = args.width
    fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
    bf = scene.make_bitfield()
    ro, rd = scene.camera_rays(W, H)
    return fp, bf, ro, rd, W, H


def cpu_baseline(fp, bf, ro, rd, args, indir, n_rays, rot=None):
    """Oracle (CPU port) on a strided sample of the same rays, all host threads.  Returns (rays/s, seconds, samples)."""
    import numpy as np
    import torch
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    N = ro.shape[0]
    stride = max(1, N // n_rays)
    sel = np.arange(0, N, stride)[:n_rays]
    P = fp.to_oracle()
    st = []
    t0 = time.perf_counter()
    O.render(P, ro.numpy()[sel], rd.numpy()[sel], bf, indir_ref=indir, bg_color=1.0, env_rot_radian=rot, dtype=torch.float32, stats=st)
    dt = time.perf_counter() - t0
    return len(sel) / dt, dt, sum(s["samples"] for s in st), len(sel)


def train_step_bench(fp_cpu, bf, ro, rd, dev, n_rays, steps=10, warmup=3):
    """BASELINE config 3 (auxiliary, not the headline): one training step forward + backward over `n_rays` random pixels of the
    frame through the run_cuda training branch on the library's CUDA operators (march_rays_train, hash_encode incl. second-order
    backward, composite_rays_train) with the toaster.ini loss terms; dense layers through cuBLAS fp32.  No optimizer step."""
    import torch
    from envidr_b200 import render, train
    g = torch.Generator().manual_seed(0)
    sel = torch.randperm(ro.shape[0], generator=g)[:n_rays]
    o, d = ro[sel].to(dev), rd[sel].to(dev)
    gt_rgb = torch.rand(n_rays, 3, generator=g).to(dev)
    gt_mask = (torch.rand(n_rays, generator=g) > 0.5).float().to(dev)
    r_img = torch.rand(n_rays, 4, generator=g).to(dev)
    field = train.TrainableField(fp_cpu.to(dev))
    bft = torch.from_numpy(bf).to(dev)
    cfg = render.RenderConfig()
    counter = torch.zeros(2, dtype=torch.int32, device=dev)

    def step(mean_count):
        for p in field.parameters():
            p.grad = None
        counter.zero_()
        out = train.render_train(field, bft, o, d, cfg, r_images=r_img, perturb=True, force_all_rays=mean_count <= 0,
                                 mean_count=mean_count, step_counter=counter)
        train.loss_epilogue(field, out, gt_rgb, gt_mask).backward()
        return out

    out = step(-1)                                   # first step sizes the static sample buffer, as the reference's mean_count does
    torch.cuda.synchronize()
    mean_count = int(out["xyzs"].shape[0]) - 128     # march_rays_train re-adds the alignment slack (raymarching.py:213-216)
    del out

    def timed(fn):
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in ev)[steps // 2]

    ms_eager = timed(lambda: step(mean_count))
    eager_samples = int(counter[0].item())
    graphed = train.GraphedTrainStep(field, bft, cfg, n_rays, mean_count)
    ms = timed(lambda: graphed(o, d, gt_rgb, gt_mask, r_img))
    counter = graphed.counter
    assert abs(int(counter[0].item()) - eager_samples) < 0.05 * eager_samples
    n_samples = int(counter[0].item())
    # the whole optimisation step: graph replay + Adam over the trainable tensors (hash table 12.2 M floats + sdf / env / renv MLPs),
    # fused (envidr_b200.optim.FusedAdam, one launch) and with torch.optim.Adam as the reference configures it (main_nerf.py:150)
    from envidr_b200.optim import FusedAdam
    params = [p for p in field.parameters() if p.requires_grad]
    res = {}
    for name, cls in (("fused_adam", FusedAdam), ("torch_adam", torch.optim.Adam)):
        try:
            opt = cls(params, lr=1e-4, betas=(0.9, 0.99), eps=1e-15)

            def full():
                graphed(o, d, gt_rgb, gt_mask, r_img)
                opt.step()
            res[f"ms_per_step_with_{name}"] = timed(full)
        except Exception as e:
            res[f"{name}_error"] = repr(e)[:160]
    # the same step with env_net as one launch per layer and direction (round-1 form: k_linear_tc x 8, IDE kernels, normalize, ...) instead of the
    # fused forward / backward kernels (envidr_b200/env_train.py), for the comparison in the same run
    try:
        field.fused_env = False
        graphed_pl = train.GraphedTrainStep(field, bft, cfg, n_rays, mean_count)
        res["ms_per_step_fwd_bwd_env_per_layer"] = timed(lambda: graphed_pl(o, d, gt_rgb, gt_mask, r_img))
    except Exception as e:
        res["env_per_layer_error"] = repr(e)[:160]
    finally:
        field.fused_env = True
    return {"rays": n_rays, "samples": n_samples, "ms_per_step_fwd_bwd": ms, "rays_per_sec": n_rays / (ms * 1e-3),
            "ms_per_step_fwd_bwd_eager": ms_eager, **res,
            "env_net": "one fused forward kernel (IDE -> layers -> unit norm, k_env_tc<SAVE>) + one fused backward kernel (k_env_bwd_tc) + one weight-gradient GEMM per layer",
            "mode": "CUDA graph replay of the captured step (train.GraphedTrainStep); host copies the step's inputs in",
            "what": "run_cuda train branch fwd+bwd (use_renv, r_images; colour L1 + mask BCE + Cauchy + eikonal through the fused loss "
                    "epilogue), fp32; ms_per_step_with_*: + optimizer step over all trainable tensors"}


def _timed_frames(fn, frames):
    import torch
    ms = []
    out = None
    for _ in range(frames):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2], out


def gpu_reference_bench(fp_cpu, bf, ro_d, rd_d, dev, indir, frames=2, train_rays=4096):
    """The reference's own CUDA path on this GPU: the "reference rays/sec on the same B200" of the north star.  Reported next to
    the contract's figures; the driver's reference arm stays `--impl reference`.

    Preferred: the REAL reference -- its NeRFNetwork under configs/scenes/toaster.ini (oracle/_ref/py, the unmodified Python tree
    shipped by oracle/build_ref.py) with the synthetic field loaded into its parameters, `model.render(...)` called with
    Trainer.eval_step's arguments (utils.py:857-859) on its own kernels (oracle/_ref/_*.so rebuilt for sm_100a).  On the same
    model object, the documented drop-ins are then timed: install() and install(patch_render=True), and
    one Trainer.train_step forward + backward on the reference's kernels (`train`: the comparator of the train_step key).
    Fallback when the tree was not shipped: oracle/ref_cuda.py (a restatement of the same host loop around the same kernels)."""
    import torch
    from oracle import ref_model as RM
    from oracle import ref_cuda
    N = ro_d.shape[0]
    if RM.available():
        from envidr_b200 import render
        RM.install_shims()
        model, opt = RM.build_model([], cuda_ray=True)
        RM.load_field(model, fp_cpu, bf)
        model.to(dev).eval()
        RM.use_backends("reference")
        opt.indir_ref = indir
        kw = RM.eval_kwargs(opt)
        fn = lambda: model.render(ro_d[None], rd_d[None], **kw)
        fn()                                                                     # warm-up (cuBLAS handles, allocator)
        torch.cuda.synchronize()
        t, out = _timed_frames(fn, frames)
        res = {"value": N / (t * 1e-3), "unit": "rays/s", "ms_per_frame": t, "image": out["image"].reshape(N, 3).detach(),
               "what": "the reference's own NeRFNetwork / NeRFRenderer.render / run_cuda (unmodified Python, toaster.ini, eval_step arguments) on its own "
                       "kernels rebuilt for sm_100a, synthetic field loaded into the model, same frame"}
        import nerf.render_func as RF
        import nerf.renderer as R
        dropin = {}
        try:
            RM.use_backends("envidr")
            for tag, kwargs in (("install(patch_render=False)", dict(patch_render=False)), ("install()", dict(patch_render=True, renderer_class=R.NeRFRenderer))):
                render.install(RF, **kwargs)
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                td, od = _timed_frames(fn, 5)
                e = (od["image"].reshape(N, 3) - res["image"]).abs().max(-1).values
                dropin[tag] = {"rays_per_sec": N / (td * 1e-3), "ms_per_frame": td, "pixels_over_1e-4_vs_reference": int((e > 1e-4).sum()),
                               "rgb_linf_max_vs_reference": float(e.max())}
        except Exception as e:
            dropin["error"] = repr(e)[:200]
        finally:
            render.uninstall(RF, R.NeRFRenderer)
            RM.use_backends("reference")
        res["dropin_on_the_reference_model"] = dropin
        try:                                                                     # Trainer.train_step on the reference's kernels
            g = torch.Generator().manual_seed(0)
            sel = torch.randperm(N, generator=g)[:train_rays].to(dev)
            o, d = ro_d[sel][None], rd_d[sel][None]
            images = torch.rand(1, train_rays, 4, generator=g).to(dev)
            images[..., 3] = (images[..., 3] > 0.5).float()
            opt.indir_ref = False
            opt.color_space, opt.alpha_bg_mode = "srgb", "white"
            opt.eikonal_loss = opt.cauchy_loss = opt.mask_loss = True
            opt.backsdf_loss = opt.relsdf_loss = opt.orientation_loss = opt.dist_bound = opt.diffuse_loss = False
            opt.eikonal_loss_weight, opt.cauchy_loss_weight, opt.entropy_loss_weight = 0.01, 0.001, 0
            model.train()
            def tstep():
                model.zero_grad(set_to_none=True)
                pred, gt, loss, ld = RM.train_step(model, opt, o, d, images)
                loss.backward()
                return loss
            model.mean_count, model.local_step = 0, 0
            model.step_counter.zero_()
            tstep()                                                              # sizes mean_count as the first steps of an epoch do
            torch.cuda.synchronize()
            model.mean_count = int(model.step_counter[0, 0].item())
            for _ in range(2):
                tstep()
            tt, _ = _timed_frames(tstep, 7)
            res["train"] = {"ms_per_step_fwd_bwd": tt, "rays": train_rays, "samples": int(model.step_counter[(model.local_step - 1) % 16, 0].item()),
                            "what": "Trainer.train_step (utils.py:560-808) forward + backward on the real reference model and kernels, "
                                    "single pass without r_images (the Trainer feeds r_images only from a dataset; our train_step key includes the renv "
                                    "branch, i.e. does more work per step), colour L1 + mask BCE + Cauchy + eikonal, no optimizer step"}
        except Exception as e:
            res["train"] = {"error": repr(e)[:300]}
        finally:
            model.eval()
        return res
    if not ref_cuda.available():
        return {"unavailable": "oracle/_ref/*.so not present"}
    F_ = ref_cuda.RefField(fp_cpu.to_oracle(), dev)
    bft = torch.from_numpy(bf).to(dev)
    fn = lambda: ref_cuda.render(F_, bft, ro_d, rd_d, indir_ref=indir, bg_color=1.0)
    fn()
    torch.cuda.synchronize()
    t, out = _timed_frames(fn, frames)
    return {"value": N / (t * 1e-3), "unit": "rays/s", "ms_per_frame": t, "image": out["image"],
            "what": "reference CUDA kernels (oracle/_ref, unmodified sources, sm_100a) + restated reference host loop + torch fp32 MLPs "
                    "(oracle/ref_cuda.py; the reference Python tree was not shipped), same frame"}


def _timed_region(fn, steps, world, dev, flush=None):
    """K steps bracketed by barrier + synchronize on both sides, CUDA events per step, MAX over ranks of the summed time (ms)."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for i, (a, b) in enumerate(ev):
        if flush is not None:
            flush.fill_(1)
        a.record()
        fn(i)
        b.record()
    barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def config5(args, fp, bft, dev, world, rank, steps, flush, e2e=True):
    """BASELINE config 5: toaster 1600x1600 relight sweep (env_rot 0..360 in 72 steps, utils.py:1297-1303), rays sharded over the ranks
    in interleaved 8x8-pixel tiles, one all-gather of the packed outputs per frame (envidr_b200.dist.render_sharded).
      frame : every rotation renders the FULL three-pass frame (what the reference does per rotation): the strong-scaling workload
      sweep : the rotation-independent passes (geometry, reflected-ray geometry) are computed once per camera and each rotation only
              shades (render.prepare_sweep / render_sweep_frame); timed over prepare + R rotations"""
    import numpy as np
    import torch
    from envidr_b200 import render, scene
    from envidr_b200 import dist as edist
    Wd = args.sweep_width
    ro, rd = scene.camera_rays(Wd, Wd)
    idx = edist.tile_shard_indices(Wd, Wd, rank, world)
    o_h, d_h = ro[idx].contiguous().pin_memory(), rd[idx].contiguous().pin_memory()
    o_s, d_s = o_h.to(dev), d_h.to(dev)
    N5 = Wd * Wd
    cfg = render.RenderConfig(indir_ref=True)
    rots = [2 * np.pi * k / 72 for k in range(72)]
    img_h = torch.empty(N5, 3).pin_memory()

    def frame(i, o=o_s, d=d_s):
        return edist.render_sharded(lambda a, b: render.render(fp, bft, a, b, cfg, bg_color=1.0, env_rot_radian=rots[i % 72], get_normal_image=True),
                                    o, d, Wd, Wd, presharded=True)

    def frame_e2e(i):
        # host rays of THIS rank's shard in, the assembled frame out to the host on rank 0 (the all-gather leaves the full frame on every rank;
        # one reader is what a caller does -- until the third session of round 2 every rank copied the whole frame back: 8 x 30.7 MB at N = 8)
        out = frame(i, o_h.to(dev, non_blocking=True), d_h.to(dev, non_blocking=True))
        if rank == 0:
            img_h.copy_(out["image"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    R = args.sweep_rotations

    def sweep(i):
        geom = render.prepare_sweep(fp, bft, o_s, d_s, cfg)
        for k in range(R):
            edist.render_sharded(lambda a, b: render.render_sweep_frame(fp, geom, cfg, rots[(i * R + k) % 72], bg_color=1.0), o_s, d_s, Wd, Wd,
                                 presharded=True)
    for i in range(3):
        frame(i)
    st = []
    render.render(fp, bft, o_s, d_s, cfg, bg_color=1.0, env_rot_radian=rots[0], stats=st)
    torch.cuda.synchronize()
    res = {"width": Wd, "rays_per_frame": N5, "rays_per_frame_per_rank": int(idx.numel()), "stats": st}
    res["frame_ms_total"] = _timed_region(frame, steps, world, dev, flush)
    if e2e:
        frame_e2e(0)
        res["frame_e2e_ms_total"] = _timed_region(frame_e2e, steps, world, dev, flush)
    sweep(0)
    n_sw = max(1, steps // 4)
    res["sweep_ms_total"] = _timed_region(sweep, n_sw, world, dev, flush)
    res["sweep_steps"], res["sweep_rotations"] = n_sw, R
    return res


def dp_train_bench(fp_cpu, bf, dev, world, rank, n_rays, steps):
    """Data-parallel training (optional): the 4,096-ray batch of BASELINE config 3 split over the ranks, every rank replays its
    captured forward + loss + backward on its share, ONE all-reduce over the flat gradient buffer of the replicated state (48.8 MB hash
    table + MLPs; envidr_b200.dist.allreduce_gradients), then the fused Adam step on every rank.  Time = max over ranks."""
    import torch
    from envidr_b200 import render, scene, train
    from envidr_b200 import dist as edist
    from envidr_b200.optim import FusedAdam
    n_local = n_rays // world
    ro, rd = scene.camera_rays(800, 800)
    g = torch.Generator().manual_seed(0)
    sel = torch.randperm(ro.shape[0], generator=g)[:n_rays][rank * n_local:(rank + 1) * n_local]
    o, d = ro[sel].to(dev), rd[sel].to(dev)
    gt_rgb = torch.rand(n_rays, 3, generator=g)[rank * n_local:(rank + 1) * n_local].to(dev)
    gt_mask = (torch.rand(n_rays, generator=g) > 0.5).float()[rank * n_local:(rank + 1) * n_local].to(dev)
    r_img = torch.rand(n_rays, 4, generator=g)[rank * n_local:(rank + 1) * n_local].to(dev)
    field = train.TrainableField(fp_cpu.to(dev))
    bft = torch.from_numpy(bf).to(dev)
    cfg = render.RenderConfig()
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    out = train.render_train(field, bft, o, d, cfg, r_images=r_img, perturb=True, force_all_rays=True, mean_count=-1, step_counter=counter)
    torch.cuda.synchronize()
    mean_count = int(out["xyzs"].shape[0]) - 128 + 2048              # head-room: the per-rank sample count varies with the noise
    del out
    graphed = train.GraphedTrainStep(field, bft, cfg, n_local, mean_count)
    params = [p for p in field.parameters() if p.requires_grad]
    opt = FusedAdam(params, lr=1e-4, betas=(0.9, 0.99), eps=1e-15)

    def local_only(i):
        graphed(o, d, gt_rgb, gt_mask, r_img)

    def full(i):
        graphed(o, d, gt_rgb, gt_mask, r_img)
        edist.allreduce_gradients(params)
        opt.step()
    for i in range(3):
        full(i)
    ms_local = _timed_region(local_only, steps, world, dev) / steps
    ms_full = _timed_region(full, steps, world, dev) / steps
    nbytes = sum(p.numel() for p in params) * 4
    return {"rays_global": n_local * world, "rays_per_rank": n_local, "ms_per_step_fwd_bwd": ms_local, "ms_per_step_with_allreduce_and_adam": ms_full,
            "rays_per_sec": n_local * world / (ms_full * 1e-3), "allreduce_bytes": nbytes,
            "what": "config-3 train step, batch split over the ranks; graph replay of forward + loss + backward, one NCCL all-reduce of the flat "
                    "gradient buffer, fused Adam; max over ranks"}


def run_multi(args, fp_cpu, bf, dev, world, rank, local):
    """N > 1: the timed step is one full 1600x1600 three-pass frame of the relight sweep (BASELINE config 5), its rays sharded over the N
    ranks; `value` = rays of the frame / time (max over ranks) -- strong scaling: the frame is the same whatever N is.  Extra keys:
    `config5_sweep` (the sweep with the rotation-independent passes shared between rotations) and `replica_frames` (round 1's weak
    form: one 800x800 frame per rank + all-gather of the N frames)."""
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    from envidr_b200 import _lib, render
    from envidr_b200 import dist as edist
    fp_cpu.precision = args.precision
    fp = fp_cpu.to(dev).pack()
    bft = torch.from_numpy(bf).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.lib()
    steps = args.steps
    # warm-up (>= 3 frames inside config5) + steady state, then the instrumented measurement
    warm = config5(args, fp, bft, dev, world, rank, max(3, args.warmup), flush, e2e=False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    lib.envidr_render_timing(1)
    l0 = lib.envidr_launch_count()
    t0 = time.time()
    c5 = config5(args, fp, bft, dev, world, rank, steps, flush)
    t1 = time.time()
    launches_all = int(lib.envidr_launch_count() - l0)
    lib.envidr_render_timing(0)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    # dominant kernel on this rank, timed alone over `steps` frames with the event hook on
    lib.envidr_render_timing(1)
    from envidr_b200 import scene
    Wd = args.sweep_width
    ro, rd = scene.camera_rays(Wd, Wd)
    idx = edist.tile_shard_indices(Wd, Wd, rank, world)
    o_s, d_s = ro[idx].to(dev), rd[idx].to(dev)
    cfg = render.RenderConfig(indir_ref=True)
    l1 = lib.envidr_launch_count()
    for i in range(steps):
        render.render(fp, bft, o_s, d_s, cfg, bg_color=1.0, env_rot_radian=0.1 * i)
    torch.cuda.synchronize()
    launches = int(lib.envidr_launch_count() - l1)
    fms, fl = ctypes.c_float(), ctypes.c_uint32()
    lib.envidr_render_field_time(ctypes.byref(fms), ctypes.byref(fl))
    lib.envidr_render_timing(0)
    # weak form of round 1 (one 800x800 frame of the sweep per rank)
    W8 = args.width
    ro8, rd8 = scene.camera_rays(W8, W8)
    ro8, rd8 = ro8.to(dev), rd8.to(dev)
    rot = 2 * np.pi * rank / world

    def replica(i):
        out = render.render(fp, bft, ro8, rd8, cfg, bg_color=1.0, env_rot_radian=rot, get_normal_image=True)
        edist.gather_frames(out["image"])
    for i in range(3):
        replica(i)
    rep_ms = _timed_region(replica, steps, world, dev, flush) / steps
    dp = None
    if not args.no_train:
        try:
            dp = dp_train_bench(fp_cpu, bf, dev, world, rank, args.train_rays, steps)
        except Exception as e:
            dp = {"error": repr(e)[:300]}
    if rank == 0:
        N5 = c5["rays_per_frame"]
        ms = c5["frame_ms_total"] / steps
        value = N5 / (ms * 1e-3)
        e2e_ms = c5["frame_e2e_ms_total"] / steps
        st = c5["stats"]
        shaded = st[1].get("shaded", st[1]["samples"]) + st[2]["samples"]
        flop_step = shaded * FLOP_ENV
        kernel_ms = fms.value / steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        ach = flop_step / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
        sw_ms = c5["sweep_ms_total"] / c5["sweep_steps"]
        R = c5["sweep_rotations"]
        line = {"metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (env_net: fp16 hi+lo split operands on tensor cores, fp32 accumulate)", "data": "synthetic",
                "config": dict(config_dict(args, Wd, Wd, True),
                               workload=f"BASELINE config 5: synthetic toaster-dims scene {Wd}x{Wd} relight sweep (env_rot 5-degree steps), use_renv + indir_ref (3 passes); "
                                        f"one step = one full frame, rays sharded over {world} ranks in interleaved 8x8-pixel tiles, one all-gather per frame",
                               rays_per_step_per_gpu=c5["rays_per_frame_per_rank"],
                               parallelism=f"ray sharding x{world} (interleaved 8x8 tiles) + one all_gather_into_tensor of 32 B/ray per frame",
                               n1_comparison="bench.py --gpus 1 renders the 800x800 frame of the metric; its `config5` key holds this workload on ONE GPU "
                                             "(the exact strong-scaling denominator)"),
                "samples_per_step_per_gpu": sum(s["samples"] for s in st),
                "e2e": {"value": N5 / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 2 * c5["rays_per_frame_per_rank"] * 12,
                        "d2h_bytes_per_step": N5 * 12, "d2h_on": "rank 0 (the assembled frame)", "ms_per_step": e2e_ms},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "kernel": "k_env_tc (rank 0's share of the frame)", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": ach / peak_tf, "traffic": None, "kernel_ms_per_step": kernel_ms, "kernel_launches_per_step": fl.value / steps,
                             "algorithmic_flop_per_step": flop_step, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback"},
                "cpu_baseline": None, "clocks": clocks,
                "config5_sweep": {"rays_per_sec": R * N5 / (sw_ms * 1e-3), "ms_per_sweep": sw_ms, "rotations": R, "ms_per_rotation": sw_ms / R,
                                  "what": "geometry pass + reflected-ray geometry once per camera, every rotation only shades (render.prepare_sweep / "
                                          "render_sweep_frame); sharded like the frames, one all-gather per rotation"},
                "dp_train_step": dp,
                "replica_frames": {"rays_per_sec": world * W8 * W8 / (rep_ms * 1e-3), "ms_per_step": rep_ms,
                                   "what": f"weak form (round 1): one {W8}x{W8} frame of the sweep per rank + all-gather of the {world} frames"}}
        print(json.dumps(line))
    dist.destroy_process_group()


def run_neus(args):
    """BASELINE config 4 (--config neus): 800x800 inference of the NeuS-style field (frequency encoding, 8 x 256 weight-normed Softplus
    layers with a skip connection, NeuS opacity, input_alpha compositing; no hash grid) through envidr_b200.neus_field.render_rays_neus.
    Same JSON contract; the dominant kernel is k_linear_tc (the dense layers of the geometry network, forward + reverse pass)."""
    import contextlib
    import numpy as np
    import torch
    from envidr_b200 import _lib, scene
    from envidr_b200 import neus_field as NF
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    W = H = args.width
    N = W * H
    nf_cpu = scene.make_neus_field(0)
    nf = nf_cpu.to(dev).pack()
    bf = scene.make_sphere_bitfield()
    bft = torch.from_numpy(bf).to(dev)
    ro, rd = scene.camera_rays(W, H)
    ro_d, rd_d = ro.to(dev), rd.to(dev)
    ro_h, rd_h = ro.pin_memory(), rd.pin_memory()
    img_h = torch.empty(N, 3).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.lib()
    st = {}

    def step(i, stats=None):
        return NF.render_rays_neus(nf, bft, ro_d, rd_d, bg_color=1.0, stats=stats)

    def step_e2e(i):
        out = NF.render_rays_neus(nf, bft, ro_h.to(dev, non_blocking=True), rd_h.to(dev, non_blocking=True), bg_color=1.0)
        img_h.copy_(out["image"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for i in range(max(3, args.warmup)):
        out = step(i, st)
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    time.sleep(0.5)
    NF.LINEAR_TIMING = []
    l0 = lib.envidr_launch_count()
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    launches = int(lib.envidr_launch_count() - l0)
    kt = sum(a.elapsed_time(b) for a, b, _ in NF.LINEAR_TIMING) / args.steps
    kflop = sum(f for _, _, f in NF.LINEAR_TIMING) / args.steps
    klaunch = len(NF.LINEAR_TIMING) / args.steps
    NF.LINEAR_TIMING = None
    t0 = time.time()
    ms = _timed_region(step, args.steps, 1, dev, flush) / args.steps
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    e2e_ms = _timed_region(step_e2e, args.steps, 1, dev, flush) / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    ach = kflop / (kt * 1e-3) / 1e12 if kt > 0 else 0.0
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import neus_oracle as NO
        torch.set_num_threads(os.cpu_count() or 1)
        n = max(256, args.cpu_sample // 8)
        sel = np.arange(0, N, max(1, N // n))[:n]
        ost = {}
        t = time.perf_counter()
        NO.render_rays(scene.neus_to_oracle(nf_cpu), ro.numpy()[sel], rd.numpy()[sel], bf, bg_color=1.0, dtype=torch.float32, stats=ost)
        dt = time.perf_counter() - t
        cpu = {"value": len(sel) / dt, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{len(sel)} rays (stride sample of the {W}x{H} frame, {ost['samples']} samples) in {dt:.1f} s; oracle/neus_oracle.py "
                         "(C march / composite + torch-CPU fp32 MLPs with autograd normals, all host threads)"}
    gref = None
    if not args.no_gpu_reference:
        try:
            from oracle import ref_model as RM
            if RM.available():
                with contextlib.redirect_stdout(sys.stderr):
                    RM.install_shims()
                    model, opt = RM.build_model(["--use_neus_sdf", "--encoding_pos", "frequency", "--multires", "6", "--geometric_init", "--num_layers", "8",
                                                 "--hidden_dim", "256", "--skip_layers", "4", "--init_variance", "0.6", "--geo_init_bias", "0.5"], cuda_ray=True)
                    with torch.no_grad():
                        for lin, (Wt, b) in zip(model.sdf_net, nf_cpu.sdf):
                            lin.weight_v.copy_(Wt); lin.weight_g.copy_(Wt.norm(dim=1, keepdim=True)); lin.bias.copy_(b)
                        for name in ("env", "diffuse", "color", "renv"):
                            for lin, (Wt, b) in zip(getattr(model, name + "_net"), getattr(nf_cpu.shading, name)):
                                lin.weight.copy_(Wt); lin.bias.copy_(b)
                        model.density_bitfield.copy_(torch.from_numpy(bf))
                    model.to(dev).eval()
                    RM.use_backends("reference")
                    opt.indir_ref = False
                    kw = RM.eval_kwargs(opt)
                    fn = lambda: model.render(ro_d[None], rd_d[None], **kw)
                    fn()
                    torch.cuda.synchronize()
                    t, o = _timed_frames(fn, 2)
                e = (o["image"].reshape(N, 3) - out["image"]).abs().max(-1).values
                gref = {"value": N / (t * 1e-3), "unit": "rays/s", "ms_per_frame": t, "speedup_ours_over_gpu_reference": (N / (ms * 1e-3)) / (N / (t * 1e-3)),
                        "rgb_linf_max_vs_ours": float(e.max()), "pixels_over_1e-4": int((e > 1e-4).sum()),
                        "what": "the reference's own NeRFNetwork (use_neus_sdf, frequency, geometric_init, 8 x 256, skip [4]) / NeRFRenderer.render / run_cuda on "
                                "its own kernels, same weights, same frame"}
        except Exception as e:
            gref = {"error": repr(e)[:300]}
    line = {"metric": "rays_per_sec", "value": N / (ms * 1e-3), "unit": "rays/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (dense layers: fp16 hi+lo split operands on tensor cores, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"BASELINE config 4: NeuS-style geometry without hash grid, {W}x{H} inference, single pass (materials.ini + use_neus_sdf, "
                                   "encoding_pos=frequency multires 6, geometric_init, 8 x 256 layers, skip_layers [4]; definition: SURVEY.md 8d)",
                       "scene": "geometric-init sphere SDF (radius 0.37-0.48), variance 0.6 (inv_s = 403), seeded rendering MLPs (env 256 / IDE degree 5)",
                       "rays_per_step_per_gpu": N, "schedule": "the reference's iterative schedule (n_step = N // n_alive <= 8): per iteration one fused geometry launch, "
                                                                "opacity, shading and one composite per image",
                       "cache": "L2 flushed between timed steps by writing a 256 MB buffer"},
            "samples_per_step_per_gpu": st.get("samples"), "march_iterations_per_step": st.get("iterations"), "shaded_samples_per_step": st.get("shaded"),
            "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 2 * N * 12, "d2h_bytes_per_step": N * 12, "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": ("k_neus_geom_tc (the whole geometry network per 128-sample tile: 8 forward + 7 reverse GEMMs, Softplus, skip, "
                                                     "frequency encoding and its transpose-Jacobian; fp16 hi/lo split, 3 MMAs per K step)") if nf.fused and nf.fused_supported()
                         else "k_linear_tc (dense layers of the geometry network, one launch per layer and direction)",
                         "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                         "kernel_ms_per_step": kt, "kernel_launches_per_step": klaunch, "kernel_share_of_step": kt / ms, "algorithmic_flop_per_step": kflop,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md)",
                         "note": "each launch timed with its own CUDA event pair in a separate untimed pass (events perturb the step)"},
            "cpu_baseline": cpu, "clocks": clocks, "gpu_reference": gref}
    print(json.dumps(line))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fp, bf, ro, rd, W, H = workload(args)
    indir = not args.no_indir
    cores = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, samples, n = cpu_baseline(fp, bf, ro, rd, args, indir, args.cpu_sample)      # the sample of the repo arm's cpu_baseline leg
        if i >= args.warmup:
            vals.append((v, dt, samples, n))
    v = sum(x[0] for x in vals) / len(vals)
    ms = 1e3 * sum(x[1] for x in vals) / len(vals)
    sample = f"{vals[0][3]} rays (stride sample of the {W}x{H} frame), {vals[0][2]} samples per step, torch-CPU fp32 MLPs + C march/hash/composite"
    print(json.dumps({
        "impl": "reference", "metric": "rays_per_sec", "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(config_dict(args, W, H, indir), schedule="the reference's iterative schedule in all passes "
                                                   "(n_step = N // n_alive <= 8, cuda_ray.py:287), CPU port (oracle.render)", parallelism="host threads"),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def config_dict(args, W, H, indir):
    return {"workload": f"synthetic toaster-dims scene {W}x{H} inference, " + ("use_renv + indir_ref (3 passes)" if indir else "1 pass (BASELINE config 2 shape)"),
            "rays_per_step_per_gpu": W * H, "hash": "L16 C2 base16 res2048 T2^19 (48.8 MB fp32)",
            "mlps": "sdf 32-64-64-15, env IDE72-256-256-256-12 x2, diffuse 24-32-3, color 28-64-64-3, renv 4-64-64-64-12",
            "max_steps": 1024, "T_thresh": 1e-4,
            "schedule": "passes 1-2: geometry-only iterative loops on the reference's schedule n_step = N // n_alive with the cap raised from 8 to 16 and a "
                        "floor of 8 in the reflected-ray pass (batching only: same composited samples, frame bit-identical to cap 8 / floor 1, "
                        "tests/test_gpu_render.py), shading deferred to one batch per pass; main pass: one batch over the per-ray sample counts found by "
                        "the geometry pass (RenderConfig.replay_main_pass)" if indir else "geometry-only iterative loop (reference schedule, n_step cap 16), then one shading batch over the composited samples "
                        "(RenderConfig.defer_shading)",
            "parallelism": f"ray/frame sharding x{args.gpus} + all_gather",
            "cache": "inputs larger than L2 are not needed: 48.8 MB table + per-iteration sample buffers are re-written each step; "
                     "L2 is flushed between timed steps by writing a 256 MB buffer"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "neus":
        if int(os.environ.get("RANK", "0")) == 0:
            run_neus(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from envidr_b200 import _lib, render
    from envidr_b200 import dist as edist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):  # the version banner goes to stdout, ahead of the one JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    fp_cpu, bf, ro, rd, W, H = workload(args)
    if world > 1:
        return run_multi(args, fp_cpu, bf, dev, world, rank, local)
    N = W * H
    indir = not args.no_indir
    fp_cpu.precision = args.precision
    fp = fp_cpu.to(dev).pack()
    bft = torch.from_numpy(bf).to(dev)
    rot = (2 * np.pi * rank / world) if world > 1 else None
    cfg = render.RenderConfig(indir_ref=indir, defer_shading=not args.no_defer, defer_secondary_shading=not args.no_defer)
    if args.sec_floor is not None:
        cfg.secondary_n_step_floor = args.sec_floor
    ro_d, rd_d = ro.to(dev), rd.to(dev)
    ro_h, rd_h = ro.pin_memory(), rd.pin_memory()
    img_h = torch.empty(N, 3).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.lib()

    def step_resident(stats=None):
        out = render.render(fp, bft, ro_d, rd_d, cfg, bg_color=1.0, env_rot_radian=rot, get_normal_image=True, stats=stats)
        frames = edist.gather_frames(out["image"])
        return out, frames

    def step_e2e():
        o = ro_h.to(dev, non_blocking=True); d = rd_h.to(dev, non_blocking=True)
        out = render.render(fp, bft, o, d, cfg, bg_color=1.0, env_rot_radian=rot, get_normal_image=True)
        edist.gather_frames(out["image"])
        img_h.copy_(out["image"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms = []

    def timed(fn, steps, instrument=False):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in ev:
            flush.fill_(1)                     # flush L2 between timed steps (outside the event bracket)
            a.record()
            fn()
            b.record()
        barrier()
        each = [a.elapsed_time(b) for a, b in ev]
        if instrument:
            step_ms.extend(each)
        ms = sum(each)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # A fresh box needs more than 3 frames (~50 ms) to reach its steady state: 60 more untimed frames (~1 s), so that the
    # resident-input region below is measured in the same state as the e2e region after it.
    barrier()
    warm_extra = 0 if args.no_extra_warmup else 60    # a fixed count: every rank must issue the same number of all-gathers
    for _ in range(warm_extra):
        step_resident()
    barrier()
    stats = []
    out, _ = step_resident(stats)
    torch.cuda.synchronize()
    samples_per_step = sum(s["samples"] for s in stats)
    iters_per_step = sum(s["iterations"] for s in stats)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)                               # nvidia-smi start-up; the GPU keeps no work queued meanwhile, frames follow
    lib.envidr_render_timing(1)                       # untimed pass with the kernel-timing hook on: creates its CUDA events
    for _ in range(args.steps):
        step_resident()
    barrier()
    l0 = lib.envidr_launch_count()
    lib.envidr_render_timing(1)                       # reset the hook's accumulators; the events now exist
    t_wall0 = time.time()
    total_ms = timed(step_resident, args.steps, instrument=True)
    t_wall1 = time.time()
    fms, fl = __import__("ctypes").c_float(), __import__("ctypes").c_uint32()
    lib.envidr_render_field_time(__import__("ctypes").byref(fms), __import__("ctypes").byref(fl))
    lib.envidr_render_timing(0)
    launches = int(lib.envidr_launch_count() - l0)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    sharded = None
    if world > 1:
        # strong scaling of ONE frame (auxiliary): interleaved 8x8-pixel tiles per rank + one all-gather of the packed outputs
        def step_sharded():
            return edist.render_sharded(lambda o, d: render.render(fp, bft, o, d, cfg, bg_color=1.0, get_normal_image=True), ro_d, rd_d, H, W)
        for _ in range(5):
            step_sharded()
        sh_ms = timed(step_sharded, args.steps) / args.steps
        sharded = {"ms_per_frame": sh_ms, "rays_per_sec": N / (sh_ms * 1e-3), "what": f"one {W}x{H} frame sharded over {world} ranks "
                   "(envidr_b200.dist.render_sharded), all-gather inside the timed region, max over ranks"}
    ms_per_step = total_ms / args.steps
    value = world * N / (ms_per_step * 1e-3)
    e2e_value = world * N / (e2e_ms / args.steps * 1e-3)

    # roofline of the dominant kernel: algorithmic FLOP per sample x samples, over its CUDA-event time.
    #   fp32 : k_field does everything          -> all FLOPs of the pass
    #   tc   : k_env_tc does the env_net passes -> 610,304 FLOP per shaded sample (2 x 152,576 MAC), geometry pass excluded
    tcp = args.precision == "tc"
    if indir:
        if tcp:
            # env_net runs on the SHADED samples: with deferred secondary shading those are the composited ones ("shaded")
            flop_step = (stats[1].get("shaded", stats[1]["samples"]) + stats[2]["samples"]) * FLOP_ENV
        else:
            flop_step = (stats[0]["samples"] * FLOP_GEOMETRY + stats[1]["samples"] * FLOP_PER_SAMPLE
                         + stats[2]["samples"] * (FLOP_PER_SAMPLE + FLOP_RENV_EXTRA))
    else:
        flop_step = stats[0].get("shaded", stats[0]["samples"]) * (FLOP_ENV if tcp else FLOP_PER_SAMPLE)
    field_ms_per_step = fms.value / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved_tf = flop_step / (field_ms_per_step * 1e-3) / 1e12 if field_ms_per_step > 0 else 0.0
    kname = ("k_env_tc (IDE + env_net x2 on tcgen05, fp16 hi/lo split operands: 3 MMAs per K step, fp32 accumulate in TMEM)" if tcp else
             "k_field (fused per-sample field: hash gather + SDF + normal + IDE + env/diffuse/colour MLPs, fp32 FFMA)")
    traffic = None                                   # dram bytes of one captured launch of the dominant kernel (ncu --set full)
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_env_tc" if tcp else "k_field"]
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        if tcp and t.get("samples") and fl.value:
            # the capture is one launch over t["samples"] samples; per launch of THIS step = bytes / sample x shaded samples / launches
            traffic_per_sample = traffic / t["samples"]
            traffic = traffic_per_sample * (flop_step / FLOP_ENV) / (fl.value / args.steps)
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": kname,
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md)",
                "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram bytes per sample of a 1 M-sample launch) scaled to this "
                                                    "step's average launch", "kernel_ms_per_step": field_ms_per_step, "kernel_launches_per_step": fl.value / args.steps,
                "kernel_share_of_step": field_ms_per_step / ms_per_step, "algorithmic_flop_per_step": flop_step,
                "executed_frac": (3.0 if tcp else 1.0) * achieved_tf / peak_tf,
                "arithmetic": ("tcgen05.mma kind::f16, 3 MMAs per K step (hi*hi + lo*hi + hi*lo): the tensor pipe executes 3x the "
                               "algorithmic FLOPs counted here") if tcp else "fp32 FFMA (exact path)"}
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, dt, s, n = cpu_baseline(fp_cpu, bf, ro, rd, args, indir, args.cpu_sample, rot)
            cpu = {"value": v, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{n} rays (stride sample of the {W}x{H} frame, {s} samples) in {dt:.1f} s; oracle CPU port "
                             f"(C march/hash/composite + torch-CPU fp32 MLPs, all host threads)"}
        gref = None
        if world == 1 and not args.no_gpu_reference:
            try:
                import contextlib
                with contextlib.redirect_stdout(sys.stderr):          # the reference prints while it parses options / builds the model
                    gref = gpu_reference_bench(fp_cpu, bf, ro_d, rd_d, dev, indir)
                if "image" in gref:
                    gimg = gref.pop("image")
                    e = (gimg - out["image"]).abs().max(-1).values
                    gref["rgb_linf_vs_ours_p99999"] = float(torch.quantile(e[e > 0], 0.99999)) if int((e > 0).sum()) else 0.0
                    gref["pixels_over_1e-4"] = int((e > 1e-4).sum())
                    mse = float(((gimg - out["image"]) ** 2).mean())
                    gref["psnr_ours_vs_reference_image_db"] = float(-10.0 * np.log10(max(mse, 1e-20)))     # PSNRMeter formula (utils.py:296-306)
                    gref["speedup_ours_over_gpu_reference"] = value / gref["value"]
            except Exception as e:                      # auxiliary: never take the headline line down with it
                gref = {"error": repr(e)[:200]}
        trn = None
        if world == 1 and not args.no_train:
            try:
                trn = train_step_bench(fp_cpu, bf, ro, rd, dev, args.train_rays)
            except Exception as e:                      # auxiliary: never take the headline line down with it
                trn = {"error": repr(e)[:200]}
        dens = None
        if world == 1 and not args.no_density:
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("density_bench", os.path.join(ROOT, "profiles", "density_bench.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                prec = fp_cpu.precision
                dens = mod.measure(dev, fp_cpu, with_reference=not args.no_gpu_reference)
                fp_cpu.precision = prec
                dens["what"] = ("NeRFRenderer.update_extra_state (renderer.py:264-352) on the 128^3 grid: ours = envidr_density_grid_update; "
                                "reference = its own kernels + torch ops + mean().item() (oracle/ref_cuda.update_extra_state)")
            except Exception as e:                      # auxiliary: never take the headline line down with it
                dens = {"error": repr(e)[:200]}
        poses = None
        try:                                              # 8 camera poses theta = 0, 45, ..., 315 (the headline is theta = 40)
            if args.no_sweep:
                raise RuntimeError("skipped (--no-sweep)")
            from envidr_b200 import scene as _scene
            per = {}
            for th in range(0, 360, 45):
                o8, d8 = _scene.camera_rays(W, H, theta_deg=float(th))
                o8, d8 = o8.to(dev), d8.to(dev)
                fn8 = lambda i: render.render(fp, bft, o8, d8, cfg, bg_color=1.0, get_normal_image=True)
                fn8(0)
                st8 = []
                render.render(fp, bft, o8, d8, cfg, bg_color=1.0, stats=st8)
                per[str(th)] = {"ms_per_frame": _timed_region(fn8, 3, 1, dev, flush) / 3, "samples": sum(s["samples"] for s in st8)}
            v = [N / (p["ms_per_frame"] * 1e-3) for p in per.values()]
            poses = {"rays_per_sec_mean": sum(v) / len(v), "rays_per_sec_min": min(v), "rays_per_sec_max": max(v), "per_theta_deg": per,
                     "what": "the same frame from the 8 camera poses of SURVEY 8d (theta = 0..315 step 45, phi = -30, radius 4), 3 timed frames each"}
        except Exception as e:
            poses = {"error": repr(e)[:200]}
        fitted = None
        if not args.no_sweep and indir and tcp:
            # SURVEY 8d / VERDICT r1 item 9: the same frame on a scene whose geometry was FITTED through the library's operators (Adam on the
            # analytic SDF through hash_encode + sdf_net, then 1 + 16 update_extra_state calls for the occupancy bit field) instead of constructed
            try:
                from envidr_b200 import scene as _scene
                import time as _time
                t0 = _time.time()
                ffp, fbits, finfo = _scene.fit_synthetic_field(0, device=dev, steps=1000, hidden_dim_env=256, ide_degree=5)
                torch.cuda.synchronize()
                finfo["fit_seconds"] = _time.time() - t0
                per = {}
                for th in (40, 130, 220, 310):
                    of, df = _scene.camera_rays(W, H, theta_deg=float(th))
                    of, df = of.to(dev), df.to(dev)
                    fnf = lambda i: render.render(ffp, fbits, of, df, cfg, bg_color=1.0, get_normal_image=True)
                    fnf(0)
                    stf = []
                    render.render(ffp, fbits, of, df, cfg, bg_color=1.0, stats=stf)
                    per[str(th)] = {"ms_per_frame": _timed_region(fnf, 3, 1, dev, flush) / 3, "samples": [s_["samples"] for s_ in stf]}
                v = [N / (p_["ms_per_frame"] * 1e-3) for p_ in per.values()]
                fitted = {"rays_per_sec_mean": sum(v) / len(v), "rays_per_sec_min": min(v), "rays_per_sec_max": max(v), "per_theta_deg": per, "fit": finfo,
                          "what": "the three-pass 800x800 frame on the FITTED synthetic toaster (envidr_b200.scene.fit_synthetic_field: hash table + sdf_net "
                                  "trained to the analytic SDF with Adam through hash_encode forward / backward, occupancy from 17 update_extra_state calls), "
                                  "4 poses, 3 timed frames each; the headline stays on the constructed scene the parity tests are written against"}
            except Exception as e:
                fitted = {"error": repr(e)[:300]}
        c5 = None
        if not args.no_sweep and indir and tcp:
            try:
                n5 = max(2, args.steps // 2)
                r5 = config5(args, fp, bft, dev, 1, 0, n5, flush, e2e=False)
                N5 = r5["rays_per_frame"]
                c5 = {"frame": {"rays_per_sec": N5 / (r5["frame_ms_total"] / n5 * 1e-3), "ms_per_frame": r5["frame_ms_total"] / n5},
                      "sweep": {"rays_per_sec": r5["sweep_rotations"] * N5 / (r5["sweep_ms_total"] / r5["sweep_steps"] * 1e-3),
                                "ms_per_rotation": r5["sweep_ms_total"] / r5["sweep_steps"] / r5["sweep_rotations"], "rotations": r5["sweep_rotations"]},
                      "width": r5["width"], "samples": [s["samples"] for s in r5["stats"]],
                      "what": "BASELINE config 5 on ONE GPU (the workload bench.py --gpus N > 1 shards): frame = full three-pass 1600x1600 frame per "
                              "light rotation; sweep = rotation-independent passes shared between the rotations (render.prepare_sweep)"}
            except Exception as e:
                c5 = {"error": repr(e)[:300]}
        line = {"metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "warmup_extra_steps": warm_extra, "ms_each_step": [round(x, 3) for x in step_ms],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (env_net: fp16 hi+lo split operands on tensor cores, fp32 accumulate)" if tcp else "f32",
                "data": "synthetic", "config": config_dict(args, W, H, indir),
                "samples_per_sec": world * samples_per_step / (ms_per_step * 1e-3), "samples_per_step_per_gpu": samples_per_step,
                "march_iterations_per_step": iters_per_step,
                "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": 2 * N * 12, "d2h_bytes_per_step": N * 12,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "train_step": trn, "gpu_reference": gref, "density_update": dens, "config5": c5, "poses": poses, "fitted_scene": fitted}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
