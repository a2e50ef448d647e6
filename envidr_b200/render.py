"""Inference rendering on top of the fused CUDA loop (envidr_render_rays).

Mirrors the reference entry points of this part of the hot path:
  * run_cuda(model, rays_o, rays_d, **kwargs)   nerf/render_func/cuda_ray.py:15-364 (inference branch :238-359)
  * NeRFRenderer.render(...)                    nerf/renderer.py:364-531 (1 pass, or the 3-pass indir_ref scheme)
`install(render_func_module)` assigns `nerf.render_func.run_cuda = run_cuda` -- the reference looks the
function up as a module attribute at call time (renderer.py:368-369), so main_nerf.py needs no change.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream
from .field import FieldParams

SQRT3 = 3 ** 0.5


@dataclasses.dataclass
class RenderConfig:
    bound: float = 1.0
    cascade: int = 1
    grid_size: int = 128
    min_near: float = 0.2
    dt_gamma: float = 0.0
    max_steps: int = 1024
    T_thresh: float = 1e-4
    aabb: Optional[Sequence[float]] = None
    input_alpha: bool = False
    # indirect-reflection passes (renderer.py:439-513)
    indir_ref: bool = False
    indir_max_steps: int = 1024
    obj_aabb: Optional[Sequence[float]] = None
    # main pass of the indirect-reflection scheme as ONE batch over the sample counts the geometry pass found (render_rays_replay)
    # instead of re-discovering ray termination iteration by iteration; False = the reference's iterative schedule
    replay_main_pass: bool = True
    # with replay_main_pass and the tensor-core field: the geometry pass also logs its per-sample geometry records, and the main
    # pass shades those (env_net + shading heads) instead of marching and evaluating the hash grid + sdf_net a second time
    reuse_geometry: bool = True
    # n_step floor of the secondary (reflected-ray) pass: it runs over few rays, so the reference schedule n_step = N // n_alive
    # starts at 1 and issues ~50 launches of <= 71 k samples; a floor only changes the batching (same composited samples).  1 = reference.
    # 8 since the third session of round 2 (was 4): with deferred shading the image is bit-identical whatever the floor (run r3_22), and every
    # iteration costs ~40 us of latency-bound march / composite launches and kernel set-up that does not shrink with the ray count
    secondary_n_step_floor: int = 8
    # secondary pass with deferred shading (tensor-core field): its iterative loop runs geometry-only (ray termination depends on
    # the density alone) while logging the per-sample records, then env_net + the shading heads run ONCE over all composited
    # samples (envidr_field_forward_records) and envidr_composite_rays_replay composites -- the main pass's scheme, applied to
    # the reflected rays.  Same composited samples and per-sample math as the iterative loop; False = shade inside the loop
    defer_secondary_shading: bool = True
    # the same scheme for the single-pass render (indir_ref = False): geometry-only loop, then one shading batch + one compositing launch
    defer_shading: bool = True
    # cap of n_step in the geometry-only LOGGED passes of the batched schedule (the passes above): the reference caps n_step = N // n_alive
    # at 8 (cuda_ray.py:287), so the long tail of a pass -- a few thousand grazing rays -- runs as dozens of tiny iterations of 3 launches
    # each; 16 halves their number.  Batching only: the logged passes are geometry-only, shading and compositing run afterwards over the
    # composited samples, and the frame is BIT-IDENTICAL for caps 8 and 16 (run r3_22: max |d image| = 0; +3.8 % marched samples past ray
    # termination).  Measured: 800x800 frame 15.92 -> 15.47 ms together with the secondary floor of 8 (20 + 23 -> 14 + 12 iterations); one
    # rank's share of the 1600x1600 frame at 8 GPUs 9.43 -> 9.03 ms.  Passes on the reference schedule (replay_main_pass=False, or any
    # pass that shades inside the loop) always use the reference's 8; logged_n_step_cap=8, secondary_n_step_floor=1 reproduce its counts
    logged_n_step_cap: int = 16

    def aabb6(self):
        return list(self.aabb) if self.aabb is not None else [-self.bound] * 3 + [self.bound] * 3


_workspaces: Dict[tuple, torch.Tensor] = {}
_log_need: Dict[int, int] = {}        # rays of a frame -> samples its geometry pass marched last time (sizes the sample log)


def _workspace(N: int, device, n_step_floor: int = 1) -> torch.Tensor:
    key = (str(device), int(N), int(n_step_floor))
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = lib().envidr_render_workspace_bytes_ex(N, n_step_floor)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        if len(_workspaces) >= 4:    # the three passes of a frame alternate between a few sizes: keep those alive
            _workspaces.pop(next(iter(_workspaces)))
        _workspaces[key] = ws
    return ws


def render_rays(field: FieldParams, bitfield: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, cfg: RenderConfig, *,
                bg_color=1.0, r_images: Optional[torch.Tensor] = None, geometry_only: bool = False,
                env_rot_radian: Optional[float] = None, get_normal_image: bool = True, visual_items: Sequence[str] = (),
                perturb: bool = False, max_steps: Optional[int] = None, min_near: Optional[float] = None,
                sample_count: bool = False, n_step_floor: int = 1, log: Optional["SampleLogBuffers"] = None,
                n_step_cap: int = 8) -> Dict[str, torch.Tensor]:
    """One run_cuda inference pass over N rays.  Returns image [N,3], depth [N], weights_sum [N] and
    (optionally) normal_image / diffuse_image / specular_image / roughness_image, all on the device."""
    if field._packed is None:
        field.pack()
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    N, dev = rays_o.shape[0], rays_o.device
    f32 = dict(dtype=torch.float32, device=dev)
    out = _lib.RenderOut()
    res: Dict[str, torch.Tensor] = {}
    if N == 0:                                       # empty ray set (e.g. no pixel of a frame qualifies for the reflected pass): empty outputs
        res = {"depth": torch.empty(0, **f32), "weights_sum": torch.empty(0, **f32), "normal_image": torch.empty(0, 3, **f32)}
        if not geometry_only:
            res["image"] = torch.empty(0, 3, **f32)
            for k, c in (("diffuse", 3), ("specular", 3), ("roughness", 1)):
                if k in visual_items:
                    res[f"{k}_image"] = torch.empty(0, c, **f32)
        if sample_count or log is not None:
            res["sample_count"] = torch.empty(0, dtype=torch.int32, device=dev)
        return res
    res["depth"] = torch.empty(N, **f32)
    res["weights_sum"] = torch.empty(N, **f32)
    if not geometry_only:
        res["image"] = torch.empty(N, 3, **f32)
    if get_normal_image or geometry_only:
        res["normal_image"] = torch.empty(N, 3, **f32)
    if not geometry_only:
        if "diffuse" in visual_items:
            res["diffuse_image"] = torch.empty(N, 3, **f32)
        if "specular" in visual_items:
            res["specular_image"] = torch.empty(N, 3, **f32)
        if "roughness" in visual_items or "specular" in visual_items:
            res["roughness_image"] = torch.empty(N, **f32)
    if sample_count or log is not None:
        res["sample_count"] = torch.empty(N, dtype=torch.int32, device=dev)
    for k, t in res.items():
        setattr(out, k, t.data_ptr())
    if log is not None:
        assert geometry_only and field.precision == "tc", "sample log: geometry-only passes of the tensor-core field"
        log_struct = log.cstruct()
        out.log = ctypes.addressof(log_struct)
    opts = _lib.RenderOpts()
    opts.bound, opts.dt_gamma, opts.T_thresh = cfg.bound, cfg.dt_gamma, cfg.T_thresh
    opts.min_near = cfg.min_near if min_near is None else min_near
    opts.max_steps = cfg.max_steps if max_steps is None else max_steps
    opts.cascade, opts.grid_size = cfg.cascade, cfg.grid_size
    for i, v in enumerate(cfg.aabb6()):
        opts.aabb[i] = float(v)
    bg_t = None
    if torch.is_tensor(bg_color):
        if bg_color.numel() == 3:
            bgc = [float(v) for v in bg_color.reshape(-1).tolist()]
        else:
            bg_t = bg_color.float().contiguous().view(-1, 3)
            assert bg_t.shape[0] == N
            bgc = [0.0, 0.0, 0.0]
    elif isinstance(bg_color, (int, float)):
        bgc = [float(bg_color)] * 3
    else:
        bgc = [float(v) for v in bg_color]
    for i in range(3):
        opts.bg_color[i] = bgc[i]
    opts.geometry_only = int(geometry_only)
    opts.input_alpha = int(cfg.input_alpha)
    opts.n_step_floor = max(1, min(8, int(n_step_floor)))
    opts.n_step_cap = max(8, min(16, int(n_step_cap)))
    if r_images is not None:
        r_images = r_images.float().contiguous().view(-1, 4)
        assert r_images.shape[0] == N
    noises = torch.rand(N, **f32) if perturb else None
    ws = _workspace(N, dev, opts.n_step_floor)
    base = ws.data_ptr()
    aligned = (base + 255) // 256 * 256
    f = field.cstruct(env_rot_radian)
    check(lib().envidr_render_rays(ctypes.byref(f), ptr(bitfield), ptr(rays_o), ptr(rays_d), ptr(r_images), ptr(noises), ptr(bg_t), N,
                                   ctypes.byref(opts), ctypes.byref(out), ctypes.c_void_p(aligned), ws.numel() - (aligned - base),
                                   stream()), "render_rays")
    if "roughness_image" in res:
        res["roughness_image"] = res["roughness_image"][..., None]
    return res


class SampleLogBuffers:
    """Device buffers of an envidr_sample_log (one entry per marched sample of a geometry-only pass)."""

    def __init__(self, capacity: int, device):
        f32 = dict(dtype=torch.float32, device=device)
        self.capacity = int(capacity)
        self.rec = torch.empty(self.capacity, 32, **f32)
        self.sigma = torch.empty(self.capacity, **f32)
        self.delta = torch.empty(self.capacity, 2, **f32)
        self.ray = torch.empty(self.capacity, dtype=torch.int32, device=device)
        self.seq = torch.empty(self.capacity, dtype=torch.int32, device=device)

    def cstruct(self) -> _lib.SampleLog:
        s = _lib.SampleLog()
        s.rec, s.sigma, s.delta, s.ray, s.seq = (t.data_ptr() for t in (self.rec, self.sigma, self.delta, self.ray, self.seq))
        s.capacity = self.capacity
        return s


_logs: Dict[str, SampleLogBuffers] = {}


def _sample_log(device, need: int, slot: str = "primary") -> SampleLogBuffers:
    """One log per device and pass, grown geometrically (a frame's geometry pass marches a few million samples: 140 B each)."""
    key = f"{device}/{slot}"
    lg = _logs.get(key)
    if lg is None or lg.capacity < need:
        lg = SampleLogBuffers(max(need, 1 << 20), device)
        _logs[key] = lg
    return lg


def _compact(mask: torch.Tensor, n: int) -> torch.Tensor:
    """Indices of the n set entries of a boolean mask, ascending, WITHOUT a host synchronisation (mask.nonzero() has to read its output
    size back; here the caller already knows n from one combined read-back): exclusive scan + scatter."""
    N = mask.shape[0]
    pos = torch.cumsum(mask, 0) - 1
    out = torch.empty(n + 1, dtype=torch.int64, device=mask.device)
    out.scatter_(0, torch.where(mask, pos, n), torch.arange(N, dtype=torch.int64, device=mask.device))
    return out[:n]


def prepare_from_log(log: SampleLogBuffers, total: int, counts: torch.Tensor, ray_idx: Optional[torch.Tensor], M: Optional[int] = None,
                     indexed: bool = False) -> Dict:
    """Ray-contiguous copy of the logged samples of the rays `ray_idx` (int64 indices into the logged pass's rays, ascending; None =
    all of them): envidr_permute_sample_log.  Nothing here depends on colour, r_images or the light rotation, so one prepared batch
    serves every frame of a relight sweep.  M = number of samples of the selected rays if the caller already knows it (render() reads
    it back together with the pass statistics); otherwise one host synchronisation (the batch size sizes the buffers).
    indexed: leave the 128-byte records in the log and return the log position of every ray-ordered sample instead
    (envidr_permute_sample_log_index; the shading kernels then read rec[ridx[m]]): for a batch that is shaded ONCE (a frame) this saves
    reading and writing 256 B per sample; a relight sweep, which shades the same batch for every rotation, keeps the copy."""
    dev = counts.device
    f32 = dict(dtype=torch.float32, device=dev)
    n_all = counts.shape[0]
    cs = (counts if ray_idx is None else counts[ray_idx]).to(torch.int64)
    n_r = int(cs.shape[0])
    incl = torch.cumsum(cs, 0)
    if M is None:
        M = int(incl[-1].item()) if n_r else 0
    off = (incl - cs).to(torch.int32)
    if ray_idx is None:
        ray_off = off
    else:
        ray_off = torch.full((n_all,), -1, dtype=torch.int32, device=dev)
        ray_off[ray_idx] = off
    sigma, delta = torch.empty(M, **f32), torch.empty(M, 2, **f32)
    lstruct = log.cstruct()
    ridx = None
    if M == 0:                                       # no composited sample among the selected rays: nothing to gather
        rec = torch.empty(0, 32, **f32)
    elif indexed and total < 2 ** 31:
        rec = log.rec
        ridx = torch.empty(M, dtype=torch.int32, device=dev)
        check(lib().envidr_permute_sample_log_index(ctypes.byref(lstruct), total, ptr(ray_off), ptr(ridx), ptr(sigma), ptr(delta), stream()),
              "permute_sample_log_index")
    else:
        rec = torch.empty(M, 32, **f32)
        check(lib().envidr_permute_sample_log(ctypes.byref(lstruct), total, ptr(ray_off), ptr(rec), ptr(sigma), ptr(delta), stream()),
              "permute_sample_log")
    # rays of the pass in selected order: (ray id = position among the selected rays, offset, count)
    rays = torch.stack([torch.arange(n_r, dtype=torch.int32, device=dev), off, cs.to(torch.int32)], -1).contiguous()
    return dict(rec=rec, ridx=ridx, sigma=sigma, delta=delta, rays=rays, n_r=n_r, M=M)


def shade_prepared(field: FieldParams, prep: Dict, cfg: RenderConfig, *, bg_color=0.0, r_images: Optional[torch.Tensor] = None,
                   visual_items: Sequence[str] = (), env_rot_radian: Optional[float] = None, rec_unrotated: bool = False) -> Dict[str, torch.Tensor]:
    """env_net + shading heads over a prepared batch (envidr_field_forward_records) and the inference compositor over it
    (envidr_composite_rays_replay).  r_images [n_r, 4] in the order of the prepared rays.  rec_unrotated: the records were captured
    without a light rotation and `env_rot_radian` is applied inside the env_net kernel."""
    rec, sigma, delta, rays, n_r, M = (prep[k] for k in ("rec", "sigma", "delta", "rays", "n_r", "M"))
    dev = rec.device
    f32 = dict(dtype=torch.float32, device=dev)
    if M == 0:                                       # rays without a single composited sample: empty images over the background
        bg0 = bg_color if torch.is_tensor(bg_color) else torch.tensor(bg_color, **f32)
        res = {"image": torch.zeros(n_r, 3, **f32) + bg0, "depth": torch.zeros(n_r, **f32), "weights_sum": torch.zeros(n_r, **f32), "_samples": 0}
        for k, c in (("diffuse", 3), ("specular", 3), ("roughness", 1)):
            if k in visual_items or (k == "roughness" and "specular" in visual_items):
                res[f"{k}_image"] = torch.zeros(n_r, c, **f32)
        return res
    r_s = None
    if r_images is not None:
        r_s = torch.empty(M, 4, **f32)
        check(lib().envidr_scatter_ray_rows4(ptr(rays), n_r, M, ptr(r_images.float().contiguous().view(-1, 4)), ptr(r_s), stream()),
              "scatter_ray_rows4")
    want = {"rgb": torch.empty(M, 3, **f32)}
    if "diffuse" in visual_items:
        want["c_diffuse"] = torch.empty(M, 3, **f32)
    if "specular" in visual_items:
        want["c_specular"] = torch.empty(M, 3, **f32)
    ridx = prep.get("ridx")
    rough = None
    if "roughness" in visual_items or "specular" in visual_items:
        rough = (rec[:, 20] if ridx is None else rec[ridx.long(), 20]).contiguous()
    if field._scratch is None or field._scratch.numel() < 32 * (M + 2):
        field._scratch = torch.empty(32 * (M + 2), **f32)
    fo = _lib.FieldOut()
    for k, t in want.items():
        setattr(fo, k, t.data_ptr())
    f = field.cstruct(env_rot_radian if rec_unrotated else None, rec_unrotated=rec_unrotated)
    check(lib().envidr_field_forward_records_indexed(ctypes.byref(f), ptr(rec), ptr(ridx), ptr(r_s), M, ctypes.byref(fo), stream()),
          "field_forward_records")
    res = {"image": torch.empty(n_r, 3, **f32), "depth": torch.empty(n_r, **f32), "weights_sum": torch.empty(n_r, **f32)}
    if "c_diffuse" in want:
        res["diffuse_image"] = torch.empty(n_r, 3, **f32)
    if "c_specular" in want:
        res["specular_image"] = torch.empty(n_r, 3, **f32)
    if rough is not None:
        res["roughness_image"] = torch.empty(n_r, **f32)
    check(lib().envidr_composite_rays_replay(ptr(sigma), ptr(want["rgb"]), None, ptr(want.get("c_diffuse")), ptr(want.get("c_specular")),
                                             ptr(rough), ptr(delta), ptr(rays), None, M, n_r, cfg.T_thresh, int(cfg.input_alpha),
                                             ptr(res["weights_sum"]), ptr(res["depth"]), ptr(res["image"]), None, ptr(res.get("diffuse_image")),
                                             ptr(res.get("specular_image")), ptr(res.get("roughness_image")), stream()), "composite_rays_replay")
    bg = bg_color if torch.is_tensor(bg_color) else torch.tensor(float(bg_color), **f32) if not isinstance(bg_color, (list, tuple)) \
        else torch.tensor([float(v) for v in bg_color], **f32)
    res["image"] = res["image"] + (1 - res["weights_sum"]).unsqueeze(-1) * bg
    if "roughness_image" in res:
        res["roughness_image"] = res["roughness_image"][..., None]
    res["_samples"] = M
    return res


def render_rays_from_log(field: FieldParams, log: SampleLogBuffers, total: int, counts: torch.Tensor, ray_idx: Optional[torch.Tensor],
                         cfg: RenderConfig, *, bg_color=0.0, r_images: Optional[torch.Tensor] = None,
                         visual_items: Sequence[str] = (), M: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """A pass shaded from a geometry-only pass's sample log: gather ray by ray (prepare_from_log), shade from the geometry records and
    composite (shade_prepared).  counts [N_all] = samples composited per ray of the logged pass; r_images [n_selected, 4] in the order
    of ray_idx.  Returns per-selected-ray images."""
    return shade_prepared(field, prepare_from_log(log, total, counts, ray_idx, M, indexed=True), cfg, bg_color=bg_color, r_images=r_images,
                          visual_items=visual_items)


_aabb_cache: Dict[tuple, torch.Tensor] = {}
REPLAY_CHUNK = 1 << 22          # samples per field launch in render_rays_replay (bounds the tensor-core scratch: 256 B / sample = 1 GB)


def render_rays_replay(field: FieldParams, bitfield: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, counts: torch.Tensor,
                       cfg: RenderConfig, *, bg_color=1.0, r_images: Optional[torch.Tensor] = None,
                       env_rot_radian: Optional[float] = None, visual_items: Sequence[str] = (), max_steps: Optional[int] = None,
                       min_near: Optional[float] = None) -> Dict[str, torch.Tensor]:
    """One inference pass over rays whose per-ray sample counts are known (`counts`, from an earlier pass over the same
    rays with `sample_count=True`): march all samples in one launch, evaluate the field on the whole batch, composite with
    the inference compositor's arithmetic.  Same samples and the same per-sample math as the iterative loop; what changes is
    the batching (one pass over ~3 M samples instead of ~60 iterations of <= 81 k), so sample positions can differ from the
    iterative schedule by the ulp-level re-synchronisation of rays_t at iteration boundaries."""
    from . import raymarching as rm
    if field._packed is None:
        field.pack()
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    N, dev = rays_o.shape[0], rays_o.device
    f32 = dict(dtype=torch.float32, device=dev)
    key = (tuple(float(v) for v in cfg.aabb6()), str(dev))
    if key not in _aabb_cache:
        _aabb_cache[key] = torch.tensor(key[0], **f32)
    nears, fars = rm.near_far_from_aabb(rays_o, rays_d, _aabb_cache[key], cfg.min_near if min_near is None else min_near)
    counts = counts.to(torch.int32).contiguous()
    M = int(counts.sum().item()) if N else 0
    xyzs, dirs, deltas = torch.empty(M, 3, **f32), torch.empty(M, 3, **f32), torch.empty(M, 2, **f32)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    check(lib().envidr_march_rays_replay(ptr(rays_o), ptr(rays_d), ptr(bitfield), cfg.bound, cfg.dt_gamma,
                                         cfg.max_steps if max_steps is None else max_steps, N, cfg.cascade, cfg.grid_size, M, ptr(nears),
                                         ptr(fars), ptr(counts), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(counter), stream()),
          "march_rays_replay")
    r_s = None
    if r_images is not None:
        r_s = torch.empty(M, 4, **f32)
        check(lib().envidr_scatter_ray_rows4(ptr(rays), N, M, ptr(r_images.float().contiguous().view(-1, 4)), ptr(r_s), stream()),
              "scatter_ray_rows4")
    want = ["sigma", "rgb"]
    if "diffuse" in visual_items:
        want.append("c_diffuse")
    if "specular" in visual_items:
        want.append("c_specular")
    if "roughness" in visual_items or "specular" in visual_items:
        want.append("roughness")
    parts = []
    for a in range(0, M, REPLAY_CHUNK):
        b = min(M, a + REPLAY_CHUNK)
        parts.append(field.forward(xyzs[a:b], dirs[a:b], None if r_s is None else r_s[a:b], env_rot_radian=env_rot_radian, want=tuple(want)))
    so = {k: (torch.cat([p[k] for p in parts]) if len(parts) != 1 else parts[0][k]) for k in want} if parts else \
        {k: torch.empty((0, 3) if k in ("rgb", "c_diffuse", "c_specular") else (0,), **f32) for k in want}
    res = {"image": torch.empty(N, 3, **f32), "depth": torch.empty(N, **f32), "weights_sum": torch.empty(N, **f32)}
    if "c_diffuse" in so:
        res["diffuse_image"] = torch.empty(N, 3, **f32)
    if "c_specular" in so:
        res["specular_image"] = torch.empty(N, 3, **f32)
    if "roughness" in so:
        res["roughness_image"] = torch.empty(N, **f32)
    check(lib().envidr_composite_rays_replay(ptr(so["sigma"]), ptr(so["rgb"]), None, ptr(so.get("c_diffuse")), ptr(so.get("c_specular")),
                                             ptr(so.get("roughness")), ptr(deltas), ptr(rays), ptr(nears), M, N, cfg.T_thresh,
                                             int(cfg.input_alpha), ptr(res["weights_sum"]), ptr(res["depth"]), ptr(res["image"]), None,
                                             ptr(res.get("diffuse_image")), ptr(res.get("specular_image")), ptr(res.get("roughness_image")),
                                             stream()), "composite_rays_replay")
    bg = bg_color if torch.is_tensor(bg_color) else torch.tensor(float(bg_color), **f32) if not isinstance(bg_color, (list, tuple)) \
        else torch.tensor([float(v) for v in bg_color], **f32)
    res["image"] = res["image"] + (1 - res["weights_sum"]).unsqueeze(-1) * bg
    if "roughness_image" in res:
        res["roughness_image"] = res["roughness_image"][..., None]
    res["_samples"] = M
    return res


def last_stats() -> Dict[str, int]:
    """Blocks until the last render_rays finished; {'iterations', 'samples'} of that pass."""
    st = (ctypes.c_uint32 * 4)()
    check(lib().envidr_render_last_stats(st), "render_last_stats")
    return dict(iterations=int(st[0]), samples=int(st[1]) | (int(st[2]) << 32))


def reflect_dir(w_o: torch.Tensor, normals: torch.Tensor) -> torch.Tensor:
    """renderer.py:20-39."""
    return 2 * torch.sum(w_o * normals, dim=-1, keepdim=True) * normals - w_o


def render(field: FieldParams, bitfield: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, cfg: RenderConfig, *, bg_color=1.0,
           get_normal_image: bool = True, env_rot_radian: Optional[float] = None, visual_items: Sequence[str] = (),
           r_images: Optional[torch.Tensor] = None, stats: Optional[list] = None) -> Dict[str, torch.Tensor]:
    """NeRFRenderer.render for the cuda_ray inference path (renderer.py:364-531): rays [N,3] -> dict with image,
    depth, weights_sum, normal_image (+ visual items).  With cfg.indir_ref the three passes of renderer.py:439-513
    are run: geometry only -> reflected secondary rays -> main pass with r_images."""
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    N = rays_o.shape[0]
    if not cfg.indir_ref:
        results = None
        if cfg.defer_shading and field.precision == "tc" and N > 0:
            # single pass with deferred shading: geometry-only loop (logging the per-sample records), then ONE env_net + heads
            # batch over the composited samples and one compositing launch -- the scheme of the 3-pass path's main pass
            log = _sample_log(rays_o.device, _log_need.get(("one", N), 8 * 1 << 20), "primary")
            geo = render_rays(field, bitfield, rays_o, rays_d, cfg, geometry_only=True, env_rot_radian=env_rot_radian,
                              sample_count=True, log=log, n_step_cap=cfg.logged_n_step_cap)
            m_all = int(geo["sample_count"].sum().item())                  # the ONE host synchronisation of the pass: composited samples ...
            st = last_stats()                                              # ... after which the pass statistics are already on the host
            _log_need[("one", N)] = int(st["samples"] * 1.25) + 4096
            if st["samples"] <= log.capacity:
                main = render_rays_from_log(field, log, st["samples"], geo["sample_count"], None, cfg, bg_color=bg_color, r_images=r_images,
                                            visual_items=visual_items, M=m_all)
                st = dict(st, shaded=int(main.pop("_samples")))
                results = dict(main, depth=geo["depth"], normal_image=geo["normal_image"])
                if stats is not None:
                    stats.append(st)
        if results is None:
            results = render_rays(field, bitfield, rays_o, rays_d, cfg, bg_color=bg_color, r_images=r_images, env_rot_radian=env_rot_radian,
                                  get_normal_image=get_normal_image, visual_items=visual_items)
            if stats is not None:
                stats.append(last_stats())
    else:
        dt = 2 * SQRT3 / cfg.indir_max_steps
        reuse = cfg.replay_main_pass and cfg.reuse_geometry and field.precision == "tc"
        log = _sample_log(rays_o.device, _log_need.get(N, 8 * 1 << 20)) if reuse else None
        geo = render_rays(field, bitfield, rays_o, rays_d, cfg, geometry_only=True, env_rot_radian=env_rot_radian,
                          sample_count=cfg.replay_main_pass, log=log, n_step_cap=cfg.logged_n_step_cap if log is not None else 8)
        normals = geo["normal_image"]
        depth = geo["depth"] - dt
        weights_sum = geo["weights_sum"]
        ref_mask = (depth != 0) & (weights_sum > 0.9)
        ray_mask = (depth != 0) & (weights_sum > 0.3)
        ref_o = rays_o + depth[:, None] * rays_d
        ref_d = reflect_dir(-rays_d, normals)
        if cfg.obj_aabb is not None:
            ob = torch.tensor(cfg.obj_aabb, dtype=torch.float32, device=rays_o.device)
            ref_mask = ref_mask & (ref_o > ob[:3]).all(-1) & (ref_o < ob[3:]).all(-1)
        # The reference indexes with the boolean masks at every use (each one a host synchronisation); here the two index lists are
        # built once and every later gather / scatter uses them.  ref_mask is a subset of ray_mask (ws > 0.9 vs > 0.3).
        # ONE read-back for everything the host has to know after the geometry pass (third session of round 2; it used to be four:
        # the pass statistics, the two nonzero() sizes and the main pass's sample count): sizes of the two index lists and the number of
        # logged samples of the main-pass rays; the pass statistics are on the host once this returns.
        if "sample_count" in geo:
            m_main_t = (geo["sample_count"].to(torch.int64) * ray_mask).sum()
        else:
            m_main_t = ray_mask.sum() * 0
        n_ref, n_ray, m_main = (int(v) for v in torch.stack([ref_mask.sum(), ray_mask.sum(), m_main_t]).tolist())
        geo_stats = last_stats() if (stats is not None or reuse) else None
        if stats is not None:
            stats.append(geo_stats)
        if reuse:
            _log_need[N] = int(geo_stats["samples"] * 1.25) + 4096          # size the log for the next frame of this shape
            if geo_stats["samples"] > log.capacity:                          # overflowed: this frame marches the main pass again
                reuse = False
        ref_idx = _compact(ref_mask, n_ref)
        ray_idx = _compact(ray_mask, n_ray)
        pos_in_ray = torch.cumsum(ray_mask, 0) - 1                       # position of a ray among the main-pass rays
        sec_o, sec_d = ref_o[ref_idx], ref_d[ref_idx]
        n_sec = sec_o.shape[0]
        ref = None
        if cfg.defer_secondary_shading and field.precision == "tc" and n_sec > 0:
            log2 = _sample_log(rays_o.device, _log_need.get(("sec", N), 4 * 1 << 20), "secondary")
            geo2 = render_rays(field, bitfield, sec_o, sec_d, cfg, geometry_only=True, env_rot_radian=env_rot_radian,
                               max_steps=cfg.indir_max_steps, min_near=dt * 2, n_step_floor=cfg.secondary_n_step_floor,
                               sample_count=True, log=log2, n_step_cap=cfg.logged_n_step_cap)
            m_sec = int(geo2["sample_count"].sum().item())                  # the pass's one read-back; its statistics are on the host afterwards
            st2 = last_stats()
            _log_need[("sec", N)] = int(st2["samples"] * 1.25) + 4096
            if st2["samples"] <= log2.capacity:
                ref = render_rays_from_log(field, log2, st2["samples"], geo2["sample_count"], None, cfg, bg_color=0.0, M=m_sec)
                st2 = dict(st2, shaded=int(ref.pop("_samples")))
            if stats is not None and ref is not None:
                stats.append(st2)
        if ref is None:                                                     # shading inside the loop (reference-shaped)
            ref = render_rays(field, bitfield, sec_o, sec_d, cfg, bg_color=0.0, env_rot_radian=env_rot_radian,
                              get_normal_image=False, max_steps=cfg.indir_max_steps, min_near=dt * 2,
                              n_step_floor=cfg.secondary_n_step_floor)
            if stats is not None:
                stats.append(last_stats() if n_sec > 0 else dict(iterations=0, samples=0))
        ref_image = torch.cat([ref["image"], ref["weights_sum"][:, None]], -1)
        r_img = ref_image.new_zeros(ray_idx.shape[0], 4)
        r_img[pos_in_ray[ref_idx]] = ref_image                              # == r_img[ref_mask[ray_mask]] = ref_image (renderer.py:484-486)
        if n_ray == 0:                                                      # nothing on screen: the main pass has no ray
            main = {"image": normals.new_zeros(0, 3), "weights_sum": normals.new_zeros(0)}
            if stats is not None:
                stats.append(dict(iterations=0, samples=0))
        elif reuse:
            main = render_rays_from_log(field, log, geo_stats["samples"], geo["sample_count"], ray_idx, cfg, bg_color=0.0, r_images=r_img,
                                        visual_items=visual_items, M=m_main)
            if stats is not None:
                stats.append(dict(iterations=1, samples=int(main.pop("_samples"))))
            else:
                main.pop("_samples")
        elif cfg.replay_main_pass:
            # the main pass visits the rays of the geometry pass again (same origins, same density): replay its sample counts
            main = render_rays_replay(field, bitfield, rays_o[ray_idx], rays_d[ray_idx], geo["sample_count"][ray_idx], cfg, bg_color=0.0,
                                      r_images=r_img, env_rot_radian=env_rot_radian, visual_items=visual_items)
            if stats is not None:
                stats.append(dict(iterations=1, samples=int(main.pop("_samples"))))
        else:
            main = render_rays(field, bitfield, rays_o[ray_idx], rays_d[ray_idx], cfg, bg_color=0.0, r_images=r_img,
                               env_rot_radian=env_rot_radian, get_normal_image=get_normal_image, visual_items=visual_items)
            if stats is not None:
                stats.append(last_stats())
        results = {"normal_image": normals, "depth": depth}
        for k in ("image", "specular_image", "diffuse_image", "roughness_image"):
            if k in main:
                v = normals.new_zeros(N, main[k].shape[-1])
                v[ray_idx] = main[k]
                results[k] = v
        ws_full = normals.new_zeros(N)
        ws_full[ray_idx] = main["weights_sum"]
        bg = bg_color if torch.is_tensor(bg_color) else torch.tensor(bg_color, dtype=torch.float32, device=rays_o.device)
        results["image"] = (torch.zeros_like(normals) + bg) * (1 - ws_full[:, None]) + results["image"]
        results["weights_sum"] = ws_full
    if "weights_sum" in results and get_normal_image and "normal_image" in results:
        w = results["weights_sum"][..., None]
        results["normal_image"] = results["normal_image"] * w + (1 - w)
    return results


# ---------------------------------------------------------------------------------------------
# relight sweep (BASELINE config 5; Trainer.test with env_rot_degree_range, nerf/utils.py:1297-1303, 1325-1328)
# ---------------------------------------------------------------------------------------------

@dataclasses.dataclass
class SweepGeometry:
    """Everything of a three-pass frame that does not depend on the light rotation: the geometry pass (depth, normals, masks, the
    per-sample geometry records of the main-pass rays) and the geometry of the reflected secondary pass.  The reference re-renders all
    three passes for each of the 72 rotations of a sweep; here they are computed once per camera and every rotation only runs
    env_net + the shading heads + the compositor over the prepared records (the rotation is applied inside the env_net kernel)."""
    N: int
    normals: torch.Tensor
    depth: torch.Tensor
    ray_idx: torch.Tensor
    ref_idx: torch.Tensor
    pos_in_ray: torch.Tensor
    main: Dict
    secondary: Optional[Dict]
    samples: Dict[str, int]


def prepare_sweep(field: FieldParams, bitfield: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, cfg: RenderConfig) -> SweepGeometry:
    """Rotation-independent part of `render(..., cfg.indir_ref=True)` for the tensor-core field (renderer.py:439-486)."""
    assert field.precision == "tc", "the relight sweep reuses the tensor-core path's geometry records"
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    N = rays_o.shape[0]
    dt = 2 * SQRT3 / cfg.indir_max_steps
    log = _sample_log(rays_o.device, _log_need.get(N, 8 * 1 << 20), "sweep-primary")
    while True:
        geo = render_rays(field, bitfield, rays_o, rays_d, cfg, geometry_only=True, sample_count=True, log=log, n_step_cap=cfg.logged_n_step_cap)
        st = last_stats()
        if st["samples"] <= log.capacity:
            break
        log = _sample_log(rays_o.device, int(st["samples"] * 1.25) + 4096, "sweep-primary")
    _log_need[N] = int(st["samples"] * 1.25) + 4096
    normals, depth, ws = geo["normal_image"], geo["depth"] - dt, geo["weights_sum"]
    ref_mask = (depth != 0) & (ws > 0.9)
    ray_mask = (depth != 0) & (ws > 0.3)
    ref_o = rays_o + depth[:, None] * rays_d
    ref_d = reflect_dir(-rays_d, normals)
    if cfg.obj_aabb is not None:
        ob = torch.tensor(cfg.obj_aabb, dtype=torch.float32, device=rays_o.device)
        ref_mask = ref_mask & (ref_o > ob[:3]).all(-1) & (ref_o < ob[3:]).all(-1)
    ref_idx = ref_mask.nonzero().squeeze(-1)
    ray_idx = ray_mask.nonzero().squeeze(-1)
    pos_in_ray = torch.cumsum(ray_mask, 0) - 1
    main = prepare_from_log(log, st["samples"], geo["sample_count"], ray_idx)
    samples = dict(geometry=int(st["samples"]), main=int(main["M"]))
    sec = None
    if ref_idx.numel() > 0:
        sec_o, sec_d = ref_o[ref_idx], ref_d[ref_idx]
        log2 = _sample_log(rays_o.device, _log_need.get(("sec", N), 4 * 1 << 20), "sweep-secondary")
        while True:
            geo2 = render_rays(field, bitfield, sec_o, sec_d, cfg, geometry_only=True, max_steps=cfg.indir_max_steps, min_near=dt * 2,
                               n_step_floor=cfg.secondary_n_step_floor, sample_count=True, log=log2, n_step_cap=cfg.logged_n_step_cap)
            st2 = last_stats()
            if st2["samples"] <= log2.capacity:
                break
            log2 = _sample_log(rays_o.device, int(st2["samples"] * 1.25) + 4096, "sweep-secondary")
        _log_need[("sec", N)] = int(st2["samples"] * 1.25) + 4096
        sec = prepare_from_log(log2, st2["samples"], geo2["sample_count"], None)
        samples.update(secondary_marched=int(st2["samples"]), secondary=int(sec["M"]))
    return SweepGeometry(N=N, normals=normals, depth=depth, ray_idx=ray_idx, ref_idx=ref_idx, pos_in_ray=pos_in_ray, main=main, secondary=sec,
                         samples=samples)


def render_sweep_frame(field: FieldParams, geom: SweepGeometry, cfg: RenderConfig, env_rot_radian: Optional[float], *, bg_color=1.0,
                       get_normal_image: bool = True, visual_items: Sequence[str] = ()) -> Dict[str, torch.Tensor]:
    """One frame of the sweep: the three-pass frame `render()` returns for this light rotation, from the shared geometry."""
    dev = geom.normals.device
    n_main = int(geom.ray_idx.shape[0])
    r_img = geom.normals.new_zeros(n_main, 4)
    if geom.secondary is not None:
        ref = shade_prepared(field, geom.secondary, cfg, bg_color=0.0, env_rot_radian=env_rot_radian, rec_unrotated=True)
        r_img[geom.pos_in_ray[geom.ref_idx]] = torch.cat([ref["image"], ref["weights_sum"][:, None]], -1)
    main = shade_prepared(field, geom.main, cfg, bg_color=0.0, r_images=r_img, visual_items=visual_items, env_rot_radian=env_rot_radian,
                          rec_unrotated=True)
    N = geom.N
    results = {"normal_image": geom.normals, "depth": geom.depth}
    for k in ("image", "specular_image", "diffuse_image", "roughness_image"):
        if k in main:
            v = geom.normals.new_zeros(N, main[k].shape[-1])
            v[geom.ray_idx] = main[k]
            results[k] = v
    ws_full = geom.normals.new_zeros(N)
    ws_full[geom.ray_idx] = main["weights_sum"]
    bg = bg_color if torch.is_tensor(bg_color) else torch.tensor(bg_color, dtype=torch.float32, device=dev)
    results["image"] = (torch.zeros_like(geom.normals) + bg) * (1 - ws_full[:, None]) + results["image"]
    results["weights_sum"] = ws_full
    if get_normal_image:
        w = ws_full[..., None]
        results["normal_image"] = geom.normals * w + (1 - w)
    return results


# ---------------------------------------------------------------------------------------------
# drop-in for the reference model
# ---------------------------------------------------------------------------------------------

_reference_run_cuda = None
_dropin_precision = "tc"          # install(precision=...): "tc" (tcgen05, the measured path) or "fp32" (exact FFMA path, 13x slower)


def _model_field(model) -> FieldParams:
    """FieldParams of a reference model, re-read and re-packed on EVERY call: the reference Trainer changes the weights between
    evaluations without telling anybody (optimizer steps, ema.store / copy_to through `.data`, load_state_dict, env swaps), tensor
    version counters do not see `.data` writes, and a stale image would silently render old MLPs with the current hash grid.
    Re-packing costs one `beta.item()` and one launch over < 1 MB of weights; only the device buffers are kept between calls."""
    field = FieldParams.from_reference_model(model, precision=_dropin_precision)
    prev = getattr(model, "_envidr_field", None)
    if prev is not None and prev.device == field.device:
        field._packed, field._scratch = prev._packed, prev._scratch          # buffers only; pack() rewrites the image
    field.pack()
    model._envidr_field = field
    return field


def run_cuda(model, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024, T_thresh=1e-4,
             get_normal_image=False, use_specular_color=True, early_stop_steps=-1, ray_depth=None, main_pass=True, r_images=None,
             geometry_only=False, grad_ray=False, bg_sphere=True, env_rot_radian=None, **kwargs):
    """Replacement for nerf.render_func.run_cuda.  The inference branch (cuda_ray.py:238-359) runs in the fused
    loop; training / debug / ray_depth / background-sphere calls are forwarded to the reference implementation
    (which then reaches our kernels through the operator-level `_backend` modules)."""
    fused_ok = not model.training and not model.opt.debug and ray_depth is None and not (model.bg_radius > 0 and bg_sphere)
    if fused_ok and model.opt.use_neus_sdf:
        # BASELINE config 4 family (NeuS geometry, frequency encoding): envidr_b200.neus_field; other NeuS variants -> reference code
        from .neus_field import NeusField, render_rays_neus
        try:
            nf = NeusField.from_reference_model(model)
        except _lib.EnvidrError:
            nf = None
        if nf is not None and not perturb:
            prefix = rays_o.shape[:-1]
            res = render_rays_neus(nf, model.density_bitfield, rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), bound=float(model.bound),
                                   cascade=int(model.cascade), grid_size=int(model.grid_size), min_near=float(model.min_near), dt_gamma=float(dt_gamma),
                                   max_steps=int(max_steps), T_thresh=float(T_thresh), bg_color=1.0 if bg_color is None else bg_color,
                                   aabb=[float(v) for v in model.aabb_infer.tolist()], env_rot_radian=env_rot_radian,
                                   r_images=None if r_images is None else r_images.reshape(-1, 4), geometry_only=geometry_only,
                                   visual_items=tuple(model.opt.visual_items) if model.opt.use_diffuse else ())
            results = {"depth": res["depth"].view(*prefix), "weights_sum": res["weights_sum"].view(*prefix),
                       "image": None if geometry_only else res["image"].view(*prefix, 3), "normal_image": res["normal_image"].view(*prefix, 3)}
            for k in ("diffuse_image", "specular_image", "roughness_image"):
                if k in res:
                    results[k] = res[k]
            return results
        fused_ok = False
    if not fused_ok:
        if _reference_run_cuda is None:
            raise _lib.EnvidrError("run_cuda: this call needs the reference training branch; call install() first")
        return _reference_run_cuda(model, rays_o, rays_d, dt_gamma=dt_gamma, bg_color=bg_color, perturb=perturb,
                                   force_all_rays=force_all_rays, max_steps=max_steps, T_thresh=T_thresh,
                                   get_normal_image=get_normal_image, use_specular_color=use_specular_color,
                                   early_stop_steps=early_stop_steps, ray_depth=ray_depth, main_pass=main_pass, r_images=r_images,
                                   geometry_only=geometry_only, grad_ray=grad_ray, bg_sphere=bg_sphere, env_rot_radian=env_rot_radian,
                                   **kwargs)
    prefix = rays_o.shape[:-1]
    field = _model_field(model)
    cfg = RenderConfig(bound=float(model.bound), cascade=int(model.cascade), grid_size=int(model.grid_size), min_near=float(model.min_near),
                       dt_gamma=float(dt_gamma), max_steps=int(max_steps), T_thresh=float(T_thresh),
                       aabb=[float(v) for v in model.aabb_infer.tolist()])
    res = render_rays(field, model.density_bitfield, rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), cfg,
                      bg_color=1.0 if bg_color is None else bg_color, r_images=None if r_images is None else r_images.reshape(-1, 4),
                      geometry_only=geometry_only, env_rot_radian=env_rot_radian, get_normal_image=get_normal_image or geometry_only,
                      visual_items=tuple(model.opt.visual_items) if model.opt.use_diffuse else (), perturb=bool(perturb))
    results = {"depth": res["depth"].view(*prefix), "weights_sum": res["weights_sum"].view(*prefix),
               "image": None if geometry_only else res["image"].view(*prefix, 3)}
    if "normal_image" in res:
        results["normal_image"] = res["normal_image"].view(*prefix, 3)
    for k in ("diffuse_image", "specular_image", "roughness_image"):
        if k in res:
            results[k] = res[k]
    return results


_reference_render = None


def render_model(model, rays_o, rays_d, staged=False, max_ray_batch=4096, get_normal_image=False, use_specular_color=True,
                 env_net_index=None, material=None, r_images=None, env_rot_radian=None, **kwargs):
    """Replacement for NeRFRenderer.render (nerf/renderer.py:364-531) on a reference model.  The evaluation-time three-pass
    indirect-reflection frame (opt.indir_ref, renderer.py:439-513) goes through `render()` of this module -- geometry pass with sample
    log, deferred shading of the secondary pass, main pass shaded from the log -- instead of three separate run_cuda calls; every other
    call (training, single pass, staged / chunked, sphere modes, ...) is forwarded to the reference method, whose run_cuda is ours."""
    opt = model.opt
    fast = (getattr(model, "cuda_ray", False) and not model.training and opt.indir_ref and not opt.debug and not opt.use_neus_sdf
            and not opt.error_bound_sample and not opt.env_sph_mode and not opt.render_env_on_sphere and not (model.bg_radius > 0)
            and material is None and use_specular_color and r_images is None      # env_net_index only selects among env_nets in env_sph_mode (network.py:530, 590), excluded above
            and not kwargs.get("perturb", False) and kwargs.get("ray_depth") is None and int(getattr(opt, "max_ray_batch_cuda", 0)) <= 0
            and rays_o.shape[0] == 1)
    if not fast:
        if _reference_render is None:
            raise _lib.EnvidrError("render_model: this call needs the reference NeRFRenderer.render; call install(patch_render=True) first")
        return _reference_render(model, rays_o, rays_d, staged=staged, max_ray_batch=max_ray_batch, get_normal_image=get_normal_image,
                                 use_specular_color=use_specular_color, env_net_index=env_net_index, material=material, r_images=r_images,
                                 env_rot_radian=env_rot_radian, **kwargs)
    field = _model_field(model)
    ob = getattr(model, "obj_aabb", None)
    cfg = RenderConfig(bound=float(model.bound), cascade=int(model.cascade), grid_size=int(model.grid_size), min_near=float(model.min_near),
                       dt_gamma=float(kwargs.get("dt_gamma", 0)), max_steps=int(kwargs.get("max_steps", 1024)),
                       T_thresh=float(kwargs.get("T_thresh", 1e-4)), aabb=[float(v) for v in model.aabb_infer.tolist()], indir_ref=True,
                       indir_max_steps=int(opt.indir_max_steps), obj_aabb=None if ob is None else [float(v) for v in ob.tolist()])
    bg = kwargs.get("bg_color")
    bg = 0.0 if bg is None else bg                                        # renderer.py:463-466: None means black in the three-pass frame
    N = rays_o.shape[1]
    res = render(field, model.density_bitfield, rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), cfg, bg_color=bg, get_normal_image=get_normal_image,
                 env_rot_radian=env_rot_radian, visual_items=tuple(opt.visual_items) if opt.use_diffuse else ())
    out = {"image": res["image"].view(1, N, 3), "depth": res["depth"].view(1, N), "weights_sum": res["weights_sum"].view(1, N),
           "normal_image": res["normal_image"].view(1, N, 3)}
    for k in ("diffuse_image", "specular_image", "roughness_image"):
        if k in res:
            out[k] = res[k].view(1, N, -1)
    return out


def install(render_func_module=None, patch_render: bool = True, renderer_class=None, precision: str = "tc"):
    """Route the reference's renderer through this library: operator-level `_backend`s + fused inference loop behind
    `nerf.render_func.run_cuda` + (patch_render, default) `NeRFRenderer.render` replaced by `render_model`, which runs the
    evaluation-time three-pass frame as one batched schedule (16.8 ms instead of 25.0 ms per 800x800 frame) and forwards every other
    call to the reference method; renderer_class defaults to nerf.renderer.NeRFRenderer.  patch_render=False patches run_cuda only.
    precision: arithmetic of the fused field, "tc" (tensor cores, default) or "fp32"."""
    global _reference_run_cuda, _reference_render, _dropin_precision
    if precision not in ("tc", "fp32"):
        raise _lib.EnvidrError(f"unknown precision {precision!r} (tc | fp32)")
    _dropin_precision = precision
    from .backend import install_into_sys_modules
    install_into_sys_modules()
    if render_func_module is None:
        import importlib
        render_func_module = importlib.import_module("nerf.render_func")
    if _reference_run_cuda is None:
        _reference_run_cuda = render_func_module.run_cuda
    render_func_module.run_cuda = run_cuda
    if patch_render:
        if renderer_class is None:
            import importlib
            renderer_class = importlib.import_module("nerf.renderer").NeRFRenderer
        if _reference_render is None:
            _reference_render = renderer_class.render
        renderer_class.render = render_model
    return render_func_module


def uninstall(render_func_module=None, renderer_class=None):
    """Undo install(): the reference's own run_cuda / NeRFRenderer.render again (the operator-level backends stay ours)."""
    global _reference_run_cuda, _reference_render
    import importlib
    if _reference_run_cuda is not None:
        (render_func_module or importlib.import_module("nerf.render_func")).run_cuda = _reference_run_cuda
        _reference_run_cuda = None
    if _reference_render is not None:
        (renderer_class or importlib.import_module("nerf.renderer").NeRFRenderer).render = _reference_render
        _reference_render = None
