"""BASELINE config 4: the NeuS-style field (use_neus_sdf + encoding_pos=frequency + geometric_init) on libenvidr_b200.

Reference path (per render iteration, nerf/render_func/cuda_ray.py:296-318 with use_neus): FreqEncoder -> 8 weight-normed 256-wide
layers with Softplus(beta=100) and a skip connection (nerf/network.py:154-222, 415-421) -> normal by autograd.grad
(nerf/renderer.py:182-198) -> NeuSDensity alpha (network.py:69-102) -> IDE / env_net / diffuse / colour heads
(network.py:524-698) -> composite_rays(input_alpha=True).

Here:
  * the whole geometry network -- frequency encoding, the dense layers forward AND the reverse pass of the analytic normal
    g <- (g . softplus'(z)) W, Softplus, skip concat / split, head -- is ONE tensor-core kernel per batch (csrc/neus_geom_tc.cu: tcgen05,
    fp16 hi/lo split, fp32 accumulate, activations stay on the SM), weight norm folded into the weights at pack time;
    `NeusField.fused = False` (or a shape outside the kernel's range) runs the same math as a chain of envidr_linear_tc launches with
    one glue kernel between them (csrc/neus_field.cu) -- the first version, kept as the cross-check;
  * the head -> normal / roughness / shading record and the NeuS opacity are one kernel each (envidr_neus_records, envidr_neus_alpha_forward);
  * shading reuses the tensor-core env_net / heads kernels on 32-float geometry records (envidr_field_forward_records);
  * `render_rays_neus` is the inference loop with the reference's schedule (n_step = clamp(N // n_alive, 1, 8)) on the library's march
    and composite kernels.
No CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

import ctypes
import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream
from .field import FieldParams, rot_theta3
from .linear_tc import _image
from .linear_tc import _run as _linear_run

SQRT3 = 3 ** 0.5
LINEAR_TIMING: Optional[list] = None      # bench.py: when a list, every dense-layer launch appends (start event, end event, flop)


def _run(x, img, bias, N, relu):
    if LINEAR_TIMING is None:
        return _linear_run(x, img, bias, N, relu)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    y = _linear_run(x, img, bias, N, relu)
    b.record()
    LINEAR_TIMING.append((a, b, 2.0 * x.shape[0] * x.shape[1] * N))
    return y


def fold_weight_norm(lin) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Effective (W, b) of an nn.Linear, with torch.nn.utils.weight_norm folded: W = g * v / ||v||_row (network.py:218-219)."""
    if hasattr(lin, "weight_g") and hasattr(lin, "weight_v"):
        v, g = lin.weight_v.detach().float(), lin.weight_g.detach().float()
        W = v * (g / v.norm(dim=1, keepdim=True))
    elif hasattr(lin, "parametrizations") and hasattr(lin.parametrizations, "weight"):
        W = lin.weight.detach().float()
    else:
        W = lin.weight.detach().float()
    b = None if lin.bias is None else lin.bias.detach().float().contiguous()
    return W.contiguous(), b


@dataclasses.dataclass
class NeusField:
    sdf: List[Tuple[torch.Tensor, Optional[torch.Tensor]]]      # effective weights [out, in] of the geometry network
    skip_layers: Sequence[int]
    multires: int                                                # FreqEncoder degree (39 = 3 + 2 * 3 * 6 inputs for 6)
    variance: torch.Tensor                                       # NeuSDensity.variance (device scalar)
    shading: FieldParams                                         # env / diffuse / colour / renv stacks + scalars (tensor-core kernels)
    geo_feat_dim: int = 12
    cos_anneal_ratio: float = 1.0
    base_steps: int = 1024
    beta_act: float = 100.0
    _img: Optional[list] = None
    _imgT: Optional[list] = None
    _scratch: Optional[torch.Tensor] = None
    _head_row: Optional[torch.Tensor] = None
    fused: bool = True           # one kernel for the whole geometry network (csrc/neus_geom_tc.cu); False: the composed chain of dense-layer launches

    @property
    def device(self):
        return self.sdf[0][0].device

    @property
    def in_dim(self) -> int:
        return 3 + 2 * 3 * self.multires

    def to(self, device) -> "NeusField":
        mv = lambda t: None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()
        sh = self.shading.to(device)
        sh.precision = "tc"
        return dataclasses.replace(self, sdf=[(mv(W), mv(b)) for W, b in self.sdf], variance=mv(self.variance), shading=sh, _img=None, _imgT=None)

    def pack(self) -> "NeusField":
        self._img = [_image(W) for W, _ in self.sdf]
        self._imgT = [_image(W.t().contiguous()) for W, _ in self.sdf]
        self._head_row = self.sdf[-1][0][0].contiguous()
        self.shading.pack()
        return self

    def fused_supported(self) -> bool:
        """Shapes the fused kernel takes (include/envidr_b200.h, envidr_neus_geometry); anything else runs the composed chain."""
        r16 = lambda v: (v + 15) // 16 * 16
        nl = len(self.sdf)
        if not (2 <= nl <= 8 and 1 <= self.multires <= 7 and len(self.skip_layers) <= 1 and all(b is not None for _, b in self.sdf)):
            return False
        if self.skip_layers and not (1 <= self.skip_layers[0] < nl - 1):
            return False
        if self.sdf[-1][0].shape[0] > 16:
            return False
        for l, (W, _) in enumerate(self.sdf):
            n, k = W.shape
            if n > 256 or k > 256 or (l + 1 < nl and r16(n) % 32) or (l > 0 and r16(k) % 32):
                return False
        return True

    def _geometry_fused(self, xyzs: torch.Tensor):
        """-> head [M,16], grad_x [M,3] through envidr_neus_geometry."""
        M, dev = xyzs.shape[0], xyzs.device
        L = lib()
        nl = len(self.sdf)
        need = int(L.envidr_neus_geometry_scratch_bytes(nl))
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=dev)
        net = _lib.NeusNet()
        for l, (W, b) in enumerate(self.sdf):
            net.layers[l].img, net.layers[l].imgT, net.layers[l].bias = self._img[l].data_ptr(), self._imgT[l].data_ptr(), b.data_ptr()
            net.layers[l].in_dim, net.layers[l].out_dim = int(W.shape[1]), int(W.shape[0])
        net.head_row, net.n_layers = self._head_row.data_ptr(), nl
        net.skip_layer = int(self.skip_layers[0]) if self.skip_layers else -1
        net.multires, net.beta = int(self.multires), float(self.beta_act)
        head = torch.empty(M, 16, dtype=torch.float32, device=dev)
        grad_x = torch.empty(M, 3, dtype=torch.float32, device=dev)
        ev = None
        if LINEAR_TIMING is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        check(L.envidr_neus_geometry(ctypes.byref(net), ptr(xyzs), M, ptr(head), ptr(grad_x), ptr(self._scratch), self._scratch.numel(), stream()),
              "neus_geometry")
        if ev is not None:
            ev[1].record()
            flop = sum(2.0 * W.shape[0] * W.shape[1] for W, _ in self.sdf) + sum(2.0 * W.shape[0] * W.shape[1] for W, _ in self.sdf[:-1])
            LINEAR_TIMING.append((ev[0], ev[1], flop * M))
        return head, grad_x

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def shading_params(env, diffuse, color, renv, device, **scalars) -> FieldParams:
        """FieldParams that carries only the rendering MLPs: the hash grid / sdf slots get inert placeholders (never evaluated:
        only envidr_field_forward_records is called on it)."""
        G = int(scalars.get("geo_feat_dim", 12))
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        sdf = [(z(64, 2), z(64)), (z(1 + G + 2, 64), z(1 + G + 2))]
        mv = lambda st: None if st is None else [(W.detach().float().contiguous().to(device), None if b is None else b.detach().float().contiguous().to(device))
                                                 for W, b in st]
        return FieldParams(embeddings=z(8, 2), offsets=torch.tensor([0, 8], dtype=torch.int32, device=device), per_level_scale=2.0,
                           base_resolution=2, bound=float(scalars.pop("bound", 1.0)), sdf=sdf, env=mv(env), diffuse=mv(diffuse), color=mv(color),
                           renv=mv(renv), precision="tc", **scalars)

    @staticmethod
    def from_reference_model(model) -> "NeusField":
        """Read a reference NeRFNetwork built with use_neus_sdf / encoding_pos=frequency (the config-4 family)."""
        opt = model.opt
        ok = (opt.use_sdf and opt.use_neus_sdf and opt.encoding_pos == "frequency" and opt.ensemble_mlp and opt.use_env_net and opt.use_diffuse
              and opt.diffuse_with_env and opt.wo_viewdir and opt.normal_with_mlp and opt.use_n_dot_viewdir and opt.use_roughness
              and opt.geo_feat_act == "unitNorm" and opt.env_feat_act == "unitNorm" and opt.encoding_ref == "integrated_dir"
              and opt.color_act == "sigmoid" and opt.diffuse_env_fusion == "concat" and not opt.split_diffuse_env
              and float(getattr(opt, "normal_anneal_ratio", 1)) >= 1 and not getattr(model, "w_material", False))
        if not ok:
            raise _lib.EnvidrError("NeuS field: model configuration is outside the supported config-4 family")
        dev = next(model.parameters()).device
        lin = lambda net: [(l.weight.detach().float().contiguous(), None if l.bias is None else l.bias.detach().float().contiguous()) for l in net]
        shading = NeusField.shading_params(
            lin(model.env_net), lin(model.diffuse_net), lin(model.color_net),
            lin(model.renv_net) if getattr(model, "renv_net", None) is not None else None, dev,
            geo_feat_dim=int(model.geo_feat_dim), ide_degree=int(opt.sh_degree), bound=float(model.bound),
            roughness_bias=float(model.roughness_bias), roughness_act_scale=float(opt.roughness_act_scale),
            roughness_scale=float(opt.roughness_scale), diffuse_kappa_inv=float(opt.diffuse_kappa_inv),
            light_intensity_scale=float(opt.light_intensity_scale), intensity_scale=float(opt.intensity_scale),
            indir_roughness_thresh=float(opt.indir_roughness_thresh), learn_indir_blend=bool(opt.learn_indir_blend))
        act = model.sdf_act
        beta_act = float(getattr(act, "beta", 100.0)) if isinstance(act, torch.nn.Softplus) else None
        if beta_act is None:
            raise _lib.EnvidrError("NeuS field: geometric_init (Softplus activations) expected")
        return NeusField(sdf=[fold_weight_norm(l) for l in model.sdf_net], skip_layers=tuple(int(s) for s in opt.skip_layers),
                         multires=int(opt.multires), variance=model.sdf_density.variance.detach().float().reshape(1).contiguous(),
                         shading=shading, geo_feat_dim=int(model.geo_feat_dim), cos_anneal_ratio=float(opt.cos_anneal_ratio),
                         base_steps=int(model.sdf_density.base_steps), beta_act=beta_act).pack()

    # ------------------------------------------------------------------------------------------------------------
    def geometry(self, xyzs: torch.Tensor, dirs: torch.Tensor, dists: Optional[torch.Tensor] = None, env_rot_radian: Optional[float] = None,
                 want_rec: bool = True) -> Dict[str, torch.Tensor]:
        """forward_sigma of the config-4 model for M samples: sdf, unit normal, roughness, alpha (NeuSDensity) and the geometry record
        for the shading kernels.  dists: deltas[:, 0] ([M]) or None = the module's base distance."""
        if self._img is None:
            self.pack()
        xyzs = xyzs.float().contiguous().view(-1, 3)
        dirs = dirs.float().contiguous().view(-1, 3)
        M, dev = xyzs.shape[0], xyzs.device
        f32 = dict(dtype=torch.float32, device=dev)
        L = lib()
        if self.fused and M > 0 and self.fused_supported():
            head, grad_x = self._geometry_fused(xyzs)
            return self._finish_geometry(head, grad_x, dirs, dists, env_rot_radian, want_rec)
        C = self.in_dim
        enc = torch.empty(M, C, **f32)
        check(L.envidr_freq_encode_forward(ptr(xyzs), M, 3, self.multires, C, ptr(enc), stream()), "freq_encode_forward")
        nl = len(self.sdf)
        h, derivs, skip_in = enc, [], {}
        for l, (W, b) in enumerate(self.sdf):
            if l in self.skip_layers:                                    # h = cat([h, x]) / sqrt(2)   (network.py:417-418)
                Nh = h.shape[1]
                cat = torch.empty(M, Nh + C, **f32)
                check(L.envidr_skip_concat_forward(ptr(h), ptr(enc), M, Nh, C, 1.0 / math.sqrt(2.0), ptr(cat), stream()), "skip_concat_forward")
                skip_in[l] = Nh
                h = cat
            z = _run(h, self._img[l], b, W.shape[0], False)
            if l != nl - 1:
                s = torch.empty_like(z)
                check(L.envidr_softplus_forward(ptr(z), z.numel(), self.beta_act, ptr(z), ptr(s), stream()), "softplus_forward")
                derivs.append(s)
            h = z
        head = h                                                          # [M, 1 + G + 2]
        # reverse pass: g = d sdf / d (input of layer l), from the sdf row of the last layer down to the encoding
        g = None
        g_enc_skip = None
        for l in range(nl - 1, -1, -1):
            W = self.sdf[l][0]
            if l == nl - 1:
                gz = torch.empty(M, W.shape[1], **f32)                    # W_last[0, :] . softplus'(z_{last-1})
                check(L.envidr_mul_rows(None, ptr(W[0].contiguous()), ptr(derivs[l - 1]), M, W.shape[1], ptr(gz), stream()), "mul_rows")
                g = gz
                continue
            gin = _run(g, self._imgT[l], None, W.shape[1], False)         # g @ W_l  (the image of W_l^T)
            if l in skip_in:
                Nh = skip_in[l]
                gh = torch.empty(M, Nh, **f32)
                g_enc_skip = torch.empty(M, C, **f32)
                check(L.envidr_skip_concat_backward(ptr(gin), M, Nh, C, 1.0 / math.sqrt(2.0), ptr(gh), ptr(g_enc_skip), 0, stream()),
                      "skip_concat_backward")
                gin = gh
            if l == 0:
                g = gin
                break
            gz = torch.empty_like(gin)
            check(L.envidr_mul_rows(ptr(gin), None, ptr(derivs[l - 1]), M, gin.shape[1], ptr(gz), stream()), "mul_rows")
            g = gz
        if g_enc_skip is not None:
            g = g + g_enc_skip
        grad_x = torch.empty(M, 3, **f32)
        check(L.envidr_freq_encode_backward(ptr(g.contiguous()), ptr(enc), M, 3, self.multires, C, ptr(grad_x), stream()), "freq_encode_backward")
        return self._finish_geometry(head, grad_x, dirs, dists, env_rot_radian, want_rec)

    def _finish_geometry(self, head, grad_x, dirs, dists, env_rot_radian, want_rec) -> Dict[str, torch.Tensor]:
        """Head + gradient -> unit normal, roughness, geometry record, NeuS opacity."""
        M, dev = head.shape[0], head.device
        f32 = dict(dtype=torch.float32, device=dev)
        L = lib()
        sh = self.shading
        out = dict(sdf=torch.empty(M, **f32), normal=torch.empty(M, 3, **f32), roughness=torch.empty(M, **f32), grad_x=grad_x)
        rec = torch.empty(M, 32, **f32) if want_rec else None
        rot = None
        if env_rot_radian is not None:
            rot = (ctypes.c_float * 9)(*[float(v) for v in rot_theta3(float(env_rot_radian)).reshape(-1)])
        check(L.envidr_neus_records(ptr(head), head.shape[1], ptr(grad_x), ptr(dirs), M, self.geo_feat_dim, sh.roughness_bias, sh.roughness_act_scale,
                                    sh.roughness_scale, int(head.shape[1] > 2 + self.geo_feat_dim), rot, ptr(out["sdf"]), ptr(out["normal"]),
                                    ptr(out["roughness"]), ptr(rec), stream()), "neus_records")
        alpha = torch.empty(M, **f32)
        d_t = None if dists is None else dists.float().contiguous().view(-1)
        check(L.envidr_neus_alpha_forward(ptr(out["sdf"]), ptr(dirs), ptr(out["normal"]), ptr(d_t), 2 * SQRT3 / self.base_steps, ptr(self.variance),
                                          float(self.cos_anneal_ratio), M, ptr(alpha), stream()), "neus_alpha_forward")
        out["sigma"] = alpha
        if rec is not None:
            out["rec"] = rec
        return out

    def shade(self, rec: torch.Tensor, r_images: Optional[torch.Tensor] = None, want=("rgb",)) -> Dict[str, torch.Tensor]:
        """env_net + diffuse / colour (/ renv) heads on geometry records: forward_color (network.py:524-698)."""
        M, dev = rec.shape[0], rec.device
        sh = self.shading
        shapes = dict(rgb=(M, 3), c_diffuse=(M, 3), c_specular=(M, 3))
        outs = {k: torch.empty(shapes[k], dtype=torch.float32, device=dev) for k in want}
        if M == 0:
            return outs
        if sh._scratch is None or sh._scratch.numel() < 32 * (M + 2):
            sh._scratch = torch.empty(32 * (M + 2), dtype=torch.float32, device=dev)
        fo = _lib.FieldOut()
        for k, t in outs.items():
            setattr(fo, k, t.data_ptr())
        f = sh.cstruct()
        ri = None if r_images is None else r_images.float().contiguous().view(-1, 4)
        check(lib().envidr_field_forward_records(ctypes.byref(f), ptr(rec), ptr(ri), M, ctypes.byref(fo), stream()), "field_forward_records")
        return outs

    def forward(self, xyzs, dirs, dists=None, r_images=None, env_rot_radian=None, want=("sigma", "rgb", "normal")) -> Dict[str, torch.Tensor]:
        g = self.geometry(xyzs, dirs, dists, env_rot_radian)
        sw = tuple(k for k in want if k in ("rgb", "c_diffuse", "c_specular"))
        out = dict(g)
        if sw:
            out.update(self.shade(g["rec"], r_images, sw))
        return {k: out[k] for k in want}


def render_rays_neus(field: NeusField, bitfield: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, *, bound: float = 1.0,
                     cascade: int = 1, grid_size: int = 128, min_near: float = 0.2, dt_gamma: float = 0.0, max_steps: int = 1024,
                     T_thresh: float = 1e-4, bg_color=1.0, aabb: Optional[Sequence[float]] = None, env_rot_radian: Optional[float] = None,
                     r_images: Optional[torch.Tensor] = None, geometry_only: bool = False, visual_items: Sequence[str] = (),
                     stats: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """run_cuda's inference branch (cuda_ray.py:238-359) for the NeuS-style field, input_alpha compositing: the reference's schedule
    (n_step = clamp(N // n_alive, 1, 8)) and samples on the library's march / composite kernels; per iteration ONE fused geometry launch
    (csrc/neus_geom_tc.cu), the record / opacity kernels, the tensor-core shading kernels and one composite_rays per image, each on its
    own copy of the alive list exactly as the reference keeps them (cuda_ray.py:318-342).  One host synchronisation per iteration
    (the alive count sizes the next batch, as in the reference)."""
    from . import raymarching as rm
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    N, dev = rays_o.shape[0], rays_o.device
    f32 = dict(dtype=torch.float32, device=dev)
    aabb_t = torch.tensor(list(aabb) if aabb is not None else [-bound] * 3 + [bound] * 3, **f32)
    nears, fars = rm.near_far_from_aabb(rays_o, rays_d, aabb_t, min_near)
    z1, z3 = (lambda: torch.zeros(N, **f32)), (lambda: torch.zeros(N, 3, **f32))
    ws, depth, image = z1(), z1(), z3()
    n_ws, n_depth, n_img = z1(), z1(), z3()
    want = [] if geometry_only else ["rgb"]
    extra = {}
    if not geometry_only and "diffuse" in visual_items:
        want.append("c_diffuse"); extra["diffuse"] = (z1(), z1(), z3())
    if not geometry_only and "specular" in visual_items:
        want.append("c_specular"); extra["specular"] = (z1(), z1(), z3())
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    rays_t = nears.clone()
    ri = None if r_images is None else r_images.float().contiguous().view(-1, 4)
    step = iters = samples = 0
    while step < max_steps:
        n_alive = alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = rm.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, bound, bitfield, cascade, grid_size, nears, fars, -1,
                                           False, dt_gamma, max_steps)
        g = field.geometry(xyzs, dirs, deltas[:, 0], env_rot_radian, want_rec=not geometry_only)
        alpha = g["sigma"]
        if stats is not None:
            samples += int((deltas[:, 0] > 0).sum())
        if geometry_only:
            rm.composite_rays(n_alive, n_step, alive, rays_t, alpha, g["normal"], deltas, ws, depth, n_img, T_thresh, True)
        else:
            r_s = None if ri is None else ri[alive.long()][:, None, :].expand(-1, n_step, -1).reshape(-1, 4).contiguous()
            sh = field.shade(g["rec"], r_s, tuple(want))
            # the normal / diffuse / specular images are composited on copies of the alive list: same opacities, same schedule, so
            # they evolve exactly like the reference's separately kept lists
            rm.composite_rays(n_alive, n_step, alive.clone(), rays_t.clone(), alpha, g["normal"], deltas, n_ws, n_depth, n_img, T_thresh, True)
            if "diffuse" in extra:
                w_, d_, im_ = extra["diffuse"]
                rm.composite_rays(n_alive, n_step, alive.clone(), rays_t.clone(), alpha, sh["c_diffuse"], deltas, w_, d_, im_, T_thresh, True)
            if "specular" in extra:                                      # roughness composited into `depth` (cuda_ray.py:326-335)
                w_, d_, im_ = extra["specular"]
                dl = deltas.clone()
                dl[:, 1] = g["roughness"]
                rm.composite_rays(n_alive, n_step, alive.clone(), rays_t.clone(), alpha, sh["c_specular"], dl, w_, d_, im_, T_thresh, True, False)
            rm.composite_rays(n_alive, n_step, alive, rays_t, alpha, sh["rgb"], deltas, ws, depth, image, T_thresh, True)
        alive = alive[alive >= 0]
        step += n_step
        iters += 1
    if stats is not None:
        stats.update(samples=samples, iterations=iters)
    res = {"depth": depth, "weights_sum": ws, "normal_image": torch.nn.functional.normalize(n_img, dim=-1, eps=1e-10)}
    if geometry_only:
        res["image"] = None
        return res
    bg = bg_color if torch.is_tensor(bg_color) else torch.tensor(float(bg_color), **f32)
    res["image"] = image + (1 - ws).unsqueeze(-1) * bg
    if "diffuse" in extra:
        res["diffuse_image"] = extra["diffuse"][2]
    if "specular" in extra:
        res["specular_image"] = extra["specular"][2]
        res["roughness_image"] = extra["specular"][1][..., None]
    return res
