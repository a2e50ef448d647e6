"""Host-side description of the per-sample field (hash grid + SDF / env / diffuse / colour / renv MLPs)
and the call into the fused CUDA kernel (envidr_field_forward).

The reference has no operator boundary for this part of the hot path: the MLPs are nn.ModuleList[nn.Linear]
driven from Python (nerf/network.py:381-698) and the normals come from torch.autograd.grad
(nerf/renderer.py:182-198).  `FieldParams.from_reference_model(model)` reads exactly the state the reference
model holds (model.encoder.embeddings / offsets, model.sdf_net, env_net, diffuse_net, color_net, renv_net,
model.sdf_density.beta, opt.*) so the fused path is a drop-in for NeRFNetwork.forward_sigma + forward_color.
"""
from __future__ import annotations

import ctypes
import dataclasses
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream

Layer = Tuple[torch.Tensor, Optional[torch.Tensor]]


def rot_theta3(th: float) -> np.ndarray:
    """Upper-left 3x3 of rot_theta (nerf/utils.py; demo.ipynb cell 5)."""
    return np.array([[math.cos(th), 0, -math.sin(th)], [0, 1, 0], [math.sin(th), 0, math.cos(th)]], np.float32)


@dataclasses.dataclass
class FieldParams:
    embeddings: torch.Tensor            # [T, 2] fp32
    offsets: torch.Tensor               # [L+1] int32
    per_level_scale: float
    base_resolution: int
    bound: float
    sdf: List[Layer]
    env: List[Layer]
    diffuse: List[Layer]
    color: List[Layer]
    renv: Optional[List[Layer]] = None
    geo_feat_dim: int = 12
    ide_degree: int = 5
    beta: float = 0.01
    beta_min: float = 0.0005
    beta_max: float = 1.0
    density_scale: float = 1.0
    roughness_bias: float = -1.0
    roughness_act_scale: float = 0.2
    roughness_scale: float = 1.0
    diffuse_kappa_inv: float = 0.64
    light_intensity_scale: float = 1.0
    intensity_scale: float = 1.0
    indir_roughness_thresh: float = 0.1
    learn_indir_blend: bool = True
    enabled_levels: int = -1
    precision: str = "tc"               # "tc": tcgen05 tensor cores, fp16 hi/lo split operands (3 MMAs), the default of every drop-in
                                        # entry (from_reference_model, checkpoint loader); "fp32": the exact FFMA path (13x slower)
    _packed: Optional[torch.Tensor] = None
    _scratch: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------------------------------
    @property
    def device(self):
        return self.embeddings.device

    @property
    def num_levels(self) -> int:
        return int(self.offsets.shape[0] - 1)

    def stacks(self) -> Dict[str, Optional[List[Layer]]]:
        return dict(sdf=self.sdf, env=self.env, diffuse=self.diffuse, color=self.color, renv=self.renv)

    def to(self, device) -> "FieldParams":
        mv = lambda t: None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()
        kw = {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}
        kw["embeddings"] = mv(self.embeddings)
        kw["offsets"] = self.offsets.detach().to(device=device, dtype=torch.int32).contiguous()
        for name, st in self.stacks().items():
            kw[name] = None if st is None else [(mv(W), mv(b)) for W, b in st]
        kw["_packed"] = None
        kw["_scratch"] = None
        return FieldParams(**kw)

    def clamped_beta(self) -> float:
        """LaplaceDensity.get_beta (nerf/network.py:39-44)."""
        return min(max(float(self.beta), float(self.beta_min)), float(self.beta_max))

    def to_oracle(self) -> Dict:
        """Plain numpy/scalar dict consumed by oracle.oracle.field_forward (tests / CPU baseline only)."""
        npy = lambda t: None if t is None else t.detach().cpu().numpy().astype(np.float32)
        P = dict(embeddings=npy(self.embeddings), offsets=self.offsets.detach().cpu().numpy().astype(np.int32),
                 per_level_scale=float(self.per_level_scale), base_resolution=int(self.base_resolution), bound=float(self.bound),
                 geo_feat_dim=self.geo_feat_dim, ide_degree=self.ide_degree, beta=self.beta, beta_min=self.beta_min,
                 beta_max=self.beta_max, density_scale=self.density_scale, roughness_bias=self.roughness_bias,
                 roughness_act_scale=self.roughness_act_scale, roughness_scale=self.roughness_scale,
                 diffuse_kappa_inv=self.diffuse_kappa_inv, light_intensity_scale=self.light_intensity_scale,
                 intensity_scale=self.intensity_scale, indir_roughness_thresh=self.indir_roughness_thresh,
                 learn_indir_blend=self.learn_indir_blend, enabled_levels=self.enabled_levels)
        for name, st in self.stacks().items():
            P[name] = None if st is None else [(npy(W), npy(b)) for W, b in st]
        return P

    # ------------------------------------------------------------------------------------------
    def cstruct(self, env_rot_radian: Optional[float] = None, rec_unrotated: bool = False) -> _lib.Field:
        f = _lib.Field()
        f.rec_unrotated = int(rec_unrotated)
        f.embeddings = self.embeddings.data_ptr()
        f.offsets = self.offsets.data_ptr()
        f.num_levels = self.num_levels
        f.level_dim = int(self.embeddings.shape[1])
        f.base_resolution = int(self.base_resolution)
        f.log2_per_level_scale = float(np.log2(self.per_level_scale))
        f.bound = float(self.bound)
        f.enabled_levels = int(self.enabled_levels)
        for name, st in self.stacks().items():
            n = 0 if st is None else len(st)
            if n > _lib.ENVIDR_MAX_LAYERS:
                raise _lib.EnvidrError(f"{name}_net has {n} layers; the fused field supports at most {_lib.ENVIDR_MAX_LAYERS}")
            setattr(f, f"n_{name}", n)
            arr = getattr(f, name)
            for i in range(n):
                W, b = st[i]
                assert W.is_cuda and W.dtype == torch.float32 and W.is_contiguous()
                arr[i].weight = W.data_ptr()
                arr[i].bias = 0 if b is None else b.data_ptr()
                arr[i].in_dim, arr[i].out_dim = int(W.shape[1]), int(W.shape[0])
        f.geo_feat_dim = self.geo_feat_dim
        f.ide_degree = self.ide_degree
        f.beta = self.clamped_beta()
        f.density_scale = self.density_scale
        f.roughness_bias, f.roughness_act_scale, f.roughness_scale = self.roughness_bias, self.roughness_act_scale, self.roughness_scale
        f.diffuse_kappa_inv, f.light_intensity_scale, f.intensity_scale = self.diffuse_kappa_inv, self.light_intensity_scale, self.intensity_scale
        f.indir_roughness_thresh = self.indir_roughness_thresh
        f.learn_indir_blend = int(self.learn_indir_blend)
        if env_rot_radian is not None:
            f.has_env_rot = 1
            R = rot_theta3(float(env_rot_radian)).reshape(-1)
            for i in range(9):
                f.env_rot[i] = float(R[i])
        if self._packed is not None:
            f.packed = self._packed.data_ptr()
            f.packed_bytes = self._packed.numel() * 4
        if self.precision not in ("fp32", "tc"):
            raise _lib.EnvidrError(f"unknown precision {self.precision!r} (fp32 | tc)")
        f.precision = 1 if self.precision == "tc" else 0
        if self._scratch is not None:
            f.scratch = self._scratch.data_ptr()
            f.scratch_samples = self._scratch.numel() // 64
        return f

    def pack(self) -> "FieldParams":
        """(Re)build the K-major padded weight images the kernels read.  Call after any weight update."""
        f = self.cstruct()
        nbytes = lib().envidr_field_pack_bytes(ctypes.byref(f))
        if nbytes == 0:
            raise _lib.EnvidrError("field rejected: " + lib().envidr_last_error().decode())
        if self._packed is None or self._packed.numel() * 4 < nbytes:
            self._packed = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
        check(lib().envidr_field_pack(ctypes.byref(f), ptr(self._packed), self._packed.numel() * 4, stream()), "field_pack")
        return self

    def forward(self, xyzs: torch.Tensor, dirs: torch.Tensor, r_images: Optional[torch.Tensor] = None, *, geometry_only: bool = False,
                env_rot_radian: Optional[float] = None, want=("sigma", "rgb", "normal")) -> Dict[str, torch.Tensor]:
        """forward_sigma + get_color_mlp_extra_params + forward_color for M samples, one kernel.
        want: subset of sigma, rgb, normal, sdf, c_diffuse, c_specular, roughness, grad_x."""
        if self._packed is None:
            self.pack()
        xyzs = xyzs.float().contiguous().view(-1, 3)
        dirs = dirs.float().contiguous().view(-1, 3)
        M = xyzs.shape[0]
        shapes = dict(sigma=(M,), rgb=(M, 3), normal=(M, 3), sdf=(M,), c_diffuse=(M, 3), c_specular=(M, 3), roughness=(M,), grad_x=(M, 3))
        outs = {k: torch.empty(shapes[k], dtype=torch.float32, device=xyzs.device) for k in want}
        fo = _lib.FieldOut()
        for k, t in outs.items():
            setattr(fo, k, t.data_ptr())
        if r_images is not None:
            r_images = r_images.float().contiguous().view(-1, 4)
            assert r_images.shape[0] == M
        if self.precision == "tc" and not geometry_only and (self._scratch is None or self._scratch.numel() < 64 * M):
            self._scratch = torch.empty(64 * max(M, 1), dtype=torch.float32, device=xyzs.device)
        f = self.cstruct(env_rot_radian)
        check(lib().envidr_field_forward(ctypes.byref(f), ptr(xyzs), ptr(dirs), ptr(r_images), M, 1 if geometry_only else 0,
                                         ctypes.byref(fo), stream()), "field_forward")
        return outs

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def from_reference_model(model, precision: str = "tc") -> "FieldParams":
        """Read the state of a reference NeRFNetwork (nerf/network.py) built for the shipped scene configs
        (encoding_pos=hashgrid_diff, ensemble_mlp, use_env_net, diffuse_with_env, wo_viewdir, ...)."""
        opt = model.opt
        ok = (opt.encoding_pos == "hashgrid_diff" and opt.use_sdf and not opt.use_neus_sdf and opt.ensemble_mlp and opt.use_env_net
              and opt.use_diffuse and opt.diffuse_with_env and opt.wo_viewdir and opt.normal_with_mlp and opt.use_n_dot_viewdir
              and opt.use_roughness and opt.geo_feat_act == "unitNorm" and opt.env_feat_act == "unitNorm"
              and opt.encoding_ref == "integrated_dir" and opt.color_act == "sigmoid" and not opt.geometric_init
              and not opt.skip_layers and opt.diffuse_env_fusion == "concat" and not opt.split_diffuse_env
              and float(getattr(opt, "normal_anneal_ratio", 1)) >= 1)          # the fused kernels take the normal from the SDF gradient alone
        if not ok:
            raise _lib.EnvidrError("fused field: model configuration is outside the fused path (use the operator-level modules)")
        lin = lambda net: [(l.weight.detach().float().contiguous(), None if l.bias is None else l.bias.detach().float().contiguous())
                           for l in net]
        enc = model.encoder
        return FieldParams(
            embeddings=enc.embeddings.detach().float().contiguous(), offsets=enc.offsets.int().contiguous(),
            per_level_scale=float(enc.per_level_scale), base_resolution=int(enc.base_resolution), bound=float(model.bound),
            sdf=lin(model.sdf_net), env=lin(model.env_net), diffuse=lin(model.diffuse_net), color=lin(model.color_net),
            renv=lin(model.renv_net) if getattr(model, "renv_net", None) is not None else None,
            geo_feat_dim=int(model.geo_feat_dim), ide_degree=int(opt.sh_degree), beta=float(model.sdf_density.beta.item()),
            beta_min=float(model.sdf_density.beta_min), beta_max=float(model.sdf_density.beta_max),
            density_scale=float(model.density_scale), roughness_bias=float(model.roughness_bias),
            roughness_act_scale=float(opt.roughness_act_scale), roughness_scale=float(opt.roughness_scale),
            diffuse_kappa_inv=float(opt.diffuse_kappa_inv), light_intensity_scale=float(opt.light_intensity_scale),
            intensity_scale=float(opt.intensity_scale), indir_roughness_thresh=float(opt.indir_roughness_thresh),
            learn_indir_blend=bool(opt.learn_indir_blend), enabled_levels=int(opt.enabled_levels), precision=precision)
