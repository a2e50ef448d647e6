"""Checkpoint / weight-format loader and EMA: the reference's on-disk formats in, FieldParams out (and back).

Formats (all `torch.save`d dicts, nerf/utils.py:1478-1562 save, :1564-1666 load):
  * full checkpoint        {'model': state_dict, 'epoch', 'global_step', 'stats', 'mean_count', 'mean_density', 'optimizer',
                            'lr_scheduler', 'scaler', 'ema'} -- or a bare state_dict (utils.py:1577-1581);
                            state_dict keys: encoder.embeddings / encoder.offsets, sdf_net.<i>.weight|bias, env_net.<i>.*,
                            diffuse_net.<i>.*, color_net.<i>.*, renv_net.<i>.*, sdf_density.beta, density_grid, density_bitfield,
                            step_counter, aabb_train, aabb_infer;
  * rendering MLPs         ckpts/rendering_mlps.pth: {'model': {diffuse_net.*, color_net.*, renv_net.*}} (opt.color_mlp_path,
                            utils.py:509-530: per-stack prefixes selected by opt.resume_mlps);
  * environment MLP        ckpts/env_ckpts/env_net_<k>.pth: {'model': {'env_net0.weight', ...}} -- written by
                            nerf/sph_loader.py:356-378 WITHOUT the dot after the stack name, so the reference's own
                            swap (utils.py:1583-1597, prefix 'env_net.') silently finds nothing in them; accepted here in both spellings.
The environment MLP fixes the directional-encoding degree: in_dim = 2 * (2^deg - 1 + deg) -> 38 = degree 4, 72 = degree 5.

EMA: torch_ema.ExponentialMovingAverage as the reference uses it (utils.py:424-427, decay 0.95, updated once per epoch :1095-1096,
swapped in for evaluation :1138-1148) -- torch-ema is an un-vendored pip dependency (requirements.txt, unpinned); restated from its
published algorithm: decay_t = min(decay, (1 + n) / (10 + n)), shadow -= (1 - decay_t) * (shadow - param); state_dict layout kept
({'decay', 'num_updates', 'shadow_params', 'collected_params'}) so checkpoints interchange.
"""
from __future__ import annotations

import collections
import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import EnvidrError
from .field import FieldParams

STACKS = ("sdf", "env", "diffuse", "color", "renv")


def load_state(path: str, map_location="cpu") -> Tuple[Dict[str, torch.Tensor], Dict]:
    """-> (model state_dict with normalised keys, the rest of the checkpoint dict)."""
    ck = torch.load(path, map_location=map_location)
    if isinstance(ck, dict) and "model" in ck:
        sd, meta = ck["model"], {k: v for k, v in ck.items() if k != "model"}
    else:
        sd, meta = ck, {}
    return normalize_keys(sd), meta


def normalize_keys(sd: Dict[str, torch.Tensor]) -> "collections.OrderedDict[str, torch.Tensor]":
    """'env_net0.weight' (sph_loader.py:356-378) -> 'env_net.0.weight'; everything else unchanged."""
    out = collections.OrderedDict()
    for k, v in sd.items():
        m = re.match(r"^(env_net|renv_net|color_net|diffuse_net|sdf_net)(\d+)\.(weight|bias)$", k)
        out[f"{m.group(1)}.{m.group(2)}.{m.group(3)}" if m else k] = v
    return out


def stack_from_state(sd: Dict[str, torch.Tensor], name: str) -> Optional[List[Tuple[torch.Tensor, Optional[torch.Tensor]]]]:
    """[(weight [out, in], bias [out] | None), ...] of `<name>_net.<i>.*`, None when the stack is absent."""
    idx = sorted({int(m.group(1)) for k in sd for m in [re.match(rf"^{name}_net\.(\d+)\.weight$", k)] if m})
    if not idx:
        return None
    if idx != list(range(len(idx))):
        raise EnvidrError(f"{name}_net: layer indices {idx} are not contiguous")
    layers = []
    for i in idx:
        W = sd[f"{name}_net.{i}.weight"].detach().float().contiguous()
        b = sd.get(f"{name}_net.{i}.bias")
        layers.append((W, None if b is None else b.detach().float().contiguous()))
    for (W0, _), (W1, _) in zip(layers, layers[1:]):
        if W1.shape[1] != W0.shape[0]:
            raise EnvidrError(f"{name}_net: layer widths do not chain ({tuple(W0.shape)} -> {tuple(W1.shape)})")
    return layers


def ide_degree_from_env(env: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]]) -> int:
    in_dim = int(env[0][0].shape[1])
    for deg in range(1, 8):
        if 2 * (2 ** deg - 1 + deg) == in_dim:
            return deg
    raise EnvidrError(f"env_net input width {in_dim} is not an integrated-directional-encoding width (38 = degree 4, 72 = degree 5)")


def field_from_state(sd: Dict[str, torch.Tensor], *, base_resolution: int = 16, desired_resolution: int = 2048, bound: float = 1.0,
                     template: Optional[FieldParams] = None, **scalars) -> FieldParams:
    """FieldParams of a full model state_dict (reference key names).  The hash-grid geometry is not stored in the checkpoint
    (hashencoder/hashgrid.py:130-146 recomputes it from the options): per_level_scale follows from base / desired resolution and
    the number of levels in `encoder.offsets`.  Stacks missing from `sd` are taken from `template` (e.g. a field whose rendering
    MLPs were loaded from ckpts/rendering_mlps.pth).  scalars: FieldParams fields that come from the options (beta_min, ...)."""
    sd = normalize_keys(sd)
    if "encoder.embeddings" not in sd or "encoder.offsets" not in sd:
        raise EnvidrError("state_dict has no hash encoder (encoder.embeddings / encoder.offsets)")
    offsets = sd["encoder.offsets"].to(torch.int32).contiguous()
    L = int(offsets.shape[0]) - 1
    pls = float(np.exp2(np.log2(desired_resolution * bound / base_resolution) / (L - 1))) if L > 1 else 1.0
    stacks = {n: stack_from_state(sd, n) for n in STACKS}
    for n in STACKS:
        if stacks[n] is None and template is not None:
            stacks[n] = getattr(template, n)
    missing = [n for n in ("sdf", "env", "diffuse", "color") if stacks[n] is None]
    if missing:
        raise EnvidrError(f"state_dict lacks the {missing} stack(s) and no template field supplies them")
    kw = dict(scalars)
    if "sdf_density.beta" in sd:
        kw.setdefault("beta", float(sd["sdf_density.beta"].reshape(-1)[0]))
    # out_dim = 1 + geo_feat_dim + use_roughness(1) + learn_indir_blend (network.py:415-448)
    geo = int(stacks["sdf"][-1][0].shape[0]) - 1 - 1 - int(bool(kw.get("learn_indir_blend", True)))
    kw.setdefault("geo_feat_dim", geo)
    env_out = int(stacks["env"][-1][0].shape[0])
    for n, extra in (("diffuse", 0), ("color", 3 + 1)):                # diffuse: [geo | f_n]; colour: [geo | n(3) | f_r | n.w_o]
        want = kw["geo_feat_dim"] + env_out + extra
        got = int(stacks[n][0][0].shape[1])
        if got != want:
            raise EnvidrError(f"{n}_net expects {got} inputs but geo_feat_dim {kw['geo_feat_dim']} + env features {env_out} (+{extra}) = {want}: "
                              "pass learn_indir_blend / geo_feat_dim matching the checkpoint's options")
    return FieldParams(embeddings=sd["encoder.embeddings"].detach().float().contiguous(), offsets=offsets, per_level_scale=pls,
                       base_resolution=base_resolution, bound=bound, sdf=stacks["sdf"], env=stacks["env"], diffuse=stacks["diffuse"],
                       color=stacks["color"], renv=stacks["renv"], ide_degree=ide_degree_from_env(stacks["env"]), **kw)


def load_rendering_mlps(field: FieldParams, path: str, resume_mlps: Sequence[str] = ("specular", "diffuse", "renv")) -> FieldParams:
    """opt.color_mlp_path (utils.py:509-530): replace the colour ('specular'), diffuse and renv stacks of `field` by the
    checkpoint's; a missing renv stack is skipped as in the reference.  Returns `field` (packed images invalidated)."""
    sd, _ = load_state(path)
    dev = field.device
    mv = lambda st: [(W.to(dev), None if b is None else b.to(dev)) for W, b in st]
    for tag, name in (("specular", "color"), ("diffuse", "diffuse"), ("renv", "renv")):
        if tag not in resume_mlps:
            continue
        st = stack_from_state(sd, name)
        if st is None:
            if tag == "renv":
                continue
            raise EnvidrError(f"{path}: no {name}_net in the checkpoint")
        setattr(field, name, mv(st))
    field._packed = None
    return field


def swap_env(field: FieldParams, path: str) -> FieldParams:
    """opt.swap_env_path (utils.py:1583-1597) / the relight sweep: replace env_net by a shipped environment MLP
    (ckpts/env_ckpts/env_net_<k>.pth, either key spelling); the IDE degree follows the MLP's input width."""
    sd, _ = load_state(path)
    st = stack_from_state(sd, "env")
    if st is None:
        raise EnvidrError(f"{path}: no env_net in the checkpoint")
    if st[-1][0].shape[0] != field.env[-1][0].shape[0]:
        raise EnvidrError("swap_env: environment feature width differs from the field's")
    dev = field.device
    field.env = [(W.to(dev), None if b is None else b.to(dev)) for W, b in st]
    field.ide_degree = ide_degree_from_env(st)
    field._packed = None
    return field


def state_from_field(field: FieldParams, density=None) -> "collections.OrderedDict[str, torch.Tensor]":
    """The reference's state_dict keys for `field` (+ the occupancy-grid buffers of an envidr_b200.density.DensityGrid):
    what `Trainer.save_checkpoint` stores under 'model', loadable by the reference's load_state_dict(strict=False)."""
    sd = collections.OrderedDict()
    sd["encoder.embeddings"] = field.embeddings.detach().cpu()
    sd["encoder.offsets"] = field.offsets.detach().cpu()
    for name, st in field.stacks().items():
        if st is None:
            continue
        for i, (W, b) in enumerate(st):
            sd[f"{name}_net.{i}.weight"] = W.detach().cpu()
            if b is not None:
                sd[f"{name}_net.{i}.bias"] = b.detach().cpu()
    sd["sdf_density.beta"] = torch.tensor(float(field.beta))
    if density is not None:
        sd["density_grid"] = density.density_grid.detach().cpu()
        sd["density_bitfield"] = density.density_bitfield.detach().cpu()
        sd["step_counter"] = density.step_counter.detach().cpu()
    return sd


def save_checkpoint(path: str, field: FieldParams, density=None, *, epoch: int = 0, global_step: int = 0, optimizer=None, ema=None,
                    stats: Optional[dict] = None) -> None:
    """utils.py:1478-1562 (the 'full' branch): a checkpoint the reference's Trainer.load_checkpoint accepts."""
    state = {"epoch": epoch, "global_step": global_step, "stats": stats or {"loss": [], "valid_loss": [], "results": [], "checkpoints": [],
                                                                          "best_result": None}}
    if density is not None:
        state["mean_count"] = int(density.mean_count)
        state["mean_density"] = float(density.mean_density)
    if optimizer is not None:
        state["optimizer"] = optimizer.state_dict()
    if ema is not None:
        state["ema"] = ema.state_dict()
    state["model"] = state_from_field(field, density)
    torch.save(state, path)


def load_checkpoint(path: str, *, device="cpu", density=None, template: Optional[FieldParams] = None, **field_kwargs):
    """-> (FieldParams on `device`, meta dict).  With `density` (a DensityGrid / reference-shaped object) its buffers and
    mean_count / mean_density are restored as Trainer.load_checkpoint does (utils.py:1623-1627)."""
    sd, meta = load_state(path)
    field = field_from_state(sd, template=template, **field_kwargs)
    if str(device) != "cpu":
        field = field.to(device)
    if density is not None:
        for k in ("density_grid", "density_bitfield", "step_counter"):
            if k in sd:
                getattr(density, k).copy_(sd[k].to(getattr(density, k).device))
        if "mean_count" in meta:
            density.mean_count = meta["mean_count"]
        if "mean_density" in meta:
            density.mean_density = meta["mean_density"]
    return field, meta


class ExponentialMovingAverage:
    """torch_ema.ExponentialMovingAverage(parameters, decay, use_num_updates=True): update / store / copy_to / restore /
    state_dict / load_state_dict with its semantics and state layout."""

    def __init__(self, parameters: Iterable[torch.Tensor], decay: float, use_num_updates: bool = True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self._params = [p for p in parameters]
        self.shadow_params = [p.clone().detach() for p in self._params]
        self.collected_params = None

    def _get(self, parameters):
        return self._params if parameters is None else list(parameters)

    @torch.no_grad()
    def update(self, parameters=None) -> None:
        params = self._get(parameters)
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        one_minus_decay = 1.0 - decay
        for s, p in zip(self.shadow_params, params):
            tmp = s - p
            tmp.mul_(one_minus_decay)
            s.sub_(tmp)

    @torch.no_grad()
    def copy_to(self, parameters=None) -> None:
        for s, p in zip(self.shadow_params, self._get(parameters)):
            p.data.copy_(s.data)

    def store(self, parameters=None) -> None:
        self.collected_params = [p.clone() for p in self._get(parameters)]

    @torch.no_grad()
    def restore(self, parameters=None) -> None:
        if self.collected_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        for c, p in zip(self.collected_params, self._get(parameters)):
            p.data.copy_(c.data)

    def state_dict(self) -> dict:
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params}

    def load_state_dict(self, state_dict: dict) -> None:
        self.decay = state_dict["decay"]
        self.num_updates = state_dict["num_updates"]
        sp = state_dict["shadow_params"]
        if len(sp) != len(self._params) or any(a.shape != b.shape for a, b in zip(sp, self._params)):
            raise ValueError("shadow_params do not match the parameters")       # the reference re-initialises on failure (utils.py:1632-1636)
        self.shadow_params = [a.to(device=p.device, dtype=p.dtype).clone() for a, p in zip(sp, self._params)]
        cp = state_dict.get("collected_params")
        self.collected_params = None if cp is None else [a.to(device=p.device, dtype=p.dtype).clone() for a, p in zip(cp, self._params)]
