"""The reference's OWN Python classes (NeRFNetwork / NeRFRenderer.render / run_cuda / Trainer.train_step) as a checker --
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/oracle.py header; nothing under envidr_b200/ may import this).

Where the reference code comes from:
  * in the build container: /root/reference, imported from where it lies;
  * on the GPU box (no /root/reference there): oracle/_ref/py/, a git-ignored copy of the reference's *.py / *.ini files made
    by oracle/build_ref.py next to the rebuilt extension modules oracle/_ref/_*.so (both travel with the snapshot, neither is
    committed).
What this module adds is only what is needed to import and drive it unmodified:
  * stubs for third-party imports the hot path never calls (trimesh, open3d, lpips, tensorboardX, ...), `numpy.math`
    (removed in numpy 2; ide_encoder.py:8,27-41 uses it), a 25-line configargparse (argparse + .ini reader);
  * `<pkg>._ext._<pkg>` resolved to oracle/_ref/_<pkg>.so (the reference's build_ext.sh layout);
  * `use_backends("reference" | "envidr")`: the wrappers look `_backend` up as a module global at call time, so the SAME
    reference model can be run on the reference's kernels and on libenvidr_b200.so in one process;
  * `build_model` = the constructor call of main_nerf.py:64-75 under configs/scenes/toaster.ini, `load_field` copies a
    synthetic field (envidr_b200.scene) into it, `eval_kwargs` = the arguments of Trainer.eval_step (utils.py:857-859),
    `train_step` = Trainer.train_step (utils.py:560-808) called on a stub `self` holding the real model.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import types
from typing import Dict, Optional

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
PKGS = ["raymarching", "hashencoder", "gridencoder", "freqencoder", "shencoder"]
_WRAPPERS = {"raymarching": "raymarching.raymarching", "hashencoder": "hashencoder.hashgrid", "gridencoder": "gridencoder.grid",
             "freqencoder": "freqencoder.freq", "shencoder": "shencoder.sphere_harmonics"}


def ref_root() -> Optional[str]:
    for p in (os.environ.get("ENVIDR_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref", "py")):
        if p and os.path.isdir(os.path.join(p, "nerf")):
            return p
    return None


def available() -> bool:
    return ref_root() is not None and os.path.exists(os.path.join(HERE, "_ref", "_raymarching.so"))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Stub(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return _Stub("call")


def _configargparse() -> types.ModuleType:
    """argparse + `--config file.ini` (key = value lines, [a, b] lists, True/False flags): what nerf/options.py needs."""
    cap = types.ModuleType("configargparse")

    class ArgumentParser(argparse.ArgumentParser):
        def add_argument(self, *a, **k):
            k.pop("is_config_file", None)
            return super().add_argument(*a, **k)

        def parse_args(self, args=None, namespace=None):
            args = list(sys.argv[1:] if args is None else args)
            extra = []
            if "--config" in args:
                path = args[args.index("--config") + 1]
                for line in open(path):
                    line = line.split("#")[0].split(";")[0].strip()
                    if not line or "=" not in line:
                        continue
                    key, val = [s.strip() for s in line.split("=", 1)]
                    if val in ("True", "true"):
                        extra.append("--" + key)
                    elif val in ("False", "false"):
                        continue
                    elif val.startswith("["):
                        extra += ["--" + key] + [v.strip() for v in val.strip("[]").split(",") if v.strip()]
                    else:
                        extra += ["--" + key, val]
            return super().parse_args(extra + args, namespace)

    cap.ArgumentParser = ArgumentParser
    return cap


_ref_backends: Dict[str, object] = {}


def install_shims(load_ref_so: bool = True) -> str:
    """Make the reference tree importable (idempotent).  Returns its root."""
    root = ref_root()
    if root is None:
        raise ImportError("reference tree not found (neither /root/reference nor oracle/_ref/py; run oracle/build_ref.py)")
    np.math = math
    for name in ["trimesh", "imageio", "tensorboardX", "mcubes", "lpips", "open3d", "open3d.visualization",
                 "open3d.visualization.rendering", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "dearpygui",
                 "dearpygui.dearpygui", "torch_ema", "cv2", "torchvision", "torchvision.transforms", "torchvision.utils", "PIL"]:
        if name in sys.modules:
            continue
        try:                                            # the real package where this image has it
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    if isinstance(sys.modules["torch_ema"], _Stub):
        sys.modules["torch_ema"].ExponentialMovingAverage = object
    if "configargparse" not in sys.modules:
        sys.modules["configargparse"] = _configargparse()
    so_dir = os.path.join(HERE, "_ref")
    for p in (so_dir, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    for pkg in PKGS:
        if f"{pkg}._ext" in sys.modules and pkg in _ref_backends:
            continue
        ext = types.ModuleType(f"{pkg}._ext")
        mod = None
        if load_ref_so:
            try:
                mod = __import__(f"_{pkg}")
            except Exception:
                mod = None
        if mod is None:
            mod = _Stub(f"_{pkg}")
        _ref_backends[pkg] = mod
        if f"{pkg}._ext" not in sys.modules:         # somebody (envidr_b200.backend.install_into_sys_modules) may have been first
            setattr(ext, f"_{pkg}", mod)
            sys.modules[f"{pkg}._ext"] = ext
            sys.modules[f"{pkg}._ext._{pkg}"] = mod
    return root


def use_backends(which: str) -> None:
    """Point the `_backend` global of every imported reference wrapper at the reference's own kernels ("reference") or at
    libenvidr_b200.so through envidr_b200.backend ("envidr")."""
    import importlib
    if which == "envidr":
        from envidr_b200 import backend as B
        table = {"raymarching": B._raymarching, "hashencoder": B._hashencoder, "gridencoder": B._gridencoder,
                 "freqencoder": B._freqencoder, "shencoder": B._shencoder}
    elif which == "reference":
        table = dict(_ref_backends)
    else:
        raise ValueError(which)
    for pkg, modname in _WRAPPERS.items():
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
        mod._backend = table[pkg]


def parse_opt(extra_args=(), config: str = "configs/scenes/toaster.ini"):
    root = install_shims()
    from nerf.options import config_parser
    argv = sys.argv
    sys.argv = ["x", "--config", os.path.join(root, config)] + list(extra_args)
    try:
        opt = config_parser()
    finally:
        sys.argv = argv
    return opt


def build_model(extra_args=(), cuda_ray: bool = True, config: str = "configs/scenes/toaster.ini", seed: int = 0):
    """main_nerf.py:64-75."""
    opt = parse_opt(extra_args, config)
    opt.cuda_ray = cuda_ray
    from nerf.network import NeRFNetwork
    torch.manual_seed(seed)
    model = NeRFNetwork(encoding="hashgrid", encoding_dir=opt.encoding_dir, bound=opt.bound, cuda_ray=cuda_ray,
                        density_scale=1, min_near=opt.min_near, density_thresh=opt.density_thresh, bg_radius=opt.bg_radius,
                        use_sdf=opt.use_sdf, hidden_dim=opt.hidden_dim, num_layers=opt.num_layers,
                        num_layers_color=opt.num_layers_color, hidden_dim_color=opt.hidden_dim_color,
                        num_layers_bg=opt.num_layers_bg, num_levels=opt.num_levels, geo_feat_dim=opt.geo_feat_dim,
                        opt=opt, env_opt=None)
    return model, opt


def load_field(model, fp, bitfield: Optional[np.ndarray] = None):
    """Copy a synthetic field (envidr_b200.field.FieldParams on the CPU) into the reference model's parameters / buffers."""
    enc = model.encoder
    assert tuple(fp.embeddings.shape) == tuple(enc.embeddings.shape), (fp.embeddings.shape, enc.embeddings.shape)
    assert np.array_equal(fp.offsets.cpu().numpy(), enc.offsets.cpu().numpy())
    with torch.no_grad():
        enc.embeddings.copy_(fp.embeddings)
        for name in ("sdf", "env", "diffuse", "color", "renv"):
            st = getattr(fp, name)
            net = getattr(model, name + "_net", None)
            if st is None or net is None:
                continue
            assert len(net) == len(st), name
            for lin, (W, b) in zip(net, st):
                lin.weight.copy_(W)
                lin.bias.copy_(b)
        model.sdf_density.beta.fill_(float(fp.beta))
        if bitfield is not None:
            model.density_bitfield.copy_(torch.from_numpy(np.asarray(bitfield)).to(model.density_bitfield.device))
    return model


def eval_kwargs(opt) -> dict:
    """Trainer.eval_step's call (utils.py:857-859): model.render(rays_o, rays_d, **eval_kwargs(opt))."""
    return dict(staged=True, bg_color=1 if opt.render_bg_color == "white" else 0, perturb=False, get_normal_image=True,
                use_specular_color=True, env_net_index=None, epoch=500, material=None, r_images=None, **vars(opt))


class _FakeTrainer:
    pass


def train_step(model, opt, rays_o, rays_d, images, epoch: int = 500):
    """Trainer.train_step (utils.py:560-808) on a stub `self` that carries the real model: returns (pred_rgb, gt_rgb, loss, loss_dict)."""
    install_shims()
    import nerf.utils as U
    fake = _FakeTrainer()
    fake.opt, fake.model, fake.device, fake.error_map = opt, model, rays_o.device, None
    fake.criterion = torch.nn.L1Loss(reduction="none") if getattr(opt, "color_l1_loss", False) else torch.nn.MSELoss(reduction="none")
    fake.epoch, fake.global_step = epoch, epoch
    data = dict(rays_o=rays_o, rays_d=rays_d, images=images.clone())
    return U.Trainer.train_step(fake, data)
