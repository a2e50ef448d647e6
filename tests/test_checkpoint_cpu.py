"""Checkpoint / weight-format loader and EMA (envidr_b200/checkpoint.py) -- host logic, runs on the CPU.

  * the reference's own state_dict contract (tests/golden/state_dict_keys.json, produced by instantiating the reference NeRFNetwork
    under configs/scenes/toaster.ini): what we save has the reference's key names and shapes, what we load is read from them;
  * the shipped published-format checkpoints (ckpts/rendering_mlps.pth, ckpts/env_ckpts/env_net_3.pth, only where /root/reference
    exists) against the golden arrays tests/golden/relight_mlps.npz made from them by the reference's own loader;
  * round trips, both spellings of the environment-MLP keys, the error behaviour, EMA against its published update rule."""
import json
import os
import types

import numpy as np
import pytest
import torch

from envidr_b200 import checkpoint as C
from envidr_b200 import scene
from envidr_b200._lib import EnvidrError

REF = os.environ.get("ENVIDR_REFERENCE", "/root/reference")


def _small_field(**kw):
    return scene.make_synthetic_field(0, hidden_dim_env=64, ide_degree=4, num_levels=4, log2_hashmap_size=10, desired_resolution=64, **kw)


def _density_like():
    return types.SimpleNamespace(density_grid=torch.rand(1, 4096), density_bitfield=torch.randint(0, 255, (512,), dtype=torch.uint8),
                                 step_counter=torch.zeros(16, 2, dtype=torch.int32), mean_count=123, mean_density=0.75)


def test_saved_state_has_the_reference_keys_and_shapes(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    fp = scene.make_synthetic_field(0)                                  # toaster.ini dimensions
    sd = C.state_from_field(fp)
    for k, v in sd.items():
        assert k in ref["keys"], k
        assert list(v.shape) == ref["keys"][k], (k, tuple(v.shape), ref["keys"][k])
    # everything the reference's state_dict holds for the path is there (aabb buffers are options, not state of the path)
    missing = set(ref["keys"]) - set(sd) - {"aabb_train", "aabb_infer", "density_grid", "density_bitfield", "step_counter"}
    assert not missing, missing
    assert abs(fp.per_level_scale - ref["per_level_scale"]) < 1e-12 and fp.base_resolution == ref["base_resolution"]
    back = C.field_from_state(sd)
    assert abs(back.per_level_scale - ref["per_level_scale"]) < 1e-12 and back.ide_degree == 5 and back.geo_feat_dim == 12
    assert torch.equal(back.embeddings, fp.embeddings) and torch.equal(back.offsets, fp.offsets)


def test_round_trip_with_density_state_optimizer_and_ema(tmp_path):
    fp = _small_field()
    fp.beta = 0.0123
    dens = _density_like()
    params = [torch.nn.Parameter(W.clone()) for W, _ in fp.sdf]
    opt = torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.99), eps=1e-15)
    for p in params:
        p.grad = torch.ones_like(p)
    opt.step()
    ema = C.ExponentialMovingAverage(params, 0.95)
    ema.update()
    path = str(tmp_path / "ngp_ep0001.pth")
    C.save_checkpoint(path, fp, dens, epoch=1, global_step=16, optimizer=opt, ema=ema)
    raw = torch.load(path, map_location="cpu")
    assert set(raw) >= {"model", "epoch", "global_step", "stats", "mean_count", "mean_density", "optimizer", "ema"}     # utils.py:1478-1505
    dens2 = _density_like()
    dens2.density_grid.zero_(); dens2.mean_count = 0; dens2.mean_density = 0
    back, meta = C.load_checkpoint(path, density=dens2, base_resolution=16, desired_resolution=64)
    assert meta["epoch"] == 1 and meta["global_step"] == 16
    for name in C.STACKS:
        for (W, b), (W2, b2) in zip(getattr(fp, name), getattr(back, name)):
            assert torch.equal(W, W2) and torch.equal(b, b2)
    assert torch.equal(back.embeddings, fp.embeddings) and abs(back.beta - 0.0123) < 1e-7
    assert abs(back.per_level_scale - fp.per_level_scale) < 1e-12 and back.num_levels == 4 and back.ide_degree == 4
    assert torch.equal(dens2.density_grid, dens.density_grid) and torch.equal(dens2.density_bitfield, dens.density_bitfield)
    assert dens2.mean_count == 123 and dens2.mean_density == 0.75
    ema2 = C.ExponentialMovingAverage([torch.nn.Parameter(p.detach().clone()) for p in params], 0.5)
    ema2.load_state_dict(meta["ema"])
    assert ema2.decay == 0.95 and ema2.num_updates == 1 and all(torch.equal(a, b) for a, b in zip(ema2.shadow_params, ema.shadow_params))
    # a bare state_dict is accepted too (utils.py:1577-1581)
    torch.save(raw["model"], str(tmp_path / "bare.pth"))
    sd, meta2 = C.load_state(str(tmp_path / "bare.pth"))
    assert meta2 == {} and "sdf_net.0.weight" in sd


def test_rendering_mlps_and_env_swap_in_both_key_spellings(tmp_path):
    fp = _small_field()
    g = torch.Generator().manual_seed(1)
    mk = lambda dims: [(torch.randn(o, i, generator=g), torch.randn(o, generator=g)) for i, o in zip(dims[:-1], dims[1:])]
    color, diffuse, renv = mk([28, 64, 64, 3]), mk([24, 32, 3]), mk([4, 64, 64, 64, 12])
    model = {}
    for name, st in (("color_net", color), ("diffuse_net", diffuse), ("renv_net", renv)):
        for i, (W, b) in enumerate(st):
            model[f"{name}.{i}.weight"], model[f"{name}.{i}.bias"] = W, b
    torch.save({"model": model}, str(tmp_path / "rendering_mlps.pth"))
    C.load_rendering_mlps(fp, str(tmp_path / "rendering_mlps.pth"), resume_mlps=("specular", "renv"))
    assert torch.equal(fp.color[1][0], color[1][0]) and torch.equal(fp.renv[3][1], renv[3][1])
    assert not torch.equal(fp.diffuse[0][0], diffuse[0][0])            # 'diffuse' was not in resume_mlps
    # without renv in the file: skipped, as the reference does (utils.py:522-529); without color: an error
    torch.save({"model": {k: v for k, v in model.items() if not k.startswith("renv")}}, str(tmp_path / "no_renv.pth"))
    C.load_rendering_mlps(fp, str(tmp_path / "no_renv.pth"))
    assert torch.equal(fp.diffuse[0][0], diffuse[0][0])
    torch.save({"model": {k: v for k, v in model.items() if k.startswith("renv")}}, str(tmp_path / "only_renv.pth"))
    with pytest.raises(EnvidrError):
        C.load_rendering_mlps(fp, str(tmp_path / "only_renv.pth"))
    # environment MLP: sph_loader.py:356-378 writes 'env_net0.weight'; a full checkpoint holds 'env_net.0.weight'
    env5 = mk([72, 96, 96, 12])
    quirk = {f"env_net{i}.{n}": t for i, (W, b) in enumerate(env5) for n, t in (("weight", W), ("bias", b))}
    dotted = {f"env_net.{i}.{n}": t for i, (W, b) in enumerate(env5) for n, t in (("weight", W), ("bias", b))}
    for j, keys in enumerate((quirk, dotted)):
        f2 = _small_field()
        assert f2.ide_degree == 4
        torch.save({"model": keys}, str(tmp_path / f"env_{j}.pth"))
        C.swap_env(f2, str(tmp_path / f"env_{j}.pth"))
        assert f2.ide_degree == 5 and torch.equal(f2.env[2][0], env5[2][0]) and f2._packed is None
    bad = {f"env_net{i}.{n}": t for i, (W, b) in enumerate(mk([40, 64, 12])) for n, t in (("weight", W), ("bias", b))}
    torch.save({"model": bad}, str(tmp_path / "env_bad.pth"))
    with pytest.raises(EnvidrError):
        C.swap_env(_small_field(), str(tmp_path / "env_bad.pth"))          # 40 is not an IDE width
    wide = {f"env_net{i}.{n}": t for i, (W, b) in enumerate(mk([38, 64, 16])) for n, t in (("weight", W), ("bias", b))}
    torch.save({"model": wide}, str(tmp_path / "env_wide.pth"))
    with pytest.raises(EnvidrError):
        C.swap_env(_small_field(), str(tmp_path / "env_wide.pth"))         # feature width 16 != 12
    with pytest.raises(EnvidrError):
        C.field_from_state({"sdf_net.0.weight": torch.zeros(4, 4)})         # no encoder


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ckpts")), reason="shipped checkpoints live in the reference tree")
def test_shipped_checkpoints_match_reference_loaded_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "relight_mlps.npz"))
    fp = _small_field()
    C.load_rendering_mlps(fp, os.path.join(REF, "ckpts", "rendering_mlps.pth"))
    C.swap_env(fp, os.path.join(REF, "ckpts", "env_ckpts", "env_net_3.pth"))
    assert fp.ide_degree == 4 and [tuple(W.shape) for W, _ in fp.env] == [(160, 38), (160, 160), (160, 160), (12, 160)]
    for name in ("env", "color", "diffuse", "renv"):
        for i, (W, b) in enumerate(getattr(fp, name)):
            np.testing.assert_array_equal(W.numpy(), z[f"{name}_net_{i}_weight"])
            np.testing.assert_array_equal(b.numpy(), z[f"{name}_net_{i}_bias"])


def test_ema_follows_the_published_rule():
    g = torch.Generator().manual_seed(0)
    p = [torch.nn.Parameter(torch.randn(5, 3, generator=g)), torch.nn.Parameter(torch.randn(7, generator=g))]
    ema = C.ExponentialMovingAverage(p, decay=0.95)
    shadow = [q.detach().clone().double() for q in p]
    for n in range(1, 40):
        with torch.no_grad():
            for q in p:
                q.add_(torch.randn(q.shape, generator=g) * 0.1)
        ema.update()
        d = min(0.95, (1 + n) / (10 + n))
        shadow = [s - (1 - d) * (s - q.detach().double()) for s, q in zip(shadow, p)]
        for s, e in zip(shadow, ema.shadow_params):
            assert float((s - e.double()).abs().max()) < 1e-5
    assert ema.num_updates == 39
    live = [q.detach().clone() for q in p]
    ema.store(); ema.copy_to()
    assert all(torch.equal(q.detach(), e) for q, e in zip(p, ema.shadow_params))
    ema.restore()
    assert all(torch.equal(q.detach(), l) for q, l in zip(p, live))
    with pytest.raises(ValueError):
        C.ExponentialMovingAverage([torch.nn.Parameter(torch.zeros(2))], 0.9).load_state_dict(ema.state_dict())
