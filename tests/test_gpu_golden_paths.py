"""The CUDA path against the reference's OWN Python, directly: the images / outputs / gradients that the reference's
NeRFRenderer.render -> run_cuda (inference loop and training branch) produced on the CPU with the oracle's operators injected
(tests/golden/make_golden.py::gen_infer_branch / gen_train_branch; golden infer_branch.npz, train_branch.npz) are compared with
what envidr_b200.render.render and envidr_b200.train produce on the B200 from the same weights and rays.
Tolerances: RGB 1e-4 on all but <= 2 pixels (ReLU-mask kinks of the normal), weights_sum / depth 5e-5.  The training
comparison is per SAMPLE (no compositing average) and meets the one documented arithmetic difference head on: the device's exp2f gives
per-level scales one ulp off libm's (deviation 1), worth ~1e-5 in an encoder output and ~1e-3 relative in its
derivative, and flips the ReLU mask of a sample sitting at a kink; hence sdf 5e-5, image 2e-4, sdf gradients 5e-3 on all but <= 3
samples, parameter gradients within 3e-2 of each tensor's scale (the tight
comparison, with the device's level scales installed in the oracle, is tests/test_gpu_train.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _field_from_golden(z, beta):
    from envidr_b200.field import FieldParams
    t = lambda a: torch.from_numpy(np.asarray(a, np.float32))
    n = lambda name: sorted({int(k.split("_")[2]) for k in z.files if k.startswith(name + "_") and k.endswith("_weight")})
    st = lambda name: [(t(z[f"{name}_{i}_weight"]), t(z[f"{name}_{i}_bias"])) for i in n(name)]
    return FieldParams(embeddings=t(z["embeddings"]), offsets=torch.from_numpy(z["offsets"].astype(np.int32)),
                       per_level_scale=float(z["per_level_scale"]), base_resolution=int(z["base_resolution"]), bound=1.0,
                       sdf=st("sdf_net"), env=st("env_net"), diffuse=st("diffuse_net"), color=st("color_net"), renv=st("renv_net"),
                       geo_feat_dim=12, ide_degree=5, beta=beta, beta_min=float(z["opt_beta_min"]), beta_max=float(z["opt_beta_max"]))


@pytest.mark.parametrize("tag,indir,rot", [("one", False, None), ("three", True, None), ("rot", True, 0.7)])
@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_cuda_render_matches_reference_python_images(dev, golden_dir, tag, indir, rot, precision):
    from envidr_b200 import render, scene
    z = np.load(os.path.join(golden_dir, "infer_branch.npz"))
    fp = _field_from_golden(z, float(z["beta"]))
    fp.precision = precision
    fp = fp.to(dev).pack()
    bf = torch.from_numpy(scene.make_bitfield()).to(dev)
    ro, rd = torch.from_numpy(z["rays_o"]).to(dev), torch.from_numpy(z["rays_d"]).to(dev)
    cfg = render.RenderConfig(indir_ref=indir, max_steps=256, indir_max_steps=256)
    res = render.render(fp, bf, ro, rd, cfg, bg_color=1.0, get_normal_image=True, env_rot_radian=rot,
                        visual_items=("diffuse", "specular", "roughness"))
    e = (res["image"].cpu().numpy() - z[f"{tag}_image"]).__abs__().max(-1)
    assert int((e > 1e-4).sum()) <= 2 and float(np.median(e)) <= 2e-5, (int((e > 1e-4).sum()), float(e.max()))
    np.testing.assert_allclose(res["weights_sum"].cpu().numpy(), z[f"{tag}_weights_sum"][:, 0], rtol=0, atol=5e-5)
    np.testing.assert_allclose(res["depth"].cpu().numpy(), z[f"{tag}_depth"][:, 0], rtol=0, atol=5e-5)
    en = np.abs(res["normal_image"].cpu().numpy() - z[f"{tag}_normal_image"]).max(-1)
    assert int((en > 1e-3).sum()) <= 2
    for k in ("diffuse_image", "specular_image", "roughness_image"):
        ek = np.abs(res[k].cpu().numpy().reshape(z[f"{tag}_{k}"].shape) - z[f"{tag}_{k}"]).max(-1)
        assert int((ek > 1e-4).sum()) <= 2, (k, int((ek > 1e-4).sum()), float(ek.max()))


@pytest.mark.parametrize("tag", ["single", "renv"])
def test_cuda_train_branch_matches_reference_python(dev, golden_dir, tag):
    from envidr_b200 import render, scene, train
    z = np.load(os.path.join(golden_dir, "train_branch.npz"))
    fp = _field_from_golden(z, float(z["beta"]))
    field = train.TrainableField(fp.to(dev), frozen=())
    bf = torch.from_numpy(scene.make_bitfield()).to(dev)
    ro, rd = torch.from_numpy(z["rays_o"]).to(dev), torch.from_numpy(z["rays_d"]).to(dev)
    ri = torch.from_numpy(z["r_images"]).to(dev) if tag == "renv" else None
    cfg = render.RenderConfig(max_steps=256)
    out = train.render_train(field, bf, ro, rd, cfg, bg_color=1.0, perturb=False, force_all_rays=True, r_images=ri)
    f = lambda t: t.detach().cpu().numpy()
    assert out["sdfs"].shape[0] == z[f"{tag}_sdfs"].shape[0]
    np.testing.assert_allclose(f(out["sdfs"]), z[f"{tag}_sdfs"], rtol=1e-3, atol=5e-5)
    np.testing.assert_allclose(f(out["image"]), z[f"{tag}_image"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(f(out["weights_sum"]), z[f"{tag}_weights_sum"], rtol=0, atol=1e-4)
    # a sample whose hidden pre-activation sits at a ReLU kink flips its mask under the one-ulp level-scale difference and its
    # gradient jumps (random weights here: many kinks); such samples are counted, not averaged away
    eg = np.abs(f(out["sdf_gradients"]) - z[f"{tag}_sdf_gradients"]).max(-1)
    tol = 2e-3 + 5e-3 * np.abs(z[f"{tag}_sdf_gradients"]).max(-1)
    assert int((eg > tol).sum()) <= 3, (int((eg > tol).sum()), float(eg.max()))
    sg = out["sdf_gradients"]
    loss = 0.3 * out["image"].sum() + (out["weights_sum"] ** 2).sum() + 5 * (out["sdfs"] ** 2).mean() + ((sg.norm(dim=-1) - 1) ** 2).mean()
    assert abs(float(loss.detach()) - float(z[f"{tag}_loss"])) <= 2e-3 * abs(float(z[f"{tag}_loss"]))
    loss.backward()
    names = {"sdf_w0": "sdf_net_0_weight", "sdf_b2": "sdf_net_2_bias", "env_w1": "env_net_1_weight", "renv_w0": "renv_net_0_weight",
             "embeddings": "encoder_embeddings", "beta": "sdf_density_beta"}
    for k, gk in names.items():
        ref = z[f"{tag}_grad_{gk}"]
        g = getattr(field, k).grad
        if ref.size == 0 or (tag == "single" and k.startswith("renv")):
            continue
        got = g.detach().cpu().numpy().reshape(ref.shape)
        scale = float(np.abs(ref).max())
        assert float(np.abs(got - ref).max()) <= 3e-2 * scale + 1e-9, (k, float(np.abs(got - ref).max()), scale)
