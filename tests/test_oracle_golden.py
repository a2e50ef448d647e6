"""Pin the torch layer of the oracle against golden vectors produced by the reference's own Python
code (tests/golden/make_golden.py: ide_encoder.py, demo.ipynb path, nerf/network.py + nerf/renderer.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O


def _layers(z, prefix, idxs):
    return [(z[f"{prefix}_{i}_weight"], z[f"{prefix}_{i}_bias"]) for i in idxs]


def _P_from_glue(z, deg):
    n = lambda name: sorted({int(k.split("_")[2]) for k in z.files if k.startswith(name + "_")})
    P = dict(bound=1.0, geo_feat_dim=12, roughness_act_scale=0.2, roughness_bias=-1.0, roughness_scale=1.0,
             beta=float(z["beta"]), beta_min=float(z["opt_beta_min"]) if "opt_beta_min" in z.files else 0.0005,
             beta_max=float(z["opt_beta_max"]) if "opt_beta_max" in z.files else 1.0,
             ide_degree=deg, diffuse_kappa_inv=0.64, light_intensity_scale=1.0, intensity_scale=1.0,
             indir_roughness_thresh=0.1, learn_indir_blend=True,
             sdf=_layers(z, "sdf_net", n("sdf_net")), env=_layers(z, "env_net", n("env_net")),
             diffuse=_layers(z, "diffuse_net", n("diffuse_net")), color=_layers(z, "color_net", n("color_net")),
             renv=_layers(z, "renv_net", n("renv_net")) if any(k.startswith("renv_net") for k in z.files) else None)
    return P


def _standin(z):
    xyz = torch.from_numpy(z["xyz"]).double()
    A, b = torch.from_numpy(z["A"]).double(), torch.from_numpy(z["b"]).double()
    ph = xyz @ A + b
    enc = torch.sin(ph)
    jac = torch.cos(ph)[:, :, None] * A.T[None]          # [M,F,3]
    return enc.numpy(), jac.numpy()


@pytest.mark.parametrize("deg", [4, 5])
def test_ide_matches_reference(golden_dir, deg):
    z = np.load(os.path.join(golden_dir, "ide.npz"))
    ml, mat, sigma = O.ide_tables(deg)
    np.testing.assert_array_equal(mat, z[f"mat{deg}"])
    np.testing.assert_array_equal(ml.astype(np.float32), z[f"ml{deg}"])
    np.testing.assert_array_equal(sigma, z[f"sigma{deg}"])
    d = torch.from_numpy(z[f"dirs{deg}"])
    l = np.concatenate([z[f"ml{deg}"][1]] * 2)
    # fp32 restatement (what the reference computes): the l=16 band cancels catastrophically in the
    # power basis, so even two fp32 evaluations of the same formula differ by ~1e-4 there.
    atol32 = np.where(l >= 16, 2e-4, 5e-6)[None]
    atol64 = np.where(l >= 16, 1e-3, 5e-6)[None]
    for rough, key in ((torch.from_numpy(z[f"rough{deg}"]), "var"), (0.64, "const")):
        ref = z[f"ide{deg}_{key}"]
        got = O.ide_encode(d.float(), rough, deg).numpy()
        assert (np.abs(got - ref) <= atol32 + 1e-5 * np.abs(ref)).all()
        r64 = rough.double() if torch.is_tensor(rough) else rough
        got = O.ide_encode(d.double(), r64, deg).float().numpy()
        assert (np.abs(got - ref) <= atol64 + 1e-5 * np.abs(ref)).all()
    assert got.shape[1] == (2 ** deg - 1 + deg) * 2


@pytest.mark.parametrize("tag", ["plain", "rot", "renv"])
def test_field_glue_matches_reference_network(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, "field_glue.npz"))
    P = _P_from_glue(z, int(z["opt_ide_degree"]))
    kw = {}
    if tag == "rot":
        kw["env_rot_radian"] = 0.7
    if tag == "renv":
        kw["r_images"] = z["r_images"]
    out = O.field_forward(P, z["xyz"], z["dirs"], enc_override=_standin(z), ide_dtype=torch.float32, **kw)
    tol = dict(atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(out["sdf"], z[f"{tag}_sdf"], **tol)
    np.testing.assert_allclose(out["sigma"], z[f"{tag}_sigma"], atol=1e-4, rtol=2e-4)
    np.testing.assert_allclose(out["geo_feat"], z[f"{tag}_geo"], **tol)
    np.testing.assert_allclose(out["normal"], z[f"{tag}_normal"], **tol)
    np.testing.assert_allclose(out["roughness"], z[f"{tag}_roughness"], **tol)
    # IDE l=16 band: fp32 cancellation noise of the reference formula itself (see test_ide_matches_reference)
    np.testing.assert_allclose(out["w_r_enc"], z[f"{tag}_w_r_enc"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(out["n_env_enc"], z[f"{tag}_n_env_enc"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(out["c_diffuse"], z[f"{tag}_c_diffuse"], **tol)
    np.testing.assert_allclose(out["c_specular"], z[f"{tag}_c_specular"], **tol)
    np.testing.assert_allclose(out["rgb"], z[f"{tag}_rgb"], **tol)


def test_field_relight_dims(golden_dir):
    z = np.load(os.path.join(golden_dir, "relight_mlps.npz"))
    P = _P_from_glue(z, 4)
    out = O.field_forward(P, z["xyz"], z["dirs"], enc_override=_standin(z), ide_dtype=torch.float32)
    np.testing.assert_allclose(out["rgb"], z["plain_rgb"], atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(out["normal"], z["plain_normal"], atol=3e-5, rtol=1e-4)


def test_demo_sphere_config1(golden_dir):
    """BASELINE config 1 (demo.ipynb): IDE x2 + env_net x2 + diffuse + specular on the sphere hits."""
    z = np.load(os.path.join(golden_dir, "demo_sphere.npz"))
    env = [(z[f"env_{i}_weight"], z[f"env_{i}_bias"]) for i in (0, 2, 4, 6)]
    dif = [(z[f"diffuse_{i}_weight"], z[f"diffuse_{i}_bias"]) for i in (0, 2)]
    spe = [(z[f"specular_{i}_weight"], z[f"specular_{i}_bias"]) for i in (0, 2, 4)]
    n = torch.from_numpy(z["xyz"]).double()
    d = torch.from_numpy(z["dirs"]).double()
    geo = torch.from_numpy(z["geo"]).double().expand(n.shape[0], -1)
    w_o = -d
    ndv = (n * w_o).sum(-1, keepdim=True)
    w_r = 2 * ndv * n - w_o
    f_n = O._unit(O._mlp(O.ide_encode(n, 0.64, 4), env), 1e-12)
    f_r = O._unit(O._mlp(O.ide_encode(w_r, float(z["kappa_inv"]), 4), env), 1e-12)
    c_d = torch.sigmoid(O._mlp(torch.cat([geo, f_n], -1), dif)).float().numpy()
    c_s = torch.sigmoid(O._mlp(torch.cat([geo, n, f_r, ndv], -1), spe)).float().numpy()
    np.testing.assert_allclose(c_d, z["diffuse"], atol=2e-5)
    np.testing.assert_allclose(c_s, z["specular"], atol=2e-5)


# ---------------------------------------------------------------------------------------------
# occupancy-grid maintenance: oracle vs the reference's own update_extra_state / mark_untrained_grid
# (tests/golden/make_golden.py::gen_density_grid, run on the CPU with the reference's Python code)
# ---------------------------------------------------------------------------------------------

def _density_P(z):
    n = sorted({int(k.split("_")[2]) for k in z.files if k.startswith("sdf_net_")})
    return dict(bound=float(z["bound"]), beta=float(z["beta"]), beta_min=float(z["opt_beta_min"]), beta_max=float(z["opt_beta_max"]),
                sdf=_layers(z, "sdf_net", n))


def _bits_equal_except_near_threshold(bits, ref_bits, grid, th, rtol=1e-4):
    a, b = np.unpackbits(bits, bitorder="little"), np.unpackbits(ref_bits, bitorder="little")
    bad = np.nonzero(a != b)[0]
    assert all(abs(grid.reshape(-1)[i] - th) <= rtol * max(th, 1e-6) for i in bad), (len(bad), bad[:8])
    return len(bad)


def test_density_grid_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "density_grid.npz"))
    P = _density_P(z)
    C, H = int(z["cascade"]), 128
    A, b = torch.from_numpy(z["A"]), torch.from_numpy(z["b"])
    enc_fn = lambda x: torch.sin(x @ A + b)
    kw = dict(cascade=C, grid_size=H, density_thresh=float(z["density_thresh"]), density_scale=float(z["density_scale"]), enc_fn=enc_fn)
    # mark_untrained_grid: exact except where a comparison is decided at rounding level
    count, margin = O.mark_untrained_grid(z["poses"], z["intrinsic"], bound=P["bound"], cascade=C, grid_size=H)
    marked_ref = np.unpackbits(z["marked"]).astype(bool)[: C * H ** 3]
    sure = margin.reshape(-1) > 1e-5
    assert sure.mean() > 0.999
    np.testing.assert_array_equal((count.reshape(-1) == 0)[sure], marked_ref[sure])
    grid = np.zeros((C, H ** 3), np.float32)
    grid[marked_ref.reshape(C, -1)] = -1
    torch.manual_seed(int(z["seed"]))
    sub = slice(None, None, 16)
    for stage in range(2):                                             # iter_density 0, 1: full updates
        noise = np.stack([torch.rand(H ** 3, 3).numpy() for _ in range(C)])
        grid, mean, th, bits, _ = O.update_extra_state(P, grid, noise=noise, **kw)
        np.testing.assert_allclose(grid.reshape(-1)[sub], z[f"full{stage}_grid_sub"], rtol=2e-5, atol=1e-6)
        assert abs(mean - float(z[f"full{stage}_mean"])) <= 1e-5 * abs(mean)
        _bits_equal_except_near_threshold(bits, z[f"full{stage}_bits"], grid, th)
    assert (grid.reshape(-1)[marked_ref] == -1).all()                 # untrained cells never come back
    # partial update: replay the reference's random draws (renderer.py:310-331) from the recorded generator state
    torch.set_rng_state(torch.from_numpy(z["part_rng_state"]))
    N = H ** 3 // 4
    coords, noise = [], []
    for cas in range(C):
        c = torch.randint(0, H, (N, 3))
        occ = torch.nonzero(torch.from_numpy(grid[cas]) > 0).squeeze(-1)
        rm = torch.randint(0, occ.shape[0], [N], dtype=torch.long)
        oc = torch.from_numpy(O.morton3D_invert(occ[rm].numpy().astype(np.int32)))
        coords.append(torch.cat([c.int(), oc.int()], 0).numpy())
        noise.append(torch.rand(2 * N, 3).numpy())
    grid_p, mean, th, bits, tmp = O.update_extra_state(P, grid, noise=np.stack(noise), coords=np.stack(coords), **kw)
    # duplicated cells: index_put keeps one of the candidates (which one is unspecified); compare the cells visited once
    idx = O.morton3D(np.stack(coords)[0])
    uniq, cnt = np.unique(idx, return_counts=True)
    once = np.zeros(H ** 3, bool); once[uniq[cnt == 1]] = True
    untouched = np.ones(H ** 3, bool); untouched[uniq] = False
    cmp_mask = (once | untouched)[sub]
    np.testing.assert_allclose(grid_p.reshape(-1)[sub][cmp_mask], z["part_grid_sub"][cmp_mask], rtol=2e-5, atol=1e-6)
    assert abs(mean - float(z["part_mean"])) <= 1e-4 * abs(mean)


# ---------------------------------------------------------------------------------------------
# ray generation + loss epilogue: oracle vs the reference's own get_rays, auxiliary block and
# Trainer.train_step (tests/golden/make_golden.py::gen_train_epilogue)
# ---------------------------------------------------------------------------------------------

def test_get_rays_matches_reference(golden_dir):
    from oracle import train_oracle as TO
    z = np.load(os.path.join(golden_dir, "train_epilogue.npz"))
    H, W = [int(v) for v in z["rays_HW"]]
    ro, rd = TO.get_rays(z["rays_poses"], z["rays_intrinsics"], H, W, z["rays_inds"])
    np.testing.assert_allclose(ro.numpy(), z["rays_o"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(rd.numpy(), z["rays_d"], rtol=0, atol=1e-6)
    ro, rd = TO.get_rays(z["rays_poses"][:1], z["rays_intrinsics"], H, W, None)
    np.testing.assert_allclose(rd.numpy(), z["rays_full_d"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(ro.numpy(), z["rays_full_o"], rtol=0, atol=1e-6)


def loss_case_from_golden(z):
    w = z["loss_w"]
    kw = dict(color_l1=True, color_w=float(w[0]), mask_w=float(w[1]), cauchy_w=float(w[2]), eikonal_w=float(w[3]), backsdf_w=float(w[4]),
              backsdf_thresh=float(w[5]), backsdf_mean=bool(z["loss_backsdf_mean"]))
    args = (z["loss_image"], z["loss_weights_sum"], z["loss_gt_rgb"], z["loss_gt_mask"], z["loss_sdfs"], z["loss_sdf_gradients"],
            z["loss_weights"], z["loss_deltas"], z["loss_rays"], float(z["loss_beta"]))
    return args, kw


def test_train_loss_matches_reference(golden_dir):
    from oracle import train_oracle as TO
    z = np.load(os.path.join(golden_dir, "train_epilogue.npz"))
    args, kw = loss_case_from_golden(z)
    pm, ds = TO.aux_point_mask(torch.from_numpy(z["loss_deltas"]), torch.from_numpy(z["loss_rays"]).long(), z["loss_sdfs"].shape[0])
    assert int(pm.sum()) == int(z["aux_point_count"])
    sd = torch.from_numpy(z["loss_sdfs"])
    np.testing.assert_array_equal((torch.roll(sd, -1, 0) - sd)[pm].numpy(), z["aux_relsdf"])
    np.testing.assert_array_equal(ds[pm, 1].numpy(), z["aux_sdf_dist"])
    terms, grads = TO.train_loss(*args, **kw)
    assert abs(terms["total"] - float(z["loss_total"])) <= 2e-6 * abs(terms["total"])
    for k in ("color", "mask", "backsdf", "cauchy", "eikonal"):
        assert abs(terms[k] - float(z[f"loss_term_{k}"])) <= 3e-6 * abs(terms[k]), k
    for k in ("image", "weights_sum", "sdfs", "sdf_gradients"):
        ref = z[f"grad_{k}"]
        np.testing.assert_allclose(grads[k], ref, rtol=2e-4, atol=2e-6 * float(np.abs(ref).max()), err_msg=k)


# ---------------------------------------------------------------------------------------------
# NeuS-style opacity: oracle vs the reference's NeuSDensity (tests/golden/make_golden.py::gen_neus)
# ---------------------------------------------------------------------------------------------

def neus_case(z, tag):
    dists = z[f"{tag}_dists"]
    dists = float(dists) if dists.ndim == 0 else dists
    grads = z[f"{tag}_grads"] if int(z[f"{tag}_with_grad"]) else None
    return z[f"{tag}_sdf"], z[f"{tag}_dirs"], dists, grads, float(z[f"{tag}_var"]), float(z[f"{tag}_ratio"]), z[f"{tag}_ga"]


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_neus_alpha_matches_reference(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, "neus.npz"))
    sdf, dirs, dists, grads, var, ratio, ga = neus_case(z, tag)
    alpha, g_sdf, g_grads, g_var = O.neus_alpha(sdf, dirs, dists, grads, var, ratio, grad_alpha=ga)
    np.testing.assert_allclose(alpha, z[f"{tag}_alpha"], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(g_sdf, z[f"{tag}_g_sdf"], rtol=5e-4, atol=5e-5 * float(np.abs(z[f"{tag}_g_sdf"]).max()) + 1e-30)
    assert abs(g_var - float(z[f"{tag}_g_var"])) <= 1e-3 * abs(float(z[f"{tag}_g_var"])) + 1e-6
    if grads is not None:
        np.testing.assert_allclose(g_grads, z[f"{tag}_g_grads"], rtol=5e-4, atol=5e-5 * float(np.abs(z[f"{tag}_g_grads"]).max()) + 1e-12)


# ---------------------------------------------------------------------------------------------
# training branch: the train oracle vs the reference's OWN run_cuda training branch + NeRFNetwork methods run on
# the CPU with the oracle's operators injected (tests/golden/make_golden.py::gen_train_branch)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag", ["single", "renv"])
def test_train_oracle_matches_reference_run_cuda_training_branch(golden_dir, tag):
    from envidr_b200 import scene
    from oracle import train_oracle as TO
    z = np.load(os.path.join(golden_dir, "train_branch.npz"))
    P = _P_from_glue(z, 5)
    P.update(embeddings=z["embeddings"], offsets=z["offsets"].astype(np.int32), per_level_scale=float(z["per_level_scale"]),
             base_resolution=int(z["base_resolution"]), density_scale=1.0, enabled_levels=-1)
    th = TO.params_from_dict(P, torch.float64, frozen=())
    bf = scene.make_bitfield()
    out = TO.render_train(th, P, z["rays_o"], z["rays_d"], bf, max_steps=256, bg_color=1.0,
                          r_images=z["r_images"] if tag == "renv" else None)
    f = lambda t: t.detach().to(torch.float32).numpy()
    assert out["sdfs"].shape[0] == z[f"{tag}_sdfs"].shape[0]                      # same samples, same padding
    np.testing.assert_allclose(f(out["sdfs"]), z[f"{tag}_sdfs"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(f(out["sigmas"]), z[f"{tag}_sigmas"], rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(f(out["sdf_gradients"]), z[f"{tag}_sdf_gradients"], rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(f(out["image"]), z[f"{tag}_image"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(f(out["weights_sum"]), z[f"{tag}_weights_sum"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(f(out["depth"]), z[f"{tag}_depth"], rtol=0, atol=2e-5)
    sg = out["sdf_gradients"]
    loss = 0.3 * out["image"].sum() + (out["weights_sum"] ** 2).sum() + 5 * (out["sdfs"] ** 2).mean() + ((sg.norm(dim=-1) - 1) ** 2).mean()
    assert abs(float(loss) - float(z[f"{tag}_loss"])) <= 2e-5 * abs(float(z[f"{tag}_loss"]))
    names = {"sdf.0.weight": "sdf_net_0_weight", "sdf.2.bias": "sdf_net_2_bias", "env.1.weight": "env_net_1_weight",
             "renv.0.weight": "renv_net_0_weight", "embeddings": "encoder_embeddings", "beta": "sdf_density_beta"}
    grads = torch.autograd.grad(loss, [th[k] for k in names], allow_unused=True)
    for (k, gk), g in zip(names.items(), grads):
        ref = z[f"{tag}_grad_{gk}"]
        if ref.size == 0 or (tag == "single" and k.startswith("renv")):
            assert g is None or float(g.abs().max()) == 0.0 or ref.size == 0, k
            continue
        got = g.detach().to(torch.float32).numpy().reshape(ref.shape)
        scale = float(np.abs(ref).max())
        assert float(np.abs(got - ref).max()) <= 2e-3 * scale + 1e-9, (k, float(np.abs(got - ref).max()), scale)


# ---------------------------------------------------------------------------------------------
# inference path: the render oracle vs the reference's OWN NeRFRenderer.render + run_cuda inference loop run on
# the CPU with the oracle's operators injected (tests/golden/make_golden.py::gen_infer_branch): single pass and three-pass frame
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag,indir,rot", [("one", False, None), ("three", True, None), ("rot", True, 0.7)])
def test_render_oracle_matches_reference_inference_path(golden_dir, tag, indir, rot):
    from envidr_b200 import scene
    z = np.load(os.path.join(golden_dir, "infer_branch.npz"))
    P = _P_from_glue(z, 5)
    P.update(embeddings=z["embeddings"], offsets=z["offsets"].astype(np.int32), per_level_scale=float(z["per_level_scale"]),
             base_resolution=int(z["base_resolution"]), density_scale=1.0, enabled_levels=-1)
    bf = scene.make_bitfield()
    res = O.render(P, z["rays_o"], z["rays_d"], bf, indir_ref=indir, indir_max_steps=256, max_steps=256, bg_color=1.0, env_rot_radian=rot)
    if rot is not None:
        assert float(np.abs(z["rot_image"] - z["three_image"]).max()) > 1e-2      # the rotation does change the frame
    assert int((z[f"{tag}_weights_sum"] > 0.5).sum()) >= 40                       # the frame sees the object
    e = np.abs(res["image"] - z[f"{tag}_image"]).max(-1)
    assert int((e > 1e-4).sum()) <= 2 and float(np.median(e)) <= 1e-5, (int((e > 1e-4).sum()), float(e.max()))
    np.testing.assert_allclose(res["weights_sum"], z[f"{tag}_weights_sum"][:, 0], rtol=0, atol=2e-5)
    np.testing.assert_allclose(res["depth"], z[f"{tag}_depth"][:, 0], rtol=0, atol=2e-5)
    en = np.abs(res["normal_image"] - z[f"{tag}_normal_image"]).max(-1)
    assert int((en > 1e-3).sum()) <= 2, (int((en > 1e-3).sum()), float(en.max()))


def test_neus_oracle_matches_reference_config4_model(golden_dir):
    """BASELINE config 4: oracle/neus_oracle.py against the reference's OWN NeRFNetwork (use_neus_sdf, frequency encoding, geometric_init,
    8 x 256 weight-normed Softplus layers with skip_layers [4]) run on the CPU by tests/golden/make_golden.py::gen_neus_field:
    per-sample sdf / normal / NeuS alpha / geo_feat / roughness / rgb, and the 16 x 16 frame of NeRFRenderer.render -> run_cuda with
    input_alpha compositing.  Also: the torch formula of the frequency encoding equals the C restatement of freqencoder.cu."""
    import os
    from envidr_b200 import scene
    from oracle import neus_oracle as NO
    from oracle import oracle as O
    z = np.load(os.path.join(golden_dir, "neus_field.npz"))
    x = z["x"]
    a = NO.freq_encode(torch.from_numpy(x).double(), 6).numpy()
    b = O.freq_encode_forward(x, 6)
    np.testing.assert_allclose(a, b, atol=3e-6)               # sinf of the fp32 argument 2^f x (+ pi/2) vs float64: argument rounding up to 32 * 6e-8
    nf = scene.make_neus_field(0, hidden_dim_env=64, ide_degree=4)
    P = scene.neus_to_oracle(nf)
    out = NO.field_forward(P, x, z["d"], z["dists"])
    np.testing.assert_allclose(out["sdf"], z["sdf"], atol=2e-6)
    np.testing.assert_allclose(out["normal"], z["normal"], atol=2e-5)
    np.testing.assert_allclose(out["sigma"], z["alpha"], atol=2e-4)     # inv_s = e^6 = 403 amplifies the fp32 sdf rounding of the reference
    np.testing.assert_allclose(out["geo_feat"], z["geo"], atol=2e-6)
    np.testing.assert_allclose(out["roughness"], z["roughness"], atol=1e-6)
    np.testing.assert_allclose(out["blend"], z["blend"], atol=1e-6)
    np.testing.assert_allclose(out["c_diffuse"], z["c_diffuse"], atol=2e-5)
    np.testing.assert_allclose(out["c_specular"], z["c_specular"], atol=2e-5)
    np.testing.assert_allclose(out["rgb"], z["rgb"], atol=3e-5)
    assert 0.05 < float((z["alpha"] > 0.5).mean()) < 0.95 and float(z["alpha"].max()) == 1.0
    P256 = dict(P)
    fr = NO.render_rays(P, z["rays_o"], z["rays_d"], scene.make_sphere_bitfield(), max_steps=256, bg_color=1.0, dtype=torch.float32)
    assert int((z["frame_weights_sum"] > 0.5).sum()) >= 30
    np.testing.assert_allclose(fr["weights_sum"], z["frame_weights_sum"][:, 0], atol=2e-4)
    np.testing.assert_allclose(fr["depth"], z["frame_depth"][:, 0], atol=1e-3)
    e = np.abs(fr["image"] - z["frame_image"]).max(-1)
    assert float(e.max()) <= 3e-4 and float(np.median(e)) <= 2e-5, (float(e.max()), float(np.median(e)))
    ws = z["frame_weights_sum"]
    n_ref = (z["frame_normal_image"] - (1 - ws)) / np.maximum(ws, 1e-6)                       # undo renderer.py:529-530
    hit = ws[:, 0] > 0.5
    assert float(np.abs(fr["normal_image"][hit] - n_ref[hit]).max()) <= 2e-3
