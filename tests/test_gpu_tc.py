"""Tensor-core path on hardware: the tcgen05 probe pins the UMMA shared-memory descriptor / operand layout conventions
(csrc/tc_common.cuh) against a plain fp32 matmul of the fp16-rounded operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("N,K", [(256, 16), (256, 64), (64, 32), (16, 256), (128, 80)])
def test_tcgen05_probe_matches_matmul(dev, N, K):
    from envidr_b200._lib import check, lib, ptr, stream
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    ref = A.half().float() @ B.half().float().T
    D = torch.full((128, N), float("nan"), device=dev)
    check(lib().envidr_tc_probe(ptr(A), ptr(B), ptr(D), N, K, 0, stream()), "tc_probe")
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    if not err < 1e-3:
        D1 = torch.full((128, N), float("nan"), device=dev)
        check(lib().envidr_tc_probe(ptr(A), ptr(B), ptr(D1), N, K, 1, stream()), "tc_probe")
        torch.cuda.synchronize()
        err1 = (D1 - ref).abs().max().item()
        pytest.fail(f"variant 0 (LBO = K-chunk stride, SBO = 8-row stride) max err {err}; swapped variant max err {err1}; "
                    f"D[0,:4]={D[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}")


def _samples(M, seed):
    from envidr_b200 import scene
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.75, 0.75, size=(M * 3, 3))
    sd = scene.analytic_sdf(x)
    x = x[np.argsort(np.abs(sd))[:M]][rng.permutation(M)]
    d = rng.standard_normal((M, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return x.astype(np.float32), d.astype(np.float32)


@pytest.mark.parametrize("env_width,deg,M", [(256, 5, 3000), (64, 4, 130), (160, 4, 64 * 148 + 7)])
def test_tensor_core_env_net_matches_fp32_path(dev, env_width, deg, M):
    """precision='tc' (tcgen05, fp16 hi/lo split, 3 MMAs) vs the FFMA path of the same library and vs the CPU oracle."""
    from envidr_b200 import scene
    from oracle import oracle as O
    fp_cpu = scene.make_synthetic_field(1, hidden_dim_env=env_width, ide_degree=deg)
    x, d = _samples(M, 0)
    xt, dt = torch.from_numpy(x).to(dev), torch.from_numpy(d).to(dev)
    want = ("sigma", "rgb", "normal", "c_diffuse", "c_specular")
    ref32 = fp_cpu.to(dev).pack().forward(xt, dt, want=want)
    fp_tc = fp_cpu.to(dev)
    fp_tc.precision = "tc"
    out = fp_tc.pack().forward(xt, dt, want=want)
    torch.cuda.synchronize()
    # geometry on tensor cores (fp16 hi/lo split, fp32 accumulate): the SDF agrees to ~2e-7, which the Laplace density
    # amplifies by up to 1/(2 beta^2) = 5000
    torch.testing.assert_close(out["sigma"], ref32["sigma"], atol=5e-3, rtol=2e-4)
    assert (out["normal"] - ref32["normal"]).abs().max().item() <= 2e-5
    err = (out["rgb"] - ref32["rgb"]).abs().max().item()
    assert err <= 3e-5, f"tensor-core path vs fp32 path: rgb max err {err}"
    assert (out["c_diffuse"] - ref32["c_diffuse"]).abs().max().item() <= 2e-5
    assert (out["c_specular"] - ref32["c_specular"]).abs().max().item() <= 2e-5
    if M <= 3000:
        ref = O.field_forward(fp_cpu.to_oracle(), x, d)
        np.testing.assert_allclose(out["rgb"].cpu().numpy(), ref["rgb"], atol=1e-4)
    # r_images / env rotation flow through the split launch unchanged
    rng = np.random.default_rng(4)
    ri = torch.from_numpy(rng.uniform(0, 1, size=(M, 4)).astype(np.float32)).to(dev)
    ri[::2, 3] = 0.97
    a = fp_tc.forward(xt, dt, ri, env_rot_radian=0.4, want=("rgb",))["rgb"]
    b = fp_cpu.to(dev).pack().forward(xt, dt, ri, env_rot_radian=0.4, want=("rgb",))["rgb"]
    assert (a - b).abs().max().item() <= 3e-5
    # geometry-only mode and extra outputs of the tensor-core geometry kernel
    ga = fp_tc.forward(xt, dt, geometry_only=True, want=("sigma", "normal", "sdf", "roughness", "grad_x"))
    gb = fp_cpu.to(dev).pack().forward(xt, dt, geometry_only=True, want=("sigma", "normal", "sdf", "roughness", "grad_x"))
    torch.testing.assert_close(ga["sdf"], gb["sdf"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(ga["grad_x"], gb["grad_x"], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(ga["roughness"], gb["roughness"], atol=1e-6, rtol=1e-5)


def test_render_with_tensor_cores(dev):
    from envidr_b200 import render, scene
    fp_cpu = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
    bft = torch.from_numpy(scene.make_bitfield()).to(dev)
    ro, rd = [t.to(dev) for t in scene.camera_rays(96, 96)]
    cfg = render.RenderConfig(indir_ref=True)
    ref = render.render(fp_cpu.to(dev).pack(), bft, ro, rd, cfg)
    fp_tc = fp_cpu.to(dev)
    fp_tc.precision = "tc"
    st = []
    out = render.render(fp_tc.pack(), bft, ro, rd, cfg, stats=st)
    assert len(st) == 3
    torch.testing.assert_close(out["weights_sum"], ref["weights_sum"], atol=1e-4, rtol=0)
    torch.testing.assert_close(out["depth"], ref["depth"], atol=3e-4, rtol=0)
    err = (out["image"] - ref["image"]).abs().max().item()
    assert err <= 1e-4, f"image L-inf tensor-core vs fp32 path {err}"
    assert (out["image"] - ref["image"]).abs().mean().item() <= 2e-6


@pytest.mark.parametrize("variant", [{"ENVIDR_ENV_TC_CTAS": "2"}, {"ENVIDR_ENV_TC_MULTICAST": "2"}, {"ENVIDR_ENV_TC_MULTICAST": "4"},
                                     {"ENVIDR_ENV_TC_REORDER": "0"}, {"ENVIDR_ENV_TC_MODE": "1"}, {"ENVIDR_GEOM_STAGE": "1"}],
                         ids=lambda v: "-".join(f"{k[7:]}={x}" for k, x in v.items()))
def test_env_kernel_variants_in_subprocess(dev, variant):
    """The opt-in variants of the env_net / geometry kernels (environment switches are read once per process, hence the subprocess): CTA pair
    (tcgen05 cta_group::2, weight halves by tensor-map TMA), weight-multicast clusters of 2 / 4, plain layer order, e4m3 correction products,
    level-0 table staged in shared memory.  Same RGB as the fp32 FFMA path within 2e-5 (3e-5 with e4m3 corrections) per sample, odd and even tile
    counts, a tail tile, and a batch smaller than one pair / cluster."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from envidr_b200 import scene
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
worst = 0.0
for env, deg, M in ((256, 5, 64 * 148 * 3 + 17), (160, 4, 64 * 5 + 1), (64, 4, 40)):
    fp_cpu = scene.make_synthetic_field(3, hidden_dim_env=env, ide_degree=deg)
    x = torch.from_numpy(rng.uniform(-0.6, 0.6, (M, 3)).astype(np.float32)).to(dev)
    d = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((M, 3)).astype(np.float32)), dim=-1).to(dev)
    ri = torch.from_numpy(rng.uniform(0, 1, (M, 4)).astype(np.float32)).to(dev)
    fp_cpu.precision = "tc"; a = fp_cpu.to(dev).pack().forward(x, d, ri, want=("rgb", "sigma"))
    fp_cpu.precision = "fp32"; b = fp_cpu.to(dev).pack().forward(x, d, ri, want=("rgb", "sigma"))
    worst = max(worst, float((a["rgb"] - b["rgb"]).abs().max()))
print("WORST", worst)
assert worst <= BOUND, worst
'''.replace("BOUND", "3e-5" if "ENVIDR_ENV_TC_MODE" in variant else "2e-5")
    env = dict(os.environ, **variant)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "WORST" in r.stdout


@pytest.mark.parametrize("layers", [2, 3, 5, 6])
def test_env_kernel_layer_counts(dev, layers):
    """k_env_tc issues the NEXT tile's layer 0 ahead of the last layer and parks the last layer's accumulator in the buffer of the layer
    before it; which buffer is free depends on the parity of the layer count (field_tc.cu, acc_buf).  env_net with 2, 3, 5 and 6 layers
    (the shipped scenes have 4) against the FFMA path, several tiles per SM so that consecutive tiles of one CTA overlap."""
    from envidr_b200 import scene
    fp_cpu = scene.make_synthetic_field(2, hidden_dim_env=256, ide_degree=5)
    e = fp_cpu.env                                   # 72 -> 256 -> 256 -> 256 -> 12
    g = torch.Generator().manual_seed(layers)
    extra = lambda: ((torch.rand(256, 256, generator=g) * 2 - 1) * (6.0 / 512) ** 0.5, torch.randn(256, generator=g) * 0.05)
    mid = {2: [], 3: [e[1]], 5: [e[1], e[2], extra()], 6: [e[1], e[2], extra(), extra()]}[layers]
    fp_cpu.env = [e[0]] + mid + [e[3]]
    M = 64 * 148 * 3 + 11
    x, d = _samples(M, 1)
    xt, dt = torch.from_numpy(x).to(dev), torch.from_numpy(d).to(dev)
    fp_cpu.precision = "fp32"
    ref = fp_cpu.to(dev).pack().forward(xt, dt, want=("rgb",))["rgb"]
    fp_cpu.precision = "tc"
    out = fp_cpu.to(dev).pack().forward(xt, dt, want=("rgb",))["rgb"]
    assert (out - ref).abs().max().item() <= 3e-5
