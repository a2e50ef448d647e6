"""NeuSDensity kernels (csrc/neus.cu through envidr_b200.neus.NeuSDensity) on the B200 against the reference's own
module (golden tests/golden/neus.npz, made by running nerf/network.py::NeuSDensity) and the float64 oracle: alpha 2e-5 relative,
gradients w.r.t. sdf / normals / variance 1e-3 of their scale; plus the pipeline config 4 uses it in: freq_encode -> alpha ->
composite_rays_train(input_alpha=True)."""
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


def _run(dev, sdf, dirs, dists, grads, var, ratio, ga):
    from envidr_b200.neus import NeuSDensity
    mod = NeuSDensity(var).to(dev)
    t = lambda a, g=False: None if a is None else torch.from_numpy(np.asarray(a, np.float32)).to(dev).requires_grad_(g)
    ts, tg = t(sdf, True), t(grads, True)
    d = t(dists) if not np.isscalar(dists) else dists
    alpha = mod(ts, t(dirs), d, tg, cos_anneal_ratio=ratio)
    (alpha * t(ga)).sum().backward()
    return alpha.detach().cpu().numpy(), ts.grad.cpu().numpy(), None if tg is None else tg.grad.cpu().numpy(), float(mod.variance.grad)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_neus_alpha_vs_reference_golden(dev, golden_dir, tag):
    from test_oracle_golden import neus_case
    z = np.load(os.path.join(golden_dir, "neus.npz"))
    sdf, dirs, dists, grads, var, ratio, ga = neus_case(z, tag)
    alpha, g_sdf, g_grads, g_var = _run(dev, sdf, dirs, dists, grads, var, ratio, ga)
    np.testing.assert_allclose(alpha, z[f"{tag}_alpha"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(g_sdf, z[f"{tag}_g_sdf"], rtol=1e-3, atol=1e-3 * float(np.abs(z[f"{tag}_g_sdf"]).max()) + 1e-30)
    assert abs(g_var - float(z[f"{tag}_g_var"])) <= 2e-3 * abs(float(z[f"{tag}_g_var"])) + 1e-5
    if grads is not None:
        np.testing.assert_allclose(g_grads, z[f"{tag}_g_grads"], rtol=1e-3, atol=1e-3 * float(np.abs(z[f"{tag}_g_grads"]).max()) + 1e-12)


def test_neus_alpha_large_vs_oracle_and_composite(dev):
    from envidr_b200 import raymarching as rm
    from envidr_b200.neus import NeuSDensity
    from oracle import oracle as O
    g = torch.Generator().manual_seed(9)
    M = 300_001
    sdf = torch.randn(M, generator=g) * 0.02
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    grads = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    dists = torch.rand(M, generator=g) * 0.004 + 0.002
    ga = torch.randn(M, generator=g)
    a_o, gs_o, gg_o, gv_o = O.neus_alpha(sdf.numpy(), dirs.numpy(), dists.numpy(), grads.numpy(), 0.4, 0.7, grad_alpha=ga.numpy())
    alpha, g_sdf, g_grads, g_var = _run(dev, sdf.numpy(), dirs.numpy(), dists.numpy(), grads.numpy(), 0.4, 0.7, ga.numpy())
    np.testing.assert_allclose(alpha, a_o, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(g_sdf, gs_o, rtol=1e-3, atol=1e-3 * float(np.abs(gs_o).max()))
    np.testing.assert_allclose(g_grads, gg_o, rtol=1e-3, atol=1e-3 * float(np.abs(gg_o).max()))
    assert abs(g_var - gv_o) <= 2e-3 * abs(gv_o)
    # alpha feeds the compositor with input_alpha = True (cuda_ray.py:121): weights_sum = 1 - prod(1 - alpha) per ray
    n_rays = 1000
    cnt = torch.full((n_rays,), 64, dtype=torch.int32)
    rays = torch.stack([torch.arange(n_rays, dtype=torch.int32), (torch.cumsum(cnt, 0) - cnt).int(), cnt], -1).to(dev)
    Mr = n_rays * 64
    mod = NeuSDensity(0.3).to(dev)
    a = mod(sdf[:Mr].to(dev), dirs[:Mr].to(dev), mod.base_dist, grads[:Mr].to(dev))
    rgbs = torch.rand(Mr, 3, generator=g).to(dev)
    deltas = torch.full((Mr, 2), mod.base_dist).to(dev)
    ws, depth, image, _ = rm.composite_rays_train(a, rgbs, deltas, rays, 1e-4, False, True)
    ref_ws = 1 - torch.prod(1 - a.view(n_rays, 64).double(), dim=1)
    assert float((ws.double() - ref_ws).abs().max()) <= 2e-4            # T_thresh early-out leaves at most 1e-4 of transmittance
    with pytest.raises(RuntimeError):
        NeuSDensity(0.3)(sdf[:4], dirs[:4], 0.01, None)                   # CPU tensors: no fallback
