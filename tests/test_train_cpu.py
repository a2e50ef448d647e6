"""Train branch (run_cuda training path) on the CPU:
  * the train oracle's forward equals the pinned inference field oracle (oracle.field_forward, itself pinned to the
    reference's network.py by tests/golden/field_glue.npz);
  * the oracle's hash_encode double-backward Function is consistent with finite differences of its own backward;
  * the product's train glue (envidr_b200/train.py), driven with the oracle's operators injected in place of the CUDA
    ones, gives the same loss and parameter gradients as the independent oracle restatement.
The CUDA operators themselves are checked in tests/test_gpu_train.py."""
import numpy as np
import pytest
import torch

from envidr_b200 import scene, train
from envidr_b200.render import RenderConfig
from oracle import oracle as O
from oracle import train_oracle as TO


def small_field(seed=0, env=64, deg=4):
    return scene.make_synthetic_field(seed, hidden_dim_env=env, ide_degree=deg, num_levels=8, log2_hashmap_size=14, desired_resolution=256)


def oracle_ops(dtype):
    """TrainOps built from the oracle's CPU operators (tests only)."""
    def near_far(ro, rd, aabb, min_near):
        n, f = O.near_far_from_aabb(ro.numpy(), rd.numpy(), aabb.numpy(), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)

    def march(ro, rd, bound, bitfield, C, H, nears, fars, counter, mean_count, perturb, align, force_all, dt_gamma, max_steps, early):
        N = ro.shape[0]
        x, d, dl, rays, cnt = O.march_rays_train(ro.numpy(), rd.numpy(), bound, bitfield.numpy(), C, H, nears.numpy(), fars.numpy(),
                                                 N * max_steps, dt_gamma=dt_gamma, max_steps=max_steps, early_stop_steps=early)
        m = int(cnt[0])
        m += align - m % align
        return torch.from_numpy(x[:m]).to(dtype), torch.from_numpy(d[:m]).to(dtype), torch.from_numpy(dl[:m]), torch.from_numpy(rays)

    def composite(sig, rgb, deltas, rays, T, ret_w, input_alpha):
        return TO.composite_rays_train(sig, rgb, deltas.numpy(), rays.numpy(), T, ret_w, input_alpha)

    def scatter(rays, src):
        return torch.from_numpy(O.get_scatter_idx(rays.numpy(), src.shape[0]))

    def henc(x01, emb, offsets, pls, H, calc):
        return TO.hash_encode(x01, emb, offsets.numpy(), pls, H, calc)

    return train.TrainOps(near_far, march, composite, scatter, henc, lambda d, k, deg: O.ide_encode(d, k, deg))


def test_train_oracle_forward_equals_pinned_field_oracle():
    fp = small_field()
    P = fp.to_oracle()
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(300, 3, generator=g) * 1.2 - 0.6).numpy().astype(np.float32)
    d = torch.nn.functional.normalize(torch.randn(300, 3, generator=g), dim=-1).numpy().astype(np.float32)
    ri = torch.rand(300, 4, generator=g).numpy().astype(np.float32)
    ref = O.field_forward(P, x, d, r_images=ri)
    th = TO.params_from_dict(P)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    sdf, sigma, geo, normals, grad_x, rough, blend = TO.forward_sigma(th, P, xt)
    rgb = TO.forward_color(th, P, geo, torch.from_numpy(d).double(), normals, rough, blend, torch.from_numpy(ri).double())
    np.testing.assert_allclose(sdf.detach().numpy(), ref["sdf"], atol=1e-6)
    np.testing.assert_allclose(sigma.detach().numpy(), ref["sigma"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(normals.detach().numpy(), ref["normal"], atol=2e-5)     # autograd vs analytic reverse pass
    np.testing.assert_allclose(rough.detach().numpy(), ref["roughness"], atol=1e-6)
    np.testing.assert_allclose(rgb.detach().numpy(), ref["rgb"], atol=2e-5)


def test_oracle_hash_encode_second_backward_matches_finite_differences():
    rng = np.random.default_rng(0)
    offsets, pls = O.hash_offsets(3, 4, 8, 10, 64)
    T = int(offsets[-1])
    emb = torch.from_numpy(rng.uniform(-1, 1, (T, 2))).double().requires_grad_(True)
    x = torch.from_numpy(rng.uniform(0.05, 0.95, (40, 3))).double().requires_grad_(True)
    w = torch.from_numpy(rng.standard_normal(8)).double()
    v = torch.from_numpy(rng.standard_normal((40, 3))).double()

    def grad_x_of(e):
        y = TO.hash_encode(x, e, offsets, pls, 8, True)
        return torch.autograd.grad((y * w).sum(), x, create_graph=True)[0]

    gx = grad_x_of(emb)
    g_emb = torch.autograd.grad((gx * v).sum(), emb)[0]               # second-order path: d(grad_x . v) / d emb
    idx = rng.choice(T, 12, replace=False)
    for i in idx:
        for c in range(2):
            e2 = emb.detach().clone(); e2[i, c] += 1e-2
            e1 = emb.detach().clone(); e1[i, c] -= 1e-2
            fd = ((grad_x_of(e2.requires_grad_(True)) * v).sum() - (grad_x_of(e1.requires_grad_(True)) * v).sum()) / 2e-2
            assert abs(float(fd) - float(g_emb[i, c])) <= 2e-3 * max(1.0, abs(float(fd))), (i, c, float(fd), float(g_emb[i, c]))


@pytest.mark.parametrize("with_r_images", [False, True])
def test_product_train_glue_with_oracle_ops_matches_train_oracle(with_r_images):
    fp = small_field()
    P = fp.to_oracle()
    bf = scene.make_bitfield()
    ro, rd = scene.camera_rays(12, 12)
    N = ro.shape[0]
    g = torch.Generator().manual_seed(2)
    gt_rgb = torch.rand(N, 3, generator=g)
    gt_mask = (torch.rand(N, generator=g) > 0.5).float()
    ri = torch.rand(N, 4, generator=g) if with_r_images else None
    # independent oracle restatement, float64
    loss_o, grads_o, out_o = TO.train_step(P, ro.numpy(), rd.numpy(), bf, gt_rgb.numpy(), gt_mask.numpy(),
                                           r_images=None if ri is None else ri.numpy(), max_steps=256)
    assert out_o["num_samples"] > 128
    # product glue, float64, oracle operators injected
    field = train.TrainableField(fp).double()
    cfg = RenderConfig(max_steps=256)
    out = train.render_train(field, torch.from_numpy(bf), ro, rd, cfg, ops=oracle_ops(torch.float64),
                             r_images=None if ri is None else ri.double())
    loss = train.loss_epilogue(field, out, gt_rgb.double(), gt_mask.double())
    loss.backward()
    assert out["xyzs"].shape[0] == out_o["xyzs"].shape[0]
    assert abs(float(loss) - loss_o) <= 1e-9 * max(1.0, abs(loss_o))
    names = {"embeddings": "embeddings", "beta": "beta"}
    for st in ("sdf", "env", "renv"):
        for i in range(field.n_layers[st]):
            names[f"{st}_w{i}"] = f"{st}.{i}.weight"
            names[f"{st}_b{i}"] = f"{st}.{i}.bias"
    checked = 0
    for pn, on in names.items():
        gp = getattr(field, pn).grad
        go = grads_o.get(on)
        if go is None:
            assert gp is None or float(gp.abs().max()) == 0.0
            continue
        assert gp is not None, pn
        scale = max(1e-12, float(np.abs(go).max()))
        assert float(np.abs(gp.numpy() - go).max()) <= 1e-5 * scale + 1e-12, pn
        checked += 1
    assert checked >= 12
    for st in ("diffuse", "color"):                                   # frozen MLPs (toaster.ini frozen_mlps)
        assert getattr(field, f"{st}_w0").grad is None


def test_compute_normal_options_detach_and_anneal():
    """compute_normal's two options (renderer.py:191-196) in the train-branch glue: detach_normal cuts the path from the normal to
    the SDF network; normal_anneal_ratio blends the SDF normal with the radial direction and renormalises."""
    import torch.nn.functional as F
    from envidr_b200 import scene, train
    fp = small_field()
    ops = oracle_ops(torch.float64)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(64, 3, generator=g) * 1.0 - 0.5).double()
    base = train.TrainableField(fp).double()
    xa = x.clone().requires_grad_(True)
    _, _, _, n0, gx0, _, _ = base.forward_sigma(ops, xa)
    ann = train.TrainableField(fp, normal_anneal_ratio=0.25).double()
    xb = x.clone().requires_grad_(True)
    _, _, _, n1, _, _, _ = ann.forward_sigma(ops, xb)
    want = F.normalize(n0.detach() * 0.25 + 0.75 * F.normalize(x, dim=-1, eps=1e-10), dim=-1, eps=1e-10)
    assert float((n1.detach() - want).abs().max()) < 1e-6
    det = train.TrainableField(fp, detach_normal=True).double()
    xc = x.clone().requires_grad_(True)
    _, _, _, n2, gx2, _, _ = det.forward_sigma(ops, xc)
    assert torch.allclose(n2, n0.detach(), atol=1e-7) and not n2.requires_grad and gx2.requires_grad     # eikonal gradient keeps its graph
