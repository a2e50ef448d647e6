"""GPU parity of the TRAIN branch (run_cuda training path, BASELINE config 3 shape: toaster dims, use_renv, r_images,
forward + backward) through the library's CUDA operators -- march_rays_train, hash_encode (forward / backward /
second-order backward), composite_rays_train (forward / backward), get_scatter_idx -- against the CPU train oracle
(oracle/train_oracle.py: float64 MLPs over the C restatements of the reference kernels).
Tolerances: loss 1e-5 relative; parameter gradients 2e-3 of the tensor's max |g| (fp32 cuBLAS layers, __expf in the
compositor, float atomics in the encoder backward vs a float64 oracle)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _reset_scales():
    yield
    from oracle import oracle as O
    O.set_level_scales(None)


def _pin_scales(fp, dev):
    from envidr_b200._lib import check, lib, ptr, stream
    from oracle import oracle as O
    L = fp.num_levels
    sc = torch.empty(L, device=dev)
    check(lib().envidr_debug_level_scales(float(np.log2(fp.per_level_scale)), int(fp.base_resolution), L, ptr(sc), stream()))
    O.set_level_scales(sc.cpu().numpy())


def _setup(W, env, deg, seed=2, **kw):
    from envidr_b200 import scene
    fp = scene.make_synthetic_field(0, hidden_dim_env=env, ide_degree=deg, **kw)
    bf = scene.make_bitfield()
    ro, rd = scene.camera_rays(W, W)
    N = ro.shape[0]
    g = torch.Generator().manual_seed(seed)
    gt_rgb = torch.rand(N, 3, generator=g)
    gt_mask = (torch.rand(N, generator=g) > 0.5).float()
    ri = torch.rand(N, 4, generator=g)
    return fp, bf, ro, rd, gt_rgb, gt_mask, ri


@pytest.mark.parametrize("W,env,deg,with_r", [(16, 64, 4, False), (20, 256, 5, True)])
def test_train_step_matches_oracle(dev, W, env, deg, with_r):
    from envidr_b200 import train
    from envidr_b200.render import RenderConfig
    from oracle import train_oracle as TO
    fp, bf, ro, rd, gt_rgb, gt_mask, ri = _setup(W, env, deg)
    _pin_scales(fp.to(dev), dev)
    loss_o, grads_o, out_o = TO.train_step(fp.to_oracle(), ro.numpy(), rd.numpy(), bf, gt_rgb.numpy(), gt_mask.numpy(),
                                           r_images=ri.numpy() if with_r else None, max_steps=512)
    field = train.TrainableField(fp.to(dev))
    cfg = RenderConfig(max_steps=512)
    out = train.render_train(field, torch.from_numpy(bf).to(dev), ro.to(dev), rd.to(dev), cfg, r_images=ri.to(dev) if with_r else None)
    loss = train.loss_epilogue(field, out, gt_rgb.to(dev), gt_mask.to(dev))
    loss.backward()
    # march: same sample count (bit-exact kernel), same padded length
    assert out["xyzs"].shape[0] == out_o["xyzs"].shape[0]
    np.testing.assert_allclose(out["image"].detach().cpu().numpy(), out_o["image"].detach().numpy(), atol=1e-4)
    np.testing.assert_allclose(out["weights_sum"].detach().cpu().numpy(), out_o["weights_sum"].detach().numpy(), atol=1e-4)
    assert abs(float(loss) - loss_o) <= 1e-5 * max(1.0, abs(loss_o))
    names = {"embeddings": "embeddings", "beta": "beta"}
    for st in ("sdf", "env", "renv"):
        for i in range(field.n_layers[st]):
            names[f"{st}_w{i}"] = f"{st}.{i}.weight"
            names[f"{st}_b{i}"] = f"{st}.{i}.bias"
    worst = 0.0
    for pn, on in names.items():
        gp, go = getattr(field, pn).grad, grads_o.get(on)
        if go is None or (not with_r and pn.startswith("renv")):
            continue
        assert gp is not None, pn
        scale = float(np.abs(go).max())
        if scale < 1e-12:
            continue
        err = float(np.abs(gp.cpu().numpy() - go).max()) / scale
        worst = max(worst, err)
        assert err <= 2e-3, (pn, err)
    print(f"[train parity] worst relative gradient error {worst:.2e}; samples {out['xyzs'].shape[0]}")


def test_train_step_mean_count_static_buffers(dev):
    """Steady-state training shape (cuda_ray.py:64-79): force_all_rays=False with mean_count fixes M, rays that do not fit
    are dropped by the march and composite to zero; gradients stay finite."""
    from envidr_b200 import train
    from envidr_b200.render import RenderConfig
    fp, bf, ro, rd, gt_rgb, gt_mask, ri = _setup(24, 64, 4)
    field = train.TrainableField(fp.to(dev))
    cfg = RenderConfig(max_steps=512)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    out = train.render_train(field, torch.from_numpy(bf).to(dev), ro.to(dev), rd.to(dev), cfg, force_all_rays=False, mean_count=4096,
                             step_counter=counter, perturb=True)
    assert out["xyzs"].shape[0] == 4096 + 128          # raymarching.py:213-216: mean_count += align - mean_count % align
    loss = train.loss_epilogue(field, out, gt_rgb.to(dev), gt_mask.to(dev))
    loss.backward()
    assert int(counter[0]) > 0
    for p in field.parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all()


@pytest.mark.parametrize("deg", [4, 5])
def test_ide_encode_backward_kernel(dev, deg):
    """envidr_ide_encode_backward (one kernel) against autograd over the oracle's float64, exact-table formulation of
    ide_encoder.py:98-130: gradients w.r.t. the direction (not normalised: all three components free) and the per-sample
    kappa_inv.  Tolerance 2e-5 of the gradient's max (fp32 kernel vs float64)."""
    from envidr_b200.ide_encoder import IntegratedDirEncoder
    from oracle import oracle as O
    g = torch.Generator().manual_seed(deg)
    B = 777
    x = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1) * (0.9 + 0.2 * torch.rand(B, 1, generator=g))
    x[0] = torch.tensor([0.0, 0.0, 1.0])                       # the (x == 0 & y == 0) guard
    kap = torch.rand(B, 1, generator=g) * 0.3
    w = torch.randn(B, 2 * (2 ** deg - 1 + deg), generator=g)
    xr, kr = x.double().requires_grad_(True), kap.double().requires_grad_(True)
    ref = O.ide_encode(xr, kr, deg, exact_tables=True)
    (ref * w.double()).sum().backward()
    enc = IntegratedDirEncoder(deg_view=deg).to(dev)
    xg, kg = x.to(dev).requires_grad_(True), kap.to(dev).requires_grad_(True)
    out = enc(xg, kg)
    (out * w.to(dev)).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), atol=2e-5)
    for a, b, name in ((xg.grad, xr.grad, "dirs"), (kg.grad, kr.grad, "kappa")):
        scale = float(b.abs().max())
        err = float((a.cpu().double() - b).abs().max()) / scale
        assert err <= 2e-5, (name, err)
    # scalar kappa (the diffuse branch: IDE(n, diffuse_kappa_inv))
    xr2 = x.double().requires_grad_(True)
    (O.ide_encode(xr2, 0.64, deg, exact_tables=True) * w.double()).sum().backward()
    xg2 = x.to(dev).requires_grad_(True)
    (enc(xg2, 0.64) * w.to(dev)).sum().backward()
    assert float((xg2.grad.cpu().double() - xr2.grad).abs().max()) / float(xr2.grad.abs().max()) <= 2e-5


def test_graphed_train_step_equals_eager(dev):
    """train.GraphedTrainStep (whole forward + backward replayed from one CUDA graph) leaves the same loss and gradients as
    the eager step on the same inputs (perturb off: the march is then deterministic)."""
    from envidr_b200 import train
    from envidr_b200.render import RenderConfig
    fp, bf, ro, rd, gt_rgb, gt_mask, ri = _setup(32, 64, 4)
    cfg = RenderConfig(max_steps=512)
    bft = torch.from_numpy(bf).to(dev)
    field = train.TrainableField(fp.to(dev))
    N = ro.shape[0]
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    out = train.render_train(field, bft, ro.to(dev), rd.to(dev), cfg, force_all_rays=False, mean_count=8192, step_counter=counter,
                             perturb=False, r_images=ri.to(dev))
    loss = train.loss_epilogue(field, out, gt_rgb.to(dev), gt_mask.to(dev))
    loss.backward()
    eager = {n: p.grad.clone() for n, p in field.named_parameters() if p.grad is not None}
    loss_eager, n_eager = float(loss.detach()), int(counter[0])
    del out, loss                                              # drop the eager autograd graph (its AccumulateGrad nodes live on the default stream)
    g = train.GraphedTrainStep(field, bft, cfg, N, 8192, perturb=False)
    for _ in range(2):                                         # replay twice: static buffers must not accumulate
        lg = g(ro.to(dev), rd.to(dev), gt_rgb.to(dev), gt_mask.to(dev), ri.to(dev))
    torch.cuda.synchronize()
    assert abs(float(lg) - loss_eager) <= 1e-6 * max(1.0, abs(loss_eager))
    assert int(g.counter[0]) == n_eager
    for n, p in field.named_parameters():
        if n in eager:
            scale = float(eager[n].abs().max()) + 1e-20
            assert float((p.grad - eager[n]).abs().max()) <= 1e-4 * scale, n      # float atomics reorder between runs


@pytest.mark.parametrize("M,K,N,relu,bias", [(1000, 72, 256, True, True), (140000, 256, 256, True, True), (777, 256, 12, False, True),
                                             (513, 28, 64, True, True), (300, 64, 3, False, False), (128, 4, 64, True, True),
                                             (1, 24, 32, True, True)])
def test_linear_tc_forward_backward(dev, M, K, N, relu, bias):
    """csrc/linear_tc.cu (tcgen05, fp16 hi/lo split operands) against float64 torch: forward, data gradient (same kernel on W^T),
    weight / bias gradients.  Tolerance: 5e-6 of the output scale (fp32-level accuracy: 22-bit operands, fp32 accumulation)."""
    from envidr_b200.linear_tc import linear_tc
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    gy = torch.randn(M, N, generator=g)
    xd, Wd = x.double().requires_grad_(True), W.double().requires_grad_(True)
    bd = None if b is None else b.double().requires_grad_(True)
    yd = torch.nn.functional.linear(xd, Wd, bd)
    if relu:
        yd = torch.relu(yd)
    yd.backward(gy.double())
    xg, Wg = x.to(dev).requires_grad_(True), W.to(dev).requires_grad_(True)
    bg = None if b is None else b.to(dev).requires_grad_(True)
    y = linear_tc(xg, Wg, bg, relu)
    y.backward(gy.to(dev))
    scale = float(yd.abs().max())
    assert float((y.detach().cpu().double() - yd.detach()).abs().max()) <= 5e-6 * scale
    kink = (yd.detach().abs() < 1e-5 * scale) if relu else torch.zeros_like(yd, dtype=torch.bool)     # ReLU mask may flip at 0
    if not bool(kink.any()):
        assert float((xg.grad.cpu().double() - xd.grad).abs().max()) <= 5e-6 * float(xd.grad.abs().max() + 1e-30)
        assert float((Wg.grad.cpu().double() - Wd.grad).abs().max()) <= 2e-5 * float(Wd.grad.abs().max() + 1e-30)
        if bias:
            assert float((bg.grad.cpu().double() - bd.grad).abs().max()) <= 2e-5 * float(bd.grad.abs().max() + 1e-30)


@pytest.mark.parametrize("M,N,K", [(70_000, 256, 256), (4097, 64, 28), (300, 32, 24), (513, 12, 256), (2, 4, 4)])
def test_wgrad_tc_atomic_accumulate_and_fused_bias(dev, M, N, K):
    """envidr_pow2_scales (operand scales + bias column sums in one pass) and k_wgrad_tc's atomic accumulation (variant bit 1)
    against float64 torch; N = 12 takes the torch column-sum path.  Also against the partials + torch.sum formulation."""
    from envidr_b200 import linear_tc as L
    g = torch.Generator().manual_seed(M + N + K)
    gy = (torch.randn(M, N, generator=g) * 1e-7).to(dev)                     # loss-gradient magnitudes
    x = torch.randn(M, K, generator=g).to(dev)
    dW, db = L.wgrad_tc(gy, x, with_bias=True)
    dW_ref = gy.double().t() @ x.double()
    db_ref = gy.double().sum(0)
    assert float((dW.double() - dW_ref).abs().max()) <= 2e-5 * float(dW_ref.abs().max())
    assert float((db.double() - db_ref).abs().max()) <= 2e-5 * float(db_ref.abs().max()) + 1e-6 * float(gy.abs().max()) * M ** 0.5
    L.WGRAD_ATOMIC = False
    try:
        dW2, db2 = L.wgrad_tc(gy, x, with_bias=True)
    finally:
        L.WGRAD_ATOMIC = True
    assert float((dW - dW2).abs().max()) <= 2e-5 * float(dW_ref.abs().max())
    assert float((db - db2).abs().max()) <= 2e-5 * float(db_ref.abs().max()) + 1e-6 * float(gy.abs().max()) * M ** 0.5


def test_mm_family_double_backward_matches_float64(dev):
    """linear_tc.mm_nt / mm_nn / mm_tn: a 3-layer ReLU net whose input gradient (create_graph=True) enters the loss, as the normals do
    (renderer.py:182-198): first- and second-order gradients against float64 torch."""
    from envidr_b200.linear_tc import linear_tc_nd
    g = torch.Generator().manual_seed(4)
    M = 3000
    x0 = torch.randn(M, 32, generator=g)
    Ws = [torch.randn(64, 32, generator=g) / 32 ** 0.5, torch.randn(64, 64, generator=g) / 8, torch.randn(15, 64, generator=g) / 8]
    bs = [torch.randn(64, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1, torch.randn(15, generator=g) * 0.1]

    def run(dtype, device, lin):
        x = x0.to(device=device, dtype=dtype).requires_grad_(True)
        W = [w.to(device=device, dtype=dtype).requires_grad_(True) for w in Ws]
        b = [v.to(device=device, dtype=dtype).requires_grad_(True) for v in bs]
        h = x
        for i in range(3):
            h = lin(h, W[i], b[i])
            if i < 2:
                h = torch.relu(h)
        sdf = h[:, 0]
        n = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True)[0]
        loss = ((n.norm(dim=-1) - 1) ** 2).mean() + (h[:, 1:] ** 2).mean() + (n * x).sum(-1).mean()
        grads = torch.autograd.grad(loss, [x] + W + b)
        return float(loss), [t.detach().double().cpu() for t in grads], n.detach().double().cpu()

    loss_r, g_r, n_r = run(torch.float64, "cpu", torch.nn.functional.linear)
    loss_t, g_t, n_t = run(torch.float32, dev, linear_tc_nd)
    assert abs(loss_t - loss_r) <= 1e-5 * max(1.0, abs(loss_r))
    assert float((n_t - n_r).abs().max()) <= 2e-5 * float(n_r.abs().max())
    for a, b in zip(g_t, g_r):
        assert float((a - b).abs().max()) <= 2e-4 * float(b.abs().max()) + 1e-9
